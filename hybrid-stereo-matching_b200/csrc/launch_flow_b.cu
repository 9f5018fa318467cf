// Second half of the lane-layout list of the persistent dataflow kernels (see matcher.hpp).
#define HSM_ST_HALF 1
#include "launch_flow.cu"
