// Second half of the lane-layout list of the k_iter_tmast kernels (see matcher.hpp: compiled in parallel).
#define HSM_ST_HALF 1
#include "launch_tmast.cu"
