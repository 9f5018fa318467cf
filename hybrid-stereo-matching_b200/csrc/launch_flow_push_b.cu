// Second half of the lane-layout list of the persistent dataflow kernels with the halo push compiled in.
#define HSM_PUSH true
#define HSM_FLOW_ENTRY launch_flow_push_any
#define HSM_FLOW_ENTRY_B launch_flow_push_b
#define HSM_ST_HALF 1
#include "launch_flow.cu"
