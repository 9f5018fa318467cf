// Second half of the lane-layout list of the k_iter_tmast kernels with the halo push compiled in.
#define HSM_PUSH true
#define HSM_TMAST_ENTRY launch_tmast_push_any
#define HSM_TMAST_ENTRY_B launch_tmast_push_b
#define HSM_ST_HALF 1
#include "launch_tmast.cu"
