// The k_iter_tmast kernels with the peer-to-peer halo push compiled in (row-band mode, hsm_peer_attach).
#define HSM_PUSH true
#define HSM_TMAST_ENTRY launch_tmast_push_any
#include "launch_tmast.cu"
