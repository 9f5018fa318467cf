// The k_iter_tma kernels with the peer-to-peer halo push compiled in (row-band mode, hsm_peer_attach).
#define HSM_PUSH true
#define HSM_TMA_ENTRY launch_tma_push_any
#include "launch_tma.cu"
