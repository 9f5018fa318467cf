// The persistent dataflow kernels with the peer-to-peer halo push and cross-GPU progress counters compiled in
// (row-band mode, hsm_peer_attach).
#define HSM_PUSH true
#define HSM_FLOW_ENTRY launch_flow_push_any
#define HSM_FLOW_ENTRY_B launch_flow_push_b
#include "launch_flow.cu"
