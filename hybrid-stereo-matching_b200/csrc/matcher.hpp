// Internal header shared by the translation units of libhsm_b200 (not part of the C ABI).
//
// The library is split so that the template-heavy iteration kernels compile in parallel:
//   hsm_api.cu       C ABI, state management, small kernels (cost volume, sweeps, readout, ...)
//   launch_fused.cu  k_iter_fused   (algorithm 1)      launch_tma.cu    k_iter_tma   (algorithm 2)
//   launch_tma2.cu   k_iter2_tma    (algorithm 3)      launch_ws.cu     k_iter_ws    (algorithm 4)
//   launch_tmast.cu  k_iter_tmast   (algorithm 5, default)
#pragma once

#include "../../include/hsm.h"

#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "kernels.cuh"

namespace hsm {

int fail(int code, const char *fmt, ...);       // records the thread-local error message, returns `code`

#define CUDA_TRY(expr)                                                                            \
    do {                                                                                          \
        cudaError_t err__ = (expr);                                                               \
        if (err__ != cudaSuccess)                                                                 \
            return hsm::fail(HSM_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__), \
                             __FILE__, __LINE__);                                                 \
    } while (0)

#define HSM_TRY(expr)                    \
    do {                                 \
        int rc__ = (expr);               \
        if (rc__ != HSM_OK) return rc__; \
    } while (0)

// Lane configuration for a disparity range: LPP lanes per pixel, VPL float4 per lane,
// NT threads per CTA of the fused kernels.
struct LaneCfg {
    int lpp, vpl, nt;
};

// Run `body` with LPP, VPL, NT as compile-time constants.
#define HSM_DISPATCH_CASE(l, v, n, ...) \
    if ((cfg__).lpp == l && (cfg__).vpl == v && (cfg__).nt == n) { constexpr int LPP = l, VPL = v, NT = n; __VA_ARGS__ }
#define HSM_DISPATCH(cfg, ...)                                  \
    do {                                                        \
        const hsm::LaneCfg cfg__ = (cfg);                       \
        HSM_DISPATCH_CASE(4, 1, 256, __VA_ARGS__)               \
        else HSM_DISPATCH_CASE(8, 1, 256, __VA_ARGS__)          \
        else HSM_DISPATCH_CASE(16, 1, 512, __VA_ARGS__)         \
        else HSM_DISPATCH_CASE(32, 1, 512, __VA_ARGS__)         \
        else HSM_DISPATCH_CASE(32, 2, 512, __VA_ARGS__)         \
        else HSM_DISPATCH_CASE(32, 4, 256, __VA_ARGS__)         \
        else HSM_DISPATCH_CASE(4, 2, 256, __VA_ARGS__)          \
        else HSM_DISPATCH_CASE(8, 2, 256, __VA_ARGS__)          \
        else HSM_DISPATCH_CASE(16, 2, 512, __VA_ARGS__)         \
    } while (0)

// The default kernel family (k_iter_tmast / k_iter_flow) also has lane layouts whose lanes-per-pixel count is
// not a power of two (PixelMap in kernels.cuh); the other families and the plane kernels only have the list above.
// The list is compiled in two halves (launch_tmast.cu / launch_tmast_b.cu etc.: HSM_ST_HALF = 0 / 1) so that the
// ~100 instantiations build in parallel; DT_LOAD (the T28 option) only exists for the first nine entries.
#define HSM_DISPATCH_ST_A(cfg, ...)                             \
    do {                                                        \
        const hsm::LaneCfg cfg__ = (cfg);                       \
        HSM_DISPATCH_CASE(4, 1, 256, __VA_ARGS__)               \
        else HSM_DISPATCH_CASE(8, 1, 256, __VA_ARGS__)          \
        else HSM_DISPATCH_CASE(16, 1, 512, __VA_ARGS__)         \
        else HSM_DISPATCH_CASE(32, 1, 512, __VA_ARGS__)         \
        else HSM_DISPATCH_CASE(32, 2, 512, __VA_ARGS__)         \
        else HSM_DISPATCH_CASE(32, 4, 256, __VA_ARGS__)         \
    } while (0)
#define HSM_DISPATCH_ST_B(cfg, ...)                             \
    do {                                                        \
        const hsm::LaneCfg cfg__ = (cfg);                       \
        HSM_DISPATCH_CASE(4, 2, 256, __VA_ARGS__)               \
        else HSM_DISPATCH_CASE(8, 2, 256, __VA_ARGS__)          \
        else HSM_DISPATCH_CASE(16, 2, 512, __VA_ARGS__)         \
        else HSM_DISPATCH_CASE(5, 1, 256, __VA_ARGS__)          \
        else HSM_DISPATCH_CASE(6, 1, 256, __VA_ARGS__)          \
        else HSM_DISPATCH_CASE(5, 2, 256, __VA_ARGS__)          \
        else HSM_DISPATCH_CASE(6, 2, 256, __VA_ARGS__)          \
        else HSM_DISPATCH_CASE(5, 1, 224, __VA_ARGS__)          \
        else HSM_DISPATCH_CASE(6, 1, 288, __VA_ARGS__)          \
        else HSM_DISPATCH_CASE(8, 1, 320, __VA_ARGS__)          \
    } while (0)
inline bool st_cfg_in_first_half(const hsm::LaneCfg &c) { return c.vpl == 1 ? (c.lpp == 4 || c.lpp == 16 || c.lpp == 32 || (c.lpp == 8 && c.nt == 256)) : (c.lpp == 32); }
// entries that also have the T28 (data plane) variants: power-of-two lanes, 8- or 16-warp CTAs
constexpr bool st_cfg_has_dt_load(int lpp, int nt) { return (lpp & (lpp - 1)) == 0 && (nt == 256 || nt == 512); }

// Resident CTAs per SM the default kernel family is compiled for (register budget) and sized for (shared memory)
constexpr int tmast_ctas_per_sm(int nt, int vpl) { return nt <= 320 ? (vpl == 1 ? 3 : (vpl == 2 ? 2 : 1)) : (vpl == 1 ? 2 : 1); }

struct ImageMaps {
    CUtensorMap L, R, P;
};

// tensor maps of the default kernel family for its own lane layout (boxes of its PX / TX pixels)
struct StMaps {
    CUtensorMap tmW[2], tmN[2], tmE[2], tmD, tsW[2], tsN[2], tsE[2];
    ImageMaps tmI;
};

}  // namespace hsm

struct hsm_matcher {
    int H = 0, W = 0, Wp = 0, D = 0, Dp = 0, cap = 0, device = 0;
    hsm::LaneCfg cfg{};          // lane layout of the plane kernels and the alternative iteration kernels (power of two)
    hsm::LaneCfg cfg5{};         // lane layout of the default kernel family (may be 5 or 6 lanes per pixel)
    bool ready = false;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    void *arena = nullptr;
    size_t arena_bytes = 0;
    float *S = nullptr, *Wm[2] = {nullptr, nullptr}, *Nm[2] = {nullptr, nullptr}, *Em[2] = {nullptr, nullptr};
    float *Dt = nullptr;
    float *L = nullptr, *R = nullptr;
    int *P = nullptr;
    char *stage[3] = {nullptr, nullptr, nullptr};
    char *host_stage = nullptr;                  // page-locked staging of the small-input path of hsm_lbp
    size_t host_stage_bytes = 0;
    long long *labels = nullptr;
    int cur = 0;          // which of the double buffers holds the current W, N, E
    int frames = 0;       // frames of the current batch
    bool have_fields = false, have_prior = false, dt_valid = false, s_valid = true;
    hsm::Blend blend{hsm::kPriorNone, 0.0f, 0.0};
    // options
    int algorithm = 5, band_rows = 0, data_term = 1, stages = 0, lane_config = 0;
    // TMA descriptors of the double-buffered planes (algorithm 2)
    CUtensorMap tmW[2], tmN[2], tmE[2], tmD;
    CUtensorMap tsW[2], tsN[2], tsE[2];          // store maps (box = TX output pixels)
    hsm::ImageMaps tmI;
    hsm::StMaps st;
    bool maps_valid = false;
    // hsm_init_fields(reinit) does not clear the message planes: it records that they are LOGICALLY zero.  The
    // first iteration then reads no plane at all (zero-filled by the TMA unit), and whoever needs real zeros in
    // memory (read-backs, the alternative kernels, n_iter = 0) materialises them first.
    bool zero_messages = false;
    // the finalising launch of the default kernel also wrote the argmin labels of the batch into `labels`
    // (fused readout): hsm_readout only has to copy them
    bool labels_valid = false;
    bool labels_fused = false;                   // set by the launch code when the last launch carried the readout
    int vlo = 0, vhi = 0, olo = 0, ohi = 0;
    // row-band mode with peer-to-peer halo push (hsm_peer_*): side 0 = band above, 1 = band below
    struct Peer {
        bool attached = false;
        void *ipc_base = nullptr;                // mapping opened with cudaIpcOpenMemHandle (null if same process)
        float *Wm[2] = {nullptr, nullptr}, *Nm[2] = {nullptr, nullptr}, *Em[2] = {nullptr, nullptr};
        unsigned *mailbox = nullptr;             // the slot of the neighbour's mailbox this matcher writes
        unsigned ghost_off = 0;                  // element offset of the neighbour's ghost row inside its planes
        unsigned *ghost_flags = nullptr;         // dataflow kernel: per-strip progress counters in the neighbour's
                                                 // mailbox region that this matcher's boundary tiles publish to
        CUtensorMap psW[2], psN[2], psE[2];      // store maps over the neighbour's ghost row, per plane generation
    } peer[2];
    // mailbox region in the arena (visible to the neighbours through the IPC mapping):
    //   [0], [1]                     whole-iteration counters written by the band above / below (stream-ordered)
    //   [kGhostBase + side * kGhostStride + strip]   per-strip progress of the neighbour's boundary tiles
    static constexpr int kGhostBase = 64, kGhostStride = 1024, kMailboxWords = kGhostBase + 2 * kGhostStride;
    unsigned *mailbox = nullptr;
    unsigned p2p_iterations = 0;
    // Frame groups: an iteration over >= 2 frames is launched as two kernels (first / second half of the
    // frames) on two streams, so that one group's tail (SMs idling while the last CTAs finish) is filled by
    // the other group's CTAs.  Set per launch by hsm_iterate; read by the launch_* functions.
    int frame_groups = 0;                        // HSM_OPT_FRAME_GROUPS: 0 = automatic, 1 = one launch per iteration
    int launch_frame0 = 0, launch_frames = 0;
    cudaStream_t launch_stream = nullptr, aux_stream = nullptr;
    unsigned *flow_state = nullptr;              // persistent dataflow kernel: task counter + per-tile progress
    size_t flow_bytes = 0;
    int sm_count = 148;
    bool launch_pdl = false;                     // programmatic dependent launch (one kernel per iteration only)
    cudaEvent_t fork_event = nullptr, join_event = nullptr;                 // iterations completed since hsm_peer_reset
    // statistics
    hsm_stats stats{};
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending, pool;
};


namespace hsm {

// shared host helpers (defined in hsm_api.cu)
Geo geometry(const hsm_matcher *m);
// Launch configuration of an iteration kernel on m->launch_stream; with m->launch_pdl the launch may
// overlap its prologue with the previous kernel's tail (see pdl_wait in iter_tma.cuh).
inline cudaLaunchConfig_t launch_config(const hsm_matcher *m, dim3 grid, int threads, size_t smem,
                                        cudaLaunchAttribute *attr)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = m->launch_stream;
    attr->id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr->val.programmaticStreamSerializationAllowed = m->launch_pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cfg;
}
int choose_band_rows(const hsm_matcher *m, int strips, int ctas_per_sm);
int check_launch(hsm_matcher *m, const char *what);
int ensure_maps(hsm_matcher *m);
bool tma_supported(const hsm_matcher *m);
IterArgs iter_args(hsm_matcher *m, int strips, bool store_s, int ctas_per_sm);   // everything but the tensor maps

// one BP iteration (or two, for algorithm 3) with the given kernel family; each lives in its own
// translation unit.  `*launched` = false when the configuration cannot use that kernel.
int launch_fused_any(hsm_matcher *m, bool store_s);
int launch_tma_any(hsm_matcher *m, bool store_s);
int launch_tma_push_any(hsm_matcher *m, bool store_s);                 // same, halo push compiled in
int launch_tma2_any(hsm_matcher *m, bool store_s, bool *launched);
int launch_ws_any(hsm_matcher *m, bool store_s);
int launch_tmast_any(hsm_matcher *m, bool store_s, bool *launched);
int launch_tmast_b(hsm_matcher *m, bool store_s, bool *launched);          // second half of the lane-layout list
int launch_tmast_push_b(hsm_matcher *m, bool store_s, bool *launched);
int launch_flow_b(hsm_matcher *m, int n_iter, bool finalize, bool *launched);
int launch_flow_push_b(hsm_matcher *m, int n_iter, bool finalize, bool *launched);
// all n_iter iterations in one persistent launch (algorithm 6); *launched = false -> caller falls back
int launch_flow_any(hsm_matcher *m, int n_iter, bool finalize, bool *launched);
int launch_flow_push_any(hsm_matcher *m, int n_iter, bool finalize, bool *launched);   // with neighbours attached
int choose_flow_band_rows(const hsm_matcher *m, int strips, int ctas_per_sm);
int launch_tmast_push_any(hsm_matcher *m, bool store_s, bool *launched);

}  // namespace hsm
