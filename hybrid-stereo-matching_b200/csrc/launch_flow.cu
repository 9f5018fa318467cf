// Launch code of the persistent dataflow kernel (see iter_flow.cuh, matcher.hpp).
// Compiled twice: as is, and through launch_flow_push.cu with HSM_PUSH = true for row bands whose neighbours
// live on other GPUs (boundary rows and progress counters go through peer-mapped memory).
#ifndef HSM_PUSH
#define HSM_PUSH false
#define HSM_FLOW_ENTRY launch_flow_any
#define HSM_FLOW_ENTRY_B launch_flow_b
#endif
#ifndef HSM_ST_HALF
#define HSM_ST_HALF 0
#endif
#include "matcher.hpp"
#include "iter_flow.cuh"

namespace hsm {

namespace {

template <int LPP, int VPL, int NT>
int launch_flow(hsm_matcher *m, int n_iter, bool finalize, bool *launched)
{
    *launched = false;
    using PM = PixelMap<LPP, NT>;
    constexpr int TX = PM::PX - 2, SLOTS = PM::LANES + 2 * LPP;
    const bool dt_load = (m->data_term == 0);
    // (layouts without the T28 variants: not launched, the caller goes on with one launch per iteration)
    if (dt_load && !st_cfg_has_dt_load(LPP, NT)) return HSM_OK;
    const size_t stage_bytes = StageLayout<LPP, VPL, NT>::stage_bytes(m->Dp, dt_load);
    const size_t ost_bytes = ((size_t)TX * m->Dp * 4 + 127) & ~(size_t)127;
    const size_t fixed_bytes = 6 * ost_bytes + (size_t)2 * VPL * SLOTS * 16 + 128;
    const int target_ctas = tmast_ctas_per_sm(NT, VPL);
    int stages = m->stages;
    if (stages <= 0) {
        const size_t budget = (size_t)227 * 1024 / target_ctas - 1024;
        stages = 3;
        while (stages > 2 && fixed_bytes + stages * stage_bytes > budget) --stages;
    }
    const size_t smem = fixed_bytes + (size_t)stages * stage_bytes;
    if (smem > (size_t)227 * 1024) return HSM_OK;
    HSM_TRY(ensure_maps(m));
    const Geo g = geometry(m);
    const int strips = (m->W + TX - 1) / TX;
    m->launch_frame0 = 0; m->launch_frames = m->frames; m->launch_stream = m->stream;
    IterArgs a = iter_args(m, strips, false, target_ctas);
    a.band_rows = choose_flow_band_rows(m, strips, target_ctas);
    if (a.l2_hints < 0) a.l2_hints = 0;           // (measured: the hints cost the persistent kernel 1-3 %)
    const int rows = m->ohi - m->olo + 1;
    const int bands = (rows + a.band_rows - 1) / a.band_rows;
    const size_t tiles = (size_t)m->frames * bands * strips;
    if (tiles * (size_t)n_iter >= 0x7fffffffull) return HSM_OK;
    // counter + per-tile progress, zeroed for every launch
    const size_t need = (tiles + 1) * sizeof(unsigned);
    if (m->flow_bytes < need) {
        if (m->flow_state) CUDA_TRY(cudaFree(m->flow_state));
        m->flow_state = nullptr; m->flow_bytes = 0;
        CUDA_TRY(cudaMalloc(&m->flow_state, need));
        m->flow_bytes = need;
    }
    CUDA_TRY(cudaMemsetAsync(m->flow_state, 0, need, m->stream));
    FlowArgs f;
    f.counter = m->flow_state; f.done = m->flow_state + 1;
    f.n_iter = n_iter; f.frames = m->frames; f.bands = bands; f.strips = strips;
    f.cur = m->cur; f.finalize = finalize ? 1 : 0;
    f.base = m->p2p_iterations;
    f.mail_in = m->mailbox;
    if (HSM_PUSH && strips > hsm_matcher::kGhostStride) return HSM_OK;
    for (int side = 0; side < 2; ++side) {
        const hsm_matcher::Peer &p = m->peer[side];
        const bool on = HSM_PUSH && p.attached;
        f.ghost_in[side] = on ? m->mailbox + hsm_matcher::kGhostBase + side * hsm_matcher::kGhostStride : nullptr;
        f.ghost_out[side] = on ? p.ghost_flags : nullptr;
        for (int gen = 0; gen < 2; ++gen) {
            f.pushW[side][gen] = on ? p.Wm[gen] : nullptr;
            f.pushN[side][gen] = on ? p.Nm[gen] : nullptr;
            f.pushE[side][gen] = on ? p.Em[gen] : nullptr;
        }
        f.push_off[side] = p.ghost_off;
    }
    FlowMaps maps;
    for (int i = 0; i < 2; ++i) {
        maps.tmW[i] = m->st.tmW[i]; maps.tmN[i] = m->st.tmN[i]; maps.tmE[i] = m->st.tmE[i];
        maps.tsW[i] = m->st.tsW[i]; maps.tsN[i] = m->st.tsN[i]; maps.tsE[i] = m->st.tsE[i];
    }
    memset(maps.push, 0, sizeof maps.push);
    for (int side = 0; side < 2; ++side)
        if (HSM_PUSH && m->peer[side].attached)
            for (int gen = 0; gen < 2; ++gen) {
                maps.push[gen].W[side] = m->peer[side].psW[gen]; maps.push[gen].N[side] = m->peer[side].psN[gen];
                maps.push[gen].E[side] = m->peer[side].psE[gen];
            }
    maps.tmD = m->st.tmD; maps.tmL = m->st.tmI.L; maps.tmR = m->st.tmI.R; maps.tmP = m->st.tmI.P;
    const bool full = (m->D == 4 * LPP * VPL);
#define HSM_LAUNCH_FLOW(T, F)                                                                                \
    do {                                                                                                     \
        auto kernel = k_iter_flow<LPP, VPL, NT, T, F, HSM_PUSH>;                                                       \
        CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));      \
        CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));         \
        int per_sm = 0;                                                                                      \
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, NT, smem));                  \
        const size_t resident = (size_t)std::max(per_sm, 1) * (size_t)m->sm_count;                           \
        const unsigned grid = (unsigned)std::min(resident, tiles * (size_t)n_iter);                          \
        if (getenv("HSM_FLOW_DEBUG"))                                                                        \
            fprintf(stderr, "flow: grid %u (%d per SM), %d bands of %d rows x %d strips x %d frames, smem %zu, stages %d\n", \
                    grid, per_sm, bands, a.band_rows, strips, m->frames, smem, stages);                      \
        kernel<<<grid, NT, smem, m->stream>>>(maps, g, a, f, stages);                                        \
    } while (0)
    if constexpr (st_cfg_has_dt_load(LPP, NT)) {
        if (full) { if (dt_load) HSM_LAUNCH_FLOW(true, true); else HSM_LAUNCH_FLOW(false, true); }
        else { if (dt_load) HSM_LAUNCH_FLOW(true, false); else HSM_LAUNCH_FLOW(false, false); }
    } else {
        if (full) HSM_LAUNCH_FLOW(false, true); else HSM_LAUNCH_FLOW(false, false);
    }
#undef HSM_LAUNCH_FLOW
    HSM_TRY(check_launch(m, "k_iter_flow"));
    m->stats.iteration_launches++;
    m->cur = (m->cur + n_iter) & 1;
    *launched = true;
    return HSM_OK;
}

}  // namespace

#if HSM_ST_HALF == 0
int HSM_FLOW_ENTRY(hsm_matcher *m, int n_iter, bool finalize, bool *launched)
{
    if (!st_cfg_in_first_half(m->cfg5)) return HSM_FLOW_ENTRY_B(m, n_iter, finalize, launched);
    HSM_DISPATCH_ST_A(m->cfg5, { return launch_flow<LPP, VPL, NT>(m, n_iter, finalize, launched); });
    return fail(HSM_ERR_ARG, "no lane configuration");
}
#else
int HSM_FLOW_ENTRY_B(hsm_matcher *m, int n_iter, bool finalize, bool *launched)
{
    HSM_DISPATCH_ST_B(m->cfg5, { return launch_flow<LPP, VPL, NT>(m, n_iter, finalize, launched); });
    return fail(HSM_ERR_ARG, "no lane configuration");
}
#endif

}  // namespace hsm
