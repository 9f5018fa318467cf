// Launch code of one iteration-kernel family (see matcher.hpp).
// Compiled twice: as is (no peers attached: the halo-push code is compiled out of the kernels), and
// through launch_tmast_push.cu with HSM_PUSH = true for row bands with peer-to-peer halo push.
#ifndef HSM_PUSH
#define HSM_PUSH false
#define HSM_TMAST_ENTRY launch_tmast_any
#define HSM_TMAST_ENTRY_B launch_tmast_b
#endif
#ifndef HSM_ST_HALF
#define HSM_ST_HALF 0
#endif
#include "matcher.hpp"
#include "iter_tma_st.cuh"

namespace hsm {

namespace {

template <int LPP, int VPL, int NT>
int launch_tmast(hsm_matcher *m, bool store_s, bool *launched)
{
    *launched = false;
    using PM = PixelMap<LPP, NT>;
    constexpr int TX = PM::PX - 2, SLOTS = PM::LANES + 2 * LPP;
    const bool dt_load = (m->data_term == 0);
    // (layouts without the T28 variants: not launched, the caller falls back to the TMA-load kernel)
    if (dt_load && !st_cfg_has_dt_load(LPP, NT)) return HSM_OK;
    const size_t stage_bytes = StageLayout<LPP, VPL, NT>::stage_bytes(m->Dp, dt_load);
    const size_t ost_bytes = ((size_t)TX * m->Dp * 4 + 127) & ~(size_t)127;
    const size_t fixed_bytes = 6 * ost_bytes + (size_t)2 * VPL * SLOTS * 16 + 128;
    int stages = m->stages;
    if (stages <= 0) {
        const int target_ctas = tmast_ctas_per_sm(NT, VPL);
        const size_t budget = (size_t)227 * 1024 / target_ctas - 1024;
        stages = 3;
        while (stages > 2 && fixed_bytes + stages * stage_bytes > budget) --stages;
    }
    const size_t smem = fixed_bytes + (size_t)stages * stage_bytes;
    if (smem > (size_t)227 * 1024) return HSM_OK;
    HSM_TRY(ensure_maps(m));
    const Geo g = geometry(m);
    const int strips = (m->W + TX - 1) / TX;
    IterArgs a = iter_args(m, strips, store_s, (tmast_ctas_per_sm(NT, VPL)));
    // One launch per iteration: the first reader of the two rows a band shares with the band above asks L2 to keep
    // them, the second lets them go, and the new planes (read again a whole iteration later) are stored as streaming
    // data.  Worth +0.7 ... +2.2 % where the kernel is HBM-bound (240x180xD32 x 128: 0.968 -> 0.974 of peak, D = 20:
    // 0.886 -> 0.908, 640x480xD64 x 4: 0.960 -> 0.976); where it is issue-bound -- the layouts with idle chunk slots
    // -- the extra instructions of the issuing thread cost ~1 % instead (profiles/r2_probe_l2_hints.txt), so only
    // full layouts use the hints, and only their instantiations contain the code (tmast_task: L2_HINTS).  (Row bands
    // over several GPUs: not measured, off.)
    if (a.l2_hints < 0) a.l2_hints = (!HSM_PUSH && m->D == 4 * LPP * VPL && !dt_load) ? 5 : 0;
    const int bands = (m->ohi - m->olo + 1 + a.band_rows - 1) / a.band_rows;
    dim3 grid(strips, bands, m->launch_frames);
    a.labels = store_s ? m->labels : nullptr;            // the finalising launch also writes the argmin labels
    m->labels_fused = store_s;
    const bool full = (m->D == 4 * LPP * VPL);
    const int in = m->cur, out = m->cur ^ 1;
    PushMaps pmaps;
    memset(&pmaps, 0, sizeof pmaps);
    for (int side = 0; side < 2; ++side)
        if (HSM_PUSH && m->peer[side].attached) {
            pmaps.W[side] = m->peer[side].psW[out]; pmaps.N[side] = m->peer[side].psN[out];
            pmaps.E[side] = m->peer[side].psE[out];
        }
    cudaLaunchAttribute pdl_attr;
    cudaLaunchConfig_t cfg = launch_config(m, grid, NT, smem, &pdl_attr);
#define HSM_LAUNCH_TMAST(T, F)                                                                               \
    do {                                                                                                     \
        auto kernel = store_s ? k_iter_tmast<LPP, VPL, NT, T, F, HSM_PUSH, true>                             \
                              : k_iter_tmast<LPP, VPL, NT, T, F, HSM_PUSH, false>;                           \
        CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));      \
        CUDA_TRY(cudaLaunchKernelEx(&cfg, kernel, m->st.tmW[in], m->st.tmN[in], m->st.tmE[in], m->st.tmD,           \
                                    m->st.tmI.L, m->st.tmI.R, m->st.tmI.P, m->st.tsW[out], m->st.tsN[out],        \
                                    m->st.tsE[out], pmaps, g, a, stages));                                     \
    } while (0)
    if constexpr (st_cfg_has_dt_load(LPP, NT)) {
        if (full) { if (dt_load) HSM_LAUNCH_TMAST(true, true); else HSM_LAUNCH_TMAST(false, true); }
        else { if (dt_load) HSM_LAUNCH_TMAST(true, false); else HSM_LAUNCH_TMAST(false, false); }
    } else {
        if (full) HSM_LAUNCH_TMAST(false, true); else HSM_LAUNCH_TMAST(false, false);
    }
#undef HSM_LAUNCH_TMAST
    HSM_TRY(check_launch(m, "k_iter_tmast"));
    m->stats.iteration_launches++;
    m->cur ^= 1;
    *launched = true;
    return HSM_OK;
}

}  // namespace

#if HSM_ST_HALF == 0
int HSM_TMAST_ENTRY(hsm_matcher *m, bool store_s, bool *launched)
{
    if (!st_cfg_in_first_half(m->cfg5)) return HSM_TMAST_ENTRY_B(m, store_s, launched);
    HSM_DISPATCH_ST_A(m->cfg5, { return launch_tmast<LPP, VPL, NT>(m, store_s, launched); });
    return fail(HSM_ERR_ARG, "no lane configuration");
}
#else
int HSM_TMAST_ENTRY_B(hsm_matcher *m, bool store_s, bool *launched)
{
    HSM_DISPATCH_ST_B(m->cfg5, { return launch_tmast<LPP, VPL, NT>(m, store_s, launched); });
    return fail(HSM_ERR_ARG, "no lane configuration");
}
#endif

}  // namespace hsm
