// Launch code of one iteration-kernel family (see matcher.hpp).
// Compiled twice: as is (no peers attached: the halo-push code is compiled out of the kernels), and
// through launch_tma_push.cu with HSM_PUSH = true for row bands with peer-to-peer halo push.
#ifndef HSM_PUSH
#define HSM_PUSH false
#define HSM_TMA_ENTRY launch_tma_any
#endif
#include "matcher.hpp"
#include "iter_tma.cuh"

namespace hsm {

namespace {

template <int LPP, int VPL, int NT>
int launch_tma(hsm_matcher *m, bool store_s)
{
    HSM_TRY(ensure_maps(m));
    const Geo g = geometry(m);
    constexpr int PX = NT / LPP, TX = PX - 2, SLOTS = NT + 2 * LPP;
    const int strips = (m->W + TX - 1) / TX;
    IterArgs a = iter_args(m, strips, store_s, ((NT == 256) ? (VPL == 1 ? 3 : 2) : (VPL == 1 ? 2 : 1)));
    const int bands = (m->ohi - m->olo + 1 + a.band_rows - 1) / a.band_rows;
    dim3 grid(strips, bands, m->launch_frames);
    const bool dt_load = (m->data_term == 0);
    const bool full = (m->D == 4 * LPP * VPL);
    (void)PX;
    const size_t stage_bytes = StageLayout<LPP, VPL, NT>::stage_bytes(m->Dp, dt_load);
    const size_t fixed_bytes = (size_t)2 * VPL * SLOTS * 16 + 128;   // exchange slots + mbarriers
    int stages = m->stages;
    if (stages <= 0) {      // as many stages (2..3) as fit the per-CTA share of shared memory
        const int target_ctas = (NT == 256) ? (VPL == 1 ? 4 : 2) : (VPL == 1 ? 2 : 1);
        const size_t budget = (size_t)227 * 1024 / target_ctas - 1024;
        stages = 3;
        while (stages > 2 && fixed_bytes + stages * stage_bytes > budget) --stages;
    }
    const size_t smem = fixed_bytes + (size_t)stages * stage_bytes;
    if (smem > (size_t)227 * 1024) return fail(HSM_ERR_ARG, "fused TMA kernel needs %zu bytes of shared memory", smem);
    const CUtensorMap &tw = m->tmW[m->cur], &tn = m->tmN[m->cur], &te = m->tmE[m->cur], &td = m->tmD;
    cudaLaunchAttribute pdl_attr;
    cudaLaunchConfig_t cfg = launch_config(m, grid, NT, smem, &pdl_attr);
#define HSM_LAUNCH_TMA(T, F)                                                                                 \
    do {                                                                                                     \
        auto kernel = k_iter_tma<LPP, VPL, NT, T, F, HSM_PUSH>;                                                        \
        CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));      \
        CUDA_TRY(cudaLaunchKernelEx(&cfg, kernel, tw, tn, te, td, m->tmI.L, m->tmI.R, m->tmI.P, g, a, stages));   \
    } while (0)
    if (full) { if (dt_load) HSM_LAUNCH_TMA(true, true); else HSM_LAUNCH_TMA(false, true); }
    else { if (dt_load) HSM_LAUNCH_TMA(true, false); else HSM_LAUNCH_TMA(false, false); }
#undef HSM_LAUNCH_TMA
    HSM_TRY(check_launch(m, "k_iter_tma"));
    m->stats.iteration_launches++;
    m->cur ^= 1;
    return HSM_OK;
}

// Two iterations per launch (algorithm 3).  Returns HSM_OK and sets *launched = false when the
// configuration cannot use it (caller falls back to one iteration per launch).
}  // namespace

int HSM_TMA_ENTRY(hsm_matcher *m, bool store_s)
{
    HSM_DISPATCH(m->cfg, { return launch_tma<LPP, VPL, NT>(m, store_s); });
    return fail(HSM_ERR_ARG, "no lane configuration");
}

}  // namespace hsm
