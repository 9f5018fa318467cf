// Launch code of one iteration-kernel family (see matcher.hpp).
#include "matcher.hpp"
#include "iter2_tma.cuh"

namespace hsm {

namespace {

template <int LPP, int VPL, int NT>
int launch_tma2(hsm_matcher *m, bool store_s, bool *launched)
{
    *launched = false;
    constexpr int PX = NT / LPP, TX = PX - 4, SLOTS = NT + LPP;
    if (TX < 4) return HSM_OK;
    const size_t stage_bytes = StageLayout<LPP, VPL, NT>::stage_bytes(m->Dp, false);
    const size_t fixed_bytes = Ring2Layout<LPP, VPL, NT>::ring_bytes(m->Dp) + (size_t)2 * VPL * SLOTS * 16 + 128;
    int stages = m->stages > 0 ? m->stages : 3;
    while (stages > 2 && fixed_bytes + stages * stage_bytes > (size_t)113 * 1024) --stages;
    const size_t smem = fixed_bytes + (size_t)stages * stage_bytes;
    if (smem > (size_t)227 * 1024) return HSM_OK;
    HSM_TRY(ensure_maps(m));
    const Geo g = geometry(m);
    const int strips = (m->W + TX - 1) / TX;
    IterArgs a = iter_args(m, strips, store_s, ((NT == 256) ? (VPL == 1 ? 3 : 2) : (VPL == 1 ? 2 : 1)));
    const int bands = (m->ohi - m->olo + 1 + a.band_rows - 1) / a.band_rows;
    dim3 grid(strips, bands, m->launch_frames);
    const bool full = (m->D == 4 * LPP * VPL);
    const CUtensorMap &tw = m->tmW[m->cur], &tn = m->tmN[m->cur], &te = m->tmE[m->cur];
#define HSM_LAUNCH_TMA2(F)                                                                                   \
    do {                                                                                                     \
        auto kernel = k_iter2_tma<LPP, VPL, NT, F>;                                                          \
        CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));      \
        kernel<<<grid, NT, smem, m->launch_stream>>>(tw, tn, te, m->tmI.L, m->tmI.R, m->tmI.P, g, a, stages);       \
    } while (0)
    if (full) HSM_LAUNCH_TMA2(true); else HSM_LAUNCH_TMA2(false);
#undef HSM_LAUNCH_TMA2
    HSM_TRY(check_launch(m, "k_iter2_tma"));
    m->stats.iteration_launches++;
    m->cur ^= 1;
    *launched = true;
    return HSM_OK;
}

// Warp-specialised variant (algorithm 4): NT consumer threads + one producer warp.
}  // namespace

int launch_tma2_any(hsm_matcher *m, bool store_s, bool *launched)
{
    HSM_DISPATCH(m->cfg, { return launch_tma2<LPP, VPL, NT>(m, store_s, launched); });
    return fail(HSM_ERR_ARG, "no lane configuration");
}

}  // namespace hsm
