// Launch code of one iteration-kernel family (see matcher.hpp).
#include "matcher.hpp"
#include "iter_ws.cuh"

namespace hsm {

namespace {

template <int LPP, int VPL, int NT>
int launch_ws(hsm_matcher *m, bool store_s)
{
    HSM_TRY(ensure_maps(m));
    const Geo g = geometry(m);
    constexpr int PX = NT / LPP, TX = PX - 2, NW = NT / 32, XR = 4;
    const int strips = (m->W + TX - 1) / TX;
    IterArgs a = iter_args(m, strips, store_s, ((NT == 256) ? (VPL == 1 ? 3 : 2) : (VPL == 1 ? 2 : 1)));
    const int bands = (m->ohi - m->olo + 1 + a.band_rows - 1) / a.band_rows;
    dim3 grid(strips, bands, m->launch_frames);
    const bool dt_load = (m->data_term == 0);
    const bool full = (m->D == 4 * LPP * VPL);
    const size_t stage_bytes = StageLayout<LPP, VPL, NT>::stage_bytes(m->Dp, dt_load);
    const size_t fixed_bytes = (size_t)NW * XR * VPL * LPP * 16 + (size_t)(2 * 3 + NW * XR) * 8 + 128;
    int stages = m->stages > 0 ? std::min(m->stages, 3) : 3;      // the exchange ring holds stages + 1 rows
    {
        const int target_ctas = (NT == 256) ? (VPL == 1 ? 4 : 2) : (VPL == 1 ? 2 : 1);
        const size_t budget = (size_t)227 * 1024 / target_ctas - 1024;
        while (stages > 2 && fixed_bytes + stages * stage_bytes > budget) --stages;
    }
    const size_t smem = fixed_bytes + (size_t)stages * stage_bytes;
    if (smem > (size_t)227 * 1024) return fail(HSM_ERR_ARG, "warp-specialised kernel needs %zu bytes of shared memory", smem);
    const CUtensorMap &tw = m->tmW[m->cur], &tn = m->tmN[m->cur], &te = m->tmE[m->cur], &td = m->tmD;
#define HSM_LAUNCH_WS(T, F)                                                                                  \
    do {                                                                                                     \
        auto kernel = k_iter_ws<LPP, VPL, NT, T, F>;                                                         \
        CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));      \
        kernel<<<grid, NT + 32, smem, m->launch_stream>>>(tw, tn, te, td, m->tmI.L, m->tmI.R, m->tmI.P, g, a, stages); \
    } while (0)
    if (full) { if (dt_load) HSM_LAUNCH_WS(true, true); else HSM_LAUNCH_WS(false, true); }
    else { if (dt_load) HSM_LAUNCH_WS(true, false); else HSM_LAUNCH_WS(false, false); }
#undef HSM_LAUNCH_WS
    HSM_TRY(check_launch(m, "k_iter_ws"));
    m->stats.iteration_launches++;
    m->cur ^= 1;
    return HSM_OK;
}

// TMA-load + TMA-store variant (algorithm 5).  *launched = false if its shared memory does not fit.
}  // namespace

int launch_ws_any(hsm_matcher *m, bool store_s)
{
    HSM_DISPATCH(m->cfg, { return launch_ws<LPP, VPL, NT>(m, store_s); });
    return fail(HSM_ERR_ARG, "no lane configuration");
}

}  // namespace hsm
