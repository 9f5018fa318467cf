// Launch code of one iteration-kernel family (see matcher.hpp).
#include "matcher.hpp"
#include "kernels.cuh"

namespace hsm {

namespace {

template <int LPP, int VPL, int NT>
int launch_fused(hsm_matcher *m, bool store_s)
{
    const Geo g = geometry(m);
    constexpr int TX = NT / LPP - 2;
    const int strips = (m->W + TX - 1) / TX;
    IterArgs a = iter_args(m, strips, store_s, ((NT == 256) ? (VPL == 1 ? 3 : 2) : (VPL == 1 ? 2 : 1)));
    const int bands = (m->ohi - m->olo + 1 + a.band_rows - 1) / a.band_rows;
    dim3 grid(strips, bands, m->launch_frames);
    const bool dt_load = (m->data_term == 0);
    const bool full = (m->D == 4 * LPP * VPL);
    if (full) {
        if (dt_load) k_iter_fused<LPP, VPL, NT, true, true><<<grid, NT, 0, m->launch_stream>>>(g, a);
        else k_iter_fused<LPP, VPL, NT, false, true><<<grid, NT, 0, m->launch_stream>>>(g, a);
    } else {
        if (dt_load) k_iter_fused<LPP, VPL, NT, true, false><<<grid, NT, 0, m->launch_stream>>>(g, a);
        else k_iter_fused<LPP, VPL, NT, false, false><<<grid, NT, 0, m->launch_stream>>>(g, a);
    }
    HSM_TRY(check_launch(m, "k_iter_fused"));
    m->stats.iteration_launches++;
    m->cur ^= 1;
    return HSM_OK;
}

}  // namespace

int launch_fused_any(hsm_matcher *m, bool store_s)
{
    HSM_DISPATCH(m->cfg, { return launch_fused<LPP, VPL, NT>(m, store_s); });
    return fail(HSM_ERR_ARG, "no lane configuration");
}

}  // namespace hsm
