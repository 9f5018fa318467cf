// Fused iteration kernel, warp-specialised variant (algorithm 4).
//
// Same arithmetic and row stream as k_iter_tma, but without any CTA-wide barrier in the steady
// state:
//   * one extra PRODUCER warp issues the TMA loads; consumer warps hand a stage back through an
//     `empty` mbarrier (one arrival per warp) as soon as they have copied their chunks to
//     registers, so the producer refills while the consumers are still computing;
//   * the only cross-thread dependency of a row, W'(y, x) from the right-hand neighbour pixel, is
//     resolved inside the warp with SHFL for all but the warp's last pixel; that one takes the value
//     of the NEXT warp's first pixel from a small shared-memory ring guarded by a per-warp
//     mbarrier.  A warp therefore waits only for its right-hand neighbour warp, never for the CTA.
// Skew between warps is bounded by the TMA ring (a warp cannot start row y + stages before every
// warp released row y), so an exchange ring of stages + 1 slots is never overrun.
#pragma once

#include "iter_tma.cuh"

namespace hsm {

#ifndef HSM_WS_MIN_BLOCKS
#define HSM_WS_MIN_BLOCKS 3
#endif

template <int LPP, int VPL, int NT, bool DT_LOAD, bool FULL>
__global__ void __launch_bounds__(NT + 32, (NT == 256 && VPL == 1) ? HSM_WS_MIN_BLOCKS : 1)
    k_iter_ws(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmN,
              const __grid_constant__ CUtensorMap tmE, const __grid_constant__ CUtensorMap tmD,
              const __grid_constant__ CUtensorMap tmL, const __grid_constant__ CUtensorMap tmR,
              const __grid_constant__ CUtensorMap tmP, Geo g, IterArgs a, int stages)
{
    using SL = StageLayout<LPP, VPL, NT>;
    constexpr int PX = NT / LPP;
    constexpr int TX = PX - 2;
    constexpr int NW = NT / 32;            // consumer warps
    constexpr int PPW = 32 / LPP;          // pixels per warp
    constexpr int XR = 4;                  // exchange ring slots (>= stages + 1)
    constexpr int NPL = DT_LOAD ? 4 : 3;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const bool STORE_S = a.store_s != 0;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;

    const unsigned row_floats = (unsigned)PX * (unsigned)g.Dp;
    const unsigned stage_bytes = SL::stage_bytes(g.Dp, DT_LOAD);
    const unsigned planes_bytes = SL::planes_bytes(g.Dp, DT_LOAD);
    float4 *xslot = reinterpret_cast<float4 *>(smem_raw + (size_t)stages * stage_bytes);   // [NW][XR][VPL][LPP]
    uint64_t *full = reinterpret_cast<uint64_t *>(xslot + NW * XR * VPL * LPP);
    uint64_t *empty = full + stages;
    uint64_t *xb = empty + stages;                                                          // [NW][XR]

    const int xs0 = (int)blockIdx.x * TX;
    const int shift_l = (xs0 - 1) & 3, shift_r = (xs0 - SL::DCOV) & 3;
    const int yb0 = a.olo + (int)blockIdx.y * a.band_rows;
    const int yb1 = min(yb0 + a.band_rows - 1, a.ohi);
    const int frame = a.frame0 + (int)blockIdx.z;
    const int n_rows = yb1 - yb0 + 3;                                    // rows yb0-1 .. yb1+1

    if (tid == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NW); }
        for (int i = 0; i < NW * XR; ++i) mbar_init(&xb[i], 1);
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == NW) {
        // ------------------------------ producer warp ------------------------------------------
        if (lane != 0) return;
        for (int r = 0; r < n_rows; ++r) {
            const int s = r % stages;
            if (r >= stages) mbar_wait(&empty[s], (uint32_t)((r / stages - 1) & 1));
            const int yy = yb0 - 1 + r - g.vlo;
            unsigned char *stage_ptr = smem_raw + (size_t)s * stage_bytes;
            float *dst = reinterpret_cast<float *>(stage_ptr);
            unsigned tx = NPL * row_floats * 4u;
            if (!DT_LOAD) tx += SL::LBOX * 4u + SL::NRB * SL::RBOX * 4u + (a.has_prior ? SL::LBOX * 4u : 0u);
            mbar_expect_tx(&full[s], tx);
            tma_load_4d(dst, &tmW, &full[s], 0, xs0 - 1, yy, frame);
            tma_load_4d(dst + row_floats, &tmN, &full[s], 0, xs0 - 1, yy, frame);
            tma_load_4d(dst + 2 * row_floats, &tmE, &full[s], 0, xs0 - 1, yy, frame);
            if (DT_LOAD) {
                tma_load_4d(dst + 3 * row_floats, &tmD, &full[s], 0, xs0 - 1, yy, frame);
            } else {
                unsigned char *img = stage_ptr + planes_bytes;
                tma_load_3d(img, &tmL, &full[s], xs0 - 1 - shift_l, yy, frame);
#pragma unroll
                for (int b = 0; b < SL::NRB; ++b)
                    tma_load_3d(img + SL::L_BYTES + b * SL::RBOX * 4, &tmR, &full[s],
                                xs0 - SL::DCOV - shift_r + b * SL::RBOX, yy, frame);
                if (a.has_prior)
                    tma_load_3d(img + SL::L_BYTES + SL::R_BYTES, &tmP, &full[s], xs0 - 1 - shift_l, yy, frame);
            }
        }
        return;
    }

    // ---------------------------------- consumer warps -----------------------------------------
    const int px = tid / LPP;
    Lane<LPP, VPL> ln;
    ln.init(tid % LPP, g.D, g.Dp);
    const int x = xs0 - 1 + px;
    const bool in_x = (x >= 0 && x < g.W);
    const size_t fplane = (size_t)frame * g.plane_stride;
    const unsigned row_stride = (unsigned)g.W * (unsigned)g.Dp;
    const int pix_in_warp = lane / LPP;
    const bool first_pixel = (pix_in_warp == 0), last_pixel = (pix_in_warp == PPW - 1);

    const bool st_w = px >= 2 && (x - 1) < g.W;
    const bool st_n = in_x && px >= 1 && px <= PX - 2;
    const bool st_e = px <= PX - 3 && (x + 1) < g.W;
    const bool mask_w = (x == 0);                    // W'(y, -1) does not exist: the pixel left of the image reads 0

    float *Wout = a.Wout + fplane, *Nout = a.Nout + fplane, *Eout = a.Eout + fplane, *Sout = a.Sout + fplane;

    Vec<VPL> S0, S1, X0, X1, D0, D1;
#pragma unroll
    for (int v = 0; v < VPL; ++v) S0.q[v] = S1.q[v] = X0.q[v] = X1.q[v] = D0.q[v] = D1.q[v] = zero4();

    const unsigned lane_f4 = (unsigned)px * (unsigned)(g.Dp / 4) + (unsigned)ln.c;
    const unsigned plane_f4 = row_floats / 4;
    const float4 *lane_base = reinterpret_cast<const float4 *>(smem_raw) + lane_f4;
    const unsigned char *img_base = smem_raw + planes_bytes;
    float4 *my_slot = xslot + (size_t)warp * XR * VPL * LPP + ln.c;                         // + (k * VPL + v) * LPP
    const float4 *next_slot = xslot + (size_t)(warp + 1 < NW ? warp + 1 : warp) * XR * VPL * LPP + ln.c;
    uint64_t *my_xb = xb + warp * XR;
    uint64_t *next_xb = xb + (warp + 1 < NW ? warp + 1 : warp) * XR;
    const bool has_next = (warp + 1 < NW);

    int stage = 0;
    uint32_t phase = 0;
    int xk = 0;                 // exchange ring slot of the current row
    uint32_t xphase = 0;
    unsigned row_off = (unsigned)yb0 * row_stride + (unsigned)x * (unsigned)g.Dp + 4u * (unsigned)ln.c;

    auto fetch = [&](Vec<VPL> &w, Vec<VPL> &n, Vec<VPL> &e, Vec<VPL> &d) {
        mbar_wait(&full[stage], phase);
        const unsigned stage_off = (unsigned)stage * stage_bytes;
        const float4 *sp = reinterpret_cast<const float4 *>(reinterpret_cast<const unsigned char *>(lane_base) + stage_off);
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            if (FULL || ln.active[v]) {
                w.q[v] = sp[v * LPP];
                n.q[v] = sp[plane_f4 + v * LPP];
                e.q[v] = sp[2 * plane_f4 + v * LPP];
                if (DT_LOAD) d.q[v] = sp[3 * plane_f4 + v * LPP];
            } else {
                w.q[v] = n.q[v] = e.q[v] = zero4();
                if (DT_LOAD) d.q[v] = zero4();
            }
        }
        if (!DT_LOAD) {
            const unsigned char *img = img_base + stage_off;
            data_vec_smem<LPP, VPL, FULL>(d, ln, reinterpret_cast<const float *>(img) + shift_l,
                                          reinterpret_cast<const float *>(img + SL::L_BYTES) + shift_r,
                                          reinterpret_cast<const int *>(img + SL::L_BYTES + SL::R_BYTES) + shift_l, px,
                                          x, in_x, g.D, a.blend, a.has_prior != 0);
        }
        // this warp is done with the stage: hand it back to the producer
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
        if (++stage == stages) { stage = 0; phase ^= 1u; }
    };
    auto store_at = [&](float *plane, unsigned off, const Vec<VPL> &val) {
#pragma unroll
        for (int v = 0; v < VPL; ++v)
            if (FULL || ln.active[v]) st_plain(plane + off + 4 * LPP * v, val.q[v]);
    };

    // row yb0-1 only contributes S'(yb0)
    {
        Vec<VPL> w, n, e, d;
#pragma unroll
        for (int v = 0; v < VPL; ++v) d.q[v] = zero4();
        fetch(w, n, e, d);
#pragma unroll
        for (int v = 0; v < VPL; ++v) S0.q[v] = sum4(w.q[v], n.q[v], e.q[v], d.q[v]);
        normalise<LPP, VPL, FULL>(S0, ln);
        if (STORE_S && st_n) store_at(Sout, row_off, S0);
    }

    auto step = [&](int y, Vec<VPL> &Sin, Vec<VPL> &Snext, const Vec<VPL> &Xin, Vec<VPL> &Xout, const Vec<VPL> &Din,
                    Vec<VPL> &Dout) {
        Vec<VPL> w, n, e;
        fetch(w, n, e, Dout);
        if (y > g.vhi) {
#pragma unroll
            for (int v = 0; v < VPL; ++v) Sin.q[v] = zero4();
        }
        // W'(y, x-1) = norm(U_W(y, x))
        Vec<VPL> t;
#pragma unroll
        for (int v = 0; v < VPL; ++v) t.q[v] = sum4(Sin.q[v], n.q[v], e.q[v], Dout.q[v]);
        normalise<LPP, VPL, FULL>(t, ln);
        if (mask_w) {
#pragma unroll
            for (int v = 0; v < VPL; ++v) t.q[v] = zero4();
        }
        // publish the warp's first pixel for the warp on the left
        if (warp > 0) {
            if (first_pixel) {
#pragma unroll
                for (int v = 0; v < VPL; ++v) my_slot[(xk * VPL + v) * LPP] = t.q[v];
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&my_xb[xk]);
        }
        if (st_w && y <= yb1) store_at(Wout, row_off - (unsigned)g.Dp, t);
        // S'(y+1, x) = norm(U_S(y, x))  -- independent of the neighbours
#pragma unroll
        for (int v = 0; v < VPL; ++v) Snext.q[v] = sum4(w.q[v], n.q[v], e.q[v], Dout.q[v]);
        normalise<LPP, VPL, FULL>(Snext, ln);
        if (STORE_S && st_n && y + 1 <= yb1) store_at(Sout, row_off + row_stride, Snext);
        // W'(y, x): from the next pixel of this warp (SHFL) or the first pixel of the next warp (ring)
        Vec<VPL> wr;
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            wr.q[v].x = __shfl_down_sync(0xffffffffu, t.q[v].x, LPP);
            wr.q[v].y = __shfl_down_sync(0xffffffffu, t.q[v].y, LPP);
            wr.q[v].z = __shfl_down_sync(0xffffffffu, t.q[v].z, LPP);
            wr.q[v].w = __shfl_down_sync(0xffffffffu, t.q[v].w, LPP);
        }
        if (has_next) {
            mbar_wait(&next_xb[xk], xphase);
            if (last_pixel) {
#pragma unroll
                for (int v = 0; v < VPL; ++v) wr.q[v] = next_slot[(xk * VPL + v) * LPP];
            }
        } else if (last_pixel) {
#pragma unroll
            for (int v = 0; v < VPL; ++v) wr.q[v] = zero4();      // strip's last halo pixel: result unused
        }
        if (++xk == XR) { xk = 0; xphase ^= 1u; }
        // X = S'(y, x) + W'(y, x)
#pragma unroll
        for (int v = 0; v < VPL; ++v) Xout.q[v] = add4(Sin.q[v], wr.q[v]);
        // N'(y-1, x) = norm(U_N(y, x))
#pragma unroll
        for (int v = 0; v < VPL; ++v) t.q[v] = add4(add4(Xout.q[v], e.q[v]), Dout.q[v]);
        normalise<LPP, VPL, FULL>(t, ln);
        const bool upper = (y > yb0);
        if (upper && st_n) store_at(Nout, row_off - row_stride, t);
        // E'(y-1, x+1) = norm(U_E(y-1, x))
#pragma unroll
        for (int v = 0; v < VPL; ++v) t.q[v] = add4(add4(Xin.q[v], t.q[v]), Din.q[v]);
        normalise<LPP, VPL, FULL>(t, ln);
        if (upper && st_e) store_at(Eout, row_off - row_stride + (unsigned)g.Dp, t);
        row_off += row_stride;
    };

    for (int y = yb0;; y += 2) {
        step(y, S0, S1, X0, X1, D0, D1);
        if (y == yb1 + 1) break;
        step(y + 1, S1, S0, X1, X0, D1, D0);
        if (y == yb1) break;
    }
}

}  // namespace hsm
