// Two BP iterations per launch (algorithm 3): k_iter_tma's row stream with a second stage.
//
// Stage A is exactly k_iter_tma: at step y it consumes the old W, N, E of image row y (TMA) and
// produces the iteration-(t+1) values W'(y), N'(y-1), E'(y-1), S'(y+1).  Instead of going to HBM
// they are parked in a 4-row shared-memory ring.  Stage B trails three rows behind (row y-3 is the
// newest row whose three mailboxes are complete AND separated from their writers by a barrier) and
// runs the same recurrence on the ring to produce the iteration-(t+2) values, which are stored.
// HBM traffic per iteration halves (12 B per element-iteration), every per-row cost that does not
// scale with the arithmetic (TMA waits, barrier, addressing, data term) is paid once per two
// iterations, and the two stages give the scheduler two independent dependency chains.
//
// Price: two halo columns per side instead of one (TX = PX - 4 output columns per strip) and
// bands that start two rows early / end two rows late.
//
// Out-of-image rule (see kernels.cuh): any value LOCATED at a pixel outside the image is zero.
// For HBM-resident state the TMA zero-fill guarantees it; for the ring the producer masks the three
// cases where an in-image source writes to an out-of-image target (W' of column 0, E' of column
// W-1, N' of the first valid row).
#pragma once

#include "iter_tma.cuh"

namespace hsm {

template <int LPP, int VPL, int NT>
struct Ring2Layout {
    static constexpr int PX = NT / LPP;
    static constexpr int ROWS = 4;
    __host__ __device__ static unsigned plane_row_bytes(int Dp) { return (unsigned)PX * Dp * 4u; }
    __host__ __device__ static unsigned ring_bytes(int Dp) { return 3u * ROWS * plane_row_bytes(Dp) + ROWS * NT * VPL * 16u; }
};

template <int LPP, int VPL, int NT, bool FULL>
__global__ void __launch_bounds__(NT, NT == 256 ? 2 : 1)
    k_iter2_tma(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmN,
                const __grid_constant__ CUtensorMap tmE, const __grid_constant__ CUtensorMap tmL,
                const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmP, Geo g, IterArgs a,
                int stages)
{
    using SL = StageLayout<LPP, VPL, NT>;
    using RL = Ring2Layout<LPP, VPL, NT>;
    constexpr int PX = NT / LPP;
    constexpr int TX = PX - 4;
    constexpr int SLOTS = NT + LPP;
    constexpr bool DT_LOAD = false;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const bool STORE_S = a.store_s != 0;

    const int tid = threadIdx.x;
    const int px = tid / LPP;
    Lane<LPP, VPL> ln;
    ln.init(tid % LPP, g.D, g.Dp);

    const unsigned row_floats = (unsigned)PX * (unsigned)g.Dp;
    const unsigned stage_bytes = SL::stage_bytes(g.Dp, DT_LOAD);
    const unsigned planes_bytes = SL::planes_bytes(g.Dp, DT_LOAD);
    const unsigned prow = RL::plane_row_bytes(g.Dp);                       // one ring row of one plane
    unsigned char *ring = smem_raw + (size_t)stages * stage_bytes;         // [plane 0..2][row 0..3][PX][Dp]
    unsigned char *ring_d = ring + 3u * RL::ROWS * prow;                   // [row 0..3][VPL][NT] float4 (data term)
    float4 *exch = reinterpret_cast<float4 *>(ring_d + RL::ROWS * NT * VPL * 16u);   // stage B: W'' exchange
    uint64_t *full = reinterpret_cast<uint64_t *>(exch + 2 * VPL * SLOTS);
    uint64_t *xbar = full + stages;
    auto ex = [&](int par, int v, int slot) -> float4 & { return exch[(par * VPL + v) * SLOTS + slot]; };

    const int xs0 = (int)blockIdx.x * TX;
    const int x0 = xs0 - 2;                                                // column of px == 0
    const int shift_l = x0 & 3, shift_r = (x0 + 1 - SL::DCOV) & 3;
    const int x = x0 + px;
    const bool in_x = (x >= 0 && x < g.W);
    const int yb0 = a.olo + (int)blockIdx.y * a.band_rows;
    const int yb1 = min(yb0 + a.band_rows - 1, a.ohi);
    const int frame = a.frame0 + (int)blockIdx.z;
    const size_t fplane = (size_t)frame * g.plane_stride;
    const unsigned row_stride = (unsigned)g.W * (unsigned)g.Dp;
    const int y_first = yb0 - 2, y_last_a = yb1 + 2, y_last = yb1 + 4;
    const int n_rows = y_last_a - y_first + 1;                             // rows stage A streams

    // stage B store ownership (target pixel inside this strip's TX output columns and inside the image)
    const bool st_w = px >= 3 && px <= PX - 2 && (x - 1) < g.W;            // W''(r, x-1)
    const bool st_n = in_x && px >= 2 && px <= PX - 3;                     // N''(r-1, x), S''(r+1, x)
    const bool st_e = px >= 1 && px <= PX - 4 && (x + 1) < g.W;            // E''(r-1, x+1)
    const bool mask_w = (x == 0);                                          // W' of column 0 lands at column -1
    const bool mask_e = (x == g.W - 1);                                    // E' of the last column lands at column W

    if (tid == 0) {
        for (int s = 0; s < stages; ++s) mbar_init(&full[s], 1);
        mbar_init(xbar, NT);
        mbar_fence_init();
    }
    // ring and exchange start as zeros (slots nobody produces -- W' slot PX-1, E' slot 0 -- stay zero)
    {
        float4 *z = reinterpret_cast<float4 *>(ring);
        const unsigned n16 = (RL::ring_bytes(g.Dp) + 2u * VPL * SLOTS * 16u) / 16u;
        for (unsigned i = tid; i < n16; i += NT) z[i] = zero4();
    }
    __syncthreads();

    auto issue = [&](int r, int s) {
        const int yy = y_first + r - g.vlo;
        unsigned char *stage_ptr = smem_raw + (size_t)s * stage_bytes;
        float *dst = reinterpret_cast<float *>(stage_ptr);
        const unsigned tx = 3u * row_floats * 4u + SL::LBOX * 4u + SL::NRB * SL::RBOX * 4u + (a.has_prior ? SL::LBOX * 4u : 0u);
        mbar_expect_tx(&full[s], tx);
        tma_load_4d(dst, &tmW, &full[s], 0, x0, yy, frame);
        tma_load_4d(dst + row_floats, &tmN, &full[s], 0, x0, yy, frame);
        tma_load_4d(dst + 2 * row_floats, &tmE, &full[s], 0, x0, yy, frame);
        unsigned char *img = stage_ptr + planes_bytes;
        tma_load_3d(img, &tmL, &full[s], x0 - shift_l, yy, frame);
#pragma unroll
        for (int b = 0; b < SL::NRB; ++b)
            tma_load_3d(img + SL::L_BYTES + b * SL::RBOX * 4, &tmR, &full[s], x0 + 1 - SL::DCOV - shift_r + b * SL::RBOX,
                        yy, frame);
        if (a.has_prior) tma_load_3d(img + SL::L_BYTES + SL::R_BYTES, &tmP, &full[s], x0 - shift_l, yy, frame);
    };
    if (tid == 0) {
        const int first = n_rows < stages ? n_rows : stages;
        for (int s = 0; s < first; ++s) issue(s, s);
    }

    float *Wout = a.Wout + fplane, *Nout = a.Nout + fplane, *Eout = a.Eout + fplane, *Sout = a.Sout + fplane;

    // carried state of both stages, two copies each (alternate per row: no register moves)
    Vec<VPL> SA0, SA1, XA0, XA1, DA0, DA1, SB0, SB1, XB0, XB1, DB0, DB1;
#pragma unroll
    for (int v = 0; v < VPL; ++v)
        SA0.q[v] = SA1.q[v] = XA0.q[v] = XA1.q[v] = DA0.q[v] = DA1.q[v] = SB0.q[v] = SB1.q[v] = XB0.q[v] = XB1.q[v] =
            DB0.q[v] = DB1.q[v] = zero4();

    const unsigned lane_f4 = (unsigned)px * (unsigned)(g.Dp / 4) + (unsigned)ln.c;
    const unsigned plane_f4 = row_floats / 4;
    const float4 *lane_base = reinterpret_cast<const float4 *>(smem_raw) + lane_f4;
    const unsigned char *img_base = smem_raw + planes_bytes;
    // ring addressing: plane p, row slot s, pixel q, this lane's chunk v
    auto ring_at = [&](int plane, int slot, int q, int v) -> float4 & {
        return *reinterpret_cast<float4 *>(ring + (size_t)(plane * RL::ROWS + slot) * prow +
                                           ((size_t)q * (g.Dp / 4) + ln.c + v * LPP) * 16u);
    };
    auto ringd_at = [&](int slot, int v) -> float4 & {
        return *reinterpret_cast<float4 *>(ring_d + ((size_t)(slot * VPL + v) * NT + tid) * 16u);
    };
    int stage = 0;
    uint32_t phase = 0, xphase = 0;
    // element offset of (row r = y - 3, this lane's pixel and chunk) inside the frame, advanced per step
    unsigned row_off = (unsigned)(y_first - 3) * row_stride + (unsigned)x * (unsigned)g.Dp + 4u * (unsigned)ln.c;

    auto store_at = [&](float *plane, unsigned off, const Vec<VPL> &val) {
#pragma unroll
        for (int v = 0; v < VPL; ++v)
            if (FULL || ln.active[v]) st_plain(plane + off + 4 * LPP * v, val.q[v]);
    };
    auto zero_vec = [&](Vec<VPL> &t) {
#pragma unroll
        for (int v = 0; v < VPL; ++v) t.q[v] = zero4();
    };

    // One row step.  Stage A works on image row y, stage B on row r = y - 3.
    auto step = [&](int y, const int par, Vec<VPL> &SAin, Vec<VPL> &SAnext, const Vec<VPL> &XAin, Vec<VPL> &XAout,
                    const Vec<VPL> &DAin, Vec<VPL> &DAout, Vec<VPL> &SBin, Vec<VPL> &SBnext, const Vec<VPL> &XBin,
                    Vec<VPL> &XBout, const Vec<VPL> &DBin, Vec<VPL> &DBout) {
        const int r = y - 3;
        const bool a_on = (y <= y_last_a);
        const bool b_on = (r >= yb0 - 1);
        const int slot_y = y & 3, slot_ym1 = (y - 1) & 3, slot_r = r & 3;
        Vec<VPL> w, n, e, t;

        // ---------------- stage A, part 1: fetch row y, publish W'_A(y, x-1), S'_A(y+1) ----------
        if (a_on) {
            mbar_wait(&full[stage], phase);
            const unsigned stage_off = (unsigned)stage * stage_bytes;
            const float4 *sp = reinterpret_cast<const float4 *>(reinterpret_cast<const unsigned char *>(lane_base) + stage_off);
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                if (FULL || ln.active[v]) {
                    w.q[v] = sp[v * LPP];
                    n.q[v] = sp[plane_f4 + v * LPP];
                    e.q[v] = sp[2 * plane_f4 + v * LPP];
                } else {
                    w.q[v] = n.q[v] = e.q[v] = zero4();
                }
            }
            const unsigned char *img = img_base + stage_off;
            data_vec_smem<LPP, VPL, FULL>(DAout, ln, reinterpret_cast<const float *>(img) + shift_l,
                                          reinterpret_cast<const float *>(img + SL::L_BYTES) + shift_r,
                                          reinterpret_cast<const int *>(img + SL::L_BYTES + SL::R_BYTES) + shift_l, px,
                                          x, in_x, g.D, a.blend, a.has_prior != 0);
#pragma unroll
            for (int v = 0; v < VPL; ++v) ringd_at(slot_y, v) = DAout.q[v];
            if (y > g.vhi) zero_vec(SAin);
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                t.q[v] = sum4(SAin.q[v], n.q[v], e.q[v], DAout.q[v]);
                SAnext.q[v] = sum4(w.q[v], n.q[v], e.q[v], DAout.q[v]);
            }
            normalise_pair<LPP, VPL, FULL>(t, SAnext, ln);
            if (mask_w) zero_vec(t);
            if (px >= 1) {
#pragma unroll
                for (int v = 0; v < VPL; ++v)
                    if (FULL || ln.active[v]) ring_at(0, slot_y, px - 1, v) = t.q[v];
            }
        }
        // ---------------- stage B, part 1: row r from the ring, publish W''(r, x-1), S''(r+1) -----
        Vec<VPL> wb, nb, eb;
        if (b_on) {
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                if (FULL || ln.active[v]) {
                    wb.q[v] = ring_at(0, slot_r, px, v);
                    nb.q[v] = ring_at(1, slot_r, px, v);
                    eb.q[v] = ring_at(2, slot_r, px, v);
                } else {
                    wb.q[v] = nb.q[v] = eb.q[v] = zero4();
                }
                DBout.q[v] = ringd_at(slot_r, v);
            }
            if (r > g.vhi) zero_vec(SBin);
#pragma unroll
            for (int v = 0; v < VPL; ++v) SBnext.q[v] = sum4(wb.q[v], nb.q[v], eb.q[v], DBout.q[v]);
            if (r >= yb0) {
#pragma unroll
                for (int v = 0; v < VPL; ++v) t.q[v] = sum4(SBin.q[v], nb.q[v], eb.q[v], DBout.q[v]);
                normalise_pair<LPP, VPL, FULL>(t, SBnext, ln);
                if (mask_w) zero_vec(t);
#pragma unroll
                for (int v = 0; v < VPL; ++v) ex(par, v, tid) = t.q[v];
                if (st_w && r <= yb1) store_at(Wout, row_off - (unsigned)g.Dp, t);
            } else {
                normalise_checked<LPP, VPL, FULL>(SBnext, ln);
            }
            if (STORE_S && st_n && r + 1 >= yb0 && r + 1 <= yb1) store_at(Sout, row_off + row_stride, SBnext);
        }
        // ---------------- one barrier for both exchanges -----------------------------------------
        mbar_arrive(xbar);
        mbar_wait(xbar, xphase);
        xphase ^= 1u;
        if (a_on) {
            if (tid == 0 && (y - y_first) + stages < n_rows) issue((y - y_first) + stages, stage);
            if (++stage == stages) { stage = 0; phase ^= 1u; }
        }
        // ---------------- stage A, part 2: N'_A(y-1, x), E'_A(y-1, x+1) into the ring --------------
        if (a_on && y > y_first) {
#pragma unroll
            for (int v = 0; v < VPL; ++v)
                XAout.q[v] = (FULL || ln.active[v]) ? add4(SAin.q[v], ring_at(0, slot_y, px, v)) : zero4();
            Vec<VPL> tn;
#pragma unroll
            for (int v = 0; v < VPL; ++v) tn.q[v] = add4(add4(XAout.q[v], e.q[v]), DAout.q[v]);
            if (y - 1 < g.vlo) {                            // block-uniform: the row above the image receives nothing
                zero_vec(tn);
#pragma unroll
                for (int v = 0; v < VPL; ++v) t.q[v] = add4(add4(XAin.q[v], tn.q[v]), DAin.q[v]);
                normalise<LPP, VPL, FULL>(t, ln);
            } else {
                normalise_chain<LPP, VPL, FULL>(tn, t, XAin, DAin, ln);
            }
#pragma unroll
            for (int v = 0; v < VPL; ++v)
                if (FULL || ln.active[v]) ring_at(1, slot_ym1, px, v) = tn.q[v];
            if (mask_e) zero_vec(t);
            if (px + 1 < PX) {
#pragma unroll
                for (int v = 0; v < VPL; ++v)
                    if (FULL || ln.active[v]) ring_at(2, slot_ym1, px + 1, v) = t.q[v];
            }
        }
        // ---------------- stage B, part 2: N''(r-1, x), E''(r-1, x+1) to HBM ------------------------
        if (b_on && r >= yb0) {
            const int rd = (px + 1 < PX) ? tid + LPP : NT + ln.c;      // last pixel group reads the zero slot
#pragma unroll
            for (int v = 0; v < VPL; ++v) XBout.q[v] = add4(SBin.q[v], ex(par, v, rd));
            Vec<VPL> tnb;
#pragma unroll
            for (int v = 0; v < VPL; ++v) tnb.q[v] = add4(add4(XBout.q[v], eb.q[v]), DBout.q[v]);
            normalise_chain<LPP, VPL, FULL>(tnb, t, XBin, DBin, ln);
            const bool upper = (r > yb0);
            if (upper && st_n) store_at(Nout, row_off - row_stride, tnb);
            if (upper && st_e) store_at(Eout, row_off - row_stride + (unsigned)g.Dp, t);
        }
        row_off += row_stride;
    };

    for (int y = y_first;; y += 2) {
        step(y, 0, SA0, SA1, XA0, XA1, DA0, DA1, SB0, SB1, XB0, XB1, DB0, DB1);
        if (y == y_last) break;
        step(y + 1, 1, SA1, SA0, XA1, XA0, DA1, DA0, SB1, SB0, XB1, XB0, DB1, DB0);
        if (y + 1 == y_last) break;
    }
}

}  // namespace hsm
