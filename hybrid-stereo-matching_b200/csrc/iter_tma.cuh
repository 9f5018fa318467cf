// Fused iteration kernel, TMA-fed variant (algorithm 2; the default, algorithm 5 in iter_tma_st.cuh, adds TMA
// stores to this design and has received the later instruction-level work).
//
// Same arithmetic and the same row-streaming structure as k_iter_fused in
// kernels.cuh (see the derivation there), but the old W / N / E (/ data) rows are
// brought in by the Tensor Memory Accelerator instead of per-thread LDG:
//
//   * each plane is described by a 4-D tensor map (Dp, cols, valid rows, frames);
//     one cp.async.bulk.tensor per plane moves the CTA's row segment
//     (PX pixels x Dp floats, contiguous in HBM) into a shared-memory stage;
//   * out-of-image coordinates (column -1, column >= W, rows outside the valid
//     range) are ZERO-FILLED by the hardware, which is exactly the "a border line
//     that receives nothing stays zero" rule of the reference (mrf.py:83-90);
//   * a ring of `stages` row stages with one mbarrier each keeps several rows in
//     flight without holding them in registers (the LDG variant needs a second
//     register set for its one-row prefetch and loses occupancy to it);
//   * a stage is refilled right after the row's only __syncthreads (every thread
//     copied its chunk of the stage into registers before that barrier).
#pragma once

#include <cuda.h>

#include "kernels.cuh"

namespace hsm {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
#ifndef HSM_WAIT_HINT_NS
#define HSM_WAIT_HINT_NS 20000
#endif
// try_wait suspends the thread in hardware until the phase completes or the time hint expires;
// without a hint it returns after a short system limit and the retry loop burns issue slots.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity), "r"((uint32_t)HSM_WAIT_HINT_NS)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2,
                                            int c3)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
// The same load with an L2 eviction-priority hint (the 64-bit policy words createpolicy.fractional produces for a
// fraction of 1.0).  A band's first two rows are also the last two rows of the band above, which reaches them one
// band's duration later: the first reader asks L2 to keep them (evict_last), the second one lets them go, and the
// rows in between are streamed (evict_first) so that they do not push the kept rows out.
constexpr uint64_t kL2EvictFirst = 0x12F0000000000000ull, kL2EvictLast = 0x14F0000000000000ull;
__device__ __forceinline__ void tma_load_4d_hint(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2,
                                                 int c3, uint64_t policy)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "l"(policy)
        : "memory");
}

// Programmatic dependent launch: when the launch carries the programmatic-stream-serialization
// attribute, the CTAs of iteration k+1 may become resident while iteration k drains and run their
// prologue (barrier init, shared-memory setup); pdl_wait() blocks until iteration k has completed and
// its stores are visible.  Without the attribute both are no-ops.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

#ifndef HSM_TMA_MIN_BLOCKS
#define HSM_TMA_MIN_BLOCKS 3
#endif

// Byte size of one pipeline stage and where the image segments sit inside it (T24 only).
template <int LPP, int VPL, int NT>
struct StageLayout {
    static constexpr int PX = PixelMap<LPP, NT>::PX;
    static constexpr int DCOV = 4 * LPP * VPL;
    // The innermost TMA coordinate must be 16-byte aligned, so image segments start at a column
    // that is a multiple of 4 (up to 3 columns early) and are 4 columns longer than needed.
    // (box widths are multiples of 4 elements = 16 bytes, a TMA requirement; PX is not one for the 7- and 9-warp CTAs)
    static constexpr int PX4 = (PX + 3) / 4 * 4;
    static constexpr int LBOX = PX4 + 4;                    // left image / prior: columns xs0 - 1 ..
    static constexpr int RSEG = PX4 + DCOV + 4;             // right image: columns xs0 - DCOV .. xs0 + PX - 1
    static constexpr int NRB = RSEG > 256 ? 2 : 1;          // TMA boxes are at most 256 elements wide
    // two boxes: each a multiple of 32 floats so that the second one lands 128-byte aligned
    static constexpr int RBOX = NRB == 1 ? RSEG : ((RSEG / 2 + 31) / 32 * 32);
    __host__ __device__ static constexpr unsigned align128(unsigned v) { return (v + 127u) & ~127u; }
    static constexpr unsigned L_BYTES = align128(LBOX * 4), R_BYTES = align128(NRB * RBOX * 4),
                              P_BYTES = align128(LBOX * 4);
    // one plane's row segment in a stage: PX x Dp floats, padded to the 128-byte alignment bulk tensor copies need
    // of their shared-memory address (only the 7- and 9-warp CTAs have a PX that needs padding)
    __host__ __device__ static unsigned seg_bytes(int Dp) { return align128((unsigned)PX * (unsigned)Dp * 4u); }
    __host__ __device__ static unsigned planes_bytes(int Dp, bool dt_load) { return (dt_load ? 4u : 3u) * seg_bytes(Dp); }
    __host__ __device__ static unsigned stage_bytes(int Dp, bool dt_load)
    {
        return planes_bytes(Dp, dt_load) + (dt_load ? 0u : (L_BYTES + R_BYTES + P_BYTES));
    }
};

template <int LPP, int VPL, int NT, bool DT_LOAD, bool FULL, bool PUSH>
__global__ void __launch_bounds__(NT, (NT == 256 && VPL == 1) ? HSM_TMA_MIN_BLOCKS : 1)
    k_iter_tma(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmN,
               const __grid_constant__ CUtensorMap tmE, const __grid_constant__ CUtensorMap tmD,
               const __grid_constant__ CUtensorMap tmL, const __grid_constant__ CUtensorMap tmR,
               const __grid_constant__ CUtensorMap tmP, Geo g, IterArgs a, int stages)
{
    using SL = StageLayout<LPP, VPL, NT>;
    const bool STORE_S = a.store_s != 0;
    constexpr int PX = NT / LPP;
    constexpr int TX = PX - 2;
    constexpr int SLOTS = NT + 2 * LPP;
    constexpr int NPL = DT_LOAD ? 4 : 3;
    extern __shared__ __align__(1024) unsigned char smem_raw[];

    const int tid = threadIdx.x;
    const int px = tid / LPP;
    Lane<LPP, VPL> ln;
    ln.init(tid % LPP, g.D, g.Dp);

    const unsigned row_floats = (unsigned)PX * (unsigned)g.Dp;          // one plane's row segment
    const unsigned stage_bytes = SL::stage_bytes(g.Dp, DT_LOAD);
    const unsigned planes_bytes = SL::planes_bytes(g.Dp, DT_LOAD);
    float4 *exch = reinterpret_cast<float4 *>(smem_raw + (size_t)stages * stage_bytes);
    uint64_t *full = reinterpret_cast<uint64_t *>(exch + 2 * VPL * SLOTS);
    uint64_t *xbar = full + stages;                                      // W' exchange barrier (all NT threads)
    auto ex = [&](int par, int v, int slot) -> float4 & { return exch[(par * VPL + v) * SLOTS + slot]; };

    const int xs0 = (int)blockIdx.x * TX;
    const int shift_l = (xs0 - 1) & 3, shift_r = (xs0 - SL::DCOV) & 3;    // see StageLayout
    const int x = xs0 - 1 + px;
    const bool in_x = (x >= 0 && x < g.W);
    const int yb0 = a.olo + (int)blockIdx.y * a.band_rows;
    const int yb1 = min(yb0 + a.band_rows - 1, a.ohi);
    const int frame = a.frame0 + (int)blockIdx.z;
    const size_t fplane = (size_t)frame * g.plane_stride;
    const unsigned row_stride = (unsigned)g.W * (unsigned)g.Dp;
    const int n_rows = yb1 - yb0 + 3;                                    // rows yb0-1 .. yb1+1

    const bool st_w = px >= 2 && (x - 1) < g.W;
    const bool st_n = in_x && px >= 1 && px <= PX - 2;
    const bool st_e = px <= PX - 3 && (x + 1) < g.W;
    const int rd_slot = (x < 0) ? (NT + LPP + ln.c) : (tid + LPP);

    if (tid == 0) {
        for (int s = 0; s < stages; ++s) mbar_init(&full[s], 1);
        mbar_init(xbar, NT);
        mbar_fence_init();
    }
    if (tid < 2 * LPP) {
#pragma unroll
        for (int v = 0; v < VPL; ++v) ex(0, v, NT + tid) = ex(1, v, NT + tid) = zero4();
    }
    __syncthreads();

    // producer side (thread 0): one bulk tensor copy per plane into stage s for row index r
    auto issue = [&](int r, int s) {
        const int yy = yb0 - 1 + r - g.vlo;            // row coordinate inside the valid-row tensor
        unsigned char *stage_ptr = smem_raw + (size_t)s * stage_bytes;
        float *dst = reinterpret_cast<float *>(stage_ptr);
        unsigned tx = NPL * row_floats * 4u;
        if (!DT_LOAD) tx += SL::LBOX * 4u + SL::NRB * SL::RBOX * 4u + (a.has_prior ? SL::LBOX * 4u : 0u);
        mbar_expect_tx(&full[s], tx);
        const int frame_ld = a.zero_in ? -1 : frame;      // out-of-tensor frame: zero fill, no memory read
        tma_load_4d(dst, &tmW, &full[s], 0, xs0 - 1, yy, frame_ld);
        tma_load_4d(dst + row_floats, &tmN, &full[s], 0, xs0 - 1, yy, frame_ld);
        tma_load_4d(dst + 2 * row_floats, &tmE, &full[s], 0, xs0 - 1, yy, frame_ld);
        if (DT_LOAD) {
            tma_load_4d(dst + 3 * row_floats, &tmD, &full[s], 0, xs0 - 1, yy, frame);
        } else {
            unsigned char *img = stage_ptr + planes_bytes;
            tma_load_3d(img, &tmL, &full[s], xs0 - 1 - shift_l, yy, frame);
#pragma unroll
            for (int b = 0; b < SL::NRB; ++b)
                tma_load_3d(img + SL::L_BYTES + b * SL::RBOX * 4, &tmR, &full[s],
                            xs0 - SL::DCOV - shift_r + b * SL::RBOX, yy, frame);
            if (a.has_prior) tma_load_3d(img + SL::L_BYTES + SL::R_BYTES, &tmP, &full[s], xs0 - 1 - shift_l, yy, frame);
        }
    };
    pdl_launch_dependents();
    pdl_wait();                                 // everything above overlaps the previous iteration's tail
    if (tid == 0) {
        const int first = n_rows < stages ? n_rows : stages;
        for (int s = 0; s < first; ++s) issue(s, s);
    }

    float *Wout = a.Wout + fplane, *Nout = a.Nout + fplane, *Eout = a.Eout + fplane;
    float *Sout = a.Sout + fplane;

    // ---- consumer side ------------------------------------------------------------------------
    // Carried between rows (two copies each, alternating per row so that no register moves are
    // needed): S'(y, x), X(y-1, x) = S'(y-1, x) + W'(y-1, x), data term of row y-1.
    Vec<VPL> S0, S1, X0, X1, D0, D1;
#pragma unroll
    for (int v = 0; v < VPL; ++v) S0.q[v] = S1.q[v] = X0.q[v] = X1.q[v] = D0.q[v] = D1.q[v] = zero4();

    const unsigned lane_f4 = (unsigned)px * (unsigned)(g.Dp / 4) + (unsigned)ln.c;   // float4 index in a row segment
    const unsigned plane_f4 = row_floats / 4;
    const float4 *lane_base = reinterpret_cast<const float4 *>(smem_raw) + lane_f4;
    const unsigned char *img_base = smem_raw + planes_bytes;
    int stage = 0;
    uint32_t phase = 0;
    // element offset of (row y, this lane's pixel and chunk) inside the frame; advanced by one row per step
    // (x may be -1 or W for halo lanes: the wrap-around cancels when the neighbour offset is added)
    unsigned row_off = (unsigned)yb0 * row_stride + (unsigned)x * (unsigned)g.Dp + 4u * (unsigned)ln.c;

    // wait for the stage holding row y, copy this lane's chunks to registers, build the data term
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
    };
    // every thread has copied the current stage: refill it with row index r + stages
    auto refill = [&](int r) {
        if (tid == 0 && r + stages < n_rows) issue(r + stages, stage);
        if (++stage == stages) { stage = 0; phase ^= 1u; }
    };
    auto store_at = [&](float *plane, unsigned off, const Vec<VPL> &val) {
#pragma unroll
        for (int v = 0; v < VPL; ++v)
            if (FULL || ln.active[v]) st_plain(plane + off + 4 * LPP * v, val.q[v]);
    };
    // peer-to-peer halo push: row `row` of a new plane also goes to the neighbouring GPU's ghost row
    const unsigned lane_col = (unsigned)x * (unsigned)g.Dp + 4u * (unsigned)ln.c;
    auto push_row = [&](float *const (&peer)[2], int row, int dx, bool mine, const Vec<VPL> &val) {
        if (!mine) return;
        const unsigned col = lane_col + (unsigned)(dx * g.Dp);
        if (row == a.olo && peer[0] != nullptr) store_at(peer[0], a.push_off[0] + col, val);
        if (row == a.ohi && peer[1] != nullptr) store_at(peer[1], a.push_off[1] + col, val);
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
        __syncthreads();
        refill(0);
    }

    // One row: Sin = S'(y, x); Xin, Din belong to row y-1; Sout_/Xout/Dout are produced for the next row.
    // The exchange of W' uses a split barrier: arrive after publishing W'(y, x-1), compute S'(y+1)
    // (which does not depend on the neighbour), then wait and pick up W'(y, x).
    auto step = [&](int y, const int par, Vec<VPL> &Sin, Vec<VPL> &Snext, const Vec<VPL> &Xin, Vec<VPL> &Xout,
                    const Vec<VPL> &Din, Vec<VPL> &Dout) {
        Vec<VPL> w, n, e;
        fetch(w, n, e, Dout);
        if (y > g.vhi) {                       // block-uniform: the row below the image sends nothing
#pragma unroll
            for (int v = 0; v < VPL; ++v) Sin.q[v] = zero4();
        }
        // W'(y, x-1) = norm(U_W(y, x)) and S'(y+1, x) = norm(U_S(y, x)): independent of each other,
        // normalised branch-free; one warp-uniform fallback if any lane saw an out-of-range operand.
        Vec<VPL> t;
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            t.q[v] = sum4(Sin.q[v], n.q[v], e.q[v], Dout.q[v]);
            Snext.q[v] = sum4(w.q[v], n.q[v], e.q[v], Dout.q[v]);
        }
        float mok_w, mok_s;
        bool bad = normalise_fast<LPP, VPL, FULL>(t, ln, &mok_w);
        bad |= normalise_fast<LPP, VPL, FULL>(Snext, ln, &mok_s);
        if (__any_sync(0xffffffffu, bad)) {                        // warp-uniform, uncommon
            Vec<VPL> uw, us;
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                uw.q[v] = sum4(Sin.q[v], n.q[v], e.q[v], Dout.q[v]);
                us.q[v] = sum4(w.q[v], n.q[v], e.q[v], Dout.q[v]);
            }
            // zeros (border columns, halo lanes) are fine for the fast path; anything else is redone exactly
            const bool exact = lane_needs_exact<LPP, VPL, FULL>(uw, ln, mok_w) ||
                               lane_needs_exact<LPP, VPL, FULL>(us, ln, mok_s);
            if (__any_sync(0xffffffffu, exact)) {
                normalise<LPP, VPL, FULL>(uw, ln);
                normalise<LPP, VPL, FULL>(us, ln);
                t = uw;
                Snext = us;
            }
        }
#pragma unroll
        for (int v = 0; v < VPL; ++v) ex(par, v, tid) = t.q[v];
        if (st_w && y <= yb1) store_at(Wout, row_off - (unsigned)g.Dp, t);
        if constexpr (PUSH) push_row(a.pushW, y, -1, st_w && y <= yb1, t);
        if (STORE_S && st_n && y + 1 <= yb1) store_at(Sout, row_off + row_stride, Snext);
        __syncthreads();                       // the row's only CTA barrier: W' published, stage consumed
        refill(y - yb0 + 1);
        // X = S'(y, x) + W'(y, x)
#pragma unroll
        for (int v = 0; v < VPL; ++v) Xout.q[v] = add4(Sin.q[v], ex(par, v, rd_slot));
        // N'(y-1, x) = norm(U_N(y, x));  E'(y-1, x+1) = norm(U_E(y-1, x)) needs N'(y-1, x)
        Vec<VPL> tn;
#pragma unroll
        for (int v = 0; v < VPL; ++v) tn.q[v] = add4(add4(Xout.q[v], e.q[v]), Dout.q[v]);
        float mok_n, mok_e;
        bad = normalise_fast<LPP, VPL, FULL>(tn, ln, &mok_n);
#pragma unroll
        for (int v = 0; v < VPL; ++v) t.q[v] = add4(add4(Xin.q[v], tn.q[v]), Din.q[v]);
        bad |= normalise_fast<LPP, VPL, FULL>(t, ln, &mok_e);
        if (__any_sync(0xffffffffu, bad)) {
            Vec<VPL> un, ue;
#pragma unroll
            for (int v = 0; v < VPL; ++v) un.q[v] = add4(add4(Xout.q[v], e.q[v]), Dout.q[v]);
            bool exact = lane_needs_exact<LPP, VPL, FULL>(un, ln, mok_n);
            // U_E is built from the (fast) N'; if that one is kept, so is the sum
#pragma unroll
            for (int v = 0; v < VPL; ++v) ue.q[v] = add4(add4(Xin.q[v], tn.q[v]), Din.q[v]);
            exact = exact || lane_needs_exact<LPP, VPL, FULL>(ue, ln, mok_e);
            if (__any_sync(0xffffffffu, exact)) {
                normalise<LPP, VPL, FULL>(un, ln);
                tn = un;
#pragma unroll
                for (int v = 0; v < VPL; ++v) ue.q[v] = add4(add4(Xin.q[v], tn.q[v]), Din.q[v]);
                normalise<LPP, VPL, FULL>(ue, ln);
                t = ue;
            }
        }
        const bool upper = (y > yb0);          // row y-1 belongs to this band
        if (upper && st_n) store_at(Nout, row_off - row_stride, tn);
        if (upper && st_e) store_at(Eout, row_off - row_stride + (unsigned)g.Dp, t);
        if (upper) {
            if constexpr (PUSH) push_row(a.pushN, y - 1, 0, st_n, tn);
            if constexpr (PUSH) push_row(a.pushE, y - 1, +1, st_e, t);
        }
        row_off += row_stride;
    };

    for (int y = yb0;; y += 2) {
        step(y, 0, S0, S1, X0, X1, D0, D1);
        if (y == yb1 + 1) break;
        step(y + 1, 1, S1, S0, X1, X0, D1, D0);
        if (y == yb1) break;
    }
}

}  // namespace hsm
