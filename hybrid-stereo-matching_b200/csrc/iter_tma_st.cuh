// k_iter_tma with TMA STORES (algorithm 5): results are staged in shared memory and leave the SM
// as one cp.async.bulk.tensor store per plane-row (issued by thread 0 after the row's barrier)
// instead of per-thread STG.128 with 64-bit address arithmetic and edge predicates.  The TMA unit
// clips columns >= W.  Arithmetic, pipeline and barrier structure are k_iter_tma's; this is the default kernel,
// so the instruction-level work went here: shared-space addresses kept per lane, the data term's border / halo /
// prior handling off the common path, warp-uniform branches instead of per-row selects.
// The body is tmast_task(): one (strip, band, frame) tile of one iteration.  k_iter_tmast runs one task per CTA,
// k_iter_flow (iter_flow.cuh) runs task after task in a persistent CTA.
#pragma once

#include "iter_tma.cuh"

namespace hsm {

__device__ __forceinline__ void tma_store_4d(const CUtensorMap *map, const void *src, int c0, int c1, int c2, int c3)
{
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void tma_store_4d_hint(const CUtensorMap *map, const void *src, int c0, int c1, int c2, int c3,
                                                  uint64_t policy)
{
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3, %4, %5}], [%1], %6;"
                 ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned *p, unsigned v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Dataflow kernel: a finished tile waiting to be published -- its local progress counter and, for a tile on
// the band's boundary in row-band mode, the counters in the neighbouring GPUs' memory.
struct FlowPending {
    unsigned *local;            // nullptr = nothing pending
    unsigned local_value;
    unsigned *remote[2];        // nullptr = no neighbour on that side / not a boundary tile
    unsigned remote_value;
};
__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Called by the thread that issued the tile's bulk stores, once they are no longer needed in flight: wait for
// their completion, then release the counters.  (Peer pushes were plain stores of all threads, fenced at
// system scope by each thread before the CTA barrier that precedes this call.)
template <bool PUSH>
__device__ __forceinline__ void flow_publish(const FlowPending &p)
{
    tma_store_wait_all();
    asm volatile("fence.proxy.async;" ::: "memory");
    if (PUSH && (p.remote[0] != nullptr || p.remote[1] != nullptr)) __threadfence_system();
    else __threadfence();
    st_release_gpu(p.local, p.local_value);
    if (PUSH) {
        if (p.remote[0]) st_release_sys(p.remote[0], p.remote_value);
        if (p.remote[1]) st_release_sys(p.remote[1], p.remote_value);
    }
}

// Row-band mode (PUSH kernels): tensor maps over the ghost rows of the neighbouring GPUs' planes (side 0 = the
// band above, 1 = the band below; dims (Dp, cols, 1, 1) based at the ghost row, peer-mapped memory).  The first /
// last owned row of a new plane is staged in shared memory for its local bulk store anyway, so the halo push is
// ONE MORE bulk tensor store of the same staging buffer, issued by the same thread: no other thread of the CTA
// executes a single instruction for it (per-thread peer stores cost ~10 % of the kernel).
struct PushMaps {
    CUtensorMap W[2], N[2], E[2];
};

// One task of the fused iteration: strip `bx`, band `by` of frame `frame`.  Shared by the one-launch-per-
// iteration kernel (k_iter_tmast: one task per CTA) and the persistent dataflow kernel (iter_flow.cuh: a CTA
// runs task after task; FLOW = true).  `stage` / `phase` are the position in the ring of row stages; they carry
// over from task to task (every load issued by a task is consumed by it, so the ring stays consistent).
// S_MODE: whether the tile also stores S' with plain stores (only the last iteration of a call does):
// 0 = never, 1 = always, 2 = per call argument `store_s_arg` (the persistent kernel decides per task).  Compiled
// out of the 49 of 50 launches that do not need it, the test costs nothing in their rows.
template <int LPP, int VPL, int NT, bool DT_LOAD, bool FULL, bool PUSH, bool FLOW, int S_MODE>
__device__ __forceinline__ void tmast_task(const CUtensorMap *tmW, const CUtensorMap *tmN, const CUtensorMap *tmE,
                                           const CUtensorMap *tmD, const CUtensorMap *tmL, const CUtensorMap *tmR,
                                           const CUtensorMap *tmP, const CUtensorMap *tsW, const CUtensorMap *tsN,
                                           const CUtensorMap *tsE, const Geo &g, const IterArgs &a, const int stages,
                                           const int bx, const int by, const int frame, const bool store_s_arg,
                                           unsigned char *smem_raw, int &stage, uint32_t &phase, const bool first_task,
                                           const PushPtrs &pp, const bool zero_in,
                                           const FlowPending *publish = nullptr, const PushMaps *pmaps = nullptr)
{
    using SL = StageLayout<LPP, VPL, NT>;
    using PM = PixelMap<LPP, NT>;
    constexpr int PX = PM::PX;
    constexpr int TX = PX - 2;
    constexpr int SLOTS = PM::LANES + 2 * LPP;   // [0, LANES): real lanes; then a scratch pixel and a zero pixel
    constexpr int NPL = DT_LOAD ? 4 : 3;
    // L2 eviction hints (IterArgs::l2_hints) exist only in the instantiations that gain from them: the code of the
    // issuing thread sits on the CTA's critical path, and merely compiling the (disabled) hint branches in cost the
    // persistent kernel 1.5-6 % and the issue-bound layouts 2-13 % (profiles/r2_probe_ab_hint_code.txt)
    constexpr bool L2_HINTS = FULL && !FLOW && !PUSH && !DT_LOAD;
    // steady-state loads issued by four warps instead of thread 0 (see issue_split): +2-4 % for the persistent kernel
    // on one frame (640x480xD64 0.844 -> 0.870 of peak), nothing for batches of full layouts, and -1.5 ... -3.5 % where
    // the kernel is issue-bound (every warp then evaluates the dispatch: D = 22, 50, the narrow frames) --
    // profiles/r2_probe_ab_split_issue.txt, same box, with the split compiled into every instantiation
#ifndef HSM_SPLIT_ISSUE
#define HSM_SPLIT_ISSUE (FLOW && FULL && !PUSH)
#endif
    constexpr bool SPLIT_ISSUE = HSM_SPLIT_ISSUE;
    const bool STORE_S = S_MODE == 2 ? store_s_arg : (S_MODE == 1);

    const int tid = threadIdx.x;
    PM pm;
    pm.init(tid);
    const int px = pm.px;
    Lane<LPP, VPL> ln;
    // a full layout (D == 4 * LPP * VPL) knows its pitch at compile time: stage sizes, slot offsets and row
    // pitches fold into immediates instead of being rebuilt from the runtime value in every row
    // (not in the persistent kernel with the halo push compiled in: there the constant costs 8 % -- 1080p D256 on two
    // GPUs 0.988 -> 1.075 ms per iteration, same box, profiles/r2_probe_ab_push_flow.txt -- while the other
    // instantiations are equal or up to 2 % faster with it)
    const int Dp = (FULL && !(FLOW && PUSH)) ? 4 * LPP * VPL : g.Dp;
    ln.init(pm.c, g.D, Dp, pm.seg0);

    const unsigned row_floats = (unsigned)PX * (unsigned)Dp;          // one plane's row segment
    const unsigned seg_floats = SL::seg_bytes(Dp) / 4u;                  // ... padded to 128 bytes inside a stage
    const unsigned stage_bytes = SL::stage_bytes(Dp, DT_LOAD);
    const unsigned planes_bytes = SL::planes_bytes(Dp, DT_LOAD);
    // output staging for the TMA stores: [plane W, N, E][2 buffers] of TX pixels, 128-byte aligned
    const unsigned ost_bytes = (((unsigned)TX * (unsigned)Dp * 4u) + 127u) & ~127u;
    unsigned char *ost = smem_raw + (size_t)stages * stage_bytes;
    float4 *exch = reinterpret_cast<float4 *>(ost + 6u * ost_bytes);
    uint64_t *full = reinterpret_cast<uint64_t *>(exch + 2 * VPL * SLOTS);
    auto ex = [&](int par, int v, int slot) -> float4 & { return exch[(par * VPL + v) * SLOTS + slot]; };

#ifdef HSM_EXPERIMENT_ALIAS_FRAMES
    const int frame_mem = frame & 1;        // timing experiment only: every frame reads / writes frames 0, 1
#else
    const int frame_mem = frame;
#endif
    // The first iteration after a message reset reads all-zero W, N, E (mrf.py:43-48).  Instead of zeroing the
    // planes and reading them back, its loads name a frame outside the tensor: the TMA unit fills the stage with
    // zeros without touching memory.
    const int frame_ld = zero_in ? -1 : frame_mem;
    const int xs0 = bx * TX;
    const int shift_l = (xs0 - 1) & 3, shift_r = (xs0 - SL::DCOV) & 3;    // see StageLayout
    const int x = xs0 - 1 + px;
    const bool in_x = (x >= 0 && x < g.W);
    const int yb0 = a.olo + by * a.band_rows;
    const int yb1 = min(yb0 + a.band_rows - 1, a.ohi);
    const size_t fplane = (size_t)frame * g.plane_stride;
    const unsigned row_stride = (unsigned)g.W * (unsigned)Dp;
    const int n_rows = yb1 - yb0 + 3;                                    // rows yb0-1 .. yb1+1

    const bool st_w = pm.real && px >= 2;                      // W'(y, x-1)  -> staging slot px-2
    const bool st_n = pm.real && px >= 1 && px <= PX - 2;      // N'(y-1, x)  -> staging slot px-1   (and S')
    const bool st_e = pm.real && px <= PX - 3;                 // E'(y-1, x+1)-> staging slot px
    const bool st_s = in_x && st_n;                            // S' goes out with plain stores (last iteration only)
    const int rd_slot = (x < 0) ? (PM::LANES + LPP + ln.c) : (pm.slot + LPP);
    // W' exchange slots of this lane as shared-space addresses: [parity][v] are constant offsets from these
    const unsigned ex_wr_addr = smem_u32(exch) + 16u * (unsigned)pm.slot;
    unsigned ex_rd_addr = smem_u32(exch) + 16u * (unsigned)rd_slot;
    // keep it in a register: left alone, the compiler re-derives the slot from the thread index in every row
    asm volatile("" : "+r"(ex_rd_addr));
    constexpr unsigned EX_V = SLOTS * 16u, EX_PAR = VPL * SLOTS * 16u;

    if (first_task) {
        if (tid == 0) {
            for (int s = 0; s < stages; ++s) mbar_init(&full[s], 1);
            mbar_fence_init();
        }
        if (tid < 2 * LPP) {
#pragma unroll
            for (int v = 0; v < VPL; ++v) ex(0, v, PM::LANES + tid) = ex(1, v, PM::LANES + tid) = zero4();
        }
        __syncthreads();
    }

    // producer side (thread 0): one bulk tensor copy per plane into stage s for row index r
    auto issue = [&](int r, int s) {
        const int yy = yb0 - 1 + r - g.vlo;            // row coordinate inside the valid-row tensor
        unsigned char *stage_ptr = smem_raw + (size_t)s * stage_bytes;
        float *dst = reinterpret_cast<float *>(stage_ptr);
        unsigned tx = NPL * row_floats * 4u;
        if (!DT_LOAD) tx += SL::LBOX * 4u + SL::NRB * SL::RBOX * 4u + (a.has_prior ? SL::LBOX * 4u : 0u);
        mbar_expect_tx(&full[s], tx);
        if (L2_HINTS && (a.l2_hints & 3) != 0) {
            // rows shared with the band above (read again by it one band later) stay in L2; the rest streams
            const bool keep = r < 2 && yb0 > a.olo;
            const bool shared_below = r >= n_rows - 2 && yb1 < a.ohi;
            if (keep || shared_below || (a.l2_hints & 2) != 0) {
                const uint64_t policy = keep ? kL2EvictLast : kL2EvictFirst;
                tma_load_4d_hint(dst, tmW, &full[s], 0, xs0 - 1, yy, frame_ld, policy);
                tma_load_4d_hint(dst + seg_floats, tmN, &full[s], 0, xs0 - 1, yy, frame_ld, policy);
                tma_load_4d_hint(dst + 2 * seg_floats, tmE, &full[s], 0, xs0 - 1, yy, frame_ld, policy);
            } else {
                tma_load_4d(dst, tmW, &full[s], 0, xs0 - 1, yy, frame_ld);
                tma_load_4d(dst + seg_floats, tmN, &full[s], 0, xs0 - 1, yy, frame_ld);
                tma_load_4d(dst + 2 * seg_floats, tmE, &full[s], 0, xs0 - 1, yy, frame_ld);
            }
        } else {
            tma_load_4d(dst, tmW, &full[s], 0, xs0 - 1, yy, frame_ld);
            tma_load_4d(dst + seg_floats, tmN, &full[s], 0, xs0 - 1, yy, frame_ld);
            tma_load_4d(dst + 2 * seg_floats, tmE, &full[s], 0, xs0 - 1, yy, frame_ld);
        }
        if (DT_LOAD) {
            tma_load_4d(dst + 3 * seg_floats, tmD, &full[s], 0, xs0 - 1, yy, frame);
        } else {
            unsigned char *img = stage_ptr + planes_bytes;
            tma_load_3d(img, tmL, &full[s], xs0 - 1 - shift_l, yy, frame);
#pragma unroll
            for (int b = 0; b < SL::NRB; ++b)
                tma_load_3d(img + SL::L_BYTES + b * SL::RBOX * 4, tmR, &full[s],
                            xs0 - SL::DCOV - shift_r + b * SL::RBOX, yy, frame);
            if (a.has_prior) tma_load_3d(img + SL::L_BYTES + SL::R_BYTES, tmP, &full[s], xs0 - 1 - shift_l, yy, frame);
        }
    };
    // The same loads in the steady state, dealt to lane 0 of warps 1..4 (warp 0 keeps the stores): the issuing code --
    // ~20 instructions per bulk copy through the uniform datapath -- follows the row's barrier, and with one thread
    // doing all of it (~130 instructions against ~230 of a row's arithmetic) warp 0 reached every barrier last.
    // The barrier's single arrival (arrive.expect_tx with the stage's byte count) is warp 1's; a copy issued by
    // another warp may complete before it -- the transaction count simply goes negative until then.
    auto issue_split = [&](int r, int s) {
        const int w = tid >> 5;
        if ((tid & 31) != 0 || w < 1 || w > 4) return;
        const int yy = yb0 - 1 + r - g.vlo;
        unsigned char *stage_ptr = smem_raw + (size_t)s * stage_bytes;
        float *dst = reinterpret_cast<float *>(stage_ptr);
        if (w == 1) {
            unsigned tx = NPL * row_floats * 4u;
            if (!DT_LOAD) tx += SL::LBOX * 4u + SL::NRB * SL::RBOX * 4u + (a.has_prior ? SL::LBOX * 4u : 0u);
            mbar_expect_tx(&full[s], tx);
            if (L2_HINTS && (a.l2_hints & 3) != 0) {
                const bool keep = r < 2 && yb0 > a.olo;
                const bool shared_below = r >= n_rows - 2 && yb1 < a.ohi;
                if (keep || shared_below || (a.l2_hints & 2) != 0)
                    tma_load_4d_hint(dst, tmW, &full[s], 0, xs0 - 1, yy, frame_ld, keep ? kL2EvictLast : kL2EvictFirst);
                else
                    tma_load_4d(dst, tmW, &full[s], 0, xs0 - 1, yy, frame_ld);
            } else {
                tma_load_4d(dst, tmW, &full[s], 0, xs0 - 1, yy, frame_ld);
            }
        } else if (w == 2 || w == 3) {
            const CUtensorMap *map = w == 2 ? tmN : tmE;
            float *to = dst + (w == 2 ? seg_floats : 2 * seg_floats);
            if (L2_HINTS && (a.l2_hints & 3) != 0) {
                const bool keep = r < 2 && yb0 > a.olo;
                const bool shared_below = r >= n_rows - 2 && yb1 < a.ohi;
                if (keep || shared_below || (a.l2_hints & 2) != 0)
                    tma_load_4d_hint(to, map, &full[s], 0, xs0 - 1, yy, frame_ld, keep ? kL2EvictLast : kL2EvictFirst);
                else
                    tma_load_4d(to, map, &full[s], 0, xs0 - 1, yy, frame_ld);
            } else {
                tma_load_4d(to, map, &full[s], 0, xs0 - 1, yy, frame_ld);
            }
        } else {
            if (DT_LOAD) {
                tma_load_4d(dst + 3 * seg_floats, tmD, &full[s], 0, xs0 - 1, yy, frame);
            } else {
                unsigned char *img = stage_ptr + planes_bytes;
                tma_load_3d(img, tmL, &full[s], xs0 - 1 - shift_l, yy, frame);
#pragma unroll
                for (int b = 0; b < SL::NRB; ++b)
                    tma_load_3d(img + SL::L_BYTES + b * SL::RBOX * 4, tmR, &full[s],
                                xs0 - SL::DCOV - shift_r + b * SL::RBOX, yy, frame);
                if (a.has_prior) tma_load_3d(img + SL::L_BYTES + SL::R_BYTES, tmP, &full[s], xs0 - 1 - shift_l, yy, frame);
            }
        }
    };
    if (!FLOW) {
        pdl_launch_dependents();
        pdl_wait();                             // everything above overlaps the previous iteration's tail
    }
    if (tid == 0) {
        const int first = n_rows < stages ? n_rows : stages;
        int s_at = stage;
        for (int s = 0; s < first; ++s) {
            issue(s, s_at);
            if (++s_at == stages) s_at = 0;
        }
        // dataflow kernel: the previous tile of this CTA is published only now, so that the completion
        // of its stores overlaps this tile's dependency wait and first loads
        if (FLOW && publish && publish->local) flow_publish<PUSH>(*publish);
    }

    float *Sout = a.Sout + fplane;
    // this lane's float4 slot inside a staging buffer (pixel slot 0 <-> column xs0)
    float4 *st_wp = reinterpret_cast<float4 *>(ost) + (px - 2) * (Dp / 4) + ln.c;
    float4 *st_np = reinterpret_cast<float4 *>(ost + 2u * ost_bytes) + (px - 1) * (Dp / 4) + ln.c;
    float4 *st_ep = reinterpret_cast<float4 *>(ost + 4u * ost_bytes) + px * (Dp / 4) + ln.c;
    const unsigned ost_f4 = ost_bytes / 16u;
    auto stage_out = [&](float4 *slot, int buf, const Vec<VPL> &val) {
#pragma unroll
        for (int v = 0; v < VPL; ++v)
            if (FULL || ln.active[v]) slot[buf * ost_f4 + v * LPP] = val.q[v];
    };
    // thread 0: one bulk tensor store per plane-row; out-of-image columns are clipped by the TMA unit
    // `push` = the neighbours' ghost-row maps of the same plane: a boundary row also goes there
    auto tma_store_row = [&](const CUtensorMap *map, const CUtensorMap *push, int plane, int buf, int y) {
        // (l2_hints bit 2: the new planes are read again a whole iteration later -- let them leave L2 early)
        if (L2_HINTS && (a.l2_hints & 4)) tma_store_4d_hint(map, ost + (2u * plane + buf) * ost_bytes, 0, xs0, y - g.vlo, frame_mem, kL2EvictFirst);
        else tma_store_4d(map, ost + (2u * plane + buf) * ost_bytes, 0, xs0, y - g.vlo, frame_mem);
#ifndef HSM_PUSH_PLAIN_STORES
        if constexpr (PUSH) {
            if (y == a.olo && pp.W[0] != nullptr) tma_store_4d(&push[0], ost + (2u * plane + buf) * ost_bytes, 0, xs0, 0, 0);
            if (y == a.ohi && pp.W[1] != nullptr) tma_store_4d(&push[1], ost + (2u * plane + buf) * ost_bytes, 0, xs0, 0, 0);
        }
#endif
    };

    // ---- consumer side ------------------------------------------------------------------------
    // Carried between rows (two copies each, alternating per row so that no register moves are
    // needed): S'(y, x), X(y-1, x) = S'(y-1, x) + W'(y-1, x), data term of row y-1.
    Vec<VPL> S0, S1, X0, X1, D0, D1;
#pragma unroll
    for (int v = 0; v < VPL; ++v) S0.q[v] = S1.q[v] = X0.q[v] = X1.q[v] = D0.q[v] = D1.q[v] = zero4();

    const unsigned lane_f4 = (unsigned)px * (unsigned)(Dp / 4) + (unsigned)ln.c;   // float4 index in a row segment
    const unsigned plane_f4 = seg_floats / 4;
    const float4 *lane_base = reinterpret_cast<const float4 *>(smem_raw) + lane_f4;
    // shared-space addresses of this lane's image / prior entries inside stage 0 (see data_vec_lds)
    const unsigned img_addr = smem_u32(smem_raw) + planes_bytes;
    const unsigned l_addr = img_addr + 4u * (unsigned)(shift_l + px);
    const unsigned r_addr = img_addr + SL::L_BYTES + 4u * (unsigned)(shift_r + px - 1 + SL::DCOV - ln.chunk_offset(0));
    const unsigned p_addr = img_addr + SL::L_BYTES + SL::R_BYTES + 4u * (unsigned)(shift_l + px);
    // element offset of (row y, this lane's pixel and chunk) inside the frame; advanced by one row per step
    // (x may be -1 or W for halo lanes: the wrap-around cancels when the neighbour offset is added)
    unsigned row_off = (unsigned)yb0 * row_stride + (unsigned)x * (unsigned)Dp + 4u * (unsigned)ln.c;

    const unsigned border_warp = DT_LOAD ? 0u : data_border_mask<LPP, VPL, FULL>(ln, in_x ? x : -1, g.D);
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
        if (!DT_LOAD)
            data_vec_lds<LPP, VPL, FULL>(d, ln, l_addr + stage_off, r_addr + stage_off, p_addr + stage_off,
                                         in_x ? x : -1, g.D, a.blend, a.has_prior != 0, border_warp);
    };
    // every thread has copied the current stage: refill it with row index r + stages
    auto refill = [&](int r) {
        if (SPLIT_ISSUE) { if (r + stages < n_rows) issue_split(r + stages, stage); }
        else if (tid == 0 && r + stages < n_rows) issue(r + stages, stage);
        if (++stage == stages) { stage = 0; phase ^= 1u; }
    };
    auto store_at = [&](float *plane, unsigned off, const Vec<VPL> &val) {
#pragma unroll
        for (int v = 0; v < VPL; ++v)
            if (FULL || ln.active[v]) st_plain(plane + off + 4 * LPP * v, val.q[v]);
    };
    // peer-to-peer halo push: row `row` of a new plane also goes to the neighbouring GPU's ghost row
    // (plain stores through the peer mapping; the TMA store maps only cover this GPU's planes)
#ifdef HSM_PUSH_PLAIN_STORES                // emergency build: per-thread peer stores instead of the bulk tensor push
    const unsigned lane_col = (unsigned)x * (unsigned)Dp + 4u * (unsigned)ln.c;
    auto push_row = [&](float *const (&peer)[2], int row, int dx, bool mine, const Vec<VPL> &val) {
        if (!mine) return;
        const unsigned col = lane_col + (unsigned)(dx * Dp);
        if (row == a.olo && peer[0] != nullptr) store_at(peer[0], pp.off[0] + col, val);
        if (row == a.ohi && peer[1] != nullptr) store_at(peer[1], pp.off[1] + col, val);
    };
#endif

    // row yb0-1 only contributes S'(yb0)
    {
        Vec<VPL> w, n, e, d;
#pragma unroll
        for (int v = 0; v < VPL; ++v) d.q[v] = zero4();
        fetch(w, n, e, d);
#pragma unroll
        for (int v = 0; v < VPL; ++v) S0.q[v] = sum4(w.q[v], n.q[v], e.q[v], d.q[v]);
        normalise<LPP, VPL, FULL>(S0, ln);
        if (STORE_S && st_s) store_at(Sout, row_off, S0);
        __syncthreads();
        refill(0);
    }

    // Fused readout (finalising launch of the one-launch-per-iteration kernel only, S_MODE == 1).  The energy of
    // mrf.py:103, (((S + W) + N) + E) + Dt, of pixel (y-1, x) is ((X(y-1) + N'(y-1, x)) + E'(y-1, x)) + Dt(y-1): the
    // first three terms are in this thread's registers at step y (X = S' + W' is the carried sum, and X + N' is an
    // intermediate of U_E anyway); E'(y-1, x) is computed at the same step by the thread of the pixel to the left
    // and passes through the TMA staging buffer, readable after the next row's barrier.  So the partial sum and the
    // data term wait one step, then argmin -> int64 label, and the separate pass over all four planes
    // (k_readout: 16 B per element read again) disappears.
    constexpr bool READOUT = (S_MODE == 1) && !FLOW;
    Vec<VPL> Ppend, Dpend;
#pragma unroll
    for (int v = 0; v < VPL; ++v) Ppend.q[v] = Dpend.q[v] = zero4();
    long long *label_row = nullptr;
    if (READOUT && a.labels != nullptr)
        label_row = a.labels + ((size_t)frame * g.H + (size_t)(yb0 - 2)) * g.W + x;     // row of step yb0's (absent) label
    // label of the pixel whose partial energy is pending: E' from staging buffer `buf`, slot of column x
    // (called by EVERY thread of the CTA: the argmin exchanges values with full-warp shuffles; a lane whose pixel
    // is not owned -- the halo column px = 0 has no slot to its left -- reads its own slot and stores nothing)
    const int e_back = px >= 1 ? Dp / 4 : 0;
    auto emit_label = [&](int buf) {
        Vec<VPL> en;
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            const float4 e4 = (FULL || ln.active[v]) ? st_ep[buf * ost_f4 + v * LPP - e_back] : zero4();
            en.q[v] = add4(add4(Ppend.q[v], e4), Dpend.q[v]);
        }
        float best;
        int arg;
        lane_argmin<LPP, VPL>(en, ln, best, arg);
        pixel_argmin<LPP, VPL>(best, arg, ln);
        if (ln.c == 0 && st_s) *label_row = (long long)arg;
    };

    // One row: Sin = S'(y, x); Xin, Din belong to row y-1; Sout_/Xout/Dout are produced for the next row.
    // The exchange of W' uses a split barrier: arrive after publishing W'(y, x-1), compute S'(y+1)
    // (which does not depend on the neighbour), then wait and pick up W'(y, x).
    auto step = [&](int y, const int par, Vec<VPL> &Sin, Vec<VPL> &Snext, const Vec<VPL> &Xin, Vec<VPL> &Xout,
                    const Vec<VPL> &Din, Vec<VPL> &Dout) {
        Vec<VPL> w, n, e;
        fetch(w, n, e, Dout);
        if (y > g.vhi) Sin = zero_vec_call<VPL>();     // block-uniform, one step per frame: the row below the image sends nothing
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
        for (int v = 0; v < VPL; ++v)
            if (PM::POW2 || pm.real) sts_f4(ex_wr_addr + par * EX_PAR + v * EX_V, t.q[v]);
        if (st_w) stage_out(st_wp, par, t);
#ifdef HSM_PUSH_PLAIN_STORES
        if constexpr (PUSH) push_row(pp.W, y, -1, st_w && (x - 1) < g.W && y <= yb1, t);
#endif
        if (STORE_S) {                         // block-uniform: the last iteration only
            if (st_s && y + 1 <= yb1) store_at(Sout, row_off + row_stride, Snext);
        }
        // staging writes (this row's W', the previous row's N', E') -> visible to the async proxy
        tma_store_fence();
        if (tid == 0) tma_store_wait_read();   // the buffers about to be reused have been read by the TMA unit
        __syncthreads();                       // the row's only CTA barrier: W' published, stage consumed
        // thread 0: refill the stage every thread has copied (row index r + stages), send this row's results
        if (SPLIT_ISSUE) {
            const int r = y - yb0 + 1;
            if (r + stages < n_rows) issue_split(r + stages, stage);
        }
        if (tid == 0) {
            const int r = y - yb0 + 1;
            if (!SPLIT_ISSUE && r + stages < n_rows) issue(r + stages, stage);
            if (y <= yb1) tma_store_row(tsW, PUSH ? pmaps->W : nullptr, 0, par, y);
            if (y > yb0 + 1) {
                tma_store_row(tsN, PUSH ? pmaps->N : nullptr, 1, par ^ 1, y - 2);
                tma_store_row(tsE, PUSH ? pmaps->E : nullptr, 2, par ^ 1, y - 2);
            }
            tma_store_commit();
        }
        if (++stage == stages) { stage = 0; phase ^= 1u; }
        if (READOUT) {                         // row y-2: its E' was staged at the previous step (buffer par ^ 1)
            if (label_row != nullptr && y >= yb0 + 2) emit_label(par ^ 1);      // (CTA-uniform condition)
        }
        // X = S'(y, x) + W'(y, x)
#pragma unroll
        for (int v = 0; v < VPL; ++v) Xout.q[v] = add4(Sin.q[v], lds_f4(ex_rd_addr + par * EX_PAR + v * EX_V));
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
        if (upper) {
            if (st_n) stage_out(st_np, par, tn);
            if (st_e) stage_out(st_ep, par, t);
#ifdef HSM_PUSH_PLAIN_STORES
            if constexpr (PUSH) push_row(pp.N, y - 1, 0, st_n && in_x, tn);
            if constexpr (PUSH) push_row(pp.E, y - 1, +1, st_e && (x + 1) < g.W, t);
#endif
        }
        if (READOUT) {                         // partial energy of row y-1, completed at the next step
#pragma unroll
            for (int v = 0; v < VPL; ++v) { Ppend.q[v] = add4(Xin.q[v], tn.q[v]); Dpend.q[v] = Din.q[v]; }
            if (label_row != nullptr) label_row += g.W;
        }
        row_off += row_stride;
    };

    for (int y = yb0;; y += 2) {
        step(y, 0, S0, S1, X0, X1, D0, D1);
        if (y == yb1 + 1) break;
        step(y + 1, 1, S1, S0, X1, X0, D1, D0);
        if (y == yb1) break;
    }
    // the last step staged N'(yb1), E'(yb1) in buffer parity of row yb1+1
    tma_store_fence();
    if (tid == 0) tma_store_wait_read();
    __syncthreads();
    if (READOUT) {                             // row yb1: its E' was staged by the last step
        if (label_row != nullptr) emit_label((yb1 + 1 - yb0) & 1);
    }
    if (tid == 0) {
        const int last_par = (yb1 + 1 - yb0) & 1;
        tma_store_row(tsN, PUSH ? pmaps->N : nullptr, 1, last_par, yb1);
        tma_store_row(tsE, PUSH ? pmaps->E : nullptr, 2, last_par, yb1);
        tma_store_commit();
        tma_store_wait_read();                  // (dataflow kernel: completion is awaited by flow_publish)
    }
}

template <int LPP, int VPL, int NT, bool DT_LOAD, bool FULL, bool PUSH, bool STORE_S>
__global__ void __launch_bounds__(NT, NT <= 320 ? (VPL == 1 ? HSM_TMA_MIN_BLOCKS : (VPL == 2 ? 2 : 1)) : 1)
    k_iter_tmast(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmN,
               const __grid_constant__ CUtensorMap tmE, const __grid_constant__ CUtensorMap tmD,
               const __grid_constant__ CUtensorMap tmL, const __grid_constant__ CUtensorMap tmR,
               const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tsW,
               const __grid_constant__ CUtensorMap tsN, const __grid_constant__ CUtensorMap tsE,
               const __grid_constant__ PushMaps pmaps, Geo g, IterArgs a, int stages)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    int stage = 0;
    uint32_t phase = 0;
    PushPtrs pp;
    for (int side = 0; side < 2; ++side) {
        pp.W[side] = a.pushW[side]; pp.N[side] = a.pushN[side]; pp.E[side] = a.pushE[side];
        pp.off[side] = a.push_off[side];
    }
    tmast_task<LPP, VPL, NT, DT_LOAD, FULL, PUSH, false, STORE_S ? 1 : 0>(
        &tmW, &tmN, &tmE, &tmD, &tmL, &tmR, &tmP, &tsW, &tsN, &tsE, g, a, stages, (int)blockIdx.x, (int)blockIdx.y,
        a.frame0 + (int)blockIdx.z, STORE_S, smem_raw, stage, phase, true, pp, a.zero_in != 0, nullptr, &pmaps);
    // a CTA that pushed rows to a neighbour leaves only when those stores have completed
    if (PUSH && threadIdx.x == 0) tma_store_wait_all();
}

}  // namespace hsm
