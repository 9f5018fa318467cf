// Persistent dataflow kernel (algorithm 6): ALL iterations of an hsm_iterate call in one launch.
//
// The unit of work is a task = (iteration, frame, band, strip) of the fused iteration (tmast_task in
// iter_tma_st.cuh).  Tasks are handed out in iteration-major order by an atomic counter; a CTA runs task after
// task.  Task (k, f, b, s) reads rows/columns one beyond its tile from the planes written by iteration k-1, so
// it waits until the up to nine tiles (b-1..b+1, s-1..s+1) of frame f have published iteration k-1 (a per-tile
// counter in global memory, release/acquire at gpu scope); the same nine tasks are the last readers of the
// planes the task overwrites (double buffering), so the wait also covers the write-after-read hazard.  A task
// only ever waits for tasks that were handed out before it, so the scheme cannot deadlock whatever the number
// of resident CTAs.  What it buys: no launch per iteration, no drain at the end of every iteration (the first
// bands of iteration k+1 start while the last bands of iteration k still run) and no CTA start-up per tile.
#pragma once

#include "iter_tma_st.cuh"

namespace hsm {

struct FlowMaps {                               // 16 tensor maps = 2 KB of kernel parameters
    CUtensorMap tmW[2], tmN[2], tmE[2];         // loads from plane generation 0 / 1
    CUtensorMap tsW[2], tsN[2], tsE[2];         // stores into plane generation 0 / 1
    CUtensorMap tmD, tmL, tmR, tmP;
    PushMaps push[2];                           // the neighbours' ghost rows, per plane generation (PUSH kernels)
};

struct FlowArgs {
    unsigned *counter;                          // next task index
    unsigned *done;                             // [frames][bands][strips]: iterations published by that tile
    int n_iter, frames, bands, strips;
    int cur;                                    // plane generation iteration 0 reads
    int finalize;                               // the last iteration also stores S'
    // Row-band mode with neighbours on other GPUs (PUSH kernels): the tiles across a band boundary live in
    // another launch on another GPU.  Their progress arrives in `ghost_in` (this GPU's memory, written by the
    // neighbour; absolute iteration numbers since hsm_peer_reset, `base` = this launch's first), ours goes to
    // `ghost_out` (the neighbour's memory).  `mail_in[side]` is the neighbour's whole-iteration counter of the
    // one-launch-per-iteration path, which also satisfies a dependency.
    const unsigned *ghost_in[2];
    const unsigned *mail_in;
    unsigned *ghost_out[2];
    unsigned base;
    float *pushW[2][2], *pushN[2][2], *pushE[2][2];   // [side][plane generation]
    unsigned push_off[2];
};

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

#ifndef HSM_FLOW_TIMEOUT_NS
#define HSM_FLOW_TIMEOUT_NS 30000000000ull      // the order of the tasks rules out a deadlock; should a dependency
                                                 // still take 30 s something is badly wrong: trap instead of hanging
#endif

template <int LPP, int VPL, int NT, bool DT_LOAD, bool FULL, bool PUSH>
__global__ void __launch_bounds__(NT, NT <= 320 ? (VPL == 1 ? HSM_TMA_MIN_BLOCKS : (VPL == 2 ? 2 : 1)) : 1)
    k_iter_flow(const __grid_constant__ FlowMaps maps, Geo g, IterArgs a, FlowArgs f, int stages)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    // two task slots behind the stage barriers (tmast_task's layout; the launcher reserves 128 bytes there).
    // A static __shared__ variable would cost a 1 KB alignment pad in front of smem_raw and with it the second
    // resident CTA at D = 64.
    unsigned *s_next;
    {
        using SL = StageLayout<LPP, VPL, NT>;
        using PM = PixelMap<LPP, NT>;
        constexpr int TX = PM::PX - 2, SLOTS = PM::LANES + 2 * LPP;
        const unsigned ost_bytes = (((unsigned)TX * (unsigned)g.Dp * 4u) + 127u) & ~127u;
        unsigned char *after = smem_raw + (size_t)stages * SL::stage_bytes(g.Dp, DT_LOAD) + 6u * ost_bytes +
                               (size_t)2 * VPL * SLOTS * sizeof(float4);
        s_next = reinterpret_cast<unsigned *>(after + 64);              // full[0..7] occupy the first 64 bytes
    }
    const unsigned per_frame = (unsigned)f.bands * (unsigned)f.strips;
    const unsigned per_iter = per_frame * (unsigned)f.frames;
    const unsigned total = per_iter * (unsigned)f.n_iter;
    int stage = 0;
    uint32_t phase = 0;
    bool first = true;
    // the tile this CTA finished last and has not published yet (see flow_publish); thread 0 owns it
    FlowPending pending;
    pending.local = nullptr; pending.remote[0] = pending.remote[1] = nullptr;
    pending.local_value = pending.remote_value = 0;
    int pending_it = 0;
    if (tid == 0) s_next[0] = atomicAdd(f.counter, 1u);
    __syncthreads();
    unsigned task = s_next[0];
    int slot = 1;
    while (task < total) {
        // ask for the next task now: the round trip hides behind this one
        if (tid == 0) s_next[slot] = atomicAdd(f.counter, 1u);
        const int it = (int)(task / per_iter);
        unsigned rem = task - (unsigned)it * per_iter;
        const int frame = (int)(rem / per_frame);
        rem -= (unsigned)frame * per_frame;
        const int pos = (int)(rem / (unsigned)f.strips);       // position of the band in the hand-out order
        const int bx = (int)(rem - (unsigned)pos * (unsigned)f.strips);
        // Row bands over several GPUs: the neighbouring GPUs wait for this band's first and last tile rows, so
        // those go first and the order runs from both ends inwards (0, B-1, 1, B-2, ...); a tile's neighbours of
        // the previous iteration are still handed out about one iteration's worth of tiles before it.
        const int by = !PUSH ? pos : ((pos & 1) ? f.bands - 1 - (pos >> 1) : (pos >> 1));
        unsigned *tile = f.done + (size_t)frame * per_frame;
        // a tile of a later iteration may depend on the unpublished one (possibly only on it): publish first
        if (pending.local != nullptr && it > pending_it) {
            if (tid == 0) flow_publish<PUSH>(pending);
            pending.local = nullptr;
        }
        if (tid < 32) {                          // warp 0: nine pollers, one neighbouring tile each
            const int nb = by + tid / 3 - 1, ns = bx + tid % 3 - 1;
            const bool in_s = tid < 9 && ns >= 0 && ns < f.strips;
            const bool local = in_s && it > 0 && nb >= 0 && nb < f.bands;
            // across the band boundary: the neighbouring GPU's tile (PUSH kernels only)
            const int side = nb < 0 ? 0 : 1;
            const bool ghost = PUSH && in_s && (nb < 0 || nb >= f.bands) && f.ghost_in[side] != nullptr &&
                               f.base + (unsigned)it > 0u;
            const unsigned *flag = local ? tile + nb * f.strips + ns : (ghost ? f.ghost_in[side] + ns : f.done);
            const unsigned need = local ? (unsigned)it : f.base + (unsigned)it;
            auto ready = [&]() -> bool {
                if (local) return ld_acquire_gpu(flag) >= need;
                if (ghost) return ld_acquire_sys(flag) >= need || ld_acquire_sys(f.mail_in + side) >= need;
                return true;
            };
            const bool waits = !ready();
            if (__any_sync(0xffffffffu, waits)) {
                // not ready yet: do not sit on our own finished tile while we wait
                if (tid == 0 && pending.local != nullptr) {
                    flow_publish<PUSH>(pending);
                    pending.local = nullptr;
                }
                if (waits) {
                    const unsigned long long t0 = global_ns();
                    do {
                        __nanosleep(32);
                        if (global_ns() - t0 > HSM_FLOW_TIMEOUT_NS) __trap();
                    } while (!ready());
                }
            }
        }
        __syncthreads();
        // the bulk tensor loads of this task (async proxy) are ordered after the acquires above
        // (lane 0 of the first five warps: thread 0 issues a task's first loads and stores, warps 1..4 the later loads)
        if ((tid & 31) == 0 && tid < 160) asm volatile("fence.proxy.async;" ::: "memory");
        const int in = (f.cur + it) & 1, out = in ^ 1;
        const bool store_s = f.finalize != 0 && it == f.n_iter - 1;
        PushPtrs pp;
#pragma unroll
        for (int side = 0; side < 2; ++side) {
            pp.W[side] = PUSH ? f.pushW[side][out] : nullptr;
            pp.N[side] = PUSH ? f.pushN[side][out] : nullptr;
            pp.E[side] = PUSH ? f.pushE[side][out] : nullptr;
            pp.off[side] = f.push_off[side];
        }
        tmast_task<LPP, VPL, NT, DT_LOAD, FULL, PUSH, true, 2>(&maps.tmW[in], &maps.tmN[in], &maps.tmE[in], &maps.tmD,
                                                            &maps.tmL, &maps.tmR, &maps.tmP, &maps.tsW[out],
                                                            &maps.tsN[out], &maps.tsE[out], g, a, stages, bx, by, frame,
                                                            store_s, smem_raw, stage, phase, first, pp,
                                                            a.zero_in != 0 && it == 0, &pending, &maps.push[out]);
        first = false;
        pending.local = tile + by * f.strips + bx;
        pending.local_value = (unsigned)it + 1u;
        pending.remote_value = f.base + (unsigned)it + 1u;
        pending.remote[0] = (PUSH && by == 0 && f.ghost_out[0] != nullptr) ? f.ghost_out[0] + bx : nullptr;
        pending.remote[1] = (PUSH && by == f.bands - 1 && f.ghost_out[1] != nullptr) ? f.ghost_out[1] + bx : nullptr;
        pending_it = it;
        // plain stores of every thread: S' of the last iteration, boundary rows pushed to the neighbours
        if (PUSH && (pending.remote[0] != nullptr || pending.remote[1] != nullptr)) __threadfence_system();
        else if (store_s) __threadfence();
        __syncthreads();
        task = s_next[slot];
        slot ^= 1;
    }
    if (pending.local != nullptr && tid == 0) flow_publish<PUSH>(pending);
}

}  // namespace hsm
