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
};

struct FlowArgs {
    unsigned *counter;                          // next task index
    unsigned *done;                             // [frames][bands][strips]: iterations published by that tile
    int n_iter, frames, bands, strips;
    int cur;                                    // plane generation iteration 0 reads
    int finalize;                               // the last iteration also stores S'
};

__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

#ifndef HSM_FLOW_TIMEOUT_NS
#define HSM_FLOW_TIMEOUT_NS 4000000000ull       // a dependency that takes 4 s is a bug: trap instead of hanging
#endif

template <int LPP, int VPL, int NT, bool DT_LOAD, bool FULL>
__global__ void __launch_bounds__(NT, NT == 256 ? (VPL == 1 ? HSM_TMA_MIN_BLOCKS : (VPL == 2 ? 2 : 1)) : 1)
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
        constexpr int TX = NT / LPP - 2, SLOTS = NT + 2 * LPP;
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
    // the tile this CTA finished last and has not published yet (see flow_publish)
    unsigned *pending = nullptr;
    unsigned pending_value = 0;
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
        const int by = (int)(rem / (unsigned)f.strips);
        const int bx = (int)(rem - (unsigned)by * (unsigned)f.strips);
        unsigned *tile = f.done + (size_t)frame * per_frame;
        // a tile of a later iteration may depend on the unpublished one (possibly only on it): publish first
        if (pending != nullptr && it > pending_it) {
            if (tid == 0) flow_publish(pending, pending_value);
            pending = nullptr;
        }
        if (it > 0 && tid < 32) {                // warp 0: nine pollers, one neighbouring tile each
            const int nb = by + tid / 3 - 1, ns = bx + tid % 3 - 1;
            const bool polls = tid < 9 && nb >= 0 && nb < f.bands && ns >= 0 && ns < f.strips;
            const unsigned *flag = tile + (polls ? nb * f.strips + ns : 0);
            bool waits = polls && ld_acquire_gpu(flag) < (unsigned)it;
            if (__any_sync(0xffffffffu, waits)) {
                // not ready yet: do not sit on our own finished tile while we wait (thread 0 owns `pending`)
                if (tid == 0 && pending != nullptr) {
                    flow_publish(pending, pending_value);
                    pending = nullptr;
                }
                if (waits) {
                    const unsigned long long t0 = global_ns();
                    do {
                        __nanosleep(32);
                        if (global_ns() - t0 > HSM_FLOW_TIMEOUT_NS) __trap();
                    } while (ld_acquire_gpu(flag) < (unsigned)it);
                }
            }
        }
        __syncthreads();
        // the bulk tensor loads of this task (async proxy) are ordered after the acquires above
        if (tid == 0) asm volatile("fence.proxy.async;" ::: "memory");
        const int in = (f.cur + it) & 1, out = in ^ 1;
        const bool store_s = f.finalize != 0 && it == f.n_iter - 1;
        tmast_task<LPP, VPL, NT, DT_LOAD, FULL, false, true>(&maps.tmW[in], &maps.tmN[in], &maps.tmE[in], &maps.tmD,
                                                             &maps.tmL, &maps.tmR, &maps.tmP, &maps.tsW[out],
                                                             &maps.tsN[out], &maps.tsE[out], g, a, stages, bx, by, frame,
                                                             store_s, smem_raw, stage, phase, first, pending,
                                                             pending_value);
        first = false;
        pending = tile + by * f.strips + bx;
        pending_value = (unsigned)it + 1u;
        pending_it = it;
        if (store_s) __threadfence();            // S' of the last iteration went out with plain stores
        __syncthreads();
        task = s_next[slot];
        slot ^= 1;
    }
    if (pending != nullptr && tid == 0) flow_publish(pending, pending_value);
}

}  // namespace hsm
