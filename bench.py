"""Benchmark of the frame-based MRF stereo hot path on B200 ("Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): synthetic DAVIS 240x180 frame pairs, D = 32, SNN-like event
prior blended into the data term ('adaptive', trust 1.0 -- the reference's offline call pattern,
stereo_framebased.py:84-85), 50 BP iterations.  One "step" = one full lbp (cost volume + 50
iterations + argmin readout) over a batch of independent frame pairs; with N GPUs every rank
processes its own batch of the frame sequence (frames sharded, no collective: weak scaling).

Printed JSON (rank 0, one line):
  value   frames/s, whole job, inputs (float64 images + float64 prior, the reference pipeline's
          dtypes) already resident in HBM, labels (int64) left in HBM; CUDA events, max over ranks.
  e2e     the same metric through the public API (StereoMRF.lbp_batch) with HOST buffers: H2D of the
          inputs and D2H of the labels inside the timed region.
  roofline  message kernel: algorithmic bytes per iteration (24 B x H x W x D x frames; an iteration is
          two concurrent launches, one per half of the frames, that share the HBM) / measured device
          time per iteration (CUDA events on the launching stream) against the measured HBM peak.
  cpu_baseline  the numpy oracle (a restatement of the reference's numpy code) on the host, one process.

--impl reference times the reference's CPU algorithm (the oracle port; the reference is Python 2 and
cannot travel to the GPU box) with every host core, frames spread over a process pool.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ROWS, COLS, LEVELS, N_ITER = 180, 240, 32, 50
TRUST, MODE = 1.0, "adaptive"
WORKLOAD = "synthetic DAVIS 240x180, D=32, SNN event-disparity prior ('adaptive', trust 1.0), 50 BP iterations"
METRIC = "stereo frames/sec (50 BP iters) at 240x180xD32"
FALLBACK_HBM_GBS = 6650.0


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as handle:
            return float(json.load(handle)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def recorded_traffic():
    """DRAM bytes per launch of the message kernel from the committed ncu capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as handle:
            return json.load(handle)
    except Exception:
        return None


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, device):
        self.device, self.lines, self.proc = device, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                universal_newlines=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def wait_ready(self, timeout=8.0):
        """Block until nvidia-smi delivered its first sample (it can take a second to start on an
        8-GPU box), then remember where the timed region's samples begin."""
        deadline = time.time() + timeout
        while self.proc is not None and not self.lines and time.time() < deadline:
            time.sleep(0.05)
        self.first = len(self.lines)

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, sm_max, power, reasons = [], [], [], set()
        first = getattr(self, "first", 0)
        lines = self.lines[first:] if len(self.lines) > first else self.lines[-1:]
        for line in lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); sm_max.append(float(parts[1])); power.append(float(parts[2]))
            except ValueError:
                continue
            for name, flag in zip(self.NAMES, parts[3:7]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(sm_max), "reasons": sorted(reasons),
                "power_w_max": max(power), "samples": len(sm)}


def make_inputs(n_frames, seed0):
    from importlib import import_module
    synth = import_module("hybrid-stereo-matching_b200.synth")
    return synth.synthetic_sequence(n_frames, ROWS, COLS, LEVELS, base_seed=seed0, with_prior=True)


# ---------------------------------------------------------------------------------- CPU arms
def _cpu_one_frame(args):
    from oracle import mrf_numpy
    left, right, prior = args
    matcher = mrf_numpy.OracleMRF((ROWS, COLS), LEVELS)
    return matcher.lbp(left, right, prior, prior_trust_factor=TRUST, prior_influence_mode=MODE, n_iter=N_ITER)


def cpu_baseline(n_frames=10):
    """numpy oracle, one process (numpy's element-wise kernels use one core)."""
    lefts, rights, priors = make_inputs(n_frames + 1, 1234)
    _cpu_one_frame((lefts[n_frames], rights[n_frames], priors[n_frames]))          # warm-up
    start = time.perf_counter()
    labels = [_cpu_one_frame((lefts[i], rights[i], priors[i])) for i in range(n_frames)]
    elapsed = time.perf_counter() - start
    return {"value": n_frames / elapsed, "unit": "frames/s", "cores": 1, "kind": "port",
            "sample": "%d frames of the workload, sequential, oracle/mrf_numpy.py (%.1f s)" % (n_frames, elapsed)}, \
        np.stack(labels)


def run_reference_arm(args, rank, world):
    """--impl reference: the reference's CPU algorithm on every host core."""
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    per_step = cores
    lefts, rights, priors = make_inputs(per_step, 1234)
    jobs = [(lefts[i], rights[i], priors[i]) for i in range(per_step)]
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        for _ in range(args.warmup):
            pool.map(_cpu_one_frame, jobs[:cores])
        start = time.perf_counter()
        for _ in range(args.steps):
            pool.map(_cpu_one_frame, jobs)
        elapsed = time.perf_counter() - start
    value = per_step * args.steps / elapsed
    sample = "%d frames per step (one per core), process pool over %d cores, oracle/mrf_numpy.py" % (per_step, cores)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_step": per_step},
            "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------- GPU arm
def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--gpus", type=int, default=1)
    parser.add_argument("--steps", type=int, default=10)
    parser.add_argument("--warmup", type=int, default=3)
    parser.add_argument("--impl", default="b200", choices=["b200", "reference"])
    parser.add_argument("--frames-per-step", type=int, default=256, help="frame pairs per GPU per step")
    parser.add_argument("--capacity", type=int, default=128, help="frame pairs per launch (device batch)")
    parser.add_argument("--cpu-frames", type=int, default=10, help="frames of the CPU baseline sample (0 = skip)")
    args = parser.parse_args()
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return 0

    import torch
    import torch.distributed as dist
    import hybrid_stereo_matching_b200 as hsm
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    frames, cap = args.frames_per_step, args.capacity
    chunks = [(s, min(s + cap, frames)) for s in range(0, frames, cap)]
    # frames of the sequence are dealt to ranks round-robin: rank r owns frames r, r + world, ...
    uniq = min(frames, 32)                      # distinct synthetic pairs per rank, tiled to the batch
    lefts, rights, priors = make_inputs(uniq, 1234 + rank * 100003)
    reps = -(-frames // uniq)
    tile = lambda a: np.ascontiguousarray(np.tile(a, (reps, 1, 1))[:frames])
    lefts, rights, priors = tile(lefts), tile(rights), tile(priors)

    matcher = hsm.StereoMRF((ROWS, COLS), LEVELS, capacity=cap, device=local_rank)
    stream = torch.cuda.Stream()                 # torch events below are recorded on this stream
    matcher.set_stream(stream.cuda_stream)

    # ---- device-resident arm ------------------------------------------------------------------
    d_left = torch.from_numpy(lefts).cuda()
    d_right = torch.from_numpy(rights).cuda()
    d_prior = torch.from_numpy(priors).cuda()
    d_out = torch.empty((frames, ROWS, COLS), dtype=torch.int64, device="cuda")

    def device_step():
        for start, stop in chunks:
            matcher.lbp_batch_device(d_left[start:stop], d_right[start:stop], d_prior[start:stop],
                                     d_out[start:stop], prior_trust_factor=TRUST, prior_influence_mode=MODE,
                                     n_iter=N_ITER, synchronize=False)

    for _ in range(max(args.warmup, 3)):
        device_step()
    barrier()
    matcher.reset_stats()
    begin, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        if rank == 0:
            clocks.wait_ready()
        barrier()
        begin.record(stream)
        for _ in range(args.steps):
            device_step()
        end.record(stream)
        barrier()
    elapsed_ms = begin.elapsed_time(end)
    stats = matcher.stats()
    labels_device = d_out.cpu().numpy()

    # ---- end-to-end arm: host buffers through the public API ------------------------------------
    h_left = hsm.pinned_empty(lefts.shape, lefts.dtype); h_left[...] = lefts
    h_right = hsm.pinned_empty(rights.shape, rights.dtype); h_right[...] = rights
    h_prior = hsm.pinned_empty(priors.shape, priors.dtype); h_prior[...] = priors
    h_out = hsm.pinned_empty((frames, ROWS, COLS), np.int64)
    matcher.set_stream(0)                          # back to the matcher's own stream (lanes overlap)

    def host_step():
        matcher.lbp_batch(h_left, h_right, h_prior, prior_trust_factor=TRUST, prior_influence_mode=MODE,
                          n_iter=N_ITER, out=h_out)

    for _ in range(max(args.warmup, 3)):
        host_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        host_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    same = bool(np.array_equal(h_out, labels_device))

    # ---- reduce over ranks (max time) -------------------------------------------------------------
    times = torch.tensor([elapsed_ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    elapsed_ms, e2e_ms = float(times[0]), float(times[1])

    if rank == 0:
        peak, peak_source = measured_peak()
        launches = stats["iteration_launches"]
        iterations = max(stats["iterations"], 1)
        # One BP iteration over a chunk of `cap` frames is launched as `per_iter` concurrent kernels (frame
        # groups on two streams, each filling the other's tail); they share the HBM, so the roofline unit is
        # the iteration: combined algorithmic bytes of its launches / the iteration's device time.
        per_iter = max(int(round(launches / iterations)), 1)
        frames_per_launch = cap if frames % cap == 0 else frames / len(chunks)
        algo_bytes = 24.0 * ROWS * COLS * LEVELS * frames_per_launch
        launch_ms = stats["iteration_ms"] / iterations
        achieved = algo_bytes / (launch_ms * 1e-3) / 1e9
        traffic = recorded_traffic()
        line = {
            "metric": METRIC, "value": world * frames * args.steps / (elapsed_ms * 1e-3), "unit": "frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_step_per_gpu": frames, "frames_per_launch": cap,
                       "kernels_per_iteration": per_iter,
                       "sharding": "frames of the sequence dealt to ranks, no collective",
                       "l2": "inputs larger than L2: %.1f GB of message state per launch vs 126 MB"
                             % (6 * 4.0 * ROWS * COLS * LEVELS * cap / 1e9),
                       "inputs": "float64 images + float64 prior (-1 = no event), int64 labels out"},
            "e2e": {"value": world * frames * args.steps / (e2e_ms * 1e-3), "unit": "frames/s",
                    "h2d_bytes_per_step": int(lefts.nbytes + rights.nbytes + priors.nbytes),
                    "d2h_bytes_per_step": int(h_out.nbytes), "timing": "wall clock between device synchronisations",
                    "labels_equal_device_arm": same},
            "gpu_launches": int(stats["kernel_launches"]),
            "roofline": {"bound": "hbm", "kernel": "k_iter_tmast (fused 4-sweep BP iteration, TMA loads + TMA stores); "
                                                    "unit = one iteration over frames_per_launch frames",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": peak_source, "algorithmic_bytes_per_launch": algo_bytes,
                         "launch_ms": launch_ms, "launches_timed": int(launches),
                         "concurrent_launches_per_iteration": per_iter,
                         "traffic": None if not traffic else
                         traffic["dram_bytes_per_launch"] * frames_per_launch / traffic["frames_per_launch"],
                         "traffic_source": None if not traffic else traffic.get("source")},
            "clocks": clocks.summary(),
        }
        if world == 1 and args.cpu_frames > 0:
            base, cpu_labels = cpu_baseline(args.cpu_frames)
            line["cpu_baseline"] = base
            # the first frames of rank 0's batch are the CPU sample's frames (seed 1234 + i)
            n = min(args.cpu_frames, uniq)
            line["label_agreement_pct"] = 100.0 * float((labels_device[:n] == cpu_labels[:n]).mean())
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
