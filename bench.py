"""Benchmark of the frame-based MRF stereo hot path on B200 ("Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline workload (BASELINE.json configs[1], "C2"): synthetic DAVIS 240x180 frame pairs, D = 32, SNN-like event
prior blended into the data term ('adaptive', trust 1.0 -- the reference's offline call pattern,
stereo_framebased.py:84-85), 50 BP iterations.  One "step" = one full lbp (cost volume + 50 iterations +
argmin readout) over a batch of independent frame pairs; with N GPUs every rank processes its own batch of
the frame sequence (frames sharded, no collective: weak scaling, configs[3]).

Printed JSON (rank 0, one line):
  value     frames/s, whole job, inputs (float64 images + float64 prior, the reference pipeline's dtypes)
            already resident in HBM, labels (int64) left in HBM; CUDA events, max over ranks.
  e2e       the same metric through the public API (StereoMRF.lbp_batch) with page-locked HOST buffers: H2D of
            the inputs and D2H of the labels inside the timed region.  e2e_pageable: the same call with ordinary
            numpy arrays (what the reference's callers pass), staged by the library.
  roofline  message kernel: algorithmic bytes per iteration (24 B x H x W x D x frames; an iteration is two
            concurrent launches, one per half of the frames, that share the HBM) / measured device time per
            iteration (CUDA events on the launching stream) against the measured HBM peak.
  configs   (N = 1) the other named shapes of BASELINE.json, each checked against a fixture written by the
            real reference: c1 (240x180 D32, no prior), c3 (640x480 D64, 100 iterations; one frame and a batch of
            4), c5 (1920x1080 D256, 50 iterations, one GPU), and the shipped disparity counts 20 / 22 / 27 / 50.
  bands     (N > 1) configs[4]: ONE 1920x1080 D256 frame split into row bands over the N GPUs, halo exchange
            every iteration, with the fused peer push ("p2p") and with NCCL send/recv ("nccl"): ms per
            iteration (CUDA events, max over ranks), fraction of the per-GPU roofline, and whether the
            gathered labels equal the reference's (fixture b09).
  cpu_baseline  the reference's own mrf.py (oracle/_ref mirror; else the numpy port) on one host core.

--impl reference times the reference's CPU implementation (the mirror of mrf.py written by build(); the numpy
port only if the mirror is absent) with every host core, frames spread over a process pool.
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
# identical in both arms (the driver compares the two config objects)
CONFIG = {"workload": WORKLOAD, "rows": ROWS, "cols": COLS, "n_levels": LEVELS, "n_iter": N_ITER,
          "prior": "5% event density, 20% outliers, -1 = no event; 'adaptive', trust 1.0",
          "inputs": "float64 images + float64 prior, int64 labels out",
          "sharding": "frames of the sequence dealt to ranks, no collective"}
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


def frame_bytes(rows, cols, levels, n_iter, prior=True):
    """algorithmic bytes of one lbp = n x 24 HWD + 4 HWD (final S') + images (+ prior) + labels."""
    hw = rows * cols
    return n_iter * 24.0 * hw * levels + 4.0 * hw * levels + 8.0 * hw + (4.0 * hw if prior else 0.0) + 4.0 * hw


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


def synth_module():
    from importlib import import_module
    return import_module("hybrid-stereo-matching_b200.synth")


def make_inputs(n_frames, seed0, rows=ROWS, cols=COLS, levels=LEVELS, with_prior=True):
    return synth_module().synthetic_sequence(n_frames, rows, cols, levels, base_seed=seed0, with_prior=with_prior)


# ---------------------------------------------------------------------------------- CPU arms
_CPU_CLASS = None


def cpu_matcher_class():
    """(class, kind): the reference's own StereoMRF from the oracle/_ref mirror ("reference"), else the numpy
    port ("port").  bench.py may execute oracle/ only here, for the CPU arms."""
    global _CPU_CLASS
    if _CPU_CLASS is None:
        from oracle import ref_mirror
        mirror = ref_mirror.load_mirror()
        if mirror is not None:
            _CPU_CLASS = (mirror, "reference", "reference mrf.py (oracle/_ref mirror, py3 token shim)")
        else:
            from oracle import mrf_numpy
            _CPU_CLASS = (mrf_numpy.OracleMRF, "port", "oracle/mrf_numpy.py")
    return _CPU_CLASS


def _cpu_one_frame(args):
    left, right, prior = args
    matcher = cpu_matcher_class()[0]((ROWS, COLS), LEVELS)
    return matcher.lbp(left, right, prior, prior_trust_factor=TRUST, prior_influence_mode=MODE, n_iter=N_ITER)


def cpu_baseline(n_frames=10):
    """The reference on ONE host core (numpy's element-wise kernels are single-threaded), sequential frames."""
    _, kind, what = cpu_matcher_class()
    lefts, rights, priors = make_inputs(n_frames + 1, 1234)
    _cpu_one_frame((lefts[n_frames], rights[n_frames], priors[n_frames]))          # warm-up
    start = time.perf_counter()
    labels = [_cpu_one_frame((lefts[i], rights[i], priors[i])) for i in range(n_frames)]
    elapsed = time.perf_counter() - start
    return {"value": n_frames / elapsed, "unit": "frames/s", "cores": 1, "kind": kind,
            "sample": "%d frames of the workload, sequential, %s (%.1f s)" % (n_frames, what, elapsed)}, \
        np.stack(labels)


def run_reference_arm(args, rank, world):
    """--impl reference: the reference's CPU implementation on every host core."""
    if rank != 0:
        return
    import multiprocessing as mp
    _, kind, what = cpu_matcher_class()
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
    sample = "%d frames per step (one per core), process pool over %d cores, %s" % (per_step, cores, what)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(CONFIG),
            "run": {"frames_per_step": per_step},
            "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------- fixture checks
def fixture_inputs(name):
    """(left, right, prior, kwargs, reference labels) of a golden fixture written by the real reference
    (oracle/make_golden.py).  Checker only: the GPU result of the same inputs is compared with it."""
    from oracle import cases
    case = cases.find_case(name)
    left, right, prior, kwargs = cases.case_inputs(case, 0)
    data, _ = cases.load_fixture(case)
    return left, right, prior, dict(kwargs), data["labels"].astype(np.int64)


# ---------------------------------------------------------------------------------- GPU legs
def device_leg(torch, hsm, shape, n_iter, frames, cap, steps, warmup, prior_kw, fixture=None, device=0):
    """Device-resident timing of `steps` lbp calls over `frames` frame pairs of `shape` = (rows, cols, levels):
    frames/s, ms per iteration, fraction of the HBM roofline (message kernel alone, and whole frame with
    SURVEY.md 8(d)'s A); frame 0 is a reference fixture's input when `fixture` is given and its labels are
    compared."""
    rows, cols, levels = shape
    peak, _ = measured_peak()
    with_prior = prior_kw is not None
    lefts, rights, priors = make_inputs(min(frames, 8), 4000 + rows + levels, rows, cols, levels, with_prior)
    reps = -(-frames // lefts.shape[0])
    tile = lambda a: None if a is None else np.ascontiguousarray(np.tile(a, (reps, 1, 1))[:frames])
    lefts, rights, priors = tile(lefts), tile(rights), tile(priors)
    want, fixture_iters = None, n_iter
    kw = dict(prior_kw or {})
    if fixture is not None:
        left, right, prior, fkw, want = fixture_inputs(fixture)
        lefts[0], rights[0] = left, right
        if with_prior:
            priors[0] = prior
        fixture_iters = fkw.pop("n_iter", n_iter)
        kw = fkw if with_prior else {}
    matcher = hsm.StereoMRF((rows, cols), levels, capacity=cap, device=device)
    stream = torch.cuda.Stream()
    matcher.set_stream(stream.cuda_stream)
    d_left, d_right = torch.from_numpy(lefts).cuda(), torch.from_numpy(rights).cuda()
    d_prior = torch.from_numpy(priors).cuda() if with_prior else None
    d_out = torch.empty((frames, rows, cols), dtype=torch.int64, device="cuda")
    chunks = [(s, min(s + cap, frames)) for s in range(0, frames, cap)]

    def step(iters):
        for start, stop in chunks:
            matcher.lbp_batch_device(d_left[start:stop], d_right[start:stop],
                                     None if d_prior is None else d_prior[start:stop], d_out[start:stop],
                                     n_iter=iters, synchronize=False, **kw)

    record = {"shape": "%dx%dxD%d" % (cols, rows, levels), "n_iter": n_iter, "frames_per_step": frames,
              "frames_per_launch": cap}
    if want is not None:                       # the fixture's own iteration count
        step(fixture_iters)
        torch.cuda.synchronize()
        got = d_out[0].cpu().numpy()
        record["fixture"] = fixture
        record["fixture_n_iter"] = fixture_iters
        record["labels_equal_reference"] = bool(np.array_equal(got, want))
    for _ in range(max(warmup, 3)):
        step(n_iter)
    torch.cuda.synchronize()
    matcher.reset_stats()
    begin, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    begin.record(stream)
    for _ in range(steps):
        step(n_iter)
    end.record(stream)
    torch.cuda.synchronize()
    ms = begin.elapsed_time(end) / steps
    stats = matcher.stats()
    # per-iteration device time of one launch group (all concurrent kernels of an iteration over `cap` frames)
    launch_ms = stats["iteration_ms"] / max(stats["iterations"], 1)
    per_launch_frames = frames / len(chunks)
    achieved = 24.0 * rows * cols * levels * per_launch_frames / (launch_ms * 1e-3) / 1e9
    whole = frame_bytes(rows, cols, levels, n_iter, with_prior) * frames / (ms * 1e-3) / 1e9
    record.update({"frames_per_s": frames / (ms * 1e-3), "ms_per_step": ms, "ms_per_iteration": launch_ms,
                   "kernel_gbs": achieved, "kernel_frac": achieved / peak, "frame_gbs": whole,
                   "frame_frac": whole / peak,
                   "launches_per_iteration": round(stats["iteration_launches"] / max(stats["iterations"], 1), 2)})
    del matcher, d_left, d_right, d_prior, d_out
    torch.cuda.empty_cache()
    return record


def single_gpu_configs(torch, hsm, device):
    """BASELINE.json's other named shapes on one GPU, and the reference's shipped disparity counts."""
    adaptive = dict(prior_trust_factor=1.0, prior_influence_mode="adaptive")
    out = {}
    out["c1"] = device_leg(torch, hsm, (180, 240, 32), 50, 128, 128, 5, 3, None,
                           fixture="b02_tsukuba_crop_d32_it50", device=device)
    out["c3_batch4"] = device_leg(torch, hsm, (480, 640, 64), 100, 4, 4, 5, 3, adaptive,
                                  fixture="b07_vga_d64_it100", device=device)
    out["c3_single"] = device_leg(torch, hsm, (480, 640, 64), 100, 1, 1, 5, 3, adaptive,
                                  fixture="b07_vga_d64_it100", device=device)
    out["c5_single_gpu"] = device_leg(torch, hsm, (1080, 1920, 256), 50, 1, 1, 3, 3, adaptive,
                                      fixture="b09_1080p_d256_it2", device=device)
    for levels, fixture in ((20, "b16_davis_d20_it10"), (50, "b17_davis_d50_it10")):
        out["davis_d%d_batch64" % levels] = device_leg(torch, hsm, (180, 240, levels), 50, 64, 64, 5, 3, adaptive,
                                                       fixture=fixture, device=device)
    for (rows, cols, levels), fixture in (((60, 80, 22), "b13_checkerboard_80x60_d22"),
                                          ((60, 70, 27), "b14_head_70x60_d27")):
        out["shipped_%dx%d_d%d_batch256" % (cols, rows, levels)] = device_leg(
            torch, hsm, (rows, cols, levels), 10, 256, 256, 5, 3, adaptive, fixture=fixture, device=device)
    return out


def single_frame_latency(hsm, device):
    """The reference's actual call: one lbp on host numpy arrays, wall clock, median of 20."""
    out = {}
    for name, (rows, cols, levels, n_iter) in (("online_80x60_d22_it10", (60, 80, 22, 10)),
                                               ("davis_240x180_d32_it50", (180, 240, 32, 50)),
                                               ("vga_640x480_d64_it100", (480, 640, 64, 100))):
        left, right, disparity = synth_module().textured_pair(rows, cols, levels, 7)
        prior = synth_module().sparse_prior(disparity, levels, 7)
        matcher = hsm.StereoMRF((rows, cols), levels, device=device)
        kw = dict(prior_trust_factor=1.0, prior_influence_mode="adaptive", n_iter=n_iter)
        for _ in range(5):
            matcher.lbp(left, right, prior, **kw)
        times = []
        for _ in range(20):
            t0 = time.perf_counter()
            matcher.lbp(left, right, prior, **kw)
            times.append(time.perf_counter() - t0)
        out[name] = {"lbp_ms": 1e3 * statistics.median(times), "frames_per_s": 1.0 / statistics.median(times)}
        del matcher
    return out


def band_legs(torch, dist, hsm, rank, world, local_rank, n_iter=50):
    """configs[4]: one 1920x1080 D256 frame in row bands over the ranks; both exchange schemes."""
    from importlib import import_module
    bands = import_module("hybrid-stereo-matching_b200.bands")
    rows, cols, levels = 1080, 1920, 256
    peak, _ = measured_peak()
    left, right, prior, fkw, want = fixture_inputs("b09_1080p_d256_it2")
    fixture_iters = fkw.pop("n_iter")
    first, last = bands.band_bounds(rows, world, rank)
    engine = bands.CudaBandEngine(rows, cols, levels, first, last, device=local_rank)
    results = {}
    capi = import_module("hybrid-stereo-matching_b200._capi")
    # "p2p": one launch per iteration, neighbours ordered by stream-ordered flags; "p2p_dataflow": ONE persistent
    # launch per GPU for the whole loop, boundary tiles exchange progress counters through peer memory; "nccl":
    # pack / send-recv / unpack after every iteration
    for exchange in ("p2p", "p2p_dataflow", "nccl"):
        engine.matcher.set_option(capi.OPT_ALGORITHM, 5)
        matcher = bands.RowBandMatcher(engine, rank, world, exchange=exchange)
        if exchange == "p2p_dataflow":
            matcher._attached = True                    # the neighbours are still attached from the "p2p" leg
        # parity first: the fixture's own call, gathered labels against the reference's
        labels = matcher.gather(matcher.lbp(left, right, prior, n_iter=fixture_iters, **fkw))
        identical = bool(np.array_equal(labels, want))
        # timing: the iteration loop (with its exchanges) on the engine's stream, CUDA events, max over ranks
        times = []
        for trial in range(3):
            begin, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            matcher.lbp(left, right, prior, n_iter=n_iter, _events=(begin, end), **fkw)
            torch.cuda.synchronize()
            t = torch.tensor([begin.elapsed_time(end)], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            times.append(float(t[0]))
        ms_iter = min(times[1:]) / n_iter
        algo = 24.0 * rows * cols * levels / world
        results[exchange] = {"exchange": exchange, "ms_per_iter": ms_iter, "n_iter": n_iter,
                             "achieved_gbs_per_gpu": algo / (ms_iter * 1e-3) / 1e9,
                             "frac": algo / (ms_iter * 1e-3) / 1e9 / peak,
                             "identical_to_reference": identical, "fixture": "b09_1080p_d256_it2",
                             "rows_per_gpu": last - first + 1}
        if exchange == "p2p_dataflow":
            engine.matcher.synchronize()
            dist.barrier()
            engine.detach_peers()
            engine.matcher.set_option(capi.OPT_ALGORITHM, 5)
    del engine
    torch.cuda.empty_cache()
    return {"shape": "1920x1080xD256", "n_gpus": world,
            "single_gpu_roofline_ms_per_iter": 24.0 * rows * cols * levels / peak / 1e6, **results}


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--gpus", type=int, default=1)
    parser.add_argument("--steps", type=int, default=10)
    parser.add_argument("--warmup", type=int, default=3)
    parser.add_argument("--impl", default="b200", choices=["b200", "reference"])
    parser.add_argument("--frames-per-step", type=int, default=256, help="frame pairs per GPU per step")
    parser.add_argument("--capacity", type=int, default=128, help="frame pairs per launch (device batch)")
    parser.add_argument("--cpu-frames", type=int, default=10, help="frames of the CPU baseline sample (0 = skip)")
    parser.add_argument("--no-configs", action="store_true", help="skip the other named shapes (N = 1)")
    parser.add_argument("--no-bands", action="store_true", help="skip the row-band legs (N > 1)")
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
    del d_left, d_right, d_prior, d_out

    # ---- end-to-end arm: host buffers through the public API ------------------------------------
    matcher.set_stream(0)                          # back to the matcher's own stream (lanes overlap)

    def time_host(left, right, prior, out):
        def host_step():
            matcher.lbp_batch(left, right, prior, prior_trust_factor=TRUST, prior_influence_mode=MODE,
                              n_iter=N_ITER, out=out)
        for _ in range(max(args.warmup, 3)):
            host_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            host_step()
        torch.cuda.synchronize()
        seconds = time.perf_counter() - t0
        barrier()
        return seconds

    h_left = hsm.pinned_empty(lefts.shape, lefts.dtype); h_left[...] = lefts
    h_right = hsm.pinned_empty(rights.shape, rights.dtype); h_right[...] = rights
    h_prior = hsm.pinned_empty(priors.shape, priors.dtype); h_prior[...] = priors
    h_out = hsm.pinned_empty((frames, ROWS, COLS), np.int64)
    e2e_s = time_host(h_left, h_right, h_prior, h_out)
    same = bool(np.array_equal(h_out, labels_device))
    # ... and with ordinary numpy arrays, as the reference's callers pass them
    p_out = np.empty((frames, ROWS, COLS), dtype=np.int64)
    pageable_s = time_host(lefts, rights, priors, p_out)
    same_pageable = bool(np.array_equal(p_out, labels_device))
    del h_left, h_right, h_prior

    # ---- reduce over ranks (max time) -------------------------------------------------------------
    times = torch.tensor([elapsed_ms, e2e_s * 1e3, pageable_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    elapsed_ms, e2e_ms, pageable_ms = float(times[0]), float(times[1]), float(times[2])
    del matcher
    torch.cuda.empty_cache()

    bands_record = None
    if world > 1 and not args.no_bands:
        bands_record = band_legs(torch, dist, hsm, rank, world, local_rank)

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
        value = world * frames * args.steps / (elapsed_ms * 1e-3)
        whole_frame = frame_bytes(ROWS, COLS, LEVELS, N_ITER) * value / world / 1e9
        h2d = int(lefts.nbytes + rights.nbytes + priors.nbytes)
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(CONFIG),
            "run": {"frames_per_step_per_gpu": frames, "frames_per_launch": cap, "kernels_per_iteration": per_iter,
                    "l2": "inputs larger than L2: %.1f GB of message state per launch vs 126 MB"
                          % (6 * 4.0 * ROWS * COLS * LEVELS * cap / 1e9)},
            "e2e": {"value": world * frames * args.steps / (e2e_ms * 1e-3), "unit": "frames/s",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(h_out.nbytes),
                    "timing": "wall clock between device synchronisations", "host_memory": "page-locked",
                    "labels_equal_device_arm": same},
            "e2e_pageable": {"value": world * frames * args.steps / (pageable_ms * 1e-3), "unit": "frames/s",
                             "h2d_bytes_per_step": int(lefts.nbytes // 2 + rights.nbytes // 2 + priors.nbytes),
                             "d2h_bytes_per_step": int(p_out.nbytes),
                             "host_memory": "ordinary numpy arrays, staged by lbp_batch (float64 images cast to "
                                            "float32 while staging)",
                             "labels_equal_device_arm": same_pageable},
            "gpu_launches": int(stats["kernel_launches"]),
            "roofline": {"bound": "hbm", "kernel": "k_iter_tmast (fused 4-sweep BP iteration, TMA loads + TMA stores); "
                                                    "unit = one iteration over frames_per_launch frames",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": peak_source, "algorithmic_bytes_per_launch": algo_bytes,
                         "launch_ms": launch_ms, "launches_timed": int(launches),
                         "concurrent_launches_per_iteration": per_iter,
                         "whole_frame_gbs": whole_frame, "whole_frame_frac": whole_frame / peak,
                         "traffic": None if not traffic else
                         traffic["dram_bytes_per_launch"] * frames_per_launch / traffic["frames_per_launch"],
                         "traffic_source": None if not traffic else traffic.get("source")},
            "clocks": clocks.summary(),
        }
        if bands_record is not None:
            line["bands"] = bands_record
        if world == 1 and not args.no_configs:
            line["configs"] = single_gpu_configs(torch, hsm, local_rank)
            line["single_frame_lbp"] = single_frame_latency(hsm, local_rank)
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
