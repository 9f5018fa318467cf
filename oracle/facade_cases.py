"""Inputs of the façade / prior-rasterisation fixtures (``tests/golden/f*.npz``, ``tests/golden/r*.npz``).

TEST INFRASTRUCTURE ONLY.  Shared by ``oracle/make_golden_facade.py`` (which runs the REAL reference classes
and functions in the build container) and by the tests that replay the same inputs through the product.
"""
import os

import numpy as np

from .cases import GOLDEN_DIR

# FramebasedStereoMatching.run (stereo_framebased.py:60-101): sequences at the shipped configurations' scale
FACADE_CASES = [
    dict(name="f01_run_priors_per_frame", n=5, rows=14, cols=20, levels=6, seed=50, priors="same"),
    dict(name="f02_run_more_priors_than_frames", n=5, rows=14, cols=20, levels=6, seed=60, priors="more"),
    dict(name="f03_run_without_priors", n=4, rows=14, cols=20, levels=6, seed=70, priors=None),
    dict(name="f04_run_checkerboard_scale", n=6, rows=60, cols=80, levels=22, seed=80, priors="same"),
    dict(name="f05_run_one_frame_online", n=3, rows=60, cols=70, levels=27, seed=90, priors="online"),
]

# generate_frames_from_spikes with pivots (hybrid_stereo.py:169-176) and without
RASTER_CASES = [
    dict(name="r01_pivots_small", seed=1, n=400, res=(20, 14), t_max=500, interval=100,
         pivots=[50.0, 100.0, 150.0, 400.0]),
    dict(name="r02_pivots_overlapping_windows", seed=2, n=3000, res=(80, 60), t_max=1000, interval=250,
         pivots=list(np.arange(50.0, 1000.0, 50.0))),
    dict(name="r03_time_intervals", seed=3, n=50, res=(9, 7), t_max=300, interval=100, pivots=None),
    dict(name="r04_dense_one_frame", seed=5, n=2000, res=(16, 12), t_max=100, interval=100, pivots=[100.0]),
    dict(name="r05_checkerboard_scale", seed=6, n=20000, res=(80, 60), t_max=2000, interval=50,
         pivots=list(np.arange(50.0, 2001.0, 50.0)), levels=22, unique_times=True),
    dict(name="r06_start_time_and_unsorted_pivots", seed=7, n=1500, res=(40, 30), t_max=800, interval=120,
         pivots=[600.0, 200.0, 410.5, 799.0], start_time=150, unique_times=True),
]


# OpticalFlowPixelCorrection.adjust (optical_flow_correction.py:34-88): reference events (t, x, y, polarity) per
# buffer window + prior frames to adjust
ADJUST_CASES = [
    dict(name="a01_adjust_predictive", seed=21, res=(20, 14), n_events=260, levels=9, pivots=[60.0, 120.0, 180.0],
         interval=50, time_arrow=+1, density=0.3),
    dict(name="a02_adjust_retrospective", seed=22, res=(20, 14), n_events=260, levels=9, pivots=[60.0, 120.0],
         interval=50, time_arrow=-1, density=0.3),
    dict(name="a03_adjust_dense_collisions", seed=23, res=(33, 27), n_events=200, levels=20, pivots=[80.0, 160.0],
         interval=60, time_arrow=+1, density=0.9),
    dict(name="a04_adjust_checkerboard_scale", seed=24, res=(80, 60), n_events=150, levels=22, pivots=[50.0],
         interval=50, time_arrow=+1, density=0.1),
]


# VelocityVectorField.fit_velocity_field (vvf.py:84-110): moving edges (so that planes fit and the rejection loop
# runs) plus background noise events; unique pixel/time jitter keeps the neighbourhoods well-posed
FLOW_CASES = [
    dict(name="v01_edge_moving_right", seed=31, res=(40, 30), speed=(0.05, 0.0), n_noise=150, t_max=400.0, dt=20),
    dict(name="v02_edge_moving_diagonally", seed=32, res=(40, 30), speed=(0.04, 0.03), n_noise=300, t_max=400.0, dt=50),
    dict(name="v03_sparse_noise_only", seed=33, res=(30, 20), speed=None, n_noise=400, t_max=300.0, dt=20),
]


def flow_events(case):
    """(N, 4) events (t, x, y, polarity), unsorted: an edge sweeping across the frame + uniform noise."""
    rng = np.random.default_rng(case["seed"])
    cols, rows = case["res"]
    events = []
    if case["speed"] is not None:
        vx, vy = case["speed"]                        # pixels per ms
        for t in np.arange(0.0, case["t_max"], 1.0):
            for y in range(rows):
                x = vx * t + (vy * t) * 0.0 + 0.15 * y * (1.0 if vy else 0.0)
                xi = int(np.floor(x + vy * t)) % cols
                if rng.random() < 0.6:
                    events.append((t + rng.random() * 0.9, float(xi), float(y), float(rng.integers(0, 2))))
    n = case["n_noise"]
    noise = np.stack([rng.random(n) * case["t_max"], rng.integers(0, cols, n).astype(float),
                      rng.integers(0, rows, n).astype(float), rng.integers(0, 2, n).astype(float)], axis=1)
    out = np.concatenate([np.asarray(events).reshape(-1, 4), noise], axis=0)
    return out[rng.permutation(out.shape[0])]


def adjust_inputs(case):
    """(reference_events (N, 4): t, x, y, polarity; prior frames (F, rows, cols) float64 with -1 = none)."""
    rng = np.random.default_rng(case["seed"])
    cols, rows = case["res"]
    n = case["n_events"]
    events = np.stack([np.sort(rng.random(n) * (max(case["pivots"]) + 5.0)), rng.integers(0, cols, n).astype(float),
                       rng.integers(0, rows, n).astype(float), rng.integers(0, 2, n).astype(float)], axis=1)
    events = events[rng.permutation(n)]
    frames = np.full((len(case["pivots"]), rows, cols), -1.0)
    mask = rng.random(frames.shape) < case["density"]
    frames[mask] = rng.integers(0, case["levels"], int(mask.sum()))
    return events, frames


def facade_inputs(case):
    """(lefts, rights, timestamps, prior_info or None) of a façade case."""
    import importlib
    synth = importlib.import_module("hybrid-stereo-matching_b200.synth")
    n, rows, cols, levels = case["n"], case["rows"], case["cols"], case["levels"]
    lefts, rights, priors = synth.synthetic_sequence(n, rows, cols, levels, base_seed=case["seed"])
    ts = np.arange(n) * 50.0 + 50.0
    kind = case["priors"]
    if kind is None:
        return lefts, rights, ts, None
    if kind == "same":
        return lefts, rights, ts, {"ts": ts, "priors": priors}
    if kind == "online":                       # int32 prior with -1 fill (online_processing.py:63,121)
        return lefts, rights, ts, {"ts": ts, "priors": priors.astype(np.int32)}
    many_ts = np.arange(12) * 25.0             # more priors than frames: the first one at or after each frame
    many = np.stack([np.roll(priors[i % n], i, axis=1) for i in range(12)])
    return lefts, rights, ts, {"ts": many_ts, "priors": many}


def raster_events(case):
    """(xs, ys, ts, zs) of a rasterisation case; the SNN output is not time-ordered."""
    rng = np.random.default_rng(case["seed"])
    n, (cols, rows) = case["n"], case["res"]
    xs, ys = rng.integers(0, cols, n), rng.integers(0, rows, n)
    if case.get("unique_times"):               # no two events share a timestamp: the result does not depend on
        ts = np.sort(rng.random(n) * case["t_max"])                      # the sort's handling of ties
    else:
        ts = np.sort(rng.integers(0, case["t_max"], n)).astype(np.float64)
    perm = rng.permutation(n)
    zs = rng.integers(0, case.get("levels", 20), n).astype(np.float64)
    return xs[perm], ys[perm], ts[perm], zs[perm]


def fixture_path(case):
    return os.path.join(GOLDEN_DIR, case["name"] + ".npz")
