"""Scratch timing probe (not the benchmark): message-kernel time per iteration for a
batch of synthetic frames, across algorithm / data-term / band-height variants."""
import argparse
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hybrid_stereo_matching_b200 as hsm
from importlib import import_module
capi = import_module("hybrid-stereo-matching_b200._capi")
synth = import_module("hybrid-stereo-matching_b200.synth")

PEAK = 6538.9


def run(rows, cols, levels, batch, n_iter, algorithm, data_term, band_rows, stages=0, lane=0, prior=True):
    lefts, rights, priors = synth.synthetic_sequence(min(batch, 4), rows, cols, levels, base_seed=1234, with_prior=prior)
    reps = -(-batch // lefts.shape[0])
    lefts = np.tile(lefts, (reps, 1, 1))[:batch]
    rights = np.tile(rights, (reps, 1, 1))[:batch]
    priors = np.tile(priors, (reps, 1, 1))[:batch] if prior else None
    m = hsm.StereoMRF((rows, cols), levels, capacity=batch)
    m.set_option(capi.OPT_LANE_CONFIG, lane)
    m.set_option(capi.OPT_ALGORITHM, algorithm)
    m.set_option(capi.OPT_DATA_TERM, data_term)
    m.set_option(capi.OPT_BAND_ROWS, band_rows)
    m.set_option(capi.OPT_STAGES, stages)
    groups = int(os.environ.get("HSM_FRAME_GROUPS", "0"))
    m.set_option(capi.OPT_FRAME_GROUPS, groups)
    m._init_fields(lefts, rights, priors, 1.0, "adaptive", True, _batched=True)
    capi.check(capi.lib().hsm_iterate(m._handle, int(os.environ.get("HSM_WARM", "3")), 1))
    m.synchronize()
    m.reset_stats()
    capi.check(capi.lib().hsm_iterate(m._handle, n_iter, 1))
    m.synchronize()
    st = m.stats()
    ms_iter = st["iteration_ms"] / n_iter
    hwd = rows * cols * levels
    algo_bytes = 24.0 * hwd * batch
    out = dict(shape="%dx%dxD%d" % (cols, rows, levels), batch=batch, algorithm=algorithm, data_term=data_term,
               band_rows=band_rows, stages=stages, lane=lane, groups=groups, ms_per_iter=round(ms_iter, 4),
               eff_GBs=round(algo_bytes / ms_iter / 1e6, 1), frac=round(algo_bytes / ms_iter / 1e6 / PEAK, 3),
               frames_per_s_50it=round(batch / (ms_iter * 50 / 1e3), 1))
    print(json.dumps(out), flush=True)
    del m


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None, help="rows,cols,levels,batch,iters,algorithm,data_term,band_rows")
    args = ap.parse_args()
    if args.only:
        run(*[int(v) for v in args.only.split(",")])
        sys.exit(0)
    for lane in (0, 7):
        for dt in (0, 1):
            for st in (2, 3):
                for br in (60, 30, 15):
                    run(180, 240, 32, 64, 20, 2, dt, br, st, lane)
    run(180, 240, 32, 256, 10, 2, 1, 0)
    run(180, 240, 32, 1, 50, 2, 1, 0)
    run(180, 240, 32, 4, 50, 2, 1, 0)
    for lane in (0, 8):
        run(480, 640, 64, 4, 10, 2, 0, 0, 0, lane)
        run(480, 640, 64, 4, 10, 2, 1, 0, 0, lane)
    run(1080, 1920, 256, 1, 4, 2, 0, 0)
    run(1080, 1920, 256, 1, 4, 2, 1, 0)
