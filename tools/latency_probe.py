"""Single-frame lbp latency through the public API (the reference's own call pattern)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hybrid_stereo_matching_b200 as hsm
from importlib import import_module
synth = import_module("hybrid-stereo-matching_b200.synth")
capi = import_module("hybrid-stereo-matching_b200._capi")
for (rows, cols, levels, n_iter) in [(480, 640, 64, 100), (180, 240, 32, 50), (180, 240, 32, 10), (60, 80, 22, 10), (288, 384, 15, 10)]:
    l, r, d = synth.textured_pair(rows, cols, levels, 5)
    p = synth.sparse_prior(d, levels, 5)
    m = hsm.StereoMRF((rows, cols), levels)
    for _ in range(3):
        m.lbp(l, r, p, prior_trust_factor=1.0, prior_influence_mode="adaptive", n_iter=n_iter)
    ts = []
    for _ in range(10):
        t0 = time.perf_counter()
        m.lbp(l, r, p, prior_trust_factor=1.0, prior_influence_mode="adaptive", n_iter=n_iter)
        ts.append(time.perf_counter() - t0)
    st = m.stats()
    print(json.dumps({"shape": [rows, cols, levels], "n_iter": n_iter, "lbp_ms_median": round(1e3 * float(np.median(ts)), 3),
                      "lbp_ms_min": round(1e3 * min(ts), 3), "frames_per_s": round(1.0 / float(np.median(ts)), 1),
                      "iter_kernel_ms_avg": round(st["iteration_ms"] / max(st["iterations"], 1), 4)}))
