"""Single-frame lbp latency against band height / kernel choice (development probe)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hybrid_stereo_matching_b200 as hsm
from importlib import import_module
synth = import_module("hybrid-stereo-matching_b200.synth")
capi = import_module("hybrid-stereo-matching_b200._capi")
for (rows, cols, levels, n_iter) in [(60, 80, 22, 10), (60, 70, 27, 10), (180, 240, 32, 50), (180, 240, 32, 10)]:
    l, r, d = synth.textured_pair(rows, cols, levels, 5)
    p = synth.sparse_prior(d, levels, 5)
    for algorithm in (5, 6):
        for band_rows in (0, 1, 2, 3, 4, 6, 8):
            m = hsm.StereoMRF((rows, cols), levels)
            m.set_option(capi.OPT_ALGORITHM, algorithm)
            m.set_option(capi.OPT_BAND_ROWS, band_rows)
            kw = dict(prior_trust_factor=1.0, prior_influence_mode="adaptive", n_iter=n_iter)
            for _ in range(5):
                m.lbp(l, r, p, **kw)
            m.reset_stats()
            ts = []
            for _ in range(30):
                t0 = time.perf_counter()
                m.lbp(l, r, p, **kw)
                ts.append(time.perf_counter() - t0)
            st = m.stats()
            print(json.dumps({"shape": [rows, cols, levels], "n_iter": n_iter, "algorithm": algorithm,
                              "band_rows": band_rows, "lbp_ms_median": round(1e3 * float(np.median(ts)), 4),
                              "iter_us_avg": round(1e3 * st["iteration_ms"] / max(st["iterations"], 1), 2)}), flush=True)
            del m
