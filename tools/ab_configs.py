"""Development A/B: the single-frame legs of bench.py's `configs` (same code path, same box) for whichever library
HSM_B200_LIB names."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import hybrid_stereo_matching_b200 as hsm
adaptive = dict(prior_trust_factor=1.0, prior_influence_mode="adaptive")
tag = os.environ.get("AB_TAG", "?")
for rep in range(2):
    for name, args in (("c3_single", ((480, 640, 64), 100, 1, 1, 5, 3, adaptive)),
                       ("c5_single_gpu", ((1080, 1920, 256), 50, 1, 1, 3, 3, adaptive)),
                       ("c3_batch4", ((480, 640, 64), 100, 4, 4, 5, 3, adaptive))):
        r = bench.device_leg(torch, hsm, *args, fixture=None, device=0)
        print(tag, rep, name, round(r["ms_per_iteration"], 4), round(r["kernel_frac"], 3), flush=True)
