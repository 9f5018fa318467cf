import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import hybrid_stereo_matching_b200 as hsm
from importlib import import_module
capi = import_module("hybrid-stereo-matching_b200._capi")
synth = import_module("hybrid-stereo-matching_b200.synth")
lib = capi.lib()
lib.hsm_debug_counters.argtypes = [ctypes.c_void_p, ctypes.c_int]
lefts, rights, priors = synth.synthetic_sequence(4, 180, 240, 32, base_seed=1234)
m = hsm.StereoMRF((180, 240), 32, capacity=4)
m.set_option(capi.OPT_BAND_ROWS, 30)
m._init_fields(lefts, rights, priors, 1.0, "adaptive", True, _batched=True)
out = (ctypes.c_ulonglong * 8)()
lib.hsm_debug_counters(out, 1)
for it in range(50):
    capi.check(lib.hsm_iterate(m._handle, 1, 1))
    lib.hsm_debug_counters(out, 1)
    if it < 5 or it % 5 == 4:
        print("iter %2d warp-steps %d  first-level pre %.1f%% post %.1f%%  exact pre %.2f%% post %.2f%%" % (it, out[0], 100.0 * out[1] / out[0], 100.0 * out[2] / out[0], 100.0 * out[3] / out[0], 100.0 * out[4] / out[0]))
