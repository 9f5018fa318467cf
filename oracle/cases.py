"""Parity case table shared by ``oracle/make_golden.py`` and the tests.

TEST INFRASTRUCTURE ONLY.

Each case names its inputs (seeded generator or the Tsukuba demo pair), the
call sequence applied to a fresh matcher, and whether the fixture stores the
full five planes (small cases) or only labels + sha1/sums of the planes (big
cases).  Shapes and parameters are lifted from the reference's notebooks and
call sites (SURVEY.md sections 3.3, 8c, 8d):

  * ``MRF_StereoMatching.ipynb`` cell 2: Tsukuba grey, n_levels=15, n_iter=10;
  * ``stereo_framebased.py:84-85``: prior_trust_factor=1.0, 'adaptive', n_iter=10;
  * ``mrf.py:146`` / ``online_processing.py:128``: lbp defaults 0.5 / 'const',
    int32 prior with -1 fill;
  * ``evaluation.ipynb`` cell 20: trust factors 10 and 100 (negative costs);
  * shipped YAML configs: max_disparity 20 / 22 / 27 (not multiples of 4).
"""
import hashlib
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                          "tests", "golden")

# A call = kwargs for loop_belief(); 'prior' is one of None, 'f64', 'f32', 'i32', 'nan'.
SMALL_CASES = [
    dict(name="s01_noprior_it3", rows=12, cols=16, levels=5, seed=11,
         calls=[dict(n_iter=3)]),
    dict(name="s02_noprior_it0", rows=9, cols=20, levels=7, seed=12,
         calls=[dict(n_iter=0)]),
    dict(name="s03_const05_f64", rows=16, cols=24, levels=8, seed=13,
         calls=[dict(n_iter=4, prior="f64", prior_trust_factor=0.5, prior_influence_mode="const")]),
    dict(name="s04_adaptive_f64", rows=16, cols=24, levels=8, seed=14,
         calls=[dict(n_iter=4, prior="f64", prior_trust_factor=1.0, prior_influence_mode="adaptive")]),
    dict(name="s05_const10_negative", rows=14, cols=22, levels=6, seed=15,
         calls=[dict(n_iter=5, prior="f64", prior_trust_factor=10.0, prior_influence_mode="const")]),
    dict(name="s06_const0", rows=14, cols=22, levels=6, seed=16,
         calls=[dict(n_iter=2, prior="f64", prior_trust_factor=0.0, prior_influence_mode="const")]),
    dict(name="s07_const1_zero_cost", rows=14, cols=22, levels=6, seed=17,
         calls=[dict(n_iter=3, prior="f64", prior_trust_factor=1.0, prior_influence_mode="const")]),
    dict(name="s08_int32_prior_online", rows=15, cols=20, levels=9, seed=18,
         calls=[dict(n_iter=3, prior="i32", prior_trust_factor=0.5, prior_influence_mode="const")]),
    dict(name="s09_f32_prior", rows=15, cols=20, levels=9, seed=19,
         calls=[dict(n_iter=3, prior="f32", prior_trust_factor=0.5, prior_influence_mode="const")]),
    dict(name="s10_nan_fill_prior", rows=15, cols=20, levels=9, seed=20,
         calls=[dict(n_iter=3, prior="nan", prior_trust_factor=1.0, prior_influence_mode="adaptive")]),
    dict(name="s11_warm_start", rows=13, cols=18, levels=6, seed=21,
         calls=[dict(n_iter=2),
                dict(n_iter=3, prior="f64", prior_trust_factor=0.5, prior_influence_mode="const",
                     reinit_messages=False, reseed=100)]),
    dict(name="s12_warm_start_it0", rows=13, cols=18, levels=6, seed=22,
         calls=[dict(n_iter=3), dict(n_iter=0, reinit_messages=False, reseed=5)]),
    dict(name="s13_levels15", rows=20, cols=31, levels=15, seed=23, calls=[dict(n_iter=5)]),
    dict(name="s14_levels22", rows=11, cols=37, levels=22, seed=24,
         calls=[dict(n_iter=3, prior="f64", prior_trust_factor=1.0, prior_influence_mode="adaptive")]),
    dict(name="s15_levels27", rows=10, cols=33, levels=27, seed=25, calls=[dict(n_iter=3)]),
    dict(name="s16_levels32", rows=10, cols=40, levels=32, seed=26, calls=[dict(n_iter=6)]),
    dict(name="s17_levels56", rows=6, cols=64, levels=56, seed=27, calls=[dict(n_iter=2)]),
    dict(name="s18_levels64", rows=7, cols=70, levels=64, seed=28,
         calls=[dict(n_iter=3, prior="f64", prior_trust_factor=0.5, prior_influence_mode="const")]),
    dict(name="s19_width_eq_levels", rows=8, cols=12, levels=12, seed=29, calls=[dict(n_iter=3)]),
    dict(name="s20_width_levels_minus1", rows=8, cols=11, levels=12, seed=30, calls=[dict(n_iter=2)]),
    dict(name="s21_one_row", rows=1, cols=17, levels=4, seed=31, calls=[dict(n_iter=3)]),
    dict(name="s22_one_col_pair", rows=17, cols=2, levels=2, seed=32, calls=[dict(n_iter=3)]),
    dict(name="s23_levels1", rows=6, cols=7, levels=1, seed=33, calls=[dict(n_iter=2)]),
    dict(name="s24_flat_images_all_ties", rows=9, cols=13, levels=5, seed=34, flat=True,
         calls=[dict(n_iter=2)]),
    dict(name="s25_two_strips_wide", rows=5, cols=70, levels=8, seed=35, calls=[dict(n_iter=4)]),
    dict(name="s26_levels130", rows=4, cols=140, levels=130, seed=36, calls=[dict(n_iter=2)]),
]

BIG_CASES = [
    # KAT-A: MRF_StereoMatching.ipynb cell 2
    dict(name="b01_tsukuba_full_d15_it10", source="tsukuba", levels=15,
         calls=[dict(n_iter=10)]),
    # KAT-B == BASELINE.json configs[0] (C1): 240x180 crop, D=32, 50 iterations, no prior
    dict(name="b02_tsukuba_crop_d32_it50", source="tsukuba_crop", levels=32,
         calls=[dict(n_iter=50)]),
    # BASELINE.json configs[1] (C2): synthetic DAVIS with prior, offline call pattern
    dict(name="b03_davis_synth_adaptive_it50", rows=180, cols=240, levels=32, seed=1234,
         calls=[dict(n_iter=50, prior="f64", prior_trust_factor=1.0, prior_influence_mode="adaptive")]),
    # ... and the online / lbp-default call pattern
    dict(name="b04_davis_synth_const05_it50", rows=180, cols=240, levels=32, seed=1234,
         calls=[dict(n_iter=50, prior="f64", prior_trust_factor=0.5, prior_influence_mode="const")]),
    # C3 shape at a CPU-affordable iteration count
    dict(name="b05_vga_d64_it3", rows=480, cols=640, levels=64, seed=4321,
         calls=[dict(n_iter=3, prior="f64", prior_trust_factor=1.0, prior_influence_mode="adaptive")]),
    # eval_prior_influence_on_mrf.py:45-46: n_levels=56, n_iter=15 (on a synthetic of DAVIS size)
    dict(name="b06_davis_d56_it15", rows=180, cols=240, levels=56, seed=77,
         calls=[dict(n_iter=15)]),
]

# Reference-pinned fixtures for the configurations the CPU oracles cannot re-run inside the test-suite's time
# budget: generated ONCE by the real reference in the build container (labels + per-plane sha1 / float64 sums),
# read by the GPU parity tests.  ``cpu`` = which oracles the CPU suite re-checks them with.
HUGE_CASES = [
    # BASELINE.json configs[2] (C3) at its real iteration count
    dict(name="b07_vga_d64_it100", rows=480, cols=640, levels=64, seed=4321, cpu=("c",),
         calls=[dict(n_iter=100, prior="f64", prior_trust_factor=1.0, prior_influence_mode="adaptive")]),
    # BASELINE.json configs[4] (C5): 1920x1080, D=256, event prior
    dict(name="b08_1080p_d256_it1", rows=1080, cols=1920, levels=256, seed=99, cpu=(),
         calls=[dict(n_iter=1, prior="f64", prior_trust_factor=0.5, prior_influence_mode="const")]),
    dict(name="b09_1080p_d256_it2", rows=1080, cols=1920, levels=256, seed=99, cpu=(),
         calls=[dict(n_iter=2, prior="f64", prior_trust_factor=1.0, prior_influence_mode="adaptive")]),
    # the 32-lane x 2 layout (D 129..256) at a useful size, full and not full
    dict(name="b10_mid_d256_it5", rows=24, cols=300, levels=256, seed=41, cpu=("c", "numpy"),
         calls=[dict(n_iter=5, prior="f64", prior_trust_factor=1.0, prior_influence_mode="adaptive")]),
    dict(name="b11_mid_d200_it4", rows=40, cols=260, levels=200, seed=42, cpu=("c", "numpy"),
         calls=[dict(n_iter=4, prior="i32", prior_trust_factor=0.5, prior_influence_mode="const")]),
    # the shipped hybrid configurations at their effective (down-scaled) resolutions, offline call pattern
    # (experiments/configs/hybrid/*.yaml:19-23, stereo_framebased.py:84-85)
    dict(name="b12_boxes_80x60_d20", rows=60, cols=80, levels=20, seed=51, cpu=("c", "numpy"),
         calls=[dict(n_iter=10, prior="f64", prior_trust_factor=1.0, prior_influence_mode="adaptive")]),
    dict(name="b13_checkerboard_80x60_d22", rows=60, cols=80, levels=22, seed=52, cpu=("c", "numpy"),
         calls=[dict(n_iter=10, prior="f64", prior_trust_factor=1.0, prior_influence_mode="adaptive")]),
    dict(name="b14_head_70x60_d27", rows=60, cols=70, levels=27, seed=53, cpu=("c", "numpy"),
         calls=[dict(n_iter=10, prior="f64", prior_trust_factor=1.0, prior_influence_mode="adaptive")]),
    dict(name="b15_head_115x120_d50", rows=120, cols=115, levels=50, seed=54, cpu=("c", "numpy"),
         calls=[dict(n_iter=10, prior="i32", prior_trust_factor=0.5, prior_influence_mode="const")]),
    # the same disparity counts at DAVIS size (batched bench shapes of DESIGN.md)
    dict(name="b16_davis_d20_it10", rows=180, cols=240, levels=20, seed=55, cpu=("c",),
         calls=[dict(n_iter=10, prior="f64", prior_trust_factor=1.0, prior_influence_mode="adaptive")]),
    dict(name="b17_davis_d50_it10", rows=180, cols=240, levels=50, seed=56, cpu=("c",),
         calls=[dict(n_iter=10, prior="f64", prior_trust_factor=1.0, prior_influence_mode="adaptive")]),
    dict(name="b18_davis_d40_it6", rows=180, cols=240, levels=40, seed=57, cpu=("c",),
         calls=[dict(n_iter=6)]),
    dict(name="b19_davis_d100_it6", rows=90, cols=240, levels=100, seed=58, cpu=("c",),
         calls=[dict(n_iter=6, prior="f64", prior_trust_factor=0.5, prior_influence_mode="const")]),
]


def find_case(name):
    for case in SMALL_CASES + BIG_CASES + HUGE_CASES:
        if case["name"] == name:
            return case
    raise KeyError(name)


def tsukuba_rgb():
    data = np.load(os.path.join(GOLDEN_DIR, "tsukuba_rgb.npz"))
    return data["left"], data["right"]


def tsukuba_grey_pair(crop=False):
    """Grey = 0.2125 r + 0.7154 g + 0.0721 b on uint8/255 in float64 (skimage
    ``as_grey`` weights, as in MRF_StereoMatching.ipynb cell 1).  ``crop`` gives
    the 180x240 centre crop, min-max normalised as frames_io.py:103-109."""
    pair = []
    for rgb in tsukuba_rgb():
        f = rgb.astype(np.float64) / 255.0
        grey = 0.2125 * f[..., 0] + 0.7154 * f[..., 1] + 0.0721 * f[..., 2]
        if crop:
            grey = grey[54:234, 72:312]
            span = grey.max() - grey.min()
            grey = 1.0 / span * (grey - grey.min()) if span != 0 else np.ones_like(grey)
        pair.append(np.ascontiguousarray(grey))
    return pair[0], pair[1]


def _prior_as(kind, disparity, levels, seed):
    import importlib
    synth = importlib.import_module("hybrid-stereo-matching_b200.synth")
    if kind is None:
        return None
    base = synth.sparse_prior(disparity, levels, seed, density=0.2, outliers=0.3)
    if kind == "f64":
        return base
    if kind == "f32":
        return base.astype(np.float32)
    if kind == "i32":
        return base.astype(np.int32)
    if kind == "nan":
        out = base.copy()
        out[out < 0] = np.nan
        return out
    raise ValueError(kind)


def case_inputs(case, call_index=0):
    """Inputs of ``case`` for its ``call_index``-th call: (left, right, prior, kwargs)."""
    import importlib
    synth = importlib.import_module("hybrid-stereo-matching_b200.synth")
    call = dict(case["calls"][call_index])
    levels = case["levels"]
    source = case.get("source")
    if source == "tsukuba":
        left, right = tsukuba_grey_pair(False)
        disparity = np.zeros(left.shape, dtype=np.int64)
    elif source == "tsukuba_crop":
        left, right = tsukuba_grey_pair(True)
        disparity = np.zeros(left.shape, dtype=np.int64)
    else:
        seed = case["seed"] + call.pop("reseed", 0)
        left, right, disparity = synth.textured_pair(case["rows"], case["cols"], levels, seed,
                                                     block=(5, 6) if case["rows"] < 100 else (30, 40))
        if case.get("flat"):
            left = np.full_like(left, 0.25)
            right = np.full_like(right, 0.25)
    call.pop("reseed", None)
    prior = _prior_as(call.pop("prior", None), disparity, levels, case.get("seed", 0) + call_index)
    return left, right, prior, call


def plane_sha1(plane_dhw):
    return hashlib.sha1(np.ascontiguousarray(plane_dhw, dtype=np.float32).tobytes()).hexdigest()


def run_case(matcher_factory, case):
    """Drive a matcher with the reference surface through ``case``; return
    (labels int64 (H, W), dict of five (D, H, W) float32 planes)."""
    left, right, _, _ = case_inputs(case, 0)
    matcher = matcher_factory(left.shape, case["levels"])
    for index in range(len(case["calls"])):
        left, right, prior, kwargs = case_inputs(case, index)
        matcher.loop_belief(left, right, prior=prior, **kwargs)
    result = np.asarray(matcher.get_map_belief())
    planes = {k: np.asarray(v) for k, v in matcher._message_field.items()}
    return result, planes


def fixture_path(case):
    return os.path.join(GOLDEN_DIR, case["name"] + ".npz")


def load_fixture(case):
    data = np.load(fixture_path(case), allow_pickle=False)
    meta = json.loads(str(data["meta"]))
    return data, meta
