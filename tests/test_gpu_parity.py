"""Parity of the CUDA path (through the C ABI / StereoMRF) against the golden fixtures
of the real reference and against the CPU oracles.  Bar: planes bit-identical
(np.array_equal), labels identical.  Runs on the B200 box: pytest -m gpu."""
import hashlib

import numpy as np
import pytest

from oracle import cases as C
from oracle import mrf_c

pytestmark = pytest.mark.gpu

# HUGE_CASES: reference-pinned fixtures without stored planes (sha1 only); the two 1080p ones have their own,
# memory-aware test in test_gpu_full_sizes.py
ALL = [(c, True) for c in C.SMALL_CASES] + [(c, False) for c in C.BIG_CASES] + \
      [(c, False) for c in C.HUGE_CASES if c["rows"] < 1000]
IDS = [c["name"] for c, _ in ALL]
# (algorithm, data_term): TMA-fed fused kernel with recomputed data term (default) / with the data plane,
# LDG-fed fused kernel (both), one kernel per direction
VARIANTS = [(6, 1), (6, 0), (5, 1), (5, 0), (4, 1), (4, 0), (3, 1), (2, 1), (2, 0), (1, 1), (1, 0), (0, 0)]
VARIANT_IDS = ["flow_T24", "flow_T28", "tmast_T24", "tmast_T28", "ws_T24", "ws_T28", "tma2_T12", "tma_T24", "tma_T28", "ldg_T24", "ldg_T28",
               "directional_T80"]


def _factory(hsm, algorithm, data_term, band_rows=0, capacity=1):
    from importlib import import_module
    capi = import_module("hybrid-stereo-matching_b200._capi")

    def make(dim, levels):
        m = hsm.StereoMRF(tuple(dim), levels, capacity=capacity)
        m.set_option(capi.OPT_ALGORITHM, algorithm)
        m.set_option(capi.OPT_DATA_TERM, data_term)
        m.set_option(capi.OPT_BAND_ROWS, band_rows)
        return m
    return make


@pytest.mark.parametrize("variant", VARIANTS, ids=VARIANT_IDS)
@pytest.mark.parametrize("case,small", ALL, ids=IDS)
def test_cuda_matches_reference_fixture(hsm, case, small, variant):
    labels, planes = C.run_case(_factory(hsm, *variant), case)
    data, meta = C.load_fixture(case)
    assert labels.dtype == np.int64
    mism = int((labels != data["labels"]).sum())
    assert mism == 0, "%d of %d labels differ" % (mism, labels.size)
    assert hashlib.sha1(labels.astype(np.int32).tobytes()).hexdigest() == meta["labels_sha1"]
    for name, plane in planes.items():
        assert plane.dtype == np.float32 and plane.shape == (case["levels"],) + labels.shape
        if small:
            want = data["plane_" + name]
            assert np.array_equal(plane, want, equal_nan=True), \
                "%s: max abs diff %g" % (name, float(np.nanmax(np.abs(plane - want))))
        assert C.plane_sha1(plane) == meta["plane_sha1"][name], name


@pytest.mark.parametrize("band_rows", [1, 2, 3, 4, 5, 7])
def test_band_height_does_not_change_results(hsm, band_rows):
    for case in (C.SMALL_CASES[3], C.SMALL_CASES[12], C.SMALL_CASES[24]):
        labels, planes = C.run_case(_factory(hsm, {1: 3, 2: 5, 3: 6, 4: 6, 5: 4, 7: 2}[band_rows], 1, band_rows=band_rows), case)
        data, _ = C.load_fixture(case)
        assert np.array_equal(labels, data["labels"])
        for name, plane in planes.items():
            assert np.array_equal(plane, data["plane_" + name]), (case["name"], name)


def test_random_shapes_against_c_oracle(hsm):
    rng = np.random.default_rng(99)
    for trial in range(40):
        rows, cols = int(rng.integers(1, 40)), int(rng.integers(2, 90))
        levels = int(rng.integers(1, min(cols + 1, 70) + 1))
        n_iter = int(rng.integers(0, 6))
        left, right = rng.random((rows, cols)), rng.random((rows, cols))
        if trial % 5 == 0:
            left, right = left.astype(np.float32), right.astype(np.float32)
        prior, kwargs = None, dict(n_iter=n_iter)
        if trial % 3:
            prior = rng.integers(-1, levels, size=(rows, cols)).astype(
                [np.float64, np.int32, np.float32, np.int64][trial % 4])
            kwargs.update(prior_trust_factor=[0.0, 0.5, 1.0, 10.0, 100.0][trial % 5],
                          prior_influence_mode=["const", "adaptive"][trial % 2])
        want_obj = mrf_c.OracleMRFC((rows, cols), levels)
        want = want_obj.lbp(left, right, prior, **kwargs)
        for variant in VARIANTS:
            got_obj = _factory(hsm, *variant)((rows, cols), levels)
            got = got_obj.lbp(left, right, prior, **kwargs)
            assert np.array_equal(got, want), (trial, variant, rows, cols, levels)
            want_planes = want_obj._message_field
            for name, plane in got_obj._message_field.items():
                assert np.array_equal(plane, want_planes[name]), (trial, variant, name, rows, cols, levels)
            assert np.array_equal(got_obj.energy_field(), want_obj.energy()), (trial, variant)


def test_float64_scalar_trust_factor(hsm):
    rng = np.random.default_rng(5)
    left, right = rng.random((9, 14)), rng.random((9, 14))
    prior = rng.integers(-1, 6, size=(9, 14)).astype(np.float64)
    want_obj = mrf_c.OracleMRFC((9, 14), 6)
    want = want_obj.lbp(left, right, prior, prior_trust_factor=np.float64(0.3), n_iter=2)
    got_obj = hsm.StereoMRF((9, 14), 6)
    got = got_obj.lbp(left, right, prior, prior_trust_factor=np.float64(0.3), n_iter=2)
    assert np.array_equal(got, want)
    assert np.array_equal(got_obj._message_field["data"], want_obj._message_field["data"])


def test_batch_equals_single_frames(hsm):
    from importlib import import_module
    synth = import_module("hybrid-stereo-matching_b200.synth")
    lefts, rights, priors = synth.synthetic_sequence(7, 36, 50, 12, base_seed=3)
    single = hsm.StereoMRF((36, 50), 12)
    want = np.stack([single.lbp(lefts[i], rights[i], priors[i], prior_trust_factor=1.0,
                                prior_influence_mode="adaptive", n_iter=6) for i in range(7)])
    for capacity in (1, 3, 7, 16):
        batched = hsm.StereoMRF((36, 50), 12, capacity=capacity)
        got = batched.lbp_batch(lefts, rights, priors, prior_trust_factor=1.0,
                                prior_influence_mode="adaptive", n_iter=6)
        assert got.shape == (7, 36, 50) and got.dtype == np.int64
        assert np.array_equal(got, want), capacity
    # planes of an inner frame of the batch
    batched = hsm.StereoMRF((36, 50), 12, capacity=7)
    batched.lbp_batch(lefts, rights, priors, prior_trust_factor=1.0, prior_influence_mode="adaptive", n_iter=6)
    single.lbp(lefts[4], rights[4], priors[4], prior_trust_factor=1.0, prior_influence_mode="adaptive", n_iter=6)
    for name, plane in single._message_field.items():
        assert np.array_equal(batched.message_field(4)[name], plane), name


@pytest.mark.gpu
@pytest.mark.parametrize("algorithm", [6, 5, 4, 2, 1])
def test_frame_groups_do_not_change_results(hsm, algorithm):
    """Two concurrent frame groups (HSM_OPT_FRAME_GROUPS = 2, the default for real workloads) == one
    launch per iteration == single frames, for odd and even frame counts."""
    from importlib import import_module
    capi = import_module("hybrid-stereo-matching_b200._capi")
    synth = import_module("hybrid-stereo-matching_b200.synth")
    lefts, rights, priors = synth.synthetic_sequence(5, 30, 44, 20, base_seed=11)
    single = hsm.StereoMRF((30, 44), 20)
    single.set_option(capi.OPT_ALGORITHM, algorithm)
    want = np.stack([single.lbp(lefts[i], rights[i], priors[i], prior_trust_factor=0.7,
                                prior_influence_mode="const", n_iter=7) for i in range(5)])
    for groups in (1, 2):
        for count in (2, 5):
            batched = hsm.StereoMRF((30, 44), 20, capacity=count)
            batched.set_option(capi.OPT_ALGORITHM, algorithm)
            batched.set_option(capi.OPT_FRAME_GROUPS, groups)
            got = batched.lbp_batch(lefts[:count], rights[:count], priors[:count], prior_trust_factor=0.7,
                                    prior_influence_mode="const", n_iter=7)
            assert np.array_equal(got, want[:count]), (groups, count)
            single.lbp(lefts[count - 1], rights[count - 1], priors[count - 1], prior_trust_factor=0.7,
                       prior_influence_mode="const", n_iter=7)
            for name, plane in single._message_field.items():
                assert np.array_equal(batched.message_field(count - 1)[name], plane), (groups, count, name)


def test_iteration_by_iteration_equals_one_call(hsm):
    """_update_message_fields() n times == loop_belief(n_iter=n) (mrf.py:130-133)."""
    case = C.SMALL_CASES[0]
    left, right, _, _ = C.case_inputs(case, 0)
    a = hsm.StereoMRF(left.shape, case["levels"])
    a.loop_belief(left, right, n_iter=0)
    for _ in range(3):
        a._update_message_fields()
    data, _ = C.load_fixture(case)
    assert np.array_equal(a.get_map_belief(), data["labels"])
    for name, plane in a._message_field.items():
        assert np.array_equal(plane, data["plane_" + name]), name


def test_uninitialised_matcher_reads_zero(hsm):
    m = hsm.StereoMRF((5, 6), 4)
    out = m.get_map_belief()
    assert out.dtype == np.int64 and out.shape == (5, 6) and not out.any()


def test_full_size_properties_c1_batch(hsm):
    """BASELINE C1/C2 size, 50 iterations, batch of 4: size-independent properties --
    messages normalised to max 1 per pixel (or all-zero border lines), labels in range,
    left border labelled mostly D-1 (zero-cost columns), batch == single."""
    from importlib import import_module
    synth = import_module("hybrid-stereo-matching_b200.synth")
    D = 32
    lefts, rights, priors = synth.synthetic_sequence(4, 180, 240, D, base_seed=1234)
    lefts[0], rights[0], priors[0], _ = C.case_inputs(C.BIG_CASES[2], 0)
    m = hsm.StereoMRF((180, 240), D, capacity=4)
    labels = m.lbp_batch(lefts, rights, priors, prior_trust_factor=1.0, prior_influence_mode="adaptive", n_iter=50)
    assert labels.min() >= 0 and labels.max() < D
    data, meta = C.load_fixture(C.BIG_CASES[2])          # frame 0 == b03 (seed 1234)
    assert np.array_equal(labels[0], data["labels"])
    planes = m.message_field(0)
    for name in ("south", "west", "north", "east"):
        assert C.plane_sha1(planes[name]) == meta["plane_sha1"][name], name
        peak = planes[name].max(axis=0)
        assert np.all((peak == 1.0) | (peak == 0.0)), name
    assert not planes["south"][:, 0, :].any() and not planes["north"][:, -1, :].any()
    assert not planes["west"][:, :, -1].any() and not planes["east"][:, :, 0].any()


@pytest.mark.parametrize("lane_config,max_levels", [(10, 20), (11, 24), (12, 40), (13, 48), (14, 20), (15, 24), (16, 32)],
                         ids=["5x1", "6x1", "5x2", "6x2", "5x1_7warps", "6x1_9warps", "8x1_10warps"])
@pytest.mark.parametrize("algorithm", [5, 6], ids=["tmast", "flow"])
def test_non_power_of_two_lane_layouts(hsm, lane_config, max_levels, algorithm):
    """Lane layouts with 5 / 6 lanes per pixel (HSM_OPT_LANE_CONFIG 10..13; chosen automatically for D = 17..24 and
    33..48, e.g. the shipped max_disparity 20 and 22): every fixture whose disparity count they cover, including
    counts that leave some of their chunks partly or wholly unused, against the reference's planes and labels."""
    from importlib import import_module
    capi = import_module("hybrid-stereo-matching_b200._capi")
    todo = [(c, True) for c in C.SMALL_CASES if c["levels"] <= max_levels] + \
           [(c, False) for c in C.HUGE_CASES if c["levels"] <= max_levels and c["levels"] > max_levels - 8]

    def make(dim, levels):
        m = hsm.StereoMRF(tuple(dim), levels)
        m.set_option(capi.OPT_ALGORITHM, algorithm)
        m.set_option(capi.OPT_LANE_CONFIG, lane_config)
        return m
    assert len(todo) >= 10
    for case, small in todo:
        labels, planes = C.run_case(make, case)
        data, meta = C.load_fixture(case)
        assert np.array_equal(labels, data["labels"]), case["name"]
        for name, plane in planes.items():
            assert C.plane_sha1(plane) == meta["plane_sha1"][name], (case["name"], name)
            if small:
                assert np.array_equal(plane, data["plane_" + name], equal_nan=True), (case["name"], name)


def test_first_iteration_reads_no_planes_but_state_reads_back_as_zero(hsm):
    """hsm_init_fields(reinit) only records that the messages are zero; read-backs before any iteration, and
    n_iter = 0, must still see real zeros -- also right after a run that left non-zero planes behind."""
    case = C.SMALL_CASES[3]
    left, right, prior, kwargs = C.case_inputs(case, 0)
    m = hsm.StereoMRF(left.shape, case["levels"])
    m.lbp(left, right, prior, prior_trust_factor=1.0, prior_influence_mode="adaptive", n_iter=4)
    m.loop_belief(left, right, prior, n_iter=0)
    for name, plane in m._message_field.items():
        if name != "data":
            assert not plane.any(), name
    want = mrf_c.OracleMRFC(left.shape, case["levels"])
    want_labels = want.lbp(left, right, prior, n_iter=0)
    assert np.array_equal(m.lbp(left, right, prior, n_iter=0), want_labels)
    # and a one-iteration run (the zero-input launch is also the finalising one)
    got = m.lbp(left, right, prior, prior_trust_factor=1.0, prior_influence_mode="adaptive", n_iter=1)
    assert np.array_equal(got, want.lbp(left, right, prior, prior_trust_factor=1.0,
                                        prior_influence_mode="adaptive", n_iter=1))
    for name, plane in want._message_field.items():
        assert np.array_equal(m._message_field[name], plane), name


@pytest.mark.parametrize("name", ["b12_boxes_80x60_d20", "b13_checkerboard_80x60_d22", "b14_head_70x60_d27"])
def test_shipped_small_frames_in_a_large_batch(hsm, name):
    """The shipped down-scaled configurations as a batch of 240 frames: the 7- / 9- / 10-warp CTA shapes chosen
    for 80- and 70-column frames, the 30-row bands of short frames and the two frame groups, every frame's labels
    against the reference fixture (frames differ: each is the fixture's pair rolled by a few rows, and the
    reference result of a rolled pair is not the rolled result, so only frame 0 and its copies are compared)."""
    case = [c for c in C.HUGE_CASES if c["name"] == name][0]
    left, right, prior, kwargs = C.case_inputs(case, 0)
    data, meta = C.load_fixture(case)
    n = 240
    lefts = np.ascontiguousarray(np.broadcast_to(left, (n,) + left.shape)).copy()
    rights = np.ascontiguousarray(np.broadcast_to(right, (n,) + right.shape)).copy()
    priors = np.ascontiguousarray(np.broadcast_to(prior, (n,) + prior.shape)).copy()
    # every third frame gets different content, so that a frame mix-up between CTAs cannot go unnoticed
    lefts[1::3] = np.roll(lefts[1::3], 5, axis=1)
    rights[1::3] = np.roll(rights[1::3], 5, axis=1)
    matcher = hsm.StereoMRF(left.shape, case["levels"], capacity=n)
    got = matcher.lbp_batch(lefts, rights, priors, n_iter=kwargs["n_iter"],
                            prior_trust_factor=kwargs["prior_trust_factor"],
                            prior_influence_mode=kwargs["prior_influence_mode"], lanes=1)
    for frame in list(range(0, n, 3)) + list(range(2, n, 3)):
        assert np.array_equal(got[frame], data["labels"]), frame
    single = hsm.StereoMRF(left.shape, case["levels"])
    want = single.lbp(lefts[1], rights[1], priors[1], n_iter=kwargs["n_iter"],
                      prior_trust_factor=kwargs["prior_trust_factor"],
                      prior_influence_mode=kwargs["prior_influence_mode"])
    for frame in range(1, n, 3):
        assert np.array_equal(got[frame], want), frame


@pytest.mark.parametrize("hints", ["0", "1", "7"])
def test_l2_eviction_hints_do_not_change_results(hints):
    """HSM_L2_HINTS (read once per process, hence the subprocess) only changes cache policies of the plane loads
    and stores of the default kernel family: same planes, same labels, with every hint on, off, and loads only."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r"""
import hashlib, sys
import numpy as np
sys.path.insert(0, %r)
import hybrid_stereo_matching_b200 as hsm
from importlib import import_module
from oracle import cases as C
capi = import_module("hybrid-stereo-matching_b200._capi")
names = ("s16_levels32", "s14_levels22", "s18_levels64", "b02_tsukuba_crop_d32_it50", "b16_davis_d20_it10")
todo = [c for c in C.SMALL_CASES + C.BIG_CASES + C.HUGE_CASES if c["name"] in names]
assert len(todo) == len(names)
for case in todo:
    for algorithm, band_rows in ((5, 0), (5, 3), (6, 4)):
        def make(dim, levels):
            m = hsm.StereoMRF(tuple(dim), levels)
            m.set_option(capi.OPT_ALGORITHM, algorithm)
            m.set_option(capi.OPT_BAND_ROWS, band_rows)
            return m
        labels, planes = C.run_case(make, case)
        data, meta = C.load_fixture(case)
        assert np.array_equal(labels, data["labels"]), (case["name"], algorithm)
        for name, plane in planes.items():
            assert C.plane_sha1(plane) == meta["plane_sha1"][name], (case["name"], algorithm, name)
print("ok")
""" % root
    env = dict(os.environ, HSM_L2_HINTS=hints)
    done = subprocess.run([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                          universal_newlines=True, timeout=600)
    assert done.returncode == 0 and done.stdout.strip().endswith("ok"), done.stdout[-2000:]
