"""The façade above the hot path (stereo_framebased.py:15-101): prior selection by timestamp, call
pattern, result stacking.  Pinned to outputs of the REAL ``FramebasedStereoMatching`` (tests/golden/f*.npz,
written by oracle/make_golden_facade.py, which executes the reference class with the reference matcher):
on CPU the product façade runs with the numpy oracle as its matcher, on the GPU with the CUDA matcher; in the
build container the real class is also run live."""
import sys

import numpy as np
import pytest

from oracle import mrf_numpy


def _sequence(n=5, rows=14, cols=20, levels=6):
    from importlib import import_module
    synth = import_module("hybrid-stereo-matching_b200.synth")
    lefts, rights, priors = synth.synthetic_sequence(n, rows, cols, levels, base_seed=50)
    return lefts, rights, priors, np.arange(n) * 50.0 + 50.0


def _reference_run(lefts, rights, ts, prior_info, levels):
    """What the reference's run() computes, frame by frame (stereo_framebased.py:60-101)."""
    out = []
    matcher = mrf_numpy.OracleMRF(lefts.shape[1:], levels)
    if prior_info is not None:
        if len(prior_info["ts"]) > len(ts):
            idx = [np.searchsorted(prior_info["ts"], t, side="left") for t in ts]
            priors = prior_info["priors"][idx]
        else:
            priors = prior_info["priors"]
        for l, r, p in zip(lefts, rights, priors):
            out.append(matcher.lbp(l, r, p, prior_trust_factor=1.0, prior_influence_mode="adaptive", n_iter=10))
    else:
        for l, r in zip(lefts, rights):
            out.append(matcher.lbp(l, r))
    return np.asarray(out)


def test_facade_host_logic_with_oracle_matcher(hsm):
    lefts, rights, priors, ts = _sequence()
    levels = 6
    factory = lambda dim, lv: mrf_numpy.OracleMRF(dim, lv)
    # more priors than frames: the closest later prior is picked
    many_ts = np.arange(12) * 25.0
    many = np.stack([np.roll(priors[i % len(priors)], i, axis=1) for i in range(12)])
    info = {"ts": many_ts, "priors": many}
    facade = hsm.FramebasedStereoMatching((20, 14), levels, inputs={"left": lefts, "right": rights, "ts": ts},
                                          matcher_factory=factory)
    assert facade.algorithm.dimension == (levels, 14, 20)        # (x, y) -> (rows, cols)
    facade.run(info)
    got = facade.get_output()
    assert got.shape == (5, 14, 20)
    assert np.array_equal(got, _reference_run(lefts, rights, ts, info, levels))
    assert np.array_equal(facade.get_timestamps(), ts)
    # same number of priors as frames, and no priors at all
    for info in ({"ts": ts, "priors": priors}, None):
        facade = hsm.FramebasedStereoMatching((20, 14), levels, inputs={"left": lefts, "right": rights, "ts": ts},
                                              matcher_factory=factory)
        facade.run(info)
        assert np.array_equal(facade.get_output(), _reference_run(lefts, rights, ts, info, levels))
    with pytest.raises(NotImplementedError):
        hsm.FramebasedStereoMatching((20, 14), levels, algorithm="sgm")


def test_shim_installs_reference_import_paths(hsm):
    from importlib import import_module
    shim = import_module("hybrid-stereo-matching_b200.shim")
    saved = {k: v for k, v in sys.modules.items() if k.startswith("stereovis")}
    try:
        shim.install()
        from stereovis.framed.algorithms import StereoMRF as A
        from stereovis.framed.algorithms.mrf import StereoMRF as B
        from stereovis.framed.stereo_framebased import FramebasedStereoMatching as F
        assert A is hsm.StereoMRF and B is hsm.StereoMRF and F is hsm.FramebasedStereoMatching
    finally:
        for k in [k for k in sys.modules if k.startswith("stereovis")]:
            del sys.modules[k]
        sys.modules.update(saved)


@pytest.mark.gpu
def test_facade_batched_cuda_equals_sequential_reference(hsm):
    lefts, rights, priors, ts = _sequence(n=7)
    levels = 6
    info = {"ts": ts, "priors": priors}
    for capacity in (3, 16):
        facade = hsm.FramebasedStereoMatching((20, 14), levels, inputs={"left": lefts, "right": rights, "ts": ts},
                                              capacity=capacity)
        facade.run(info)
        assert np.array_equal(facade.get_output(), _reference_run(lefts, rights, ts, info, levels))
    facade = hsm.FramebasedStereoMatching((20, 14), levels, inputs={"left": lefts, "right": rights, "ts": ts})
    facade.run(None)
    assert np.array_equal(facade.get_output(), _reference_run(lefts, rights, ts, None, levels))
    # online call pattern: int32 shared-buffer prior with -1 fill, n_iter kwarg only (online_processing.py:121-128)
    online = hsm.FramebasedStereoMatching((20, 14), levels)
    prior32 = priors[0].astype(np.int32)
    got = online.run_one_frame(lefts[0], rights[0], prior32, n_iter=4)
    want = mrf_numpy.OracleMRF((14, 20), levels).lbp(lefts[0], rights[0], prior32, n_iter=4)
    assert np.array_equal(got, want) and got.dtype == np.int64


def test_online_iteration_budget_rule():
    """next_n_iter == the expression in online_processing.py:122-127, over a sweep of states."""
    from importlib import import_module
    framed = import_module("hybrid-stereo-matching_b200.framed")
    for min_iter, max_iter in ((1, 10), (2, 50), (5, 5)):
        for n_iter in range(0, 60, 3):
            for duration in (0.0, 10.0, 39.9, 40.0, 41.0, 200.0):
                for budget in (50.0, 50.0 * 2.5):
                    t = duration / budget
                    want = int(n_iter + 0.2 * n_iter + 1 if t < 0.8 else n_iter - 0.5 * n_iter)
                    want = min(max_iter, max(min_iter, want))
                    assert framed.next_n_iter(n_iter, duration, budget, min_iter, max_iter) == want
    assert framed.next_n_iter(3, 1.0, 50.0, 1, 10, adaptive=False) == 10


# ---- against the real reference class --------------------------------------------------------------
from oracle import facade_cases as F                                   # noqa: E402
from oracle import facade_ref                                          # noqa: E402


def _run_product_facade(hsm, case, **facade_kwargs):
    lefts, rights, ts, info = F.facade_inputs(case)
    facade = hsm.FramebasedStereoMatching((case["cols"], case["rows"]), case["levels"],
                                          inputs={"left": lefts, "right": rights, "ts": ts}, **facade_kwargs)
    if case["priors"] == "online":
        for i in range(case["n"]):
            facade.run_one_frame(lefts[i], rights[i], info["priors"][i], n_iter=3 + i)
    else:
        facade.run(info)
    return facade


@pytest.mark.parametrize("case", F.FACADE_CASES, ids=[c["name"] for c in F.FACADE_CASES])
def test_facade_host_logic_matches_reference_fixture(hsm, case):
    data = np.load(F.fixture_path(case), allow_pickle=False)
    facade = _run_product_facade(hsm, case, matcher_factory=lambda dim, lv: mrf_numpy.OracleMRF(dim, lv))
    out = facade.get_output()
    assert str(out.dtype) == str(data["dtype"]) and out.shape == data["depth"].shape
    assert np.array_equal(out, data["depth"])
    assert np.array_equal(facade.get_timestamps(), data["timestamps"])


@pytest.mark.skipif(not facade_ref.available(), reason="reference checkout not present")
def test_fixtures_are_what_the_real_class_returns():
    reference = facade_ref.load()
    case = F.FACADE_CASES[1]
    lefts, rights, ts, info = F.facade_inputs(case)
    facade = reference((case["cols"], case["rows"]), case["levels"], inputs={"left": lefts, "right": rights, "ts": ts})
    facade.run(info)
    assert np.array_equal(facade.get_output(), np.load(F.fixture_path(case))["depth"])


@pytest.mark.gpu
@pytest.mark.parametrize("capacity", [2, 64])
@pytest.mark.parametrize("case", F.FACADE_CASES, ids=[c["name"] for c in F.FACADE_CASES])
def test_facade_cuda_matches_reference_fixture(hsm, case, capacity):
    data = np.load(F.fixture_path(case), allow_pickle=False)
    facade = _run_product_facade(hsm, case, capacity=capacity)
    out = facade.get_output()
    assert str(out.dtype) == str(data["dtype"]) and np.array_equal(out, data["depth"])
