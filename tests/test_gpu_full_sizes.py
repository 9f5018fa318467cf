"""BASELINE.json's larger configurations at full size, pinned to outputs of the REAL reference
(tests/golden/b07-b09, written by oracle/make_golden.py in the build container):

  * C3 (640x480, D=64) at its real 100 iterations: b07 -- labels and all five planes (sha1) for the automatic
    choice (persistent dataflow kernel), the one-launch-per-iteration default kernel, the TMA-load fallback,
    and inside a batch (frame groups);
  * C5 (1920x1080, D=256, event prior): b08 (1 iteration, 'const' 0.5) and b09 (2 iterations, 'adaptive') --
    labels and planes for the automatic choice, labels + one plane for the other kernels, and the row-band
    split with halo exchange (device copies stand in for NCCL on one GPU; the cross-GPU paths are covered by
    bench.py --gpus N, which checks the gathered labels against the same fixture).

No test here compares one CUDA kernel with another: every comparator is a reference output."""
import gc
import hashlib

import numpy as np
import pytest

from oracle import cases as C

pytestmark = pytest.mark.gpu


def _capi():
    from importlib import import_module
    return import_module("hybrid-stereo-matching_b200._capi")


def _plane_sha(matcher, which, frame=0):
    """sha1 of one plane in the reference's (D, H, W) layout; one plane at a time (2.1 GB each at 1080p)."""
    plane = matcher._plane(which, frame)
    digest = hashlib.sha1(plane.tobytes()).hexdigest()
    del plane
    gc.collect()
    return digest


def _check_planes_normalised(matcher):
    for which, name in ((0, "south"), (1, "west"), (2, "north"), (3, "east")):
        plane = matcher._plane(which)
        peak = plane.max(axis=0)
        assert np.all((peak == 1.0) | (peak == 0.0)), name
    assert not matcher._plane(0)[:, 0, :].any() and not matcher._plane(2)[:, -1, :].any()
    assert not matcher._plane(1)[:, :, -1].any() and not matcher._plane(3)[:, :, 0].any()


C3 = C.find_case("b07_vga_d64_it100")


@pytest.mark.parametrize("algorithm", [None, 5, 2, 6], ids=["auto", "tmast", "tma", "flow"])
def test_c3_vga_d64_100_iterations_matches_reference(hsm, algorithm):
    """configs[2]: synthetic 640x480, D=64, 100 BP iterations -- against the reference's own output."""
    left, right, prior, kwargs = C.case_inputs(C3, 0)
    data, meta = C.load_fixture(C3)
    matcher = hsm.StereoMRF(left.shape, C3["levels"])
    if algorithm is not None:
        matcher.set_option(_capi().OPT_ALGORITHM, algorithm)
    kwargs = dict(kwargs)
    n_iter = kwargs.pop("n_iter")
    labels = matcher.lbp(left, right, prior, n_iter=n_iter, **kwargs)
    assert labels.dtype == np.int64 and np.array_equal(labels, data["labels"])
    for which, name in enumerate(_capi().PLANES):
        assert _plane_sha(matcher, which) == meta["plane_sha1"][name], name
    if algorithm is None:
        _check_planes_normalised(matcher)


def test_c3_inside_a_batch_matches_reference(hsm):
    """The same frame as frame 1 of a batch of 3 (two concurrent frame groups): identical to the reference."""
    from importlib import import_module
    synth = import_module("hybrid-stereo-matching_b200.synth")
    left, right, prior, kwargs = C.case_inputs(C3, 0)
    data, meta = C.load_fixture(C3)
    lefts, rights, priors = synth.synthetic_sequence(3, 480, 640, 64, base_seed=5)
    lefts[1], rights[1], priors[1] = left, right, prior
    matcher = hsm.StereoMRF(left.shape, 64, capacity=3)
    labels = matcher.lbp_batch(lefts, rights, priors, **kwargs)
    assert np.array_equal(labels[1], data["labels"])
    for which, name in enumerate(_capi().PLANES):
        assert _plane_sha(matcher, which, frame=1) == meta["plane_sha1"][name], name


C5_CASES = [C.find_case("b08_1080p_d256_it1"), C.find_case("b09_1080p_d256_it2")]


@pytest.mark.parametrize("case", C5_CASES, ids=[c["name"] for c in C5_CASES])
def test_c5_1080p_d256_matches_reference(hsm, case):
    """configs[4] on one GPU: automatic kernel choice, all five planes; other kernels: labels + west plane."""
    left, right, prior, kwargs = C.case_inputs(case, 0)
    data, meta = C.load_fixture(case)
    want = data["labels"].astype(np.int64)
    kwargs = dict(kwargs)
    n_iter = kwargs.pop("n_iter")
    matcher = hsm.StereoMRF(left.shape, case["levels"])
    labels = matcher.lbp(left, right, prior, n_iter=n_iter, **kwargs)
    assert np.array_equal(labels, want)
    for which, name in enumerate(_capi().PLANES):
        assert _plane_sha(matcher, which) == meta["plane_sha1"][name], name
    del matcher
    gc.collect()
    for algorithm in (2, 6, 1):          # TMA loads + plain stores (the D=256 fallback), dataflow, LDG
        other = hsm.StereoMRF(left.shape, case["levels"])
        other.set_option(_capi().OPT_ALGORITHM, algorithm)
        assert np.array_equal(other.lbp(left, right, prior, n_iter=n_iter, **kwargs), want), algorithm
        assert _plane_sha(other, 1) == meta["plane_sha1"]["west"], algorithm
        del other
        gc.collect()


def test_c5_1080p_d256_row_bands_match_reference(hsm):
    """configs[4] split into 3 row bands with a halo exchange per iteration == the reference's un-banded run."""
    import torch
    from importlib import import_module
    bands = import_module("hybrid-stereo-matching_b200.bands")
    case = C5_CASES[1]
    left, right, prior, kwargs = C.case_inputs(case, 0)
    data, meta = C.load_fixture(case)
    rows, cols, levels, world = case["rows"], case["cols"], case["levels"], 3
    n_iter = kwargs["n_iter"]
    engines = [bands.CudaBandEngine(rows, cols, levels, *bands.band_bounds(rows, world, r)) for r in range(world)]
    for e in engines:
        with e.stream_context():
            e.init_fields(left, right, prior, kwargs["prior_trust_factor"], kwargs["prior_influence_mode"])
    tops = [e.new_halo() for e in engines]
    bottoms = [e.new_halo() for e in engines]
    for it in range(n_iter):
        last = it == n_iter - 1
        for e in engines:
            with e.stream_context():
                e.iterate(finalize=last)
        if last:
            break
        for r, e in enumerate(engines):
            with e.stream_context():
                e.pack(tops[r] if r > 0 else None, bottoms[r] if r < world - 1 else None)
        torch.cuda.synchronize()
        for r, e in enumerate(engines):
            with e.stream_context():
                e.unpack(bottoms[r - 1] if r > 0 else None, tops[r + 1] if r < world - 1 else None)
        torch.cuda.synchronize()
    got = np.concatenate([e.labels() for e in engines], axis=0)
    assert np.array_equal(got, data["labels"].astype(np.int64))
    west = np.concatenate([e.matcher._plane(1)[:, 1:e.own + 1, :] for e in engines], axis=1)
    assert hashlib.sha1(np.ascontiguousarray(west).tobytes()).hexdigest() == meta["plane_sha1"]["west"]
