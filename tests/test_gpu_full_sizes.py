"""BASELINE.json's larger configurations at full size, checked through properties that do not
need the (minutes-long) CPU oracle: two independent CUDA implementations agree bit for bit
(fused TMA kernel vs one-kernel-per-direction with stock div.rn.f32), the row-band split equals the
un-banded run, messages are max-normalised, border lines are zero.  C3 additionally matches the
golden fixture of the reference for its first iterations (b05, tests/test_gpu_parity.py)."""
import hashlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def _check_planes(planes):
    for name in ("south", "west", "north", "east"):
        peak = planes[name].max(axis=0)
        assert np.all((peak == 1.0) | (peak == 0.0)), name
    assert not planes["south"][:, 0, :].any() and not planes["north"][:, -1, :].any()
    assert not planes["west"][:, :, -1].any() and not planes["east"][:, :, 0].any()


def test_c3_vga_d64_100_iterations(hsm):
    """configs[2]: synthetic 640x480, D=64, 100 BP iterations."""
    from importlib import import_module
    capi = import_module("hybrid-stereo-matching_b200._capi")
    synth = import_module("hybrid-stereo-matching_b200.synth")
    rows, cols, levels, n_iter = 480, 640, 64, 100
    left, right, disparity = synth.textured_pair(rows, cols, levels, 4321)
    prior = synth.sparse_prior(disparity, levels, 4321)
    kw = dict(prior_trust_factor=1.0, prior_influence_mode="adaptive", n_iter=n_iter)
    fast = hsm.StereoMRF((rows, cols), levels)
    labels = fast.lbp(left, right, prior, **kw)
    slow = hsm.StereoMRF((rows, cols), levels)
    slow.set_option(capi.OPT_ALGORITHM, 0)
    want = slow.lbp(left, right, prior, **kw)
    assert labels.dtype == np.int64 and np.array_equal(labels, want)
    planes, want_planes = fast._message_field, slow._message_field
    for name in planes:
        assert _sha(planes[name]) == _sha(want_planes[name]), name
    _check_planes(planes)
    # the matcher recovers the synthetic disparity away from the zero-cost left border
    inner = (slice(None), slice(levels, None))
    assert (labels[inner] == disparity[inner]).mean() > 0.9


def test_c5_1080p_d256_bands_equal_single(hsm):
    """configs[4]: 1920x1080, D=256 -- three iterations; single matcher vs per-direction kernels vs
    a 3-band split with halo exchange (device copies stand in for NCCL on one GPU)."""
    import torch
    from importlib import import_module
    capi = import_module("hybrid-stereo-matching_b200._capi")
    bands = import_module("hybrid-stereo-matching_b200.bands")
    synth = import_module("hybrid-stereo-matching_b200.synth")
    rows, cols, levels, n_iter, world = 1080, 1920, 256, 3, 3
    left, right, disparity = synth.textured_pair(rows, cols, levels, 99)
    prior = synth.sparse_prior(disparity, levels, 99)
    kw = dict(prior_trust_factor=0.5, prior_influence_mode="const")

    single = hsm.StereoMRF((rows, cols), levels)
    labels = single.lbp(left, right, prior, n_iter=n_iter, **kw)
    west_sha = _sha(single._plane(1))
    del single
    reference = hsm.StereoMRF((rows, cols), levels)
    reference.set_option(capi.OPT_ALGORITHM, 0)
    assert np.array_equal(reference.lbp(left, right, prior, n_iter=n_iter, **kw), labels)
    assert _sha(reference._plane(1)) == west_sha
    del reference
    torch.cuda.empty_cache()

    engines = [bands.CudaBandEngine(rows, cols, levels, *bands.band_bounds(rows, world, r)) for r in range(world)]
    for e in engines:
        with e.stream_context():
            e.init_fields(left, right, prior, kw["prior_trust_factor"], kw["prior_influence_mode"])
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
    assert np.array_equal(got, labels)
