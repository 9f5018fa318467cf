"""Inputs that leave the fast-division range: tiny / denormal messages in the zero-cost border
columns (large D, many iterations), negative costs that grow (trust factors 10 and 100 of
evaluation.ipynb cell 20), non-finite values.  Every iteration kernel must still agree with the
C oracle bit for bit (NaNs compare equal to NaNs)."""
import numpy as np
import pytest

from oracle import mrf_c

pytestmark = pytest.mark.gpu

ALGOS = [5, 2, 3, 4, 1, 0]


def _run_all(hsm, shape, levels, left, right, prior, kwargs, algos=ALGOS):
    from importlib import import_module
    capi = import_module("hybrid-stereo-matching_b200._capi")
    want_obj = mrf_c.OracleMRFC(shape, levels)
    with np.errstate(all="ignore"):
        want = want_obj.lbp(left, right, prior, **kwargs)
    want_planes = want_obj._message_field
    for algo in algos:
        m = hsm.StereoMRF(shape, levels)
        m.set_option(capi.OPT_ALGORITHM, algo)
        got = m.lbp(left, right, prior, **kwargs)
        planes = m._message_field
        for name in ("south", "west", "north", "east", "data"):
            a, b = planes[name], want_planes[name]
            same = (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))
            assert same.all(), "algorithm %d plane %s: %d elements differ" % (algo, name, int((~same).sum()))
        assert np.array_equal(got, want), "algorithm %d labels" % algo
    return want_planes


def test_denormal_messages_in_border_columns(hsm):
    """D = 64, 45 iterations: values for far labels in the left columns decay below 2^-60 and into
    the denormal range, so the exact per-element path and div.rn.f32's slow path are exercised."""
    rng = np.random.default_rng(7)
    shape, levels = (10, 96), 64
    left, right = rng.random(shape), rng.random(shape)
    planes = _run_all(hsm, shape, levels, left, right, None, dict(n_iter=45))
    tiny = sum(int(((np.abs(p) < 2.0 ** -60) & (p != 0)).sum()) for k, p in planes.items() if k != "data")
    assert tiny > 0, "the case no longer produces tiny messages"


@pytest.mark.parametrize("trust", [10.0, 100.0])
def test_negative_costs_from_large_trust_factors(hsm, trust):
    rng = np.random.default_rng(8)
    shape, levels = (14, 40), 12
    left, right = rng.random(shape), rng.random(shape)
    prior = rng.integers(-1, levels, size=shape).astype(np.float64)
    _run_all(hsm, shape, levels, left, right, prior, dict(n_iter=12, prior_trust_factor=trust,
                                                          prior_influence_mode="const"))


def test_dense_negative_costs_small_label_count(hsm):
    """D = 2 with a dense prior and trust 100: whole pixels can have a negative max."""
    rng = np.random.default_rng(9)
    shape, levels = (9, 17), 2
    left, right = rng.random(shape), rng.random(shape)
    prior = rng.integers(0, levels, size=shape).astype(np.int32)
    _run_all(hsm, shape, levels, left, right, prior, dict(n_iter=8, prior_trust_factor=100.0,
                                                          prior_influence_mode="const"))


def test_exact_zero_costs_everywhere(hsm):
    """Identical, quantised images: most data costs are exactly zero, many ties."""
    rng = np.random.default_rng(10)
    shape, levels = (12, 30), 9
    base = np.round(rng.random((12, 30 + levels)) * 4) / 4
    left, right = base[:, levels:], base[:, levels:].copy()
    _run_all(hsm, shape, levels, left, right, None, dict(n_iter=10))


def test_non_finite_pixels_propagate_like_numpy(hsm):
    rng = np.random.default_rng(11)
    shape, levels = (8, 20), 6
    left, right = rng.random(shape), rng.random(shape)
    left[3, 7] = np.nan
    right[5, 11] = np.inf
    _run_all(hsm, shape, levels, left, right, None, dict(n_iter=3))


def test_huge_pixel_values(hsm):
    rng = np.random.default_rng(12)
    shape, levels = (8, 20), 6
    left, right = rng.random(shape) * 1e30, rng.random(shape) * 1e-30
    _run_all(hsm, shape, levels, left, right, None, dict(n_iter=4))
