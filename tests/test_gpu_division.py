"""The fused kernel divides a whole pixel by one max with a shared reciprocal
(kernels.cuh, `make_divisor` / `div4`).  It must give the bits of IEEE division
(div.rn.f32, what numpy's float32 `/` computes) for every operand, including zeros,
denormals, huge values, infinities and hard-to-round mantissas."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _mismatches(a4, b):
    from importlib import import_module
    capi = import_module("hybrid-stereo-matching_b200._capi")
    a4 = np.ascontiguousarray(a4, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    assert a4.shape == (b.size, 4)
    bad = ctypes.c_uint64(12345)
    capi.check(capi.lib().hsm_selftest_division(a4.ctypes.data_as(ctypes.c_void_p),
                                                b.ctypes.data_as(ctypes.c_void_p), b.size, ctypes.byref(bad)))
    return bad.value


def _random_floats(rng, n, exp_lo, exp_hi, signed=True):
    mant = rng.integers(0, 1 << 23, size=n, dtype=np.uint32)
    # a share of hard mantissas: all ones, all zeros, one bit, alternating
    special = rng.integers(0, 8, size=n)
    mant = np.where(special == 0, np.uint32(0x7fffff), mant)
    mant = np.where(special == 1, np.uint32(0), mant)
    mant = np.where(special == 2, np.uint32(1), mant)
    mant = np.where(special == 3, np.uint32(0x7ffffe), mant)
    expo = rng.integers(exp_lo + 127, exp_hi + 128, size=n, dtype=np.uint32)
    sign = rng.integers(0, 2, size=n, dtype=np.uint32) if signed else np.zeros(n, dtype=np.uint32)
    return ((sign << 31) | (expo << 23) | mant).astype(np.uint32).view(np.float32)


def test_division_matches_ieee_in_message_range(hsm):
    """Operands as the matcher sees them: numerators 0 .. a few, divisor = a positive max."""
    rng = np.random.default_rng(1)
    n = 1 << 21
    b = _random_floats(rng, n, -8, 3, signed=False)
    a = _random_floats(rng, 4 * n, -30, 3).reshape(n, 4)
    a[rng.random((n, 4)) < 0.1] = 0.0
    assert _mismatches(a, b) == 0
    # numerators <= divisor, like U / max(U)
    frac = rng.random((n, 4)).astype(np.float32)
    assert _mismatches(frac * b[:, None], b) == 0


def test_division_matches_ieee_full_exponent_range(hsm):
    rng = np.random.default_rng(2)
    n = 1 << 21
    b = _random_floats(rng, n, -127, 127)        # includes denormals (biased exponent 0)
    a = _random_floats(rng, 4 * n, -127, 127).reshape(n, 4)
    assert _mismatches(a, b) == 0


def test_fast_sequence_range(hsm):
    """The unchecked sequence of normalise_fast() claims 2^-100 <= a <= m <= 2^26; the self-test kernel runs it
    wherever that holds and compares with div.rn.f32.  Small numerators against ordinary and small maxima,
    hard mantissas, and the corners of the range."""
    rng = np.random.default_rng(7)
    n = 1 << 21
    for m_lo, m_hi in ((-2, 3), (-100, 26), (20, 26), (-100, -90)):
        m = _random_floats(rng, n, m_lo, m_hi, signed=False)
        m = np.minimum(m, np.float32(2.0 ** 26))
        a = _random_floats(rng, 4 * n, -100, 26, signed=False).reshape(n, 4)
        a = np.minimum(a, m[:, None])                      # numerators never exceed the max
        a[:, 0] = np.where(rng.random(n) < 0.5, m, a[:, 0])   # the max itself is one of the numerators
        assert _mismatches(a, m) == 0, (m_lo, m_hi)
    # quotients just above the smallest normal number: a near 2^-100, m near 2^26
    m = np.full(n, 2.0 ** 26, dtype=np.float32)
    m[::2] = np.nextafter(m[::2], np.float32(0))
    a = _random_floats(rng, 4 * n, -100, -98, signed=False).reshape(n, 4)
    a[:, 3] = np.float32(2.0 ** -100)
    assert _mismatches(a, m) == 0


def test_division_special_values(hsm):
    specials = np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, 1e-45, -1e-45, 1.1754944e-38, 3.4028235e38,
                         2.0 ** -60, 2.0 ** 60, np.nextafter(np.float32(2.0 ** -60), np.float32(0)),
                         np.nextafter(np.float32(2.0 ** 60), np.float32(np.inf)), 3.0, 1.0 / 3.0, 2.0 ** -100,
                         2.0 ** 100, 0.99999994, 1.0000001], dtype=np.float32)
    aa, bb = np.meshgrid(specials, specials, indexing="ij")
    a4 = np.repeat(aa.reshape(-1, 1), 4, axis=1)
    a4[:, 1] = -a4[:, 1]
    a4[:, 2] = 0.0
    a4[:, 3] = 1.0
    assert _mismatches(a4, bb.reshape(-1)) == 0


def test_signed_zero_bits(hsm):
    """-0 / m must stay -0 for m > 0 (bit pattern, not just value)."""
    m = hsm.StereoMRF((4, 6), 1)
    left = np.full((4, 6), 0.5)
    right = np.full((4, 6), 0.25)
    prior = np.zeros((4, 6))
    m.loop_belief(left, right, prior, prior_trust_factor=3.0, n_iter=2)       # all costs negative, D = 1
    from oracle import mrf_numpy
    o = mrf_numpy.OracleMRF((4, 6), 1)
    o.loop_belief(left, right, prior, prior_trust_factor=3.0, n_iter=2)
    for name, plane in o._message_field.items():
        assert m._message_field[name].tobytes() == plane.tobytes(), name
