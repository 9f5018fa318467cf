"""The CPU oracles against the golden fixtures the REAL reference produced
(oracle/make_golden.py), and -- when /root/reference is present, i.e. in the build
container -- against the reference itself, live.  No GPU needed."""
import hashlib

import numpy as np
import pytest

from oracle import cases as C
from oracle import mrf_c, mrf_numpy, ref_loader

ALL = [(c, True) for c in C.SMALL_CASES] + [(c, False) for c in C.BIG_CASES]
IDS = [c["name"] for c, _ in ALL]


def _check_against_fixture(factory, case, small):
    labels, planes = C.run_case(factory, case)
    data, meta = C.load_fixture(case)
    assert labels.dtype == np.int64 and labels.shape == data["labels"].shape
    assert np.array_equal(labels, data["labels"])
    assert hashlib.sha1(labels.astype(np.int32).tobytes()).hexdigest() == meta["labels_sha1"]
    for name, plane in planes.items():
        assert plane.dtype == np.float32
        assert C.plane_sha1(plane) == meta["plane_sha1"][name], name
        if small:
            assert np.array_equal(plane, data["plane_" + name], equal_nan=True), name


@pytest.mark.parametrize("case,small", ALL, ids=IDS)
def test_numpy_oracle_matches_reference_fixture(case, small):
    _check_against_fixture(lambda dim, levels: mrf_numpy.OracleMRF(tuple(dim), levels), case, small)


@pytest.mark.parametrize("case,small", ALL, ids=IDS)
def test_c_oracle_matches_reference_fixture(case, small):
    _check_against_fixture(lambda dim, levels: mrf_c.OracleMRFC(tuple(dim), levels), case, small)


HUGE_NUMPY = [c for c in C.HUGE_CASES if "numpy" in c["cpu"]]
HUGE_C = [c for c in C.HUGE_CASES if "c" in c["cpu"]]


@pytest.mark.parametrize("case", HUGE_NUMPY, ids=[c["name"] for c in HUGE_NUMPY])
def test_numpy_oracle_matches_reference_fixture_huge(case):
    _check_against_fixture(lambda dim, levels: mrf_numpy.OracleMRF(tuple(dim), levels), case, False)


@pytest.mark.parametrize("case", HUGE_C, ids=[c["name"] for c in HUGE_C])
def test_c_oracle_matches_reference_fixture_huge(case):
    """Includes C3 at its real 100 iterations (b07).  The 1080p D=256 fixtures (b08, b09) are checked by the
    GPU tests only: one CPU pass needs ~20 GB and minutes."""
    _check_against_fixture(lambda dim, levels: mrf_c.OracleMRFC(tuple(dim), levels), case, False)


def test_every_huge_fixture_exists_and_names_its_inputs():
    for case in C.HUGE_CASES:
        data, meta = C.load_fixture(case)
        assert meta["name"] == case["name"] and len(meta["inputs_sha1"]) == 40
        assert data["labels"].shape == (case["rows"], case["cols"])
        assert set(meta["plane_sha1"]) == {"south", "west", "north", "east", "data"}


def test_known_answer_sha1_from_survey():
    """KAT-A / KAT-B label hashes recorded independently in SURVEY.md section 8c."""
    _, meta_a = C.load_fixture(C.BIG_CASES[0])
    _, meta_b = C.load_fixture(C.BIG_CASES[1])
    assert meta_a["labels_sha1"] == "52fc8bcfd4002900331def289de0c1bdbb138d26"
    assert meta_b["labels_sha1"] == "578b5823ed294c4d54d4d9e5c7c73cf891dd1838"
    data, _ = C.load_fixture(C.BIG_CASES[0])
    hist = np.bincount(data["labels"].ravel().astype(np.int64), minlength=15)
    assert hist.tolist() == [1689, 72, 2108, 77, 15705, 34266, 14774, 1143, 14483, 260, 6759, 6426,
                             2943, 650, 9237]


def test_fixture_inputs_have_not_drifted():
    """The seeded generator must still produce the inputs the fixtures were made from."""
    for case in C.SMALL_CASES[:6] + C.BIG_CASES[2:3]:
        _, meta = C.load_fixture(case)
        digest = hashlib.sha1()
        for index in range(len(case["calls"])):
            left, right, prior, _ = C.case_inputs(case, index)
            digest.update(np.ascontiguousarray(left).tobytes())
            digest.update(np.ascontiguousarray(right).tobytes())
            if prior is not None:
                digest.update(np.ascontiguousarray(prior).tobytes())
        assert digest.hexdigest() == meta["inputs_sha1"], case["name"]


@pytest.mark.skipif(not ref_loader.reference_available(), reason="reference checkout not present")
@pytest.mark.parametrize("case", C.SMALL_CASES[:12], ids=[c["name"] for c in C.SMALL_CASES[:12]])
def test_oracles_match_live_reference(case):
    reference = ref_loader.reference_class()
    want_labels, want_planes = C.run_case(lambda dim, levels: reference(tuple(dim), levels), case)
    for factory in (lambda dim, levels: mrf_numpy.OracleMRF(tuple(dim), levels),
                    lambda dim, levels: mrf_c.OracleMRFC(tuple(dim), levels)):
        labels, planes = C.run_case(factory, case)
        assert np.array_equal(labels, want_labels)
        for name in want_planes:
            assert np.array_equal(planes[name], want_planes[name], equal_nan=True), name


@pytest.mark.skipif(not ref_loader.reference_available(), reason="reference checkout not present")
def test_live_reference_random_shapes():
    """Random small shapes / parameters: both oracles == reference, bit for bit."""
    reference = ref_loader.reference_class()
    rng = np.random.default_rng(2024)
    for trial in range(25):
        rows, cols = int(rng.integers(1, 14)), int(rng.integers(2, 24))
        levels = int(rng.integers(1, min(cols + 1, 20) + 1))
        n_iter = int(rng.integers(0, 5))
        left, right = rng.random((rows, cols)), rng.random((rows, cols))
        prior = None
        kwargs = dict(n_iter=n_iter)
        if trial % 3:
            prior = rng.integers(-1, levels, size=(rows, cols)).astype([np.float64, np.int32, np.float32][trial % 3])
            kwargs.update(prior_trust_factor=[0.0, 0.5, 1.0, 10.0][trial % 4],
                          prior_influence_mode=["const", "adaptive"][trial % 2])
        ref = reference((rows, cols), levels)
        want = ref.lbp(left, right, prior, **kwargs)
        for cls in (mrf_numpy.OracleMRF, mrf_c.OracleMRFC):
            got_obj = cls((rows, cols), levels)
            got = got_obj.lbp(left, right, prior, **kwargs)
            assert np.array_equal(got, want), (trial, cls.__name__)
            for name, plane in ref._message_field.items():
                assert np.array_equal(got_obj._message_field[name], plane), (trial, cls.__name__, name)


def test_float64_scalar_trust_factor_uses_float64_product():
    """numpy float64 scalar factor: product computed in float64 (NEP 50), both oracles agree."""
    rng = np.random.default_rng(5)
    left, right = rng.random((9, 14)), rng.random((9, 14))
    prior = rng.integers(-1, 6, size=(9, 14)).astype(np.float64)
    a = mrf_numpy.OracleMRF((9, 14), 6)
    b = mrf_c.OracleMRFC((9, 14), 6)
    la = a.lbp(left, right, prior, prior_trust_factor=np.float64(0.3), n_iter=2)
    lb = b.lbp(left, right, prior, prior_trust_factor=np.float64(0.3), n_iter=2)
    assert np.array_equal(la, lb)
    for name in a._message_field:
        assert np.array_equal(a._message_field[name], b._message_field[name]), name
