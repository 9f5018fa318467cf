"""Prior adjustment by event optical flow (row N4) and down-scaling (row N3).

N4 is pinned to outputs of the REAL ``OpticalFlowPixelCorrection.adjust`` (tests/golden/a*.npz, written by
oracle/make_golden_facade.py, which runs the reference class including its velocity-field fit).  N3 has no such
pin (scikit-image is not installable here): the CUDA kernel is checked against the numpy restatement of
scikit-image 0.13's algorithm, and against properties that hold for any correct bilinear down-scaler."""
import numpy as np
import pytest

from oracle import facade_cases as F
from oracle import flow_numpy, flow_ref

IDS = [c["name"] for c in F.ADJUST_CASES]


def _product():
    from importlib import import_module
    return import_module("hybrid-stereo-matching_b200.priors")


@pytest.mark.parametrize("case", F.ADJUST_CASES, ids=IDS)
def test_numpy_adjust_matches_reference_fixture(case):
    _, frames = F.adjust_inputs(case)
    data = np.load(F.fixture_path(case), allow_pickle=False)
    got = flow_numpy.adjust(frames, case["res"], -1, case["time_arrow"])
    assert str(got.dtype) == str(data["dtype"]) and np.array_equal(got, data["adjusted"])


@pytest.mark.skipif(not flow_ref.available(), reason="reference checkout not present")
def test_adjust_fixture_is_what_the_real_class_returns():
    _, correction_class = flow_ref.load()
    case = F.ADJUST_CASES[1]
    events, frames = F.adjust_inputs(case)
    correction = correction_class(case["res"], events, buffer_pivots=case["pivots"], buffer_interval=case["interval"])
    got = correction.adjust(frames, prior_nondata_value=-1, time_arrow=case["time_arrow"])
    assert np.array_equal(got, np.load(F.fixture_path(case))["adjusted"])


@pytest.mark.gpu
@pytest.mark.parametrize("case", F.ADJUST_CASES, ids=IDS)
def test_cuda_adjust_matches_reference_fixture(hsm, case):
    events, frames = F.adjust_inputs(case)
    data = np.load(F.fixture_path(case), allow_pickle=False)
    correction = _product().OpticalFlowPixelCorrection(case["res"], events, buffer_pivots=case["pivots"],
                                                       buffer_interval=case["interval"])
    assert [len(i) for i in correction.reference_frames_indices] == \
        [int(((events[:, 0] >= p - case["interval"]) & (events[:, 0] <= p)).sum()) for p in case["pivots"]]
    got = correction.adjust(frames, prior_nondata_value=-1, time_arrow=case["time_arrow"])
    assert got.dtype == np.float32 and np.array_equal(got, data["adjusted"])
    with pytest.raises(AssertionError):
        correction.adjust(frames[:-1] if len(frames) > 1 else np.concatenate([frames, frames]))


@pytest.mark.gpu
def test_cuda_adjust_other_fill_values_and_full_size(hsm):
    rng = np.random.default_rng(4)
    cols, rows = 240, 180
    frames = np.full((3, rows, cols), np.nan)
    mask = rng.random(frames.shape) < 0.4
    frames[mask] = rng.integers(0, 32, int(mask.sum()))
    events = np.stack([np.linspace(0, 160, 50), np.zeros(50), np.zeros(50), np.zeros(50)], axis=1)
    correction = _product().OpticalFlowPixelCorrection((cols, rows), events, buffer_pivots=[50.0, 100.0, 150.0])
    for arrow in (+1, -1):
        got = correction.adjust(frames, prior_nondata_value=np.nan, time_arrow=arrow)
        want = flow_numpy.adjust(frames, (cols, rows), np.nan, arrow)
        assert np.array_equal(got, want, equal_nan=True)


@pytest.mark.gpu
@pytest.mark.parametrize("shape,factor", [((180, 240), (3, 3)), ((120, 140), (2, 2)), ((61, 83), (2, 3)),
                                          ((30, 40), (1, 1)), ((45, 70), (4, 2))])
def test_cuda_rescale_matches_restated_skimage_algorithm(hsm, shape, factor):
    rng = np.random.default_rng(shape[0])
    frames = np.round(rng.random((3,) + shape) * 255.0)               # uint8-like content, as imread returns
    frames[1] += 7.0                                                  # range not containing 0: cval is preserved
    got = _product().rescale_frames(frames, factor)
    for i in range(3):
        want = flow_numpy.rescale(frames[i], 1.0 / np.asarray(factor))
        assert got[i].shape == want.shape and got[i].tobytes() == want.tobytes(), i
    # any bilinear down-scaler: an integer factor f with odd f picks pixel (f*y + f//2, f*x + f//2) exactly
    if factor == (3, 3):
        assert np.array_equal(got[0], frames[0][1::3, 1::3])
    if factor == (2, 2):                                              # even factor: the mean of a 2 x 2 block
        f = frames[0]
        assert np.array_equal(got[0], 0.5 * (0.5 * f[0::2, 0::2] + 0.5 * f[0::2, 1::2])
                              + 0.5 * (0.5 * f[1::2, 0::2] + 0.5 * f[1::2, 1::2]))


@pytest.mark.gpu
def test_prepare_frames_full_pipeline(hsm):
    """crop -> down-scale -> contrast, the array part of load_frames for the shipped *_downsampled.yaml files
    (head_downsampled: crop (55, 25), resolution (140, 120), factor (2, 2))."""
    from oracle import frames_numpy
    rng = np.random.default_rng(8)
    frames = np.round(rng.random((2, 180, 240)) * 255.0)
    got = _product().prepare_frames(frames, crop_region=(55, 25), resolution=(140, 120), scale_down_factor=(2, 2),
                                    adjust_contrast=True)
    assert got.shape == (2, 60, 70)
    for i in range(2):
        small = flow_numpy.rescale(frames[i][25:145, 55:195], 0.5)
        assert got[i].tobytes() == frames_numpy.normalise_contrast(small).tobytes()


# ---- velocity field (vvf.py) ---------------------------------------------------------------------------
FLOW_IDS = [c["name"] for c in F.FLOW_CASES]


def _polarity_groups(events):
    events = events[np.argsort(events[:, 0], kind="stable")]
    positive = events[:, 3] == 0
    return events[positive][:, :-1], events[~positive][:, :-1]


@pytest.mark.parametrize("case", F.FLOW_CASES[:1] + F.FLOW_CASES[2:], ids=[FLOW_IDS[0], FLOW_IDS[2]])
def test_numpy_velocity_field_matches_reference_fixture(case):
    """The restatement against the REAL VelocityVectorField's output (float64, SVD least squares on both sides)."""
    data = np.load(F.fixture_path(case), allow_pickle=False)
    positive, negative = _polarity_groups(F.flow_events(case))
    for group, want in ((positive, data["positive"]), (negative, data["negative"])):
        group = group[:1500]                                   # the CPU suite's time budget
        got = flow_numpy.fit_velocity_field(group, time_interval=case["dt"])
        # a prefix of the stream sees the same neighbourhoods except within dt of its end
        keep = group[:, 0] < group[-1, 0] - case["dt"] if group.shape[0] else np.zeros(0, dtype=bool)
        assert np.allclose(got[keep], want[:group.shape[0]][keep], rtol=1e-9, atol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("case", F.FLOW_CASES, ids=FLOW_IDS)
def test_cuda_velocity_field_matches_reference_fixture(hsm, case):
    """float64 plane fits on the GPU (normal equations) vs the reference's scipy lstsq: equal to rounding.
    Tolerance: 1e-6 relative + 1e-9 absolute on each velocity component."""
    from importlib import import_module
    flow = import_module("hybrid-stereo-matching_b200.flow")
    data = np.load(F.fixture_path(case), allow_pickle=False)
    field = flow.VelocityVectorField(time_interval=case["dt"])
    got = field.fit_velocity_field(F.flow_events(case), assume_sorted=False, concatenate_polarity_groups=False)
    for name in ("positive", "negative"):
        want = data[name]
        assert got[name].shape == want.shape
        assert np.isfinite(got[name]).all()
        assert np.array_equal((np.abs(got[name]) > 1e-9).any(axis=1), (np.abs(want) > 1e-9).any(axis=1))
        assert np.allclose(got[name], want, rtol=1e-6, atol=1e-9), float(np.abs(got[name] - want).max())
    stacked = field.fit_velocity_field(F.flow_events(case))
    assert stacked.shape == (data["positive"].shape[0] + data["negative"].shape[0], 2)
