"""Prior rasterisation (next row N2) and frame preparation (N3).

CPU: the numpy checker against fixtures written by the REAL reference functions (oracle/make_golden_facade.py),
and -- in the build container -- against those functions live; the product's window logic against the
reference's segmentation.  GPU: the device rasterisation against the same fixtures.

Ties: the reference orders events by ``np.argsort`` (not stable), so for two events with the SAME timestamp on
the same pixel of a frame its result depends on the sort implementation.  Such pixels are found from the event
list and left out of the comparison; everything else must match exactly."""
import numpy as np
import pytest

from oracle import facade_cases as F
from oracle import frames_numpy, frames_ref

IDS = [c["name"] for c in F.RASTER_CASES]


def _ambiguous_pixels(case, xs, ys, ts, zs, lo, hi, inclusive):
    """Mask (N, rows, cols): the latest timestamp at that pixel inside the frame's window is shared by events
    with different values."""
    cols, rows = case["res"]
    mask = np.zeros((len(lo), rows, cols), dtype=bool)
    for f, (a, b) in enumerate(zip(lo, hi)):
        inside = (ts >= a) & ((ts <= b) if inclusive else (ts < b))
        latest = {}
        for e in np.flatnonzero(inside):
            key = (ys[e], xs[e])
            if key not in latest or ts[e] > latest[key][0]:
                latest[key] = (ts[e], {zs[e]})
            elif ts[e] == latest[key][0]:
                latest[key][1].add(zs[e])
        for (y, x), (_, values) in latest.items():
            mask[f, y, x] = len(values) > 1
    return mask


def _product():
    from importlib import import_module
    return import_module("hybrid-stereo-matching_b200.priors")


def _fixture(case):
    return np.load(F.fixture_path(case), allow_pickle=False)


@pytest.mark.parametrize("case", F.RASTER_CASES, ids=IDS)
def test_numpy_checker_matches_reference_fixture(case):
    xs, ys, ts, zs = F.raster_events(case)
    data = _fixture(case)
    lo, hi, inclusive, _ = _product()._windows(ts, case.get("start_time", 0), case["interval"], case["pivots"])
    keep = ~_ambiguous_pixels(case, xs, ys, ts, zs, lo, hi, inclusive)
    for fill in (-1, "nan"):
        got, stamps = frames_numpy.rasterise(case["res"], xs, ys, ts, zs, start_time=case.get("start_time", 0),
                                             time_interval=case["interval"], pivots=case["pivots"],
                                             non_pixel_value=fill)
        want = data["frames_%s" % fill]
        assert got.shape == want.shape
        assert np.array_equal(got[keep], want[keep], equal_nan=True)
        assert np.array_equal(np.asarray(stamps, dtype=np.float64), data["stamps_%s" % fill])
    if case.get("unique_times"):
        assert keep.all()


@pytest.mark.skipif(not frames_ref.available(), reason="reference checkout not present")
@pytest.mark.parametrize("case", F.RASTER_CASES, ids=IDS)
def test_window_logic_matches_real_reference_segmentation(case):
    """split_frames_by_time of the product (binary search over time windows) selects the same events per frame,
    in the same chronological order, as the reference's function."""
    split_ref, _ = frames_ref.load()
    xs, ys, ts, zs = F.raster_events(case)
    want, want_ts = split_ref(ts, case.get("start_time", 0), case["interval"], case["pivots"])
    got, got_ts = _product().split_frames_by_time(ts, case.get("start_time", 0), case["interval"], case["pivots"])
    assert len(want) == len(got)
    for a, b in zip(want, got):
        assert sorted(a.tolist()) == sorted(b.tolist())
        assert np.array_equal(ts[a], ts[b])
    assert np.array_equal(np.asarray(want_ts), np.asarray(got_ts))


@pytest.mark.gpu
@pytest.mark.parametrize("case", F.RASTER_CASES, ids=IDS)
def test_cuda_rasterisation_matches_reference_fixture(hsm, case):
    priors = _product()
    xs, ys, ts, zs = F.raster_events(case)
    data = _fixture(case)
    lo, hi, inclusive, _ = priors._windows(ts, case.get("start_time", 0), case["interval"], case["pivots"])
    keep = ~_ambiguous_pixels(case, xs, ys, ts, zs, lo, hi, inclusive)
    for fill in (-1, "nan"):
        got, stamps = priors.generate_frames_from_spikes(case["res"], xs, ys, ts, zs,
                                                         start_time=case.get("start_time", 0),
                                                         time_interval=case["interval"], pivots=case["pivots"],
                                                         non_pixel_value=fill)
        want = data["frames_%s" % fill]
        assert got.dtype == np.float64 and got.shape == want.shape
        assert np.array_equal(got[keep], want[keep], equal_nan=True)
        assert np.array_equal(np.asarray(stamps, dtype=np.float64), data["stamps_%s" % fill])
        # tie pixels: one of the tied events' values, namely the one a stable sort puts last
        stable, _ = frames_numpy.rasterise(case["res"], xs, ys, ts, zs, start_time=case.get("start_time", 0),
                                           time_interval=case["interval"], pivots=case["pivots"],
                                           non_pixel_value=fill, sort_kind="stable")
        assert np.array_equal(got, stable, equal_nan=True)
    # the device-resident form: int32 labels, -1 where there is no event
    if case["pivots"] is not None:
        labels, _ = priors.rasterise_priors_on_device(case["res"], xs, ys, ts, zs,
                                                      start_time=case.get("start_time", 0),
                                                      time_interval=case["interval"], pivots=case["pivots"])
        host = labels.download()
        assert host.dtype == np.int32 and np.array_equal(host[keep], data["frames_-1"][keep].astype(np.int32))
    with pytest.raises(IndexError):
        priors.generate_frames_from_spikes(case["res"], [case["res"][0]], [0], [1.0], [3.0], pivots=[5.0])
    # an out-of-range event outside every window is never indexed by the reference either
    priors.generate_frames_from_spikes(case["res"], [case["res"][0]], [0], [100.0], [3.0], pivots=[5.0],
                                       time_interval=1)


@pytest.mark.gpu
def test_empty_and_odd_inputs(hsm):
    priors = _product()
    frames, stamps = priors.generate_frames_from_spikes((9, 7), [], [], [], [], pivots=[10.0, 20.0],
                                                        non_pixel_value=-1)
    assert frames.shape == (2, 7, 9) and (frames == -1).all()
    frames, _ = priors.generate_frames_from_spikes((9, 7), [-1, 2], [-7, 3], [5.0, 6.0], [4.0, 2.5],
                                                   pivots=[10.0], non_pixel_value="nan")
    assert frames[0, 0, 8] == 4.0 and frames[0, 3, 2] == 2.5 and np.isnan(frames).sum() == 61
    labels, _ = priors.rasterise_priors_on_device((9, 7), [-1, 2], [-7, 3], [5.0, 6.0], [4.0, 2.5], pivots=[10.0])
    host = labels.download()
    assert host[0, 0, 8] == 4 and host[0, 3, 2] == -1 and (host == -1).sum() == 62     # 2.5 matches no level


@pytest.mark.gpu
def test_device_resident_priors_feed_the_matcher(hsm):
    """hybrid_stereo.py:169-183 with the priors never leaving the GPU: events -> int32 label frames on the device
    -> run(prior_dict) == the same run with the reference-format float64 host frames."""
    from importlib import import_module
    priors = _product()
    synth = import_module("hybrid-stereo-matching_b200.synth")
    from oracle import mrf_numpy
    case = F.RASTER_CASES[4]
    xs, ys, ts, zs = F.raster_events(case)
    cols, rows = case["res"]
    ts_frames = np.asarray(case["pivots"][:7])
    lefts, rights, _ = synth.synthetic_sequence(len(ts_frames), rows, cols, 22, base_seed=9, with_prior=False)
    on_device, stamps = priors.rasterise_priors_on_device((cols, rows), xs, ys, ts, zs, time_interval=case["interval"],
                                                          pivots=ts_frames)
    facade = hsm.FramebasedStereoMatching((cols, rows), 22, inputs={"left": lefts, "right": rights, "ts": ts_frames},
                                          capacity=3)
    facade.run({"priors": on_device, "ts": stamps})
    want_frames = _fixture(case)["frames_-1"][:len(ts_frames)]
    oracle = mrf_numpy.OracleMRF((rows, cols), 22)
    want = np.asarray([oracle.lbp(l, r, p, prior_trust_factor=1.0, prior_influence_mode="adaptive", n_iter=10)
                       for l, r, p in zip(lefts, rights, want_frames)])
    assert np.array_equal(facade.get_output(), want)
    # and through the host-array form of the same frames
    host_frames, _ = priors.generate_frames_from_spikes((cols, rows), xs, ys, ts, zs, time_interval=case["interval"],
                                                        pivots=ts_frames, non_pixel_value=-1)
    facade = hsm.FramebasedStereoMatching((cols, rows), 22, inputs={"left": lefts, "right": rights, "ts": ts_frames})
    facade.run({"priors": host_frames, "ts": stamps})
    assert np.array_equal(facade.get_output(), want)


@pytest.mark.gpu
def test_contrast_normalisation_bits(hsm):
    """float64 results identical to the reference's expression, frame by frame."""
    priors = _product()
    rng = np.random.default_rng(3)
    frames = rng.random((5, 37, 53)) * 255.0
    frames[1] = np.round(frames[1])                 # uint8-like content
    frames[2] = 7.5                                 # flat frame -> ones
    frames[3, 4, 5] = -1e9
    got = priors.normalise_contrast(frames)
    for i in range(5):
        want = frames_numpy.normalise_contrast(frames[i])
        assert got[i].tobytes() == want.tobytes(), i
    assert priors.normalise_contrast(frames[0]).tobytes() == frames_numpy.normalise_contrast(frames[0]).tobytes()


@pytest.mark.gpu
def test_prepare_frames_crop_and_contrast(hsm):
    """crop (x, y order) + per-frame contrast stretch == frames_io.py:91-95, 103-109, 130-131."""
    priors = _product()
    rng = np.random.default_rng(5)
    frames = np.round(rng.random((3, 60, 80)) * 255.0)
    got = priors.prepare_frames(frames, crop_region=(11, 7), resolution=(40, 30), adjust_contrast=True)
    assert got.shape == (3, 30, 40)
    for i in range(3):
        want = frames_numpy.normalise_contrast(frames[i][7:37, 11:51])
        assert got[i].tobytes() == want.tobytes(), i
    plain = priors.prepare_frames(frames, crop_region=(11, 7), resolution=(40, 30))
    assert np.array_equal(plain, frames[:, 7:37, 11:51])
