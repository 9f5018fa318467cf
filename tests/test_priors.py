"""Prior rasterisation (next row N2): numpy oracle vs the real reference functions (build container
only), and the CUDA scatter vs the oracle."""
import numpy as np
import pytest

from oracle import frames_numpy, frames_ref


def _events(seed, n, cols, rows, t_max, levels=20):
    rng = np.random.default_rng(seed)
    xs = rng.integers(0, cols, n)
    ys = rng.integers(0, rows, n)
    ts = np.sort(rng.integers(0, t_max, n)).astype(np.float64)
    perm = rng.permutation(n)                      # the SNN output is not time-ordered
    return xs[perm], ys[perm], ts[perm], rng.integers(0, levels, n).astype(np.float64)[perm]


CASES = [dict(seed=1, n=400, res=(20, 14), t_max=500, interval=100, pivots=[50.0, 100.0, 150.0, 400.0]),
         dict(seed=2, n=3000, res=(80, 60), t_max=1000, interval=250, pivots=list(np.arange(50.0, 1000.0, 50.0))),
         dict(seed=3, n=50, res=(9, 7), t_max=300, interval=100, pivots=None),
         dict(seed=4, n=0, res=(9, 7), t_max=300, interval=100, pivots=[10.0, 20.0]),
         dict(seed=5, n=2000, res=(16, 12), t_max=100, interval=100, pivots=[100.0])]


@pytest.mark.skipif(not frames_ref.available(), reason="reference checkout not present")
@pytest.mark.parametrize("case", CASES, ids=[str(c["seed"]) for c in CASES])
def test_numpy_oracle_matches_real_reference(case):
    _, reference = frames_ref.load()
    xs, ys, ts, zs = _events(case["seed"], case["n"], case["res"][0], case["res"][1], case["t_max"])
    if case["n"] == 0 or case["pivots"] is None and case["n"] < 1:
        pytest.skip("the reference cannot handle an empty event list (np.max of empty)")
    for fill in (-1, "nan"):
        want, want_ts = reference(case["res"], xs, ys, ts, zs, time_interval=case["interval"], pivots=case["pivots"],
                                  non_pixel_value=fill)
        got, got_ts = frames_numpy.rasterise(case["res"], xs, ys, ts, zs, time_interval=case["interval"],
                                             pivots=case["pivots"], non_pixel_value=fill)
        assert np.array_equal(got, want, equal_nan=True)
        assert np.array_equal(np.asarray(got_ts), np.asarray(want_ts))


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[str(c["seed"]) for c in CASES])
def test_cuda_rasterisation_matches_oracle(hsm, case):
    from importlib import import_module
    priors = import_module("hybrid-stereo-matching_b200.priors")
    xs, ys, ts, zs = _events(case["seed"], case["n"], case["res"][0], case["res"][1], case["t_max"])
    for fill in (-1, "nan"):
        if case["n"] == 0 and case["pivots"] is None:
            continue
        want, want_ts = frames_numpy.rasterise(case["res"], xs, ys, ts, zs, time_interval=case["interval"],
                                               pivots=case["pivots"], non_pixel_value=fill)
        got, got_ts = priors.generate_frames_from_spikes(case["res"], xs, ys, ts, zs, time_interval=case["interval"],
                                                         pivots=case["pivots"], non_pixel_value=fill)
        assert got.dtype == np.float64 and got.shape == want.shape
        assert np.array_equal(got, want, equal_nan=True)
        assert np.array_equal(np.asarray(got_ts), np.asarray(want_ts))
    with pytest.raises(IndexError):
        priors.generate_frames_from_spikes(case["res"], [case["res"][0]], [0], [1.0], [3.0], pivots=[5.0])


@pytest.mark.gpu
def test_rasterised_prior_feeds_the_matcher(hsm):
    """hybrid_stereo.py:169-183: events -> prior frames (-1 fill) -> run(prior_dict)."""
    from importlib import import_module
    priors = import_module("hybrid-stereo-matching_b200.priors")
    synth = import_module("hybrid-stereo-matching_b200.synth")
    from oracle import mrf_numpy
    lefts, rights, _ = synth.synthetic_sequence(3, 14, 20, 6, base_seed=9, with_prior=False)
    ts_frames = np.array([100.0, 200.0, 300.0])
    xs, ys, ts, zs = _events(11, 300, 20, 14, 300, levels=6)
    frames, stamps = priors.generate_frames_from_spikes((20, 14), xs, ys, ts, zs, time_interval=100,
                                                        pivots=ts_frames, non_pixel_value=-1)
    facade = hsm.FramebasedStereoMatching((20, 14), 6, inputs={"left": lefts, "right": rights, "ts": ts_frames})
    facade.run({"priors": frames, "ts": stamps})
    want_frames, _ = frames_numpy.rasterise((20, 14), xs, ys, ts, zs, time_interval=100, pivots=ts_frames,
                                            non_pixel_value=-1)
    oracle = mrf_numpy.OracleMRF((14, 20), 6)
    want = np.asarray([oracle.lbp(l, r, p, prior_trust_factor=1.0, prior_influence_mode="adaptive", n_iter=10)
                       for l, r, p in zip(lefts, rights, want_frames)])
    assert np.array_equal(facade.get_output(), want)


@pytest.mark.gpu
def test_contrast_normalisation_bits(hsm):
    """float64 results identical to the reference's expression, frame by frame."""
    from importlib import import_module
    priors = import_module("hybrid-stereo-matching_b200.priors")
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
    from importlib import import_module
    priors = import_module("hybrid-stereo-matching_b200.priors")
    rng = np.random.default_rng(5)
    frames = np.round(rng.random((3, 60, 80)) * 255.0)
    got = priors.prepare_frames(frames, crop_region=(11, 7), resolution=(40, 30), adjust_contrast=True)
    assert got.shape == (3, 30, 40)
    for i in range(3):
        want = frames_numpy.normalise_contrast(frames[i][7:37, 11:51])
        assert got[i].tobytes() == want.tobytes(), i
    plain = priors.prepare_frames(frames, crop_region=(11, 7), resolution=(40, 30))
    assert np.array_equal(plain, frames[:, 7:37, 11:51])
    with pytest.raises(NotImplementedError):
        priors.prepare_frames(frames, scale_down_factor=(2, 2))
