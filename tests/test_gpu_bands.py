"""Row-band mode of the CUDA matcher on ONE GPU: several band engines run in lock-step and swap
their halo rows by device copies (what NCCL does between GPUs).  Must equal the un-banded CUDA
result and the oracle bit for bit."""
import numpy as np
import pytest

from oracle import mrf_c

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world,shape,n_iter", [(2, (21, 50, 12), 5), (3, (20, 70, 32), 4), (4, (9, 33, 6), 6)])
def test_cuda_band_engines_match_unbanded(hsm, world, shape, n_iter):
    import torch
    from importlib import import_module
    bands = import_module("hybrid-stereo-matching_b200.bands")
    synth = import_module("hybrid-stereo-matching_b200.synth")
    rows, cols, levels = shape
    left, right, disparity = synth.textured_pair(rows, cols, levels, 9, block=(5, 6))
    prior = synth.sparse_prior(disparity, levels, 9, density=0.2)
    kwargs = dict(prior_trust_factor=0.5, prior_influence_mode="const")

    want = mrf_c.OracleMRFC((rows, cols), levels)
    want_labels = want.lbp(left, right, prior, n_iter=n_iter, **kwargs)
    single = hsm.StereoMRF((rows, cols), levels)
    assert np.array_equal(single.lbp(left, right, prior, n_iter=n_iter, **kwargs), want_labels)

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
    assert np.array_equal(got, want_labels)
    want_planes = want._message_field
    for r, e in enumerate(engines):
        first, last_row = bands.band_bounds(rows, world, r)
        for name, plane in e.planes().items():
            assert np.array_equal(plane, want_planes[name][:, first:last_row + 1, :]), (r, name)


@pytest.mark.parametrize("world,shape,n_iter,levels_algo", [(2, (21, 50), 6, (12, 5)), (3, (20, 70), 5, (32, 5)),
                                                            (3, (12, 230), 4, (200, 2))])
def test_fused_peer_push_matches_unbanded(hsm, world, shape, n_iter, levels_algo):
    """The fused halo exchange on ONE GPU: the band engines live in one process, attach to each other by
    address and their iteration kernels push boundary rows into each other's ghost rows; ordering comes
    from stream-ordered flag waits (no spinning kernels, so two streams on one GPU cannot deadlock)."""
    import torch
    from importlib import import_module
    bands = import_module("hybrid-stereo-matching_b200.bands")
    synth = import_module("hybrid-stereo-matching_b200.synth")
    capi = import_module("hybrid-stereo-matching_b200._capi")
    rows, cols = shape
    levels, algorithm = levels_algo
    left, right, disparity = synth.textured_pair(rows, cols, levels, 19, block=(5, 6))
    prior = synth.sparse_prior(disparity, levels, 19, density=0.2)
    want = mrf_c.OracleMRFC((rows, cols), levels)
    want_labels = want.lbp(left, right, prior, prior_trust_factor=1.0, prior_influence_mode="adaptive", n_iter=n_iter)

    engines = [bands.CudaBandEngine(rows, cols, levels, *bands.band_bounds(rows, world, r)) for r in range(world)]
    for e in engines:
        e.matcher.set_option(capi.OPT_ALGORITHM, algorithm)
        with e.stream_context():
            e.init_fields(left, right, prior, 1.0, "adaptive")
    infos = [e.peer_info() for e in engines]
    for r, e in enumerate(engines):
        if r > 0:
            e.attach_peer(0, infos[r - 1], same_process=True)
        if r < world - 1:
            e.attach_peer(1, infos[r + 1], same_process=True)
        e.reset_peers()
    torch.cuda.synchronize()
    for it in range(n_iter):                       # every rank queues its iteration; flags order them on the device
        for e in engines:
            with e.stream_context():
                e.iterate(finalize=(it == n_iter - 1))
    torch.cuda.synchronize()
    got = np.concatenate([e.labels() for e in engines], axis=0)
    assert np.array_equal(got, want_labels)
    want_planes = want._message_field
    for r, e in enumerate(engines):
        first, last_row = bands.band_bounds(rows, world, r)
        for name, plane in e.planes().items():
            assert np.array_equal(plane, want_planes[name][:, first:last_row + 1, :]), (r, name)
        e.detach_peers()


@pytest.mark.parametrize("world,shape,n_iter,levels,warm", [(2, (21, 50), 6, 12, 0), (3, (20, 70), 7, 32, 0),
                                                           (3, (33, 64), 9, 20, 3), (4, (40, 90), 8, 64, 2)])
def test_persistent_kernels_exchange_across_bands(hsm, world, shape, n_iter, levels, warm):
    """Row bands with the persistent dataflow kernel: every band engine queues ALL its iterations as one
    launch; boundary rows and per-strip progress counters travel between the kernels through peer-mapped
    memory.  Several engines on ONE GPU here (the shapes are small enough for all kernels to be resident at
    once, which this single-GPU arrangement needs; across GPUs there is no such requirement).  `warm` > 0:
    the first iterations go through the one-launch-per-iteration path with stream-ordered flags, then the
    persistent kernel takes over -- the two flag schemes have to interoperate."""
    import torch
    from importlib import import_module
    bands = import_module("hybrid-stereo-matching_b200.bands")
    synth = import_module("hybrid-stereo-matching_b200.synth")
    rows, cols = shape
    left, right, disparity = synth.textured_pair(rows, cols, levels, 23, block=(5, 6))
    prior = synth.sparse_prior(disparity, levels, 23, density=0.2)
    want = mrf_c.OracleMRFC((rows, cols), levels)
    want_labels = want.lbp(left, right, prior, prior_trust_factor=1.0, prior_influence_mode="adaptive", n_iter=n_iter)

    capi = import_module("hybrid-stereo-matching_b200._capi")
    engines = [bands.CudaBandEngine(rows, cols, levels, *bands.band_bounds(rows, world, r)) for r in range(world)]
    for e in engines:
        e.matcher.set_option(capi.OPT_ALGORITHM, 6)
        with e.stream_context():
            e.init_fields(left, right, prior, 1.0, "adaptive")
    infos = [e.peer_info() for e in engines]
    for r, e in enumerate(engines):
        if r > 0:
            e.attach_peer(0, infos[r - 1], same_process=True)
        if r < world - 1:
            e.attach_peer(1, infos[r + 1], same_process=True)
        e.reset_peers()
    torch.cuda.synchronize()
    for it in range(warm):
        for e in engines:
            with e.stream_context():
                e.iterate(finalize=False)
    for e in engines:
        with e.stream_context():
            e.iterate(finalize=True, n_iter=n_iter - warm)
    torch.cuda.synchronize()
    got = np.concatenate([e.labels() for e in engines], axis=0)
    assert np.array_equal(got, want_labels)
    want_planes = want._message_field
    for r, e in enumerate(engines):
        first, last_row = bands.band_bounds(rows, world, r)
        for name, plane in e.planes().items():
            assert np.array_equal(plane, want_planes[name][:, first:last_row + 1, :]), (r, name)
        e.detach_peers()
