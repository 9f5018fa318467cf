"""Multi-rank host logic on CPU (gloo, world_size 2 and 3): the row-band protocol of
hybrid-stereo-matching_b200/bands.py driven with a numpy engine must reproduce the un-banded
oracle bit for bit; frame sharding must cover every frame exactly once."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _band_worker(rank, world, port, shape, n_iter, with_prior, queue):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from importlib import import_module
    bands = import_module("hybrid-stereo-matching_b200.bands")
    synth = import_module("hybrid-stereo-matching_b200.synth")
    from band_engine_numpy import NumpyBandEngine
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rows, cols, levels = shape
        left, right, disparity = synth.textured_pair(rows, cols, levels, 5, block=(5, 6))
        prior = synth.sparse_prior(disparity, levels, 5, density=0.2) if with_prior else None
        first, last = bands.band_bounds(rows, world, rank)
        engine = NumpyBandEngine(rows, cols, levels, first, last)
        matcher = bands.RowBandMatcher(engine, rank, world)
        local = matcher.lbp(left, right, prior, prior_trust_factor=1.0, prior_influence_mode="adaptive",
                            n_iter=n_iter)
        full = matcher.gather(local)
        planes = {k: np.ascontiguousarray(v) for k, v in engine.planes().items()}
        queue.put((rank, first, last, full, planes))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,shape,n_iter,with_prior", [
    (2, (13, 17, 6), 4, False),
    (2, (12, 20, 9), 3, True),
    (3, (11, 15, 5), 5, True),
    (2, (2, 9, 3), 3, False),          # one row per rank
])
def test_row_band_protocol_matches_unbanded_oracle(world, shape, n_iter, with_prior):
    from importlib import import_module
    sys.path.insert(0, ROOT)
    synth = import_module("hybrid-stereo-matching_b200.synth")
    from oracle import mrf_numpy
    rows, cols, levels = shape
    left, right, disparity = synth.textured_pair(rows, cols, levels, 5, block=(5, 6))
    prior = synth.sparse_prior(disparity, levels, 5, density=0.2) if with_prior else None
    want = mrf_numpy.OracleMRF((rows, cols), levels)
    want_labels = want.lbp(left, right, prior, prior_trust_factor=1.0, prior_influence_mode="adaptive", n_iter=n_iter)

    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_band_worker, args=(r, world, port, shape, n_iter, with_prior, queue))
             for r in range(world)]
    for p in procs:
        p.start()
    results = [queue.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    covered = []
    for rank, first, last, full, planes in sorted(results, key=lambda r: r[0]):
        assert np.array_equal(full, want_labels), "rank %d assembled a different map" % rank
        for name, plane in planes.items():
            assert np.array_equal(plane, want._message_field[name][:, first:last + 1, :]), (rank, name)
        covered.extend(range(first, last + 1))
    assert covered == list(range(rows))


def test_band_bounds_and_frame_sharding():
    from importlib import import_module
    sys.path.insert(0, ROOT)
    bands = import_module("hybrid-stereo-matching_b200.bands")
    for rows in (1, 2, 7, 180, 1080):
        for world in (1, 2, 3, 4, 8):
            if world > rows:
                with pytest.raises(ValueError):
                    bands.band_bounds(rows, world, 0)
                continue
            spans = [bands.band_bounds(rows, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == rows - 1
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert b0 == a1 + 1
            sizes = [b - a + 1 for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    for n in (0, 1, 10, 10000):
        for world in (1, 2, 4, 8):
            got = sorted(i for r in range(world) for i in bands.frames_for_rank(n, world, r))
            assert got == list(range(n))
