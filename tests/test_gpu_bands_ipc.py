"""Row bands across PROCESSES: two (three) ranks, each with its own CUDA context on cuda:0, neighbours attached
through CUDA IPC mappings, boundary rows pushed into each other's memory as bulk tensor stores, iterations ordered by
stream-ordered flags -- the same code path as one rank per GPU (bench.py --gpus N measures that one), minus NVLink.
Each rank calls ``lbp`` TWICE on the same ``RowBandMatcher`` and one rank is deliberately late for the second
call: a push or flag of the first call must not leak into the second (the matcher drains and meets the other ranks
before it re-initialises its planes and mailboxes).  Checked against the C oracle."""
import os
import sys
import time

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _worker(rank, world, port, shape, n_iter, exchange, queue):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from importlib import import_module
    bands = import_module("hybrid-stereo-matching_b200.bands")
    synth = import_module("hybrid-stereo-matching_b200.synth")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.cuda.set_device(0)
        rows, cols, levels = shape
        first, last = bands.band_bounds(rows, world, rank)
        engine = bands.CudaBandEngine(rows, cols, levels, first, last, device=0)
        matcher = bands.RowBandMatcher(engine, rank, world, exchange=exchange)
        results = []
        for call, seed in enumerate((5, 6)):
            left, right, disparity = synth.textured_pair(rows, cols, levels, seed, block=(5, 6))
            prior = synth.sparse_prior(disparity, levels, seed, density=0.2)
            if call == 1 and rank == world - 1:
                time.sleep(0.5)                      # the others are already waiting for this rank
            local = matcher.lbp(left, right, prior, prior_trust_factor=1.0, prior_influence_mode="adaptive",
                                n_iter=n_iter + call)
            west = engine.planes()["west"]
            results.append((np.asarray(local).copy(), np.ascontiguousarray(west)))
        if exchange == "p2p":
            engine.matcher.synchronize()
            dist.barrier()
            engine.detach_peers()
        queue.put((rank, first, last, results))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,shape,n_iter,exchange", [(2, (22, 50, 12), 5, "p2p"), (3, (21, 70, 20), 4, "p2p"),
                                                         (2, (16, 40, 32), 3, "p2p")])
def test_row_bands_across_processes(world, shape, n_iter, exchange):
    import torch.multiprocessing as mp
    from importlib import import_module
    sys.path.insert(0, ROOT)
    synth = import_module("hybrid-stereo-matching_b200.synth")
    from oracle import mrf_c
    rows, cols, levels = shape
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    port = 29700 + (os.getpid() % 1500)
    procs = [ctx.Process(target=_worker, args=(r, world, port, shape, n_iter, exchange, queue)) for r in range(world)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(world):
        rank, first, last, results = queue.get(timeout=180)
        got[rank] = (first, last, results)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for call, seed in enumerate((5, 6)):
        left, right, disparity = synth.textured_pair(rows, cols, levels, seed, block=(5, 6))
        prior = synth.sparse_prior(disparity, levels, seed, density=0.2)
        want = mrf_c.OracleMRFC((rows, cols), levels)
        want_labels = want.lbp(left, right, prior, prior_trust_factor=1.0, prior_influence_mode="adaptive",
                               n_iter=n_iter + call)
        want_west = want._message_field["west"]
        for rank in range(world):
            first, last, results = got[rank]
            labels, west = results[call]
            assert np.array_equal(labels, want_labels[first:last + 1]), (call, rank)
            assert np.array_equal(west, want_west[:, first:last + 1, :]), (call, rank)
