"""Row-band run of one large frame over the GPUs of a node (torchrun), checked against the
single-GPU result.  python -m torch.distributed.run --nproc-per-node N tools/band_demo.py [rows cols levels iters]"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import hybrid_stereo_matching_b200 as hsm
from importlib import import_module
bands = import_module("hybrid-stereo-matching_b200.bands")
synth = import_module("hybrid-stereo-matching_b200.synth")

rows, cols, levels, n_iter = [int(v) for v in (sys.argv[1:5] + [1080, 1920, 256, 10][len(sys.argv) - 1:])]
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
left, right, disparity = synth.textured_pair(rows, cols, levels, 4321)
prior = synth.sparse_prior(disparity, levels, 4321)
first, last = bands.band_bounds(rows, world, rank)
engine = bands.CudaBandEngine(rows, cols, levels, first, last, device=local)
exchange = os.environ.get("HSM_EXCHANGE", "nccl")
matcher = bands.RowBandMatcher(engine, rank, world, exchange=exchange)
kw = dict(prior_trust_factor=1.0, prior_influence_mode="adaptive", n_iter=n_iter)
matcher.lbp(left, right, prior, **kw)                                  # warm-up
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
local_labels = matcher.lbp(left, right, prior, **kw)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
elapsed = time.perf_counter() - t0
# iteration loop only (no init / readout): time n_iter iterations with their halo exchanges
engine.init_fields(left, right, prior, 1.0, "adaptive")
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t2 = time.perf_counter()
if exchange == "p2p" and world > 1:
    engine.reset_peers()
    dist.barrier()
    t2 = time.perf_counter()
with engine.stream_context():
    if exchange == "p2p" or world == 1:
        engine.iterate(finalize=True, n_iter=n_iter)
    else:
        for it in range(n_iter):
            engine.iterate(finalize=(it == n_iter - 1))
            if it < n_iter - 1:
                matcher._exchange()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
loop_s = time.perf_counter() - t2
# the same without any exchange (wrong results, timing only): what the exchange costs
if exchange == "p2p" and world > 1:
    engine.detach_peers(); matcher._attached = False
with engine.stream_context():
    # fresh messages (continuing a converged run would time the rare exact-division path, not the exchange),
    # and an untimed first use of the kernels without the push
    engine.init_fields(left, right, prior, 1.0, "adaptive")
    engine.iterate(finalize=False, n_iter=2)
    engine.init_fields(left, right, prior, 1.0, "adaptive")
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t3 = time.perf_counter()
with engine.stream_context():
    engine.iterate(finalize=True, n_iter=n_iter)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
noex_s = time.perf_counter() - t3
full = matcher.gather(local_labels)
if rank == 0:
    ref = hsm.StereoMRF((rows, cols), levels, device=local)
    t1 = time.perf_counter()
    want = ref.lbp(left, right, prior, **kw)
    single = time.perf_counter() - t1
    print(json.dumps({"exchange": exchange, "shape": [rows, cols, levels], "n_iter": n_iter, "world": world, "band_s": round(elapsed, 4), "iter_ms_with_exchange": round(1e3 * loop_s / n_iter, 4),
                      "iter_ms_without_exchange": round(1e3 * noex_s / n_iter, 4),
                      "single_gpu_s_incl_init": round(single, 4), "identical_to_single_gpu": bool(np.array_equal(full, want))}))
if world > 1:
    dist.destroy_process_group()
