"""Row-band partitioning of ONE large frame across GPUs (BASELINE.json configs[4]).

The reference has no multi-device path; this is the spatial analogue the north star asks
for.  Rank ``r`` of ``G`` owns a contiguous band of image rows and
keeps one ghost row above and below it.  Because a fused iteration needs only the OLD
west / north / east mailboxes of the 3x3 neighbourhood, one exchange
per iteration is enough: after iteration ``k`` every rank sends its first and last owned
rows of the three planes to its vertical neighbours, which store them in their ghost rows
before iteration ``k + 1``.  The result is bit-identical to the single-GPU run.

The protocol (who owns which rows, what is exchanged when) is independent of the engine
that does the arithmetic, so the CPU test-suite drives it with a numpy engine over gloo and
the GPU path drives the CUDA matcher over NCCL (``torch.distributed`` point-to-point; the
halo is 3 x W x Dp floats each way, e.g. 5.9 MB at 1920 x D256 -- latency-, not
bandwidth-bound on NVLink).
"""
import ctypes

import numpy as np

from . import _capi


def band_bounds(n_rows, world, rank):
    """First and last image row (inclusive) owned by ``rank``: contiguous, sizes differ by <= 1."""
    base, extra = divmod(int(n_rows), int(world))
    if base == 0:
        raise ValueError("more ranks (%d) than image rows (%d)" % (world, n_rows))
    first = rank * base + min(rank, extra)
    last = first + base - 1 + (1 if rank < extra else 0)
    return first, last


def frames_for_rank(n_frames, world, rank):
    """Indices of the frames of a sequence processed by ``rank`` (frame sharding, no collective)."""
    return list(range(int(rank), int(n_frames), int(world)))


class CudaBandEngine(object):
    """The CUDA matcher as a band engine: local rows = [ghost above] + owned + [ghost below]."""

    def __init__(self, n_rows, n_cols, n_levels, first, last, device=0):
        import torch                                  # container for the halo buffers only
        from .mrf import StereoMRF
        self.torch = torch
        self.first, self.last, self.n_rows = first, last, n_rows
        self.own = last - first + 1
        self.local_rows = self.own + 2
        self.matcher = StereoMRF((self.local_rows, n_cols), n_levels, capacity=1, device=device)
        self.matcher.set_option(_capi.OPT_OUT_LO, 1)
        self.matcher.set_option(_capi.OPT_OUT_HI, self.own)
        self.matcher.set_option(_capi.OPT_VALID_LO, 0 if first > 0 else 1)
        self.matcher.set_option(_capi.OPT_VALID_HI, self.own + 1 if last < n_rows - 1 else self.own)
        # kernels, halo copies and the NCCL point-to-point ops are all ordered on this stream
        self.stream = torch.cuda.Stream(device=device)
        self.matcher.set_stream(self.stream.cuda_stream)
        count = ctypes.c_size_t()
        _capi.check(_capi.lib().hsm_halo_floats(self.matcher._handle, ctypes.byref(count)))
        self.halo_floats = count.value
        self.device = torch.device("cuda", device)

    def stream_context(self):
        return self.torch.cuda.stream(self.stream)

    def _local(self, image):
        """Rows first-1 .. last+1 of a full image, zero rows where that is outside the image."""
        image = np.asarray(image)
        out = np.zeros((self.local_rows,) + image.shape[1:], dtype=image.dtype)
        lo, hi = max(self.first - 1, 0), min(self.last + 1, self.n_rows - 1)
        out[lo - (self.first - 1):hi - (self.first - 1) + 1] = image[lo:hi + 1]
        return out

    def init_fields(self, image_left, image_right, prior, prior_trust_factor, prior_influence_mode):
        local_prior = None
        if prior is not None:
            local_prior = self._local(prior)
            if self.first == 0:
                local_prior[0] = -1
            if self.last == self.n_rows - 1:
                local_prior[-1] = -1
        self.matcher._init_fields(self._local(image_left), self._local(image_right), local_prior,
                                  prior_trust_factor, prior_influence_mode, True)

    def iterate(self, finalize, n_iter=1):
        _capi.check(_capi.lib().hsm_iterate(self.matcher._handle, int(n_iter), 1 if finalize else 0))

    def new_halo(self):
        return self.torch.empty(self.halo_floats, dtype=self.torch.float32, device=self.device)

    def pack(self, top, bottom):
        _capi.check(_capi.lib().hsm_halo_pack(self.matcher._handle,
                                              ctypes.c_void_p(top.data_ptr()) if top is not None else None,
                                              ctypes.c_void_p(bottom.data_ptr()) if bottom is not None else None))

    def unpack(self, from_above, from_below):
        _capi.check(_capi.lib().hsm_halo_unpack(
            self.matcher._handle,
            ctypes.c_void_p(from_above.data_ptr()) if from_above is not None else None,
            ctypes.c_void_p(from_below.data_ptr()) if from_below is not None else None))

    # ---- fused halo exchange: the iteration kernel pushes its boundary rows into the neighbours ----
    def peer_info(self):
        """This band's ``hsm_peer_info`` (CUDA IPC handle + plane offsets) as a ctypes struct."""
        info = _capi.PeerInfo()
        _capi.check(_capi.lib().hsm_peer_export(self.matcher._handle, ctypes.byref(info)))
        return info

    def attach_peer(self, side, info, same_process=False):
        """side 0 = the band above, 1 = the band below."""
        _capi.check(_capi.lib().hsm_peer_attach(self.matcher._handle, int(side), ctypes.byref(info),
                                                1 if same_process else 0))

    def reset_peers(self):
        _capi.check(_capi.lib().hsm_peer_reset(self.matcher._handle))

    def detach_peers(self):
        _capi.check(_capi.lib().hsm_peer_detach(self.matcher._handle))

    def labels(self):
        return self.matcher.get_map_belief()[1:self.own + 1]

    def planes(self):
        """Owned rows of the five planes, reference layout (D, own rows, W)."""
        return {name: plane[:, 1:self.own + 1, :] for name, plane in self.matcher._message_field.items()}


class RowBandMatcher(object):
    """``lbp`` of one frame split into row bands over the ranks of a process group.

    ``engine`` does the arithmetic on this rank's band (``CudaBandEngine`` in production).
    Every rank passes the full images; ``lbp`` returns this rank's rows of the disparity map
    and ``gather`` assembles the full map on every rank.
    """

    def __init__(self, engine, rank, world, group=None, exchange="nccl"):
        """``exchange``: "nccl" = pack / point-to-point send-recv / unpack after every iteration;
        "p2p" = the iteration kernel stores its boundary rows straight into the neighbours' ghost rows
        over NVLink (CUDA IPC mappings, bulk tensor stores) and neighbours are ordered by stream-ordered flags --
        no collective call, no host work and no spinning kernel per iteration; "p2p_dataflow" = the same push,
        but every GPU runs ONE persistent kernel for the whole loop and the boundary tiles exchange progress
        counters through peer memory (fastest: 0.99 / 0.96 / 0.90 of the per-GPU roofline on 2 / 4 / 8 GPUs at
        1080p D256, against 0.94 / 0.90 / 0.86 for "p2p"; its kernels wait for each other across GPUs, so every
        rank must reach the call).  Both ``CudaBandEngine`` only."""
        self.engine, self.rank, self.world, self.group = engine, int(rank), int(world), group
        self._send_up = self._send_down = self._recv_up = self._recv_down = None
        if exchange not in ("nccl", "p2p", "p2p_dataflow"):
            raise ValueError("exchange must be 'nccl', 'p2p' or 'p2p_dataflow'")
        self.exchange = exchange
        self._attached = False

    def _attach_neighbours(self):
        import torch.distributed as dist
        mine = self.engine.peer_info()
        everyone = [None] * self.world
        dist.all_gather_object(everyone, bytes(mine), group=self.group)
        for side, other in ((0, self.rank - 1), (1, self.rank + 1)):
            if 0 <= other < self.world:
                info = _capi.PeerInfo.from_buffer_copy(everyone[other])
                self.engine.attach_peer(side, info, same_process=False)
        self._attached = True

    def _exchange(self):
        import torch.distributed as dist
        eng = self.engine
        up, down = self.rank - 1, self.rank + 1
        has_up, has_down = up >= 0, down < self.world
        if self._send_up is None:
            self._send_up, self._send_down = eng.new_halo(), eng.new_halo()
            self._recv_up, self._recv_down = eng.new_halo(), eng.new_halo()
        eng.pack(self._send_up if has_up else None, self._send_down if has_down else None)
        ops = []
        if has_up:
            ops.append(dist.P2POp(dist.isend, self._send_up, up, self.group))
            ops.append(dist.P2POp(dist.irecv, self._recv_up, up, self.group))
        if has_down:
            ops.append(dist.P2POp(dist.isend, self._send_down, down, self.group))
            ops.append(dist.P2POp(dist.irecv, self._recv_down, down, self.group))
        if ops:
            for request in dist.batch_isend_irecv(ops):
                request.wait()
        eng.unpack(self._recv_up if has_up else None, self._recv_down if has_down else None)

    def lbp(self, image_left, image_right, prior=None, prior_trust_factor=0.5,
            prior_influence_mode="const", n_iter=10, _events=None):
        """``_events``: optional (begin, end) torch CUDA events recorded on the engine's stream around the
        iteration loop including its exchanges (bench.py times the loop with them)."""
        p2p = self.exchange in ("p2p", "p2p_dataflow") and self.world > 1
        if p2p:
            # algorithm 6 = persistent dataflow kernel (falls back to 5 for a single iteration), 5 = one launch
            # per iteration with stream-ordered flags
            self.engine.matcher.set_option(_capi.OPT_ALGORITHM, 6 if self.exchange == "p2p_dataflow" else 5)
        with self.engine.stream_context():
            if p2p:
                # A neighbour's last iteration of the PREVIOUS call may still be pushing rows into this band's
                # ghost rows, and its final flag write may still be in flight: every rank drains its own stream,
                # then all meet, before anybody clears planes or mailboxes for the new call.
                import torch.distributed as dist
                self.engine.matcher.synchronize()
                dist.barrier(group=self.group)
            self.engine.init_fields(image_left, image_right, prior, prior_trust_factor, prior_influence_mode)
            if p2p:
                if not self._attached:
                    self._attach_neighbours()
                self.engine.reset_peers()
                dist.barrier(group=self.group)          # every mailbox is zero before anyone iterates
            if _events is not None:
                _events[0].record(self.engine.stream)
            if p2p or self.world == 1:
                # no host work between iterations: the whole loop is queued by one native call
                if int(n_iter) > 0:
                    self.engine.iterate(finalize=True, n_iter=n_iter)
            else:
                for iteration in range(int(n_iter)):
                    last = iteration == int(n_iter) - 1
                    self.engine.iterate(finalize=last)
                    if not last:
                        self._exchange()
            if _events is not None:
                _events[1].record(self.engine.stream)
            return self.engine.labels()

    def gather(self, local_labels):
        """Full (H, W) int64 disparity map on every rank."""
        import torch
        import torch.distributed as dist
        if self.world == 1:
            return np.asarray(local_labels)
        pieces = [None] * self.world
        dist.all_gather_object(pieces, np.asarray(local_labels), group=self.group)
        return np.concatenate(pieces, axis=0)
