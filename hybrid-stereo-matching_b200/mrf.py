"""``StereoMRF`` -- drop-in for ``stereovis.framed.algorithms.mrf.StereoMRF``.

Same constructor, methods, defaults and return types as the reference class
(reference mrf.py:8-167); the numpy arithmetic is replaced by the sm_100a CUDA
kernels behind the C ABI of ``include/hsm.h``.  There is no CPU fallback.

Differences that are deliberate and invisible to the reference's callers:
  * state lives on the GPU; ``_message_field`` / ``_belief_field`` are read-back
    properties that return the reference's ``(D, H, W)`` float32 layout;
  * no CUDA call happens before the first compute call, because the reference's
    online mode constructs the matcher in the parent and calls ``lbp`` in a
    ``fork()``ed child (hybrid_stereo.py:107-111 -> online_processing.py:148);
  * extra, optional surface: ``capacity`` (frames per launch) and ``lbp_batch``.
"""
import ctypes
import logging
import time
import weakref

import numpy as np

from . import _capi

logger = logging.getLogger(__name__)

_IMAGE_CODES = {np.dtype(np.float32): _capi.F32, np.dtype(np.float64): _capi.F64}
_PRIOR_CODES = {np.dtype(np.float32): _capi.F32, np.dtype(np.float64): _capi.F64,
                np.dtype(np.int32): _capi.I32, np.dtype(np.int64): _capi.I64}


def _as_image(array):
    """C-contiguous float32/float64 view or copy; other dtypes are cast like
    ``astype('float32')`` (mrf.py:41-42).  Returns (array, dtype code)."""
    array = np.asarray(array)
    if array.dtype not in _IMAGE_CODES:
        array = array.astype("float32")
    array = np.ascontiguousarray(array)
    return array, _IMAGE_CODES[array.dtype]


def _as_prior(array):
    array = np.asarray(array)
    if array.dtype not in _PRIOR_CODES:
        array = array.astype(np.float64 if array.dtype.kind == "f" else np.int64)
    array = np.ascontiguousarray(array)
    return array, _PRIOR_CODES[array.dtype]


def pinned_empty(shape, dtype):
    """numpy array backed by page-locked host memory (cudaHostAlloc through the C ABI), for
    asynchronous transfers in ``lbp_batch``.  Freed when the array is garbage collected."""
    dtype = np.dtype(dtype)
    count = int(np.prod(shape))
    pointer = ctypes.c_void_p()
    nbytes = max(count * dtype.itemsize, 1)
    _capi.check(_capi.lib().hsm_host_alloc(nbytes, ctypes.byref(pointer)))
    address = pointer.value
    buffer = (ctypes.c_char * nbytes).from_address(address)
    # every numpy view keeps `buffer` alive; release the pinned block when the last one goes
    weakref.finalize(buffer, lambda: _capi.lib().hsm_host_free(ctypes.c_void_p(address)))
    return np.frombuffer(buffer, dtype=dtype, count=count).reshape(shape)


def is_pinned(array):
    """True when the array's memory is page-locked and known to CUDA (``pinned_empty``, a pinned torch tensor,
    ``cudaHostRegister``): asynchronous copies from pageable memory block the calling thread instead."""
    flag = ctypes.c_int(0)
    _capi.check(_capi.lib().hsm_host_is_pinned(array.ctypes.data_as(ctypes.c_void_p), ctypes.byref(flag)))
    return bool(flag.value)


_CAI_CODES = {"<f4": _capi.F32, "<f8": _capi.F64, "<i4": _capi.I32, "<i8": _capi.I64}


def _device_array(obj, what):
    """(pointer, dtype code, shape) of a CUDA array (anything with __cuda_array_interface__,
    e.g. a torch tensor -- used as a container only)."""
    cai = obj.__cuda_array_interface__
    if cai.get("strides") is not None:
        raise ValueError("%s must be C-contiguous" % what)
    if cai["typestr"] not in _CAI_CODES:
        raise ValueError("%s has unsupported dtype %s" % (what, cai["typestr"]))
    return cai["data"][0], _CAI_CODES[cai["typestr"]], tuple(cai["shape"])


def _blend(prior, prior_trust_factor, prior_influence_mode):
    """(mode, keep_f32, keep_f64) for ``(1 - prior_trust_factor) * data_diff`` (mrf.py:58-66).

    A Python float factor is a weak scalar: the product is float32, i.e.
    ``float32(1 - tau) * diff``.  A numpy float64 scalar makes numpy compute the
    product in float64 before the store into the float32 plane.
    """
    if prior is None:
        return _capi.PRIOR_NONE, 0.0, 0.0
    if prior_influence_mode == "adaptive":
        return _capi.PRIOR_ADAPTIVE, 0.0, 0.0
    keep = 1 - prior_trust_factor
    if isinstance(keep, np.floating) and keep.dtype == np.float64:
        return _capi.PRIOR_CONST_F64, 0.0, float(keep)
    return _capi.PRIOR_CONST, float(np.float32(keep)), 0.0


class StereoMRF(object):
    """Markov random field stereo matcher with the reference's message-passing schedule."""

    def __init__(self, dim, n_levels, capacity=1, device=0):
        dim = tuple(int(d) for d in dim)
        self.n_levels = n_levels
        self.dimension = (n_levels,) + dim                              # mrf.py:13-14
        self.capacity = int(capacity)
        self.device = int(device)
        self._rows, self._cols = dim
        self._handle = ctypes.c_void_p()
        _capi.check(_capi.lib().hsm_create(self._rows, self._cols, int(n_levels), self.capacity,
                                           self.device, ctypes.byref(self._handle)))
        self._frames = 0
        self._batched = False

    def __del__(self):
        handle = getattr(self, "_handle", None)
        if handle is not None and handle.value:
            try:
                _capi.lib().hsm_destroy(handle)
            except Exception:
                pass
            self._handle = None

    # ------------------------------------------------------------------ options
    def set_option(self, key, value):
        _capi.check(_capi.lib().hsm_set_option(self._handle, int(key), int(value)))

    def get_option(self, key):
        out = ctypes.c_int()
        _capi.check(_capi.lib().hsm_get_option(self._handle, int(key), ctypes.byref(out)))
        return out.value

    def set_stream(self, cuda_stream):
        _capi.check(_capi.lib().hsm_set_stream(self._handle, ctypes.c_void_p(int(cuda_stream))))

    def stats(self):
        out = _capi.Stats()
        _capi.check(_capi.lib().hsm_get_stats(self._handle, ctypes.byref(out)))
        return {name: getattr(out, name) for name, _ in _capi.Stats._fields_}

    def reset_stats(self):
        _capi.check(_capi.lib().hsm_reset_stats(self._handle))

    def synchronize(self):
        _capi.check(_capi.lib().hsm_sync(self._handle))

    # ------------------------------------------------------------- reference API
    def _check_pair(self, image_left, image_right, prior, batched):
        image_left, image_right = np.asarray(image_left), np.asarray(image_right)
        assert image_left.shape == image_right.shape                     # mrf.py:40
        frame_shape = image_left.shape[1:] if batched else image_left.shape
        if frame_shape != (self._rows, self._cols):
            # the reference fails with a numpy broadcasting ValueError here (mrf.py:70)
            raise ValueError("could not broadcast input array from shape %r into shape %r"
                             % (frame_shape, (self._rows, self._cols)))
        if prior is not None:
            prior = np.asarray(prior)
            assert image_right.shape == prior.shape                      # mrf.py:53
        return image_left, image_right, prior

    def _host_inputs(self, image_left, image_right, prior, prior_trust_factor, prior_influence_mode, batched):
        """The C ABI's view of the inputs: (n_frames, keep-alive arrays, argument tuple up to the blend)."""
        image_left, image_right, prior = self._check_pair(image_left, image_right, prior, batched)
        n_frames = image_left.shape[0] if batched else 1
        left, code = _as_image(image_left)
        right, code_r = _as_image(image_right)
        if code_r != code:
            left, right = left.astype("float32"), right.astype("float32")
            code = _capi.F32
        mode, keep32, keep64 = _blend(prior, prior_trust_factor, prior_influence_mode)
        if mode == _capi.PRIOR_ADAPTIVE and self.n_levels > self._cols:
            # level l = cols slices an empty column range and the reference's data_diff.max() raises (mrf.py:59)
            raise ValueError("zero-size array to reduction operation maximum which has no identity")
        prior_ptr, prior_code = None, _capi.F32
        if prior is not None:
            prior, prior_code = _as_prior(prior)
            prior_ptr = prior.ctypes.data_as(ctypes.c_void_p)
        args = (self._handle, n_frames, left.ctypes.data_as(ctypes.c_void_p), right.ctypes.data_as(ctypes.c_void_p),
                code, prior_ptr, prior_code, mode, keep32, keep64)
        return n_frames, (left, right, prior), args

    def _init_fields(self, image_left, image_right, prior=None, prior_trust_factor=1.0,
                     prior_influence_mode="adaptive", reinit_messages=True, _batched=False):
        """Cost volume + optional prior blend + optional message reset (mrf.py:21-71)."""
        n_frames, keep_alive, args = self._host_inputs(image_left, image_right, prior, prior_trust_factor,
                                                       prior_influence_mode, _batched)
        _capi.check(_capi.lib().hsm_init_fields(*args, 1 if reinit_messages else 0, 0))
        self._remember_images(keep_alive[0], keep_alive[1])
        del keep_alive
        self._frames, self._batched = n_frames, _batched

    def _remember_images(self, image_left, image_right):
        """``reference_image`` / ``secondary_image`` of mrf.py:41-42 (float32 casts of the last pair), built when
        somebody reads them."""
        self._last_pair = (image_left, image_right)

    @property
    def reference_image(self):
        return np.asarray(self._last_pair[0]).astype("float32")

    @property
    def secondary_image(self):
        return np.asarray(self._last_pair[1]).astype("float32")

    def _update_message_fields(self):
        """One iteration: south, west, north, east sweeps (mrf.py:73-94)."""
        _capi.check(_capi.lib().hsm_iterate(self._handle, 1, 1))

    def loop_belief(self, image_left, image_right, prior=None, prior_trust_factor=1.0,
                    prior_influence_mode="const", n_iter=10, reinit_messages=True):
        """Initialise the fields and run ``n_iter`` iterations (mrf.py:106-134)."""
        self._init_fields(image_left, image_right, prior, prior_trust_factor,
                          prior_influence_mode, reinit_messages=reinit_messages)
        start = time.time()
        _capi.check(_capi.lib().hsm_iterate(self._handle, int(n_iter), 1))
        if logger.isEnabledFor(logging.DEBUG):                           # mrf.py:134 (needs the device to finish)
            self.synchronize()
            logger.debug("Mean iteration time: %ss", (time.time() - start) / max(int(n_iter), 1))

    def _readout(self):
        if self._frames == 0:
            # never initialised: all planes are zero, argmin of zeros is 0 (mrf.py:103-104)
            return np.zeros((self._rows, self._cols), dtype=np.int64)
        labels = np.empty((self._frames, self._rows, self._cols), dtype=np.int64)
        _capi.check(_capi.lib().hsm_readout(self._handle, labels.ctypes.data_as(ctypes.c_void_p), 0, 0))
        return labels if self._batched else labels[0]

    def _update_belief_field(self):
        self._belief_field_cache = self._readout()

    def get_map_belief(self):
        """MAP disparity map, int64 ``(H, W)`` like ``np.argmin`` (mrf.py:136-144)."""
        self._update_belief_field()
        return self._belief_field_cache

    def lbp(self, image_left, image_right, prior=None, prior_trust_factor=0.5,
            prior_influence_mode="const", n_iter=10):
        """``loop_belief`` (always re-initialising) + ``get_map_belief`` (mrf.py:146-167), as ONE native call:
        upload, cost volume, ``n_iter`` iterations and readout are queued back to back and the host waits
        once, for the labels."""
        n_frames, keep_alive, args = self._host_inputs(image_left, image_right, prior, prior_trust_factor,
                                                       prior_influence_mode, False)
        labels = np.empty((1, self._rows, self._cols), dtype=np.int64)
        start = time.time()
        _capi.check(_capi.lib().hsm_lbp(*args, int(n_iter), labels.ctypes.data_as(ctypes.c_void_p), 0, 0))
        logger.debug("Mean iteration time: %ss", (time.time() - start) / max(int(n_iter), 1))      # mrf.py:134
        self._remember_images(keep_alive[0], keep_alive[1])
        del keep_alive
        self._frames, self._batched = n_frames, False
        self._belief_field_cache = labels[0]
        return self._belief_field_cache

    # ----------------------------------------------------------------- extensions
    def _lanes(self, count):
        """This matcher plus helper matchers of the same geometry (own stream, own device
        buffers), so that one lane's host<->device copies overlap another lane's kernels."""
        lanes = getattr(self, "_helper_lanes", None)
        if lanes is None:
            lanes = self._helper_lanes = []
        while len(lanes) < count - 1:
            helper = StereoMRF((self._rows, self._cols), self.n_levels, self.capacity, self.device)
            for key in (_capi.OPT_LANE_CONFIG, _capi.OPT_ALGORITHM, _capi.OPT_BAND_ROWS, _capi.OPT_DATA_TERM,
                        _capi.OPT_STAGES, _capi.OPT_FRAME_GROUPS):
                helper.set_option(key, self.get_option(key))
            lanes.append(helper)
        return [self] + lanes[:count - 1]

    def _staging(self, name, shape, dtype):
        """A page-locked staging array of this lane, grown on demand and reused across calls."""
        ring = self.__dict__.setdefault("_stage_ring", {})
        need = int(np.prod(shape)) * np.dtype(dtype).itemsize
        block = ring.get(name)
        if block is None or block.nbytes < need:
            block = ring[name] = pinned_empty((max(need, 1),), np.uint8)
        return block[:need].view(dtype).reshape(shape)

    def lbp_batch(self, images_left, images_right, priors=None, prior_trust_factor=0.5,
                  prior_influence_mode="const", n_iter=10, out=None, lanes=2):
        """``lbp`` over a stack of independent frame pairs ``(N, H, W)`` -> ``(N, H, W)`` int64.

        Frames are processed ``capacity`` at a time, alternating between ``lanes`` matchers (own stream, own
        device buffers) so that transfers overlap compute.  Page-locked arrays (``pinned_empty``) are handed
        to the copy engine as they are.  Ordinary (pageable) numpy arrays -- what the reference's callers
        pass (stereo_framebased.py:83-85) -- go through a ring of page-locked staging buffers owned by the
        lanes: a small thread pool fills the next chunk's staging (float64 images are cast to float32 on the
        way, exactly ``astype('float32')`` of mrf.py:41-42) while the GPU works on the previous chunk, and
        labels come back through staging the same way, so the caller never blocks inside a copy.  Each
        frame's result equals what ``lbp`` returns for it alone (every call re-initialises the messages,
        mrf.py:165).
        """
        images_left, images_right = np.asarray(images_left), np.asarray(images_right)
        device_priors = priors is not None and hasattr(priors, "device_pointer")
        if priors is not None and not device_priors:
            priors = np.asarray(priors)
        assert images_left.ndim == 3 and images_left.shape == images_right.shape
        n_total = images_left.shape[0]
        if out is None:
            out = np.empty((n_total, self._rows, self._cols), dtype=np.int64)
        assert out.shape == (n_total, self._rows, self._cols) and out.dtype == np.int64 \
            and out.flags["C_CONTIGUOUS"]
        if n_total == 0:
            return out
        self._check_pair(images_left[:1], images_right[:1], None, True)
        if priors is not None:
            assert tuple(priors.shape) == images_right.shape                      # mrf.py:53
        # Chunk size: the capacity, but at least two chunks per lane when there is more than one capacity's
        # worth of frames -- the first chunk's upload and the last chunk's download are the only transfers
        # that cannot overlap compute, so they should be a small part of the call.
        lanes = max(1, int(lanes))
        chunk = self.capacity
        if n_total > self.capacity and lanes > 1:
            chunk = min(self.capacity, max(-(-n_total // (2 * lanes)), 8))
        chunks = [(start, min(start + chunk, n_total)) for start in range(0, n_total, chunk)]
        workers = self._lanes(max(1, min(lanes, len(chunks))))
        mode, keep32, keep64 = _blend(priors, prior_trust_factor, prior_influence_mode)
        lib = _capi.lib()

        def host_input(lane, name, array, image):
            """(pointer, dtype code, keep-alive) of one chunk input as the C ABI should see it."""
            if image and array.dtype not in _IMAGE_CODES:
                array = array.astype("float32")
            if not image and array.dtype not in _PRIOR_CODES:
                array = array.astype(np.float64 if array.dtype.kind == "f" else np.int64)
            if array.flags["C_CONTIGUOUS"] and is_pinned(array):
                return array, None
            target_dtype = np.float32 if image else array.dtype
            staged = lane._staging(name, array.shape, target_dtype)
            return staged, array                       # (to be filled from `array` by the copy pool)

        pool = self.__dict__.get("_copy_pool")
        if pool is None:
            from concurrent.futures import ThreadPoolExecutor
            pool = self._copy_pool = ThreadPoolExecutor(max_workers=4)

        def fill(pairs):
            jobs = []
            for staged, source in pairs:
                if source is None:
                    continue
                # split big copies so that the pool's threads share them (numpy releases the GIL in copyto)
                parts = max(1, min(4, staged.shape[0]))
                bounds = np.linspace(0, staged.shape[0], parts + 1).astype(int)
                for a, b in zip(bounds[:-1], bounds[1:]):
                    if b > a:
                        jobs.append(pool.submit(np.copyto, staged[a:b], source[a:b], "same_kind"))
            for job in jobs:
                job.result()

        out_pinned = is_pinned(out)
        pending = {}                                   # lane index -> (staged labels, destination slice)

        def drain(index):
            entry = pending.pop(index, None)
            if entry is not None:
                fill([(entry[1], entry[0])])

        for index, (start, stop) in enumerate(chunks):
            which = index % len(workers)
            lane = workers[which]
            lane.synchronize()                         # the lane's previous chunk (and its staging) is done
            drain(which)
            left, left_src = host_input(lane, "left", images_left[start:stop], True)
            right, right_src = host_input(lane, "right", images_right[start:stop], True)
            pairs = [(left, left_src), (right, right_src)]
            if left.dtype != right.dtype:              # mixed image dtypes: both through float32 staging
                left, right = left.astype("float32"), right.astype("float32")
                pairs = []
            prior_ptr, prior_code, on_device = None, _capi.F32, 0
            prior = None
            if device_priors:
                prior_ptr, prior_code, on_device = ctypes.c_void_p(priors.device_pointer(start)), _capi.I32, 2
            elif priors is not None:
                prior, prior_src = host_input(lane, "prior", priors[start:stop], False)
                pairs.append((prior, prior_src))
                prior_ptr, prior_code = prior.ctypes.data_as(ctypes.c_void_p), _PRIOR_CODES[prior.dtype]
            fill(pairs)
            target = out[start:stop] if out_pinned else lane._staging("labels", (stop - start, self._rows, self._cols),
                                                                      np.int64)
            _capi.check(lib.hsm_lbp(lane._handle, stop - start, left.ctypes.data_as(ctypes.c_void_p),
                                    right.ctypes.data_as(ctypes.c_void_p), _IMAGE_CODES[left.dtype], prior_ptr,
                                    prior_code, mode, keep32, keep64, int(n_iter),
                                    target.ctypes.data_as(ctypes.c_void_p), on_device, 2))
            lane._frames, lane._batched = stop - start, True
            if not out_pinned:
                pending[which] = (target, out[start:stop])
        for which, lane in enumerate(workers):
            lane.synchronize()
            drain(which)
        self._belief_field_cache = out[chunks[-1][0]:chunks[-1][1]]
        return out

    def lbp_batch_device(self, images_left, images_right, priors, out, prior_trust_factor=0.5,
                         prior_influence_mode="const", n_iter=10, synchronize=True):
        """``lbp`` for ``N <= capacity`` frame pairs that already live on this GPU.

        The arguments are CUDA arrays (``__cuda_array_interface__``: torch tensors, cupy
        arrays) used purely as containers: images float32/float64 ``(N, H, W)``, priors
        None or float/int ``(N, H, W)``, ``out`` int64 ``(N, H, W)``.  Work is queued on the
        matcher's stream (see ``set_stream``).
        """
        left_ptr, code, shape = _device_array(images_left, "images_left")
        right_ptr, code_r, shape_r = _device_array(images_right, "images_right")
        assert shape == shape_r and len(shape) == 3
        if shape[1:] != (self._rows, self._cols):
            raise ValueError("could not broadcast input array from shape %r into shape %r"
                             % (shape[1:], (self._rows, self._cols)))
        if code != code_r or code not in (_capi.F32, _capi.F64):
            raise ValueError("device images must both be float32 or both be float64")
        out_ptr, out_code, out_shape = _device_array(out, "out")
        assert out_code == _capi.I64 and out_shape == shape
        mode, keep32, keep64 = _blend(priors, prior_trust_factor, prior_influence_mode)
        prior_ptr, prior_code = None, _capi.F32
        if priors is not None:
            prior_ptr, prior_code, prior_shape = _device_array(priors, "priors")
            assert prior_shape == shape
        _capi.check(_capi.lib().hsm_lbp(
            self._handle, shape[0], ctypes.c_void_p(left_ptr), ctypes.c_void_p(right_ptr), code,
            ctypes.c_void_p(prior_ptr) if prior_ptr else None, prior_code, mode, keep32, keep64,
            int(n_iter), ctypes.c_void_p(out_ptr), 1, 0 if synchronize else 1))
        self._frames, self._batched = shape[0], True
        return out

    # ------------------------------------------------------- read-back accessors
    def _plane(self, which, frame=0):
        out = np.empty((self._rows, self._cols, self.n_levels), dtype=np.float32)
        _capi.check(_capi.lib().hsm_get_plane(self._handle, int(frame), int(which),
                                              out.ctypes.data_as(ctypes.c_void_p)))
        return np.ascontiguousarray(out.transpose(2, 0, 1))

    def message_field(self, frame=0):
        """The reference's ``_message_field`` dict for one frame: name -> (D, H, W) float32."""
        return {name: self._plane(index, frame) for index, name in enumerate(_capi.PLANES)}

    @property
    def _message_field(self):
        return self.message_field(0)

    @property
    def _belief_field(self):
        return self._belief_field_cache

    def energy_field(self, frame=0):
        """Energy of mrf.py:103 for one frame, (D, H, W) float32."""
        out = np.empty((self._rows, self._cols, self.n_levels), dtype=np.float32)
        _capi.check(_capi.lib().hsm_get_energy(self._handle, int(frame), out.ctypes.data_as(ctypes.c_void_p)))
        return np.ascontiguousarray(out.transpose(2, 0, 1))
