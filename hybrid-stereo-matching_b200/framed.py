"""``FramebasedStereoMatching`` -- the direct caller of the hot path, py3-clean and batched.

Mirrors ``stereovis/framed/stereo_framebased.py`` (reference, lines 15-101): same constructor
arguments, ``run_one_frame`` / ``run`` / ``get_output`` / ``get_timestamps``, same hard-coded call
pattern in ``run`` (trust 1.0, 'adaptive', 10 iterations with priors; ``lbp`` defaults without),
same prior selection by timestamp.  Differences: the per-frame Python loop of ``run`` is one
batched ``lbp_batch`` call (frames are independent, every ``lbp`` re-initialises: mrf.py:165), the
progress bar / debug PNG dumps of the reference (stereo_framebased.py:81,92-97) are dropped.
"""
import logging
import time

import numpy as np

from .mrf import StereoMRF

logger = logging.getLogger(__name__)


def select_priors(prior_info, frame_timestamps):
    """One prior per frame (stereo_framebased.py:71-79): if there are more priors than frames,
    pick for each frame the first prior whose timestamp is >= the frame's (np.searchsorted,
    side='left'); otherwise use them in order."""
    n_frames = len(frame_timestamps)
    priors = prior_info["priors"]
    on_device = hasattr(priors, "device_pointer")          # priors.DevicePriors: rasterised on the GPU, still there
    if len(prior_info["ts"]) > n_frames:
        indices = [np.searchsorted(prior_info["ts"], t_frame, side="left") for t_frame in frame_timestamps]
        return np.asarray(priors.download() if on_device else priors)[indices]
    return priors if on_device else np.asarray(priors)


def next_n_iter(n_iter, last_duration_ms, frame_budget_ms, min_iter=1, max_iter=10, adaptive=True):
    """Iteration budget of the online loop (online_processing.py:122-127): when the previous frame used
    less than 80 % of its time budget grow by 20 % + 1, otherwise halve; clamp to [min_iter, max_iter].
    ``frame_budget_ms`` = nominal frame length x SNN slow-down factor.  With ``adaptive=False`` the
    reference always runs ``max_iter`` iterations."""
    if not adaptive:
        return max_iter
    used = last_duration_ms / frame_budget_ms
    n_iter = int(n_iter + 0.2 * n_iter + 1 if used < 0.8 else n_iter - 0.5 * n_iter)
    return min(max_iter, max(min_iter, n_iter))


def default_capacity(rows, cols, n_levels, n_frames=None, budget_bytes=6 << 30, most=64):
    """Frames per launch for a matcher of this geometry: as many as fit a device-memory budget per lane (the arena
    holds 8 planes per frame: 44 MB at 240x180xD32, 630 MB at 640x480xD64, 17 GB at 1920x1080xD256), at most
    ``most`` and no more than the sequence has.  No CUDA call (the object may be built before a fork)."""
    import ctypes
    from . import _capi
    one = ctypes.c_size_t()
    _capi.check(_capi.lib().hsm_device_bytes(int(rows), int(cols), int(n_levels), 1, ctypes.byref(one)))
    capacity = max(1, min(int(most), int(budget_bytes // max(one.value, 1))))
    if n_frames is not None:
        capacity = max(1, min(capacity, int(n_frames)))
    return capacity


class FramebasedStereoMatching(object):
    def __init__(self, resolution, max_disparity, algorithm="mrf", inputs=None, capacity=None, device=0,
                 matcher_factory=None):
        if algorithm == "mrf":
            # (x, y) resolution -> numpy (rows, cols) = (y, x)            stereo_framebased.py:18-21
            x, y = resolution
            if capacity is None and matcher_factory is None:
                capacity = default_capacity(y, x, max_disparity, None if inputs is None else len(inputs["ts"]))
            factory = matcher_factory or (lambda dim, levels: StereoMRF(dim, levels, capacity=capacity, device=device))
            self.algorithm = factory((y, x), max_disparity)
            if inputs is not None:
                self.frames_left = np.asarray(inputs["left"])
                self.frames_right = np.asarray(inputs["right"])
                self.frames_timestamps = np.asarray(inputs["ts"])
                self.depth_frames = []
        else:
            raise NotImplementedError("Only MRF is supported.")           # stereo_framebased.py:30

    def get_timestamps(self):
        return self.frames_timestamps

    def get_output(self):
        self.depth_frames = np.asarray(self.depth_frames)
        return self.depth_frames

    def run_one_frame(self, image_left, image_right, prior=None, **kwargs):
        """One frame (online mode): ``lbp(image_left, image_right, prior, **kwargs)``, result appended."""
        depth_map = self.algorithm.lbp(image_left, image_right, prior, **kwargs)
        if not hasattr(self, "depth_frames") or not isinstance(self.depth_frames, list):
            self.depth_frames = list(getattr(self, "depth_frames", []))
        self.depth_frames.append(depth_map)
        return depth_map

    def run(self, prior_info=None):
        """All frames (offline mode), batched on the GPU."""
        n_frames = len(self.frames_timestamps)
        start_timer = time.time()
        if prior_info is not None:
            priors = select_priors(prior_info, self.frames_timestamps)
            assert len(priors) == len(self.frames_left) == len(self.frames_right)
            kwargs = dict(prior_trust_factor=1.0, prior_influence_mode="adaptive", n_iter=10)   # :84-85
        else:
            priors, kwargs = None, {}
        if hasattr(self.algorithm, "lbp_batch"):
            maps = self.algorithm.lbp_batch(self.frames_left, self.frames_right, priors, **kwargs)
            self.depth_frames = list(self.depth_frames) + [m for m in maps]
        else:                                                             # any matcher with the reference surface
            for i in range(n_frames):
                self.run_one_frame(self.frames_left[i], self.frames_right[i],
                                   None if priors is None else priors[i], **kwargs)
        elapsed = time.time() - start_timer
        logger.info("Frame-based stereo matching took %ss per image pair on average.", elapsed / max(n_frames, 1))
