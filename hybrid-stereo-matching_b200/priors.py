"""Event list -> prior frames (the producer of the matcher's ``prior`` input).

Mirrors ``split_frames_by_time`` and ``generate_frames_from_spikes`` of the reference
(stereovis/utils/frames_io.py:184-248), which ``hybrid_stereo.py:169-176`` calls with the frame
timestamps as ``pivots`` and ``non_pixel_value=-1``.  The index bookkeeping stays on the host
(numpy, same calls as the reference so that tie order matches); the scatter -- last event per pixel
wins -- runs on the GPU through ``hsm_rasterise_events``.
"""
import ctypes
import logging

import numpy as np

from . import _capi

logger = logging.getLogger(__name__)


def split_frames_by_time(ts, start_time=0, time_interval=100, pivots=None):
    """Indices of the events of each frame, chronological (frames_io.py:184-204)."""
    ts = np.asarray(ts)
    sorted_indices = np.argsort(ts)
    sorted_time = ts[sorted_indices]
    sorted_time = sorted_time[sorted_time >= start_time]
    sorted_indices = sorted_indices[sorted_indices.size - sorted_time.size:]
    if pivots is None:
        wraps = np.convolve(sorted_time % time_interval, [1, -1], mode="valid") < 0
        ticks = np.where(np.concatenate([np.array([False]), wraps]))[0]
        timestamps = sorted_time[ticks]
        indices_frames = np.split(sorted_indices, ticks)[:-1]      # the last, incomplete frame is dropped
    else:
        indices_frames = [sorted_indices[np.where(np.logical_and(sorted_time >= tick - time_interval,
                                                                 sorted_time <= tick))[0]] for tick in pivots]
        timestamps = pivots
    return indices_frames, timestamps


def generate_frames_from_spikes(resolution, xs, ys, ts, zs, start_time=0, time_interval=100, pivots=None,
                                non_pixel_value="nan", return_time_indices=False, device=0):
    """Frames ``(N, rows, cols)`` float64 buffered from events (frames_io.py:207-248)."""
    logger.info("Generating %s frames from spikes",
                len(pivots) if pivots is not None else int(np.max(ts) / time_interval))
    xs, ys = np.asarray(xs).astype(int), np.asarray(ys).astype(int)
    ts, zs = np.asarray(ts), np.asarray(zs)
    indices_frames, timestamps = split_frames_by_time(ts, start_time, time_interval, pivots)
    n_cols, n_rows = resolution
    fill = float(non_pixel_value)
    n_frames = len(indices_frames)
    order = np.concatenate(indices_frames) if n_frames else np.zeros(0, dtype=np.int64)
    order = order.astype(np.int64)
    frame_of = np.repeat(np.arange(n_frames, dtype=np.int32), [len(i) for i in indices_frames])
    ex, ey = xs[order], ys[order]
    # numpy fancy indexing semantics: negative indices wrap once, anything else is an IndexError
    for name, values, size in (("1", ey, n_rows), ("2", ex, n_cols)):
        bad = (values < -size) | (values >= size)
        if bad.any():
            raise IndexError("index %d is out of bounds for axis %s with size %d" % (values[bad][0], name, size))
    ex = np.ascontiguousarray(np.where(ex < 0, ex + n_cols, ex), dtype=np.int32)
    ey = np.ascontiguousarray(np.where(ey < 0, ey + n_rows, ey), dtype=np.int32)
    ez = np.ascontiguousarray(zs[order], dtype=np.float64)
    frames = np.empty((n_frames, n_rows, n_cols), dtype=np.float64)
    ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    _capi.check(_capi.lib().hsm_rasterise_events(order.size, ptr(ex), ptr(ey), ptr(ez), ptr(frame_of), n_frames,
                                                 n_rows, n_cols, fill, ptr(frames), int(device)))
    if return_time_indices:
        return frames, timestamps, indices_frames
    return frames, timestamps


def normalise_contrast(images, device=0):
    """Per-frame min-max normalisation to [0, 1] in float64 (frames_io.py:103-109 applied to every
    frame as load_frames does, frames_io.py:130-131).  ``images``: (N, rows, cols) or (rows, cols)."""
    images = np.asarray(images)
    single = images.ndim == 2
    stack = np.ascontiguousarray(images[None] if single else images, dtype=np.float64)
    out = np.empty_like(stack)
    ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    _capi.check(_capi.lib().hsm_normalise_contrast(stack.shape[0], stack.shape[1], stack.shape[2], ptr(stack),
                                                   ptr(out), int(device)))
    return out[0] if single else out


def prepare_frames(images, crop_region=None, resolution=None, scale_down_factor=(1, 1), adjust_contrast=False,
                   device=0):
    """The array part of ``load_frames`` (frames_io.py:91-131): crop every frame to ``resolution`` =
    (width, height) at ``crop_region`` = (column, row) of the top-left corner, then (optionally) stretch each
    frame's contrast to [0, 1] on the GPU.  Down-scaling (``skimage.transform.rescale``, frames_io.py:99-101)
    is an un-vendored third-party resampler whose output cannot be pinned here, so any
    factor other than (1, 1) is refused rather than approximated."""
    if tuple(int(f) for f in scale_down_factor) != (1, 1):
        raise NotImplementedError("scale_down_factor != (1, 1): skimage.rescale semantics are not reproduced")
    images = np.asarray(images)
    if images.ndim == 2:
        images = images[None]
    if crop_region is not None and resolution is not None:
        col_start, row_start = crop_region                      # frames_io.py:91-95: x, y order
        dx, dy = resolution
        images = images[:, row_start:row_start + dy, col_start:col_start + dx]
    if adjust_contrast:
        return normalise_contrast(images, device=device)
    return np.ascontiguousarray(images)
