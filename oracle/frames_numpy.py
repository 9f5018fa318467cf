"""numpy restatement of the reference's prior rasterisation (frames_io.py:184-248).
TEST INFRASTRUCTURE ONLY -- the checker for hybrid-stereo-matching_b200/priors.py."""
import numpy as np


def split(ts, start_time=0, time_interval=100, pivots=None, sort_kind=None):
    ts = np.asarray(ts)
    order = np.argsort(ts, kind=sort_kind)
    times = ts[order]
    keep = times >= start_time
    times = times[keep]
    order = order[order.size - times.size:]
    if pivots is None:
        wrapped = np.convolve(times % time_interval, [1, -1], mode="valid") < 0
        ticks = np.where(np.concatenate([np.array([False]), wrapped]))[0]
        return np.split(order, ticks)[:-1], times[ticks]
    groups = [order[np.where((times >= tick - time_interval) & (times <= tick))[0]] for tick in pivots]
    return groups, pivots


def rasterise(resolution, xs, ys, ts, zs, start_time=0, time_interval=100, pivots=None, non_pixel_value="nan",
              sort_kind=None):
    """``sort_kind``: None = numpy's default like the reference (ties in an implementation-defined order),
    "stable" = ties in input order (what the product defines)."""
    xs, ys = np.asarray(xs).astype(int), np.asarray(ys).astype(int)
    ts, zs = np.asarray(ts), np.asarray(zs)
    groups, timestamps = split(ts, start_time, time_interval, pivots, sort_kind)
    n_cols, n_rows = resolution
    frames = np.ones((len(groups), n_rows, n_cols)) * float(non_pixel_value)
    for i, inds in enumerate(groups):
        if inds.size:
            frames[i, ys[inds], xs[inds]] = zs[inds]
    return frames, timestamps


def normalise_contrast(img):
    """frames_io.py:103-109."""
    scale_val = 1.
    if img.max() - img.min() != 0:
        return scale_val / (img.max() - img.min()) * (img - img.min())
    return np.ones_like(img) * scale_val
