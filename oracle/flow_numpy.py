"""numpy restatement of the reference's prior adjustment and of scikit-image 0.13's down-scaling.
TEST INFRASTRUCTURE ONLY -- checkers for hybrid-stereo-matching_b200/priors.py (rows N3, N4)."""
import numpy as np

_SHIFTS = np.asarray([(1, 0), (1, 1), (0, 1), (-1, 1), (-1, 0), (-1, -1), (0, -1), (1, -1)], dtype=np.int32)


def compute_shift(x, y):
    """optical_flow_correction.py:64-65."""
    if np.linalg.norm([x, y]) > 1.:
        return _SHIFTS[int(np.floor(np.round(8 * np.arctan2(y, x) / (2 * np.pi)))) % 8]
    return np.array([0, 0])


def adjust(event_frames, resolution, prior_nondata_value=-1, time_arrow=+1):
    """optical_flow_correction.py:34-88 without the (unused, see its FIXME at :69) velocity field: every pixel
    with a value >= 0 is scattered, in row-major order, one compass step away from where it is."""
    event_frames = np.asarray(event_frames)
    adjusted_frames = np.empty_like(event_frames, dtype=np.float32)
    for i, prior_frame in enumerate(event_frames):
        adjusted = np.ones_like(prior_frame) * prior_nondata_value
        for row, col in np.argwhere(prior_frame >= 0):
            dcol, drow = time_arrow * compute_shift(col, row)
            if 0 <= col + dcol < resolution[0] and 0 <= row - drow < resolution[1]:
                adjusted[row - drow, col + dcol] = prior_frame[row, col]
        adjusted_frames[i] = adjusted
    return adjusted_frames


def rescale(image, scale):
    """scikit-image 0.13.0 ``rescale(image, scale, preserve_range=True, mode='constant')`` for a 2-D image, with
    the exact affine map (skimage estimates it by least squares).  PARITY UNPINNED: restated from the published
    algorithm (rescale -> resize -> warp -> _warp_fast / bilinear_interpolation, _clip_warp_output); scikit-image
    cannot be installed in the build container."""
    image = np.asarray(image).astype(np.double)
    scale = np.atleast_1d(scale)
    row_scale, col_scale = (scale[0], scale[1]) if scale.size > 1 else (scale[0], scale[0])
    rows, cols = image.shape
    out_rows, out_cols = int(np.round(row_scale * rows)), int(np.round(col_scale * cols))
    rs, cs = float(rows) / out_rows, float(cols) / out_cols

    def pixel(r, c):
        return 0.0 if (r < 0 or r >= rows or c < 0 or c >= cols) else image[r, c]

    out = np.empty((out_rows, out_cols))
    for y in range(out_rows):
        r = rs * y + (rs * 0.5 - 0.5)
        minr, maxr = int(np.floor(r)), int(np.ceil(r))
        dr = r - minr
        for x in range(out_cols):
            c = cs * x + (cs * 0.5 - 0.5)
            minc, maxc = int(np.floor(c)), int(np.ceil(c))
            dc = c - minc
            top = (1 - dc) * pixel(minr, minc) + dc * pixel(minr, maxc)
            bottom = (1 - dc) * pixel(maxr, minc) + dc * pixel(maxr, maxc)
            out[y, x] = (1 - dr) * top + dr * bottom
    lo, hi = image.min(), image.max()
    preserve = not (lo <= 0 <= hi)
    mask = out == 0
    np.clip(out, lo, hi, out=out)
    if preserve:
        out[mask] = 0
    return out


def fit_velocity_field(events, time_interval=20, neighbourhood_size=(3, 3), rejection_threshold=0.005,
                       convergence_threshold=1e-5, max_iter_steps=5, min_events=10):
    """vvf.py:32-82 restated for one polarity group: ``events`` (N, 3) of (t, x, y) sorted by time -> (N, 2).
    Least squares through numpy's lstsq (an SVD like the reference's scipy.linalg.lstsq)."""
    events = np.asarray(events, dtype=np.float64)
    velocities = np.zeros((events.shape[0], 2))
    times_all = events[:, 0]
    for index, event in enumerate(events):
        first = np.searchsorted(times_all, event[0] - time_interval, side="right")
        last = np.searchsorted(times_all, event[0] + time_interval, side="right")
        window = events[first:last]
        near = window[(np.abs(window[:, 1] - event[1]) <= neighbourhood_size[0])
                      & (np.abs(window[:, 2] - event[2]) <= neighbourhood_size[1])]
        count = near.shape[0]
        if count < min_events:
            continue
        positions = np.hstack([near[:, 1:], np.ones((count, 1))])
        times = near[:, 0]
        best = np.linalg.lstsq(positions, times, rcond=None)[0]
        epsilon, steps = np.inf, 0
        while steps < max_iter_steps and epsilon > convergence_threshold:
            steps += 1
            normal = np.hstack((best[2], best[:-1]))
            normal = normal / np.linalg.norm(normal)
            accepted = np.where(np.abs(np.dot(near, normal)) <= rejection_threshold)[0]
            if accepted.size < min_events or accepted.size == count:
                break
            new = np.linalg.lstsq(positions[accepted], times[accepted], rcond=None)[0]
            epsilon = np.linalg.norm(best - new)
            best = new
        velocities[index] = best[:-1]
    return velocities
