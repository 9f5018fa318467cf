"""Synthetic stereo inputs of the shapes BASELINE.json names.

The reference's DAVIS recordings are not in its checkout
(``.MISSING_LARGE_BLOBS``), so the benchmark and the parity tests use a
textured generator with a known block-wise disparity map, plus a sparse
"SNN-like" prior: the true disparity at a few percent of the pixels, some of
them wrong, and ``-1`` where there is no event -- the format
``generate_frames_from_spikes(..., non_pixel_value=-1)`` produces for the
reference (hybrid_stereo.py:168-176, frames_io.py:207-248).

Everything here is host-side numpy and deterministic in ``seed``.
"""
import numpy as np


def _box_blur(img, radius=2):
    padded = np.pad(img, radius, mode="edge")
    rows, cols = img.shape
    acc = np.zeros_like(img)
    for dy in range(2 * radius + 1):
        for dx in range(2 * radius + 1):
            acc += padded[dy:dy + rows, dx:dx + cols]
    return acc / float((2 * radius + 1) ** 2)


def textured_pair(rows, cols, n_levels, seed, noise=0.02, block=(30, 40)):
    """Return ``(left, right, true_disparity)``: float64 images in [0, 1], int64 map.

    ``left[y, x] == right[y, x - d(y, x)]`` up to the additive noise, which is
    the correspondence the data term ``|L[y, x] - R[y, x - l]|`` scores
    (reference mrf.py:69-71).
    """
    rng = np.random.default_rng(seed)
    tex = _box_blur(rng.random((rows, cols + n_levels)))
    span = tex.max() - tex.min()
    tex = (tex - tex.min()) / (span if span > 0 else 1.0)
    low = min(2, n_levels - 1)
    high = max(low + 1, n_levels - 2)
    bh, bw = block
    blocks = rng.integers(low, high, size=(-(-rows // bh), -(-cols // bw)))
    disparity = np.kron(blocks, np.ones((bh, bw), dtype=np.int64))[:rows, :cols]
    source_cols = np.arange(cols)[None, :] + n_levels - disparity
    right = tex[:, n_levels:n_levels + cols].copy()
    left = np.take_along_axis(tex, source_cols, axis=1)
    left = np.clip(left + noise * rng.standard_normal(left.shape), 0.0, 1.0)
    right = np.clip(right + noise * rng.standard_normal(right.shape), 0.0, 1.0)
    return left, right, disparity


def sparse_prior(true_disparity, n_levels, seed, density=0.05, outliers=0.2, dtype=np.float64):
    """Event-disparity prior: ``-1`` = no event, else a disparity label."""
    rng = np.random.default_rng(seed + 7919)
    shape = true_disparity.shape
    prior = np.full(shape, -1, dtype=np.int64)
    events = rng.random(shape) < density
    values = true_disparity.copy()
    wrong = rng.random(shape) < outliers
    values[wrong] = rng.integers(0, n_levels, size=int(wrong.sum()))
    prior[events] = values[events]
    return prior.astype(dtype)


def synthetic_sequence(n_frames, rows, cols, n_levels, base_seed=1234, with_prior=True,
                       dtype=np.float64):
    """``(lefts, rights, priors)`` stacks of shape ``(N, H, W)`` (priors None if unused)."""
    lefts = np.empty((n_frames, rows, cols), dtype=dtype)
    rights = np.empty((n_frames, rows, cols), dtype=dtype)
    priors = np.empty((n_frames, rows, cols), dtype=dtype) if with_prior else None
    for i in range(n_frames):
        left, right, disparity = textured_pair(rows, cols, n_levels, base_seed + i)
        lefts[i], rights[i] = left, right
        if with_prior:
            priors[i] = sparse_prior(disparity, n_levels, base_seed + i)
    return lefts, rights, priors
