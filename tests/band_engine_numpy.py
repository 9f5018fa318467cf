"""A CPU band engine for the row-band protocol tests (tests only): the numpy oracle run on
this rank's rows plus ghost rows.  One reference iteration on the local array is exact for
the owned rows as long as the ghost rows hold the neighbours' old W, N, E."""
import contextlib

import numpy as np
import torch

from oracle import mrf_numpy


class NumpyBandEngine(object):
    def __init__(self, n_rows, n_cols, n_levels, first, last):
        self.first, self.last, self.n_rows = first, last, n_rows
        self.lo, self.hi = max(first - 1, 0), min(last + 1, n_rows - 1)    # local rows incl. ghosts
        self.n_cols, self.n_levels = n_cols, n_levels
        self.state = mrf_numpy.new_state(self.hi - self.lo + 1, n_cols, n_levels)

    def stream_context(self):
        return contextlib.nullcontext()

    def init_fields(self, image_left, image_right, prior, prior_trust_factor, prior_influence_mode):
        rows = slice(self.lo, self.hi + 1)
        mrf_numpy.cost_volume(self.state, image_left[rows], image_right[rows],
                              None if prior is None else prior[rows], prior_trust_factor,
                              prior_influence_mode, reinit_messages=True)

    def iterate(self, finalize, n_iter=1):
        mrf_numpy.iterate(self.state, int(n_iter))

    def new_halo(self):
        return torch.empty(3 * self.n_cols * self.n_levels, dtype=torch.float32)

    def _row(self, image_row):
        return image_row - self.lo

    def pack(self, top, bottom):
        for buf, row in ((top, self.first), (bottom, self.last)):
            if buf is not None:
                rows = [self.state[k][:, self._row(row), :].T for k in ("west", "north", "east")]   # (W, D) each
                buf.copy_(torch.from_numpy(np.ascontiguousarray(np.stack(rows)).reshape(-1)))

    def unpack(self, from_above, from_below):
        for buf, row in ((from_above, self.first - 1), (from_below, self.last + 1)):
            if buf is not None:
                data = buf.numpy().reshape(3, self.n_cols, self.n_levels)
                for i, k in enumerate(("west", "north", "east")):
                    self.state[k][:, self._row(row), :] = data[i].T

    def labels(self):
        full = mrf_numpy.labels(self.state)
        return full[self._row(self.first):self._row(self.last) + 1]

    def planes(self):
        rows = slice(self._row(self.first), self._row(self.last) + 1)
        return {k: v[:, rows, :] for k, v in self.state.items()}
