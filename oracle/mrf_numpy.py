"""CPU oracle (numpy) for the frame-based MRF stereo matcher.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU
baseline / ``--impl reference`` legs do, and only as the checker or as the
timed CPU arm.

This is a restatement (not a copy) of the reference algorithm in
``stereovis/framed/algorithms/mrf.py`` (class ``StereoMRF``).  It keeps the
reference's memory layout -- five float32 planes of shape ``(D, H, W)``,
disparity outermost -- and uses the same numpy primitives in the same order,
so that (i) results are bit-identical to the reference and (ii) its run time
is representative of the reference's when it is used as the CPU baseline.

Parity pinning: the reference ships no tests or golden vectors for this path
(SURVEY.md section 8c).  The oracle is pinned instead against the reference
itself, executed in the build container by ``oracle/ref_loader.py``:
``oracle/make_golden.py`` writes the reference's outputs to ``tests/golden/``
and ``tests/test_oracle.py`` checks this module against them bit for bit.

Reference behaviours that are reproduced on purpose:
  * no smoothness term: a message is the sum of the sender's other mailboxes
    (mrf.py:82), shifted by one pixel (mrf.py:83-90);
  * direction order south, west, north, east, each sweep seeing the planes the
    earlier sweeps of the same iteration already rewrote (mrf.py:81);
  * the excluded mailbox is the sender's *own* ``direction`` plane (mrf.py:82);
  * normalisation divides by the per-pixel max over disparity, 0 -> 1
    (mrf.py:92-94);
  * the data term is 0 for columns ``x < l`` (mrf.py:69-71).
"""
import numpy as np

# Plane order == insertion order of the reference's dict under Python 3
# (mrf.py:15-19); np.sum(list, axis=0) folds left in this order.
PLANES = ("south", "west", "north", "east", "data")
DIRECTIONS = ("south", "west", "north", "east")          # mrf.py:81


def new_state(rows, cols, n_levels):
    """Five zero planes, (D, H, W) float32 (mrf.py:12-19)."""
    shape = (int(n_levels), int(rows), int(cols))
    return {name: np.zeros(shape, dtype=np.float32) for name in PLANES}


def cost_volume(state, image_left, image_right, prior=None, prior_trust_factor=1.0,
                prior_influence_mode="adaptive", reinit_messages=True):
    """Fill the data plane, optionally blending the prior (mrf.py:21-71).

    ``data[l, y, x] = |L[y, x] - R[y, x - l]|`` for ``x >= l`` and is left
    untouched (always zero) for ``x < l``.  Where ``prior[y, x] == l`` (and
    ``x >= l``) the cost becomes ``(1 - tau) * diff`` with tau the constant
    trust factor (``const``) or the cost itself (``adaptive``).
    Returns the float32 images (mrf.py:41-42).
    """
    assert image_left.shape == image_right.shape                     # mrf.py:40
    left32 = image_left.astype("float32")
    right32 = image_right.astype("float32")
    n_levels = state["data"].shape[0]
    if reinit_messages:                                              # mrf.py:43-48
        for name in PLANES:
            state[name] = np.zeros(state[name].shape, dtype=np.float32)
    width = left32.shape[1]
    data = state["data"]
    if prior is not None:                                            # mrf.py:52-53
        assert image_right.shape == prior.shape and prior.shape == data.shape[1:]
    for level in range(n_levels):
        diff = np.abs(left32[:, level:] - right32[:, :width - level])   # mrf.py:55 / :70
        if prior is None:
            data[level, :, level:] = diff
            continue
        hit = (prior == level)[:, level:]                            # mrf.py:57
        tau = prior_trust_factor
        if prior_influence_mode == "adaptive":                       # mrf.py:58-62
            tau = diff[hit] if diff.max() > 0 else 0.0
        view = data[level, :, level:]
        view[hit] = (1 - tau) * diff[hit]                            # mrf.py:66
        view[~hit] = diff[~hit]                                      # mrf.py:67
    return left32, right32


def sweep(state, direction):
    """One direction of message passing with max-normalisation (mrf.py:82-94)."""
    update = np.sum([plane for name, plane in state.items() if name != direction], axis=0)
    target = state[direction]
    if direction == "south":
        target[:, 1:, :] = update[:, :-1, :]
    elif direction == "west":
        target[:, :, :-1] = update[:, :, 1:]
    elif direction == "north":
        target[:, :-1, :] = update[:, 1:, :]
    elif direction == "east":
        target[:, :, 1:] = update[:, :, :-1]
    else:
        raise ValueError(direction)
    norm = np.max(target, axis=0, keepdims=True)
    norm[norm == 0] = 1.0
    target /= norm


def iterate(state, n_iter=1):
    """``n_iter`` full iterations, four sweeps each (mrf.py:73-94, 130-133)."""
    for _ in range(int(n_iter)):
        for direction in DIRECTIONS:
            sweep(state, direction)


def energy(state):
    """Sum of the five planes in dict order (mrf.py:103)."""
    return np.sum([state[name] for name in PLANES], axis=0)


def labels(state):
    """MAP labels: first minimum of the energy over disparity (mrf.py:104)."""
    return np.argmin(energy(state), axis=0)


class OracleMRF(object):
    """Same surface as the reference's ``StereoMRF`` (mrf.py:8-167)."""

    def __init__(self, dim, n_levels):
        self.n_levels = n_levels
        self.dimension = (n_levels,) + tuple(dim)
        self._message_field = new_state(dim[0], dim[1], n_levels)

    def loop_belief(self, image_left, image_right, prior=None, prior_trust_factor=1.0,
                    prior_influence_mode="const", n_iter=10, reinit_messages=True):
        self.reference_image, self.secondary_image = cost_volume(
            self._message_field, image_left, image_right, prior, prior_trust_factor,
            prior_influence_mode, reinit_messages)
        iterate(self._message_field, n_iter)

    def get_map_belief(self):
        self._belief_field = labels(self._message_field)
        return self._belief_field

    def lbp(self, image_left, image_right, prior=None, prior_trust_factor=0.5,
            prior_influence_mode="const", n_iter=10):
        self.loop_belief(image_left, image_right, prior, prior_trust_factor,
                         prior_influence_mode, n_iter, reinit_messages=True)
        return self.get_map_belief()
