"""ctypes front end of the C oracle (``oracle/mrf_c.c``).

TEST INFRASTRUCTURE ONLY (see ``oracle/mrf_numpy.py``).  Same surface as the
reference's ``StereoMRF`` (mrf.py:8-167); planes are kept disparity-innermost
inside and transposed to the reference's ``(D, H, W)`` on read-back.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libmrf_oracle.so")
_lib = None


def build(force=False):
    """Compile oracle/mrf_c.c with gcc (no GPU needed)."""
    if force or not os.path.exists(_LIB_PATH) or \
            os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "mrf_c.c")):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "_build/libmrf_oracle.so"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        fp = ctypes.POINTER(ctypes.c_float)
        _lib.mrf_c_cost_volume.argtypes = [fp, fp, ctypes.POINTER(ctypes.c_int32), ctypes.c_int,
                                           ctypes.c_float, ctypes.c_double,
                                           ctypes.c_int, ctypes.c_int, ctypes.c_int, fp]
        _lib.mrf_c_iterate.argtypes = [fp, fp, fp, fp, fp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_int]
        _lib.mrf_c_readout.argtypes = [fp, fp, fp, fp, fp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                       ctypes.POINTER(ctypes.c_int64), fp]
        for fn in (_lib.mrf_c_cost_volume, _lib.mrf_c_iterate, _lib.mrf_c_readout):
            fn.restype = None
    return _lib


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def prior_labels(prior, n_levels):
    """int32 label the prior equals exactly (``prior == l`` for some l in [0, D)), else -1."""
    p = np.asarray(prior)
    with np.errstate(invalid="ignore"):
        ok = (p >= 0) & (p < n_levels) & (p == np.floor(p)) if p.dtype.kind == "f" \
            else (p >= 0) & (p < n_levels)
    out = np.full(p.shape, -1, dtype=np.int32)
    out[ok] = p[ok].astype(np.int32)
    return out


def blend_mode(prior, prior_trust_factor, prior_influence_mode):
    """(mode, one_minus_tau_f32, one_minus_tau_f64) as mrf_c_cost_volume expects."""
    if prior is None:
        return 0, 0.0, 0.0
    if prior_influence_mode == "adaptive":
        return 2, 0.0, 0.0
    keep = 1 - prior_trust_factor                                    # mrf.py:66
    if isinstance(keep, np.floating) and keep.dtype == np.float64:
        return 3, 0.0, float(keep)                                   # numpy float64 scalar: f64 product
    return 1, float(np.float32(keep)), 0.0


class OracleMRFC(object):
    def __init__(self, dim, n_levels):
        self.n_levels = n_levels
        self.dimension = (n_levels,) + tuple(dim)
        rows, cols = dim
        self._planes = {k: np.zeros((rows, cols, n_levels), dtype=np.float32)
                        for k in ("south", "west", "north", "east", "data")}

    @property
    def _message_field(self):
        return {k: np.ascontiguousarray(v.transpose(2, 0, 1)) for k, v in self._planes.items()}

    def loop_belief(self, image_left, image_right, prior=None, prior_trust_factor=1.0,
                    prior_influence_mode="const", n_iter=10, reinit_messages=True):
        assert image_left.shape == image_right.shape
        rows, cols, levels = self._planes["data"].shape
        left = np.ascontiguousarray(image_left.astype("float32"))
        right = np.ascontiguousarray(image_right.astype("float32"))
        if reinit_messages:
            for k in ("south", "west", "north", "east"):
                self._planes[k][...] = 0
        mode, keep32, keep64 = blend_mode(prior, prior_trust_factor, prior_influence_mode)
        plabel = None
        if prior is not None:
            assert image_right.shape == prior.shape and prior.shape == (rows, cols)
            plabel = np.ascontiguousarray(prior_labels(prior, levels))
        p = self._planes
        lib().mrf_c_cost_volume(_fp(left), _fp(right),
                                plabel.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)) if plabel is not None else None,
                                mode, keep32, keep64, rows, cols, levels, _fp(p["data"]))
        lib().mrf_c_iterate(_fp(p["south"]), _fp(p["west"]), _fp(p["north"]), _fp(p["east"]),
                            _fp(p["data"]), rows, cols, levels, int(n_iter))

    def energy(self):
        rows, cols, levels = self._planes["data"].shape
        p = self._planes
        out = np.empty((rows, cols, levels), dtype=np.float32)
        lab = np.empty((rows, cols), dtype=np.int64)
        lib().mrf_c_readout(_fp(p["south"]), _fp(p["west"]), _fp(p["north"]), _fp(p["east"]),
                            _fp(p["data"]), rows, cols, levels,
                            lab.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), _fp(out))
        return np.ascontiguousarray(out.transpose(2, 0, 1))

    def get_map_belief(self):
        rows, cols, levels = self._planes["data"].shape
        p = self._planes
        lab = np.empty((rows, cols), dtype=np.int64)
        lib().mrf_c_readout(_fp(p["south"]), _fp(p["west"]), _fp(p["north"]), _fp(p["east"]),
                            _fp(p["data"]), rows, cols, levels,
                            lab.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), None)
        self._belief_field = lab
        return lab

    def lbp(self, image_left, image_right, prior=None, prior_trust_factor=0.5,
            prior_influence_mode="const", n_iter=10):
        self.loop_belief(image_left, image_right, prior, prior_trust_factor,
                         prior_influence_mode, n_iter, reinit_messages=True)
        return self.get_map_belief()
