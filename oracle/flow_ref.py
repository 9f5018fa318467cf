"""Run the REAL prior-adjustment classes of the reference in the build container:
``VelocityVectorField`` (stereovis/spiking/algorithms/vvf.py) and ``OpticalFlowPixelCorrection``
(stereovis/spiking/optical_flow_correction.py).

TEST INFRASTRUCTURE ONLY.  Neither module imports here (py2 ``xrange``, ``np.int``, ``spinn_utilities``, package
imports), so the class definitions are cut out of the sources with ``ast`` and executed in a namespace that
supplies their imports: numpy with ``np.int`` mapped to ``int``, scipy's ``lstsq``, ``xrange`` = ``range``, the
reference's own ``split_frames_by_time`` (oracle/frames_ref.py) and a no-op ``ProgressBar``.  Nothing is copied
into the repository; oracle/make_golden_facade.py stores outputs as fixtures.
"""
import ast
import logging
import os
import time
import types

import numpy as np

from .ref_loader import REFERENCE_ROOT
from . import frames_ref
from .facade_ref import _ProgressBar

_VVF = os.path.join(REFERENCE_ROOT, "stereovis", "spiking", "algorithms", "vvf.py")
_OFC = os.path.join(REFERENCE_ROOT, "stereovis", "spiking", "optical_flow_correction.py")


def available():
    return os.path.isfile(_VVF) and os.path.isfile(_OFC)


def _class_from(path, name, namespace):
    tree = ast.parse(open(path).read())
    wanted = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == name]
    assert len(wanted) == 1
    exec(compile(ast.Module(body=wanted, type_ignores=[]), path, "exec"), namespace)
    return namespace[name]


def load():
    """(VelocityVectorField, OpticalFlowPixelCorrection) of the reference."""
    from scipy.linalg import lstsq
    np_shim = types.SimpleNamespace(**{k: getattr(np, k) for k in dir(np) if not k.startswith("_")})
    np_shim.int = int
    namespace = {"np": np_shim, "lstsq": lstsq, "xrange": range, "time": time,
                 "logger": logging.getLogger("reference_flow"), "ProgressBar": _ProgressBar,
                 "split_frames_by_time": frames_ref.load()[0]}
    field = _class_from(_VVF, "VelocityVectorField", namespace)
    correction = _class_from(_OFC, "OpticalFlowPixelCorrection", namespace)
    return field, correction
