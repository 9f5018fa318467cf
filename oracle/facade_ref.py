"""Run the REAL ``FramebasedStereoMatching`` of the reference (stereovis/framed/stereo_framebased.py:15-101)
in the build container.

TEST INFRASTRUCTURE ONLY.  The module cannot be imported here (``spinn_utilities`` and ``matplotlib`` are
absent, ``stereovis.framed.algorithms`` uses a py2 implicit relative import), so the class definition is cut
out of the source with ``ast`` and executed in a namespace that provides what the module's imports would have:
``StereoMRF`` = the real reference matcher (``oracle/ref_loader.py``), ``ProgressBar`` and ``plt`` = no-op stubs
(progress output and the debug PNG dumps of ``run`` are side effects, not results).  Nothing is copied into
the repository; ``oracle/make_golden_facade.py`` stores the outputs as fixtures for the GPU box.
"""
import ast
import logging
import os
import time

import numpy as np

from .ref_loader import REFERENCE_ROOT, reference_class

_PATH = os.path.join(REFERENCE_ROOT, "stereovis", "framed", "stereo_framebased.py")


def available():
    return os.path.isfile(_PATH)


class _ProgressBar(object):
    def __init__(self, *args, **kwargs):
        pass

    def update(self):
        pass

    def end(self):
        pass


class _Plt(object):
    @staticmethod
    def imsave(*args, **kwargs):
        pass


def load():
    """The reference's ``FramebasedStereoMatching`` class, bound to the reference's ``StereoMRF``."""
    tree = ast.parse(open(_PATH).read())
    wanted = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "FramebasedStereoMatching"]
    assert len(wanted) == 1
    namespace = {"np": np, "time": time, "logger": logging.getLogger("reference_stereo_framebased"),
                 "StereoMRF": reference_class(), "ProgressBar": _ProgressBar, "plt": _Plt}
    exec(compile(ast.Module(body=wanted, type_ignores=[]), _PATH, "exec"), namespace)
    return namespace["FramebasedStereoMatching"]
