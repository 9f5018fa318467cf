"""Run the REAL ``split_frames_by_time`` / ``generate_frames_from_spikes`` of the reference
(stereovis/utils/frames_io.py:184-248) in the build container.

TEST INFRASTRUCTURE ONLY.  ``frames_io.py`` cannot be imported here (skimage, matplotlib absent;
``np.int`` was removed from numpy), so the two function definitions are cut out of the source
with ``ast`` and executed with ``np.int`` mapped to ``int`` -- nothing is copied into the repo.
"""
import ast
import logging
import os
import types

import numpy as np

from .ref_loader import REFERENCE_ROOT

_PATH = os.path.join(REFERENCE_ROOT, "stereovis", "utils", "frames_io.py")


def available():
    return os.path.isfile(_PATH)


def load():
    source = open(_PATH).read()
    tree = ast.parse(source)
    wanted = [n for n in tree.body if isinstance(n, ast.FunctionDef)
              and n.name in ("split_frames_by_time", "generate_frames_from_spikes")]
    assert len(wanted) == 2
    module = ast.Module(body=wanted, type_ignores=[])
    np_shim = types.SimpleNamespace(**{k: getattr(np, k) for k in dir(np) if not k.startswith("_")})
    np_shim.int = int
    namespace = {"np": np_shim, "logger": logging.getLogger("reference_frames_io")}
    exec(compile(module, _PATH, "exec"), namespace)
    return namespace["split_frames_by_time"], namespace["generate_frames_from_spikes"]
