"""Generate ``tests/golden/*.npz`` by running the REAL reference in this container.

TEST INFRASTRUCTURE ONLY.  Run once from the repo root in the build container
(``python -m oracle.make_golden``); the outputs are committed.  The GPU box has
no ``/root/reference``, so tests there read these fixtures.

For each case in ``oracle/cases.py`` the reference ``StereoMRF`` (loaded by
``oracle/ref_loader.py``) is driven through the case's calls and the fixture
stores: labels, and either the five planes (small cases) or per-plane sha1 +
float64 sums (big cases), together with a sha1 of the inputs so a drifting
input generator is detected rather than silently compared.
"""
import hashlib
import json
import os
import sys

import numpy as np

from . import cases as C
from .ref_loader import REFERENCE_ROOT, reference_class


def _export_tsukuba():
    from PIL import Image
    out = os.path.join(C.GOLDEN_DIR, "tsukuba_rgb.npz")
    demo = os.path.join(REFERENCE_ROOT, "data", "demo_samples")
    left = np.asarray(Image.open(os.path.join(demo, "tsukuba_l.png")).convert("RGB"))
    right = np.asarray(Image.open(os.path.join(demo, "tsukuba_r.png")).convert("RGB"))
    np.savez_compressed(out, left=left, right=right)
    return out


def _inputs_sha1(case):
    digest = hashlib.sha1()
    for index in range(len(case["calls"])):
        left, right, prior, _ = C.case_inputs(case, index)
        digest.update(np.ascontiguousarray(left).tobytes())
        digest.update(np.ascontiguousarray(right).tobytes())
        if prior is not None:
            digest.update(np.ascontiguousarray(prior).tobytes())
    return digest.hexdigest()


def main(argv=None):
    """``python -m oracle.make_golden [case name ...]``: all cases, or only the named ones."""
    names = list(sys.argv[1:] if argv is None else argv)
    os.makedirs(C.GOLDEN_DIR, exist_ok=True)
    if not names:
        print("tsukuba ->", _export_tsukuba())
    reference = reference_class()
    factory = lambda dim, levels: reference(tuple(dim), levels)
    todo = [(c, True) for c in C.SMALL_CASES] + [(c, False) for c in C.BIG_CASES + C.HUGE_CASES]
    for case, small in todo:
        if names and case["name"] not in names:
            continue
        labels, planes = C.run_case(factory, case)
        meta = dict(name=case["name"], inputs_sha1=_inputs_sha1(case), small=small,
                    plane_sha1={k: C.plane_sha1(v) for k, v in planes.items()},
                    plane_sum={k: float(np.sum(v, dtype=np.float64)) for k, v in planes.items()},
                    labels_sha1=hashlib.sha1(labels.astype(np.int32).tobytes()).hexdigest(),
                    generated_by="oracle/make_golden.py from the reference's mrf.py (py3 token shim)",
                    numpy=np.__version__)
        arrays = dict(meta=np.asarray(json.dumps(meta)),
                      labels=labels.astype(np.int16 if case["levels"] > 127 else np.int8))
        if small:
            for k, v in planes.items():
                arrays["plane_" + k] = v
        np.savez_compressed(C.fixture_path(case), **arrays)
        print("%-34s labels sha1 %s" % (case["name"], meta["labels_sha1"]), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
