"""The drop-in boundary is a C ABI: a host program written in plain C (tests/c_abi/lbp_host.c, compiled here with
gcc against include/hsm.h and linked to libhsm_b200.so) must reproduce the reference fixtures bit for bit."""
import os
import struct
import subprocess
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cases as C

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SINGLE_CALL = [c for c in C.SMALL_CASES
               if len(c["calls"]) == 1 and c["calls"][0].get("prior") in (None, "f64")]


@pytest.fixture(scope="module")
def host_program(tmp_path_factory, hsm):
    import shutil
    from importlib import import_module
    if shutil.which("gcc") is None:
        pytest.skip("no C compiler on this box")
    capi = import_module("hybrid-stereo-matching_b200._capi")
    capi.lib()                                               # builds the library if needed
    out = str(tmp_path_factory.mktemp("c_abi") / "lbp_host")
    lib_dir = os.path.dirname(capi.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "c_abi", "lbp_host.c"), "-o", out, "-L", lib_dir, "-lhsm_b200",
                    "-Wl,-rpath," + lib_dir], check=True)
    return out


@pytest.mark.parametrize("case", SINGLE_CALL, ids=[c["name"] for c in SINGLE_CALL])
def test_c_host_program_matches_fixture(host_program, tmp_path, case):
    from importlib import import_module
    mrf = import_module("hybrid-stereo-matching_b200.mrf")
    capi = import_module("hybrid-stereo-matching_b200._capi")
    left, right, prior, kwargs = C.case_inputs(case, 0)
    rows, cols, levels = case["rows"], case["cols"], case["levels"]
    # loop_belief's defaults (mrf.py:106-107); the blend parameters are host logic (mrf.py:58-66)
    mode, keep32, keep64 = mrf._blend(prior, kwargs.get("prior_trust_factor", 1.0),
                                      kwargs.get("prior_influence_mode", "const"))
    src, dst = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(src, "wb") as handle:
        handle.write(struct.pack("<6i", rows, cols, levels, kwargs.get("n_iter", 10), int(prior is not None), mode))
        handle.write(struct.pack("<f", keep32))
        handle.write(struct.pack("<d", keep64))
        handle.write(np.ascontiguousarray(left, dtype=np.float64).tobytes())
        handle.write(np.ascontiguousarray(right, dtype=np.float64).tobytes())
        if prior is not None:
            handle.write(np.ascontiguousarray(prior, dtype=np.float64).tobytes())
    done = subprocess.run([host_program, src, dst], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120)
    assert done.returncode == 0, done.stderr.decode()
    raw = open(dst, "rb").read()
    n = rows * cols
    labels = np.frombuffer(raw, dtype=np.int64, count=n).reshape(rows, cols)
    data, _ = C.load_fixture(case)
    assert np.array_equal(labels, data["labels"])
    planes = np.frombuffer(raw, dtype=np.float32, offset=8 * n).reshape(5, rows, cols, levels)
    for index, name in enumerate(capi.PLANES):
        assert np.array_equal(planes[index].transpose(2, 0, 1), data["plane_" + name]), name
