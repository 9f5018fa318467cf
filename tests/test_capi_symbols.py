"""The C-ABI library builds for sm_100a, loads on a CPU-only box and exports every
symbol include/hsm.h declares.  No compute calls (no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest


def _declared_symbols(header_path):
    text = open(header_path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hsm_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_header_symbols(hsm):
    from importlib import import_module
    capi = import_module("hybrid-stereo-matching_b200._capi")
    lib = capi.lib()
    declared = _declared_symbols(capi.HEADER_PATH)
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), "missing export: " + name
    assert sorted(lib._hsm_symbols) == declared, "ctypes signatures out of sync with include/hsm.h"
    assert lib.hsm_version() >= 100


def test_create_validates_arguments_without_cuda(hsm):
    m = hsm.StereoMRF((12, 16), 5)
    assert m.n_levels == 5 and m.dimension == (5, 12, 16)
    with pytest.raises(ValueError):
        hsm.StereoMRF((12, 16), 18)          # n_levels > cols + 1: the reference cannot slice it
    with pytest.raises(ValueError):
        hsm.StereoMRF((12, 16), 0)
    with pytest.raises(ValueError):
        hsm.StereoMRF((12, 1000), 513)


def test_shape_contract_errors_match_reference(hsm):
    m = hsm.StereoMRF((12, 16), 5)
    a, b = np.zeros((12, 16)), np.zeros((12, 15))
    with pytest.raises(AssertionError):      # mrf.py:40
        m.lbp(a, b)
    with pytest.raises(AssertionError):      # mrf.py:53
        m.lbp(a, a, prior=np.zeros((3, 3)))
    with pytest.raises(ValueError):          # numpy broadcasting error in the reference
        m.lbp(np.zeros((10, 16)), np.zeros((10, 16)))


def test_no_cpu_fallback_without_gpu(hsm):
    """On a box without a usable GPU the product path must fail loudly, not fall back."""
    from importlib import import_module
    capi = import_module("hybrid-stereo-matching_b200._capi")
    count = ctypes.c_int(0)
    status = capi.lib().hsm_device_count(ctypes.byref(count))
    if status == 0 and count.value > 0:
        pytest.skip("a GPU is present")
    m = hsm.StereoMRF((12, 16), 5)
    with pytest.raises(RuntimeError):
        m.lbp(np.zeros((12, 16)), np.zeros((12, 16)))


def test_product_package_never_imports_the_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "hybrid-stereo-matching_b200")
    for dirpath, _, files in os.walk(pkg):
        for name in files:
            if name.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, name)).read()
                assert "oracle" not in text.replace("the oracle compares", ""), os.path.join(dirpath, name)
