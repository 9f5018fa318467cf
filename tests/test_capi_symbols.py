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


def test_options_round_trip_and_validation_without_cuda(hsm):
    """hsm_set_option / hsm_get_option need no device: every option round-trips, bad values are refused."""
    from importlib import import_module
    capi = import_module("hybrid-stereo-matching_b200._capi")
    m = hsm.StereoMRF((40, 64), 20)
    good = {capi.OPT_ALGORITHM: [0, 1, 2, 3, 4, 5, 6], capi.OPT_BAND_ROWS: [0, 7, 40], capi.OPT_DATA_TERM: [0, 1],
            capi.OPT_STAGES: [0, 2, 3, 8], capi.OPT_LANE_CONFIG: [0, 2, 3, 10, 11, 12, 14, 15, 16], capi.OPT_FRAME_GROUPS: [0, 1, 2],
            capi.OPT_VALID_LO: [0, 1], capi.OPT_VALID_HI: [38, 39], capi.OPT_OUT_LO: [1], capi.OPT_OUT_HI: [38]}
    for key, values in good.items():
        for value in values:
            m.set_option(key, value)
            assert m.get_option(key) == value, (key, value)
    bad = {capi.OPT_ALGORITHM: [-1, 7], capi.OPT_BAND_ROWS: [-1], capi.OPT_STAGES: [1, 9, -2],
           capi.OPT_LANE_CONFIG: [1, 17, -1],       # entry 1 covers 16 levels only; 10..16: 5 / 6-lane layouts, other CTA widths
           capi.OPT_FRAME_GROUPS: [-1, 3], capi.OPT_VALID_LO: [-1, 40], capi.OPT_OUT_HI: [40], 99: [0]}
    for key, values in bad.items():
        for value in values:
            with pytest.raises(ValueError):
                m.set_option(key, value)


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
