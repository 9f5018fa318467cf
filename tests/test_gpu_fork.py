"""Online mode constructs the matcher in the parent and calls lbp in a fork()ed child
(hybrid_stereo.py:107-111 -> online_processing.py:148): no CUDA call may happen before the first
compute call.  Run in a fresh interpreter (the pytest process already owns a CUDA context)."""
import subprocess
import sys
import os

import pytest

pytestmark = pytest.mark.gpu

SCRIPT = r"""
import os, sys, pickle
sys.path.insert(0, %r)
import numpy as np
import multiprocessing as mp
import hybrid_stereo_matching_b200 as hsm
from importlib import import_module
synth = import_module("hybrid-stereo-matching_b200.synth")
from oracle import mrf_numpy

left, right, disparity = synth.textured_pair(20, 30, 8, 3, block=(5, 6))
prior = synth.sparse_prior(disparity, 8, 3, density=0.3).astype(np.int32)
matcher = hsm.FramebasedStereoMatching((30, 20), 8)          # constructed in the parent, as hybrid_stereo.py does
results = mp.Manager().list()                                 # online_processing.py:132 stores (prior, depth_map) here

def child():
    depth = matcher.run_one_frame(left, right, prior, n_iter=5)
    results.append((prior, depth))

ctx = mp.get_context("fork")
p = ctx.Process(target=child)
p.start(); p.join(120)
assert p.exitcode == 0, p.exitcode
got_prior, got = results[0]
want = mrf_numpy.OracleMRF((20, 30), 8).lbp(left, right, prior, n_iter=5)
assert isinstance(got, np.ndarray) and got.dtype == np.int64 and np.array_equal(got, want)
print("fork ok")
"""


def test_construct_in_parent_compute_in_forked_child():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", SCRIPT % root], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                         universal_newlines=True, timeout=300)
    assert out.returncode == 0 and "fork ok" in out.stdout, out.stdout
