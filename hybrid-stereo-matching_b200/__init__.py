"""B200-native frame-based MRF stereo matcher (drop-in for the reference's
``stereovis.framed.algorithms.StereoMRF``).

The package directory is named after the project (``hybrid-stereo-matching_b200``);
``import hybrid_stereo_matching_b200`` (the alias module at the repo root) or
``importlib.import_module("hybrid-stereo-matching_b200")`` both give this package.
"""
from ._capi import build_library, LIB_PATH      # noqa: F401
from .mrf import StereoMRF, pinned_empty        # noqa: F401
from .framed import FramebasedStereoMatching    # noqa: F401

__all__ = ["StereoMRF", "FramebasedStereoMatching", "pinned_empty", "build_library", "LIB_PATH"]
