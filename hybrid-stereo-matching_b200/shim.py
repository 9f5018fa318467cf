"""Install this package at the reference's import paths.

The reference imports the matcher as ``from stereovis.framed.algorithms import StereoMRF``
(stereo_framebased.py:5; the plugin seam is stereovis/framed/algorithms/__init__.py:1) and the
notebooks as ``stereovis.framed.algorithms.mrf.StereoMRF``.  ``install()`` registers modules of
those names in ``sys.modules`` (creating the parent packages if the reference is not importable),
so ``hybrid_stereo.py`` / ``main.py``-driven experiments pick up the CUDA matcher unchanged.
"""
import sys
import types


def install(replace_facade=True):
    from . import framed, mrf

    def ensure(name):
        module = sys.modules.get(name)
        if module is None:
            module = types.ModuleType(name)
            module.__path__ = []
            sys.modules[name] = module
            if "." in name:
                parent, _, child = name.rpartition(".")
                setattr(ensure(parent), child, module)
        return module

    algorithms = ensure("stereovis.framed.algorithms")
    algorithms.StereoMRF = mrf.StereoMRF
    mrf_module = types.ModuleType("stereovis.framed.algorithms.mrf")
    mrf_module.StereoMRF = mrf.StereoMRF
    sys.modules["stereovis.framed.algorithms.mrf"] = mrf_module
    algorithms.mrf = mrf_module
    if replace_facade:
        facade = types.ModuleType("stereovis.framed.stereo_framebased")
        facade.FramebasedStereoMatching = framed.FramebasedStereoMatching
        facade.StereoMRF = mrf.StereoMRF
        sys.modules["stereovis.framed.stereo_framebased"] = facade
        ensure("stereovis.framed").stereo_framebased = facade
    return algorithms
