"""Load the *real* reference ``StereoMRF`` in the build container.

TEST INFRASTRUCTURE ONLY (see ``oracle/mrf_numpy.py``).

The reference (``/root/reference/stereovis/framed/algorithms/mrf.py``) is
Python 2.  It runs unchanged under Python 3 / numpy 2 after exactly four token
substitutions (3 x ``xrange`` -> ``range``, 1 x ``.iteritems()`` -> ``.items()``;
SURVEY.md section 8c).  This loader reads the source where it lies, applies
those substitutions in memory and executes it as a module -- nothing is copied
into the repository.  ``/root/reference`` only exists in the build container,
so this is used by ``oracle/make_golden.py`` (fixture generation) and by the
``not gpu`` tests that validate the oracle when the reference is present; it
is never used on the GPU box.
"""
import os
import types

REFERENCE_ROOT = os.environ.get("HSM_REFERENCE_ROOT", "/root/reference")
_MRF_PATH = os.path.join(REFERENCE_ROOT, "stereovis", "framed", "algorithms", "mrf.py")
_SUBSTITUTIONS = (("xrange(", "range("), (".iteritems()", ".items()"))
_EXPECTED_COUNTS = (3, 1)


def reference_available():
    return os.path.isfile(_MRF_PATH)


def load_reference_module():
    """Return the reference's mrf module (py3-shimmed), or raise FileNotFoundError."""
    with open(_MRF_PATH, "r") as handle:
        source = handle.read()
    for (old, new), expected in zip(_SUBSTITUTIONS, _EXPECTED_COUNTS):
        found = source.count(old)
        if found != expected:
            raise RuntimeError("reference changed: %r occurs %d times, expected %d"
                               % (old, found, expected))
        source = source.replace(old, new)
    module = types.ModuleType("reference_mrf")
    module.__file__ = _MRF_PATH
    exec(compile(source, _MRF_PATH, "exec"), module.__dict__)
    return module


def reference_class():
    return load_reference_module().StereoMRF
