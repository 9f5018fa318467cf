"""ctypes binding of the C ABI in ``include/hsm.h`` (``libhsm_b200.so``).

The shared library is built in-tree by ``build_library()`` (nvcc, sm_100a) and
loaded from the package directory.  There is no CPU fallback: if the library is
missing and cannot be built, or no GPU is usable, the calls raise.
"""
import ctypes
import os
import shutil
import subprocess

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_PKG_DIR, "csrc")
LIB_PATH = os.path.join(_PKG_DIR, "libhsm_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_PKG_DIR), "include", "hsm.h")

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared"]

# status codes / enums of include/hsm.h
OK, ERR_ARG, ERR_CUDA, ERR_STATE = 0, 1, 2, 3
F32, F64, I32, I64 = 0, 1, 2, 3
PRIOR_NONE, PRIOR_CONST, PRIOR_ADAPTIVE, PRIOR_CONST_F64 = 0, 1, 2, 3
OPT_ALGORITHM, OPT_BAND_ROWS, OPT_DATA_TERM = 0, 1, 2
OPT_VALID_LO, OPT_VALID_HI, OPT_OUT_LO, OPT_OUT_HI, OPT_STAGES, OPT_LANE_CONFIG = 3, 4, 5, 6, 7, 8
OPT_FRAME_GROUPS = 9
PLANES = ("south", "west", "north", "east", "data")      # reference dict order, mrf.py:15-19


class PeerInfo(ctypes.Structure):
    _fields_ = [("ipc_handle", ctypes.c_ubyte * 64), ("off_w", ctypes.c_uint64 * 2), ("off_n", ctypes.c_uint64 * 2),
                ("off_e", ctypes.c_uint64 * 2), ("off_mailbox", ctypes.c_uint64), ("local_base", ctypes.c_uint64),
                ("rows", ctypes.c_int), ("cols", ctypes.c_int), ("dp", ctypes.c_int), ("out_lo", ctypes.c_int),
                ("out_hi", ctypes.c_int), ("device", ctypes.c_int), ("current", ctypes.c_int),
                ("reserved", ctypes.c_int)]


class Stats(ctypes.Structure):
    _fields_ = [("kernel_launches", ctypes.c_uint64), ("iteration_launches", ctypes.c_uint64),
                ("iterations", ctypes.c_uint64), ("iteration_ms", ctypes.c_double),
                ("device_bytes", ctypes.c_uint64)]


def _sources():
    return [os.path.join(_CSRC, f) for f in sorted(os.listdir(_CSRC)) if f.endswith((".cu", ".cuh", ".hpp"))] \
        + [HEADER_PATH]


STAMP_PATH = LIB_PATH + ".sources"      # sha256 of the sources the library was built from


def _source_digest():
    import hashlib
    digest = hashlib.sha256()
    for src in _sources():
        digest.update(os.path.basename(src).encode() + b"\0")
        with open(src, "rb") as handle:
            digest.update(handle.read())
    digest.update(" ".join(NVCC_FLAGS).encode())
    return digest.hexdigest()


def library_is_stale():
    """True when libhsm_b200.so is missing or was built from other sources.  The comparison is by content
    (a stamp file written next to the library), so that a copy of the tree with fresh timestamps neither
    triggers a rebuild nor hides a stale build; without a stamp, modification times decide."""
    if not os.path.exists(LIB_PATH):
        return True
    if os.path.exists(STAMP_PATH):
        with open(STAMP_PATH) as handle:
            return handle.read().strip() != _source_digest()
    built = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(src) > built for src in _sources())


def build_library(force=False, verbose=False, jobs=None):
    """Compile csrc/*.cu into libhsm_b200.so for sm_100a (works without a GPU).  The translation
    units (C ABI + one per iteration-kernel family) are compiled in parallel, then linked."""
    if not force and not library_is_stale():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build %s" % LIB_PATH)
    units = [f for f in sorted(os.listdir(_CSRC)) if f.endswith(".cu")]
    obj_dir = os.path.join(_PKG_DIR, "build")
    os.makedirs(obj_dir, exist_ok=True)
    compile_flags = [flag for flag in NVCC_FLAGS if flag != "-shared"] + (["-Xptxas", "-v"] if verbose else [])
    compile_flags += os.environ.get("HSM_NVCC_EXTRA", "").split()          # development: e.g. -DHSM_STORE_POLICY=1

    def compile_unit(unit):
        obj = os.path.join(obj_dir, unit[:-3] + ".o")
        cmd = [nvcc] + compile_flags + ["-c", "-o", obj, os.path.join(_CSRC, unit)]
        proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, universal_newlines=True)
        return unit, obj, proc.returncode, proc.stdout

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=jobs or min(len(units), os.cpu_count() or 1)) as pool:
        results = list(pool.map(compile_unit, units))
    for unit, _, code, output in results:
        if code != 0:
            raise RuntimeError("nvcc failed on %s:\n%s" % (unit, output))
        if verbose:
            print(output)
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH] + [r[1] for r in results]
    proc = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, universal_newlines=True)
    if proc.returncode != 0:
        raise RuntimeError("link failed:\n" + proc.stdout)
    if not os.environ.get("HSM_NVCC_EXTRA"):             # experiment builds are never "the" library
        with open(STAMP_PATH, "w") as handle:
            handle.write(_source_digest() + "\n")
    elif os.path.exists(STAMP_PATH):
        os.remove(STAMP_PATH)
    return LIB_PATH


_lib = None


def lib():
    """The loaded library (built first if it is missing or older than its sources)."""
    global _lib
    if _lib is not None:
        return _lib
    override = os.environ.get("HSM_B200_LIB")          # development: try an alternative build
    if override is None and library_is_stale():
        build_library()
    handle = ctypes.CDLL(override or LIB_PATH)
    vp, ci, cf, cd, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_double, ctypes.c_size_t
    sig = {
        "hsm_version": ([], ci),
        "hsm_last_error": ([], ctypes.c_char_p),
        "hsm_device_count": ([ctypes.POINTER(ci)], ci),
        "hsm_create": ([ci, ci, ci, ci, ci, ctypes.POINTER(vp)], ci),
        "hsm_destroy": ([vp], ci),
        "hsm_set_stream": ([vp, vp], ci),
        "hsm_set_option": ([vp, ci, ci], ci),
        "hsm_get_option": ([vp, ci, ctypes.POINTER(ci)], ci),
        "hsm_device_bytes": ([ci, ci, ci, ci, ctypes.POINTER(sz)], ci),
        "hsm_init_fields": ([vp, ci, vp, vp, ci, vp, ci, ci, cf, cd, ci, ci], ci),
        "hsm_iterate": ([vp, ci, ci], ci),
        "hsm_readout": ([vp, vp, ci, ci], ci),
        "hsm_lbp": ([vp, ci, vp, vp, ci, vp, ci, ci, cf, cd, ci, vp, ci, ci], ci),
        "hsm_sync": ([vp], ci),
        "hsm_get_plane": ([vp, ci, ci, vp], ci),
        "hsm_get_energy": ([vp, ci, vp], ci),
        "hsm_set_plane": ([vp, ci, ci, vp], ci),
        "hsm_halo_floats": ([vp, ctypes.POINTER(sz)], ci),
        "hsm_halo_pack": ([vp, vp, vp], ci),
        "hsm_halo_unpack": ([vp, vp, vp], ci),
        "hsm_peer_export": ([vp, ctypes.POINTER(PeerInfo)], ci),
        "hsm_peer_attach": ([vp, ci, ctypes.POINTER(PeerInfo), ci], ci),
        "hsm_peer_reset": ([vp], ci),
        "hsm_peer_detach": ([vp], ci),
        "hsm_get_stats": ([vp, ctypes.POINTER(Stats)], ci),
        "hsm_reset_stats": ([vp], ci),
        "hsm_rasterise_events": ([sz, vp, vp, vp, vp, ci, vp, vp, ci, ci, ci, cd, vp, vp, ci], ci),
        "hsm_rescale_frames": ([ci, ci, ci, vp, ci, ci, vp, ci], ci),
        "hsm_adjust_priors": ([ci, ci, ci, vp, cd, ci, vp, ci], ci),
        "hsm_fit_velocity_field": ([sz, vp, vp, vp, cd, cd, cd, cd, cd, ci, ci, vp, ci], ci),
        "hsm_device_alloc": ([sz, ci, ctypes.POINTER(vp)], ci),
        "hsm_device_free": ([vp, ci], ci),
        "hsm_device_read": ([vp, vp, sz, ci], ci),
        "hsm_normalise_contrast": ([ci, ci, ci, vp, vp, ci], ci),
        "hsm_selftest_division": ([vp, vp, sz, ctypes.POINTER(ctypes.c_uint64)], ci),
        "hsm_host_alloc": ([sz, ctypes.POINTER(vp)], ci),
        "hsm_host_free": ([vp], ci),
        "hsm_host_is_pinned": ([vp, ctypes.POINTER(ci)], ci),
    }
    for name, (argtypes, restype) in sig.items():
        fn = getattr(handle, name)
        fn.argtypes, fn.restype = argtypes, restype
    handle._hsm_symbols = tuple(sig)
    _lib = handle
    return _lib


def check(status):
    """Map a C status to the reference's error behaviour:
    shape-contract violations -> ValueError, CUDA/state errors -> RuntimeError (which
    hybrid_stereo.py:141-145 catches to fall back from online to offline mode)."""
    if status == OK:
        return
    message = lib().hsm_last_error().decode("utf-8", "replace")
    if status == ERR_ARG:
        raise ValueError(message)
    raise RuntimeError(message)
