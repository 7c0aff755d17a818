"""ctypes binding of include/mcstate_b200.h (libmcstate_b200.so).

The shared library holds every kernel and the whole engine; Python only
marshals host arrays across the C ABI.  There is deliberately no fallback: if
the library is missing, or no CUDA device is usable, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# MCSTATE_B200_LIB: load another build of the same engine (kernel tuning experiments)
LIB_PATH = os.environ.get("MCSTATE_B200_LIB") or os.path.join(_HERE, "libmcstate_b200.so")

OK, ERR_ARG, ERR_CUDA, ERR_STATE = 0, -1, -2, -3
SCRAMBLER_PLUS, SCRAMBLER_STARSTAR = 0, 1
SEED_SHARED, SEED_PER_SET = 0, 1

u64p = C.POINTER(C.c_uint64)
f64p = C.POINTER(C.c_double)
i32p = C.POINTER(C.c_int)
szp = C.POINTER(C.c_size_t)
handle_p = C.c_void_p


class McbError(RuntimeError):
    """An error reported by the native engine (R: a stop() from the model object)."""

    def __init__(self, code, message):
        super().__init__(message)
        self.code = code


# name -> (restype, argtypes); every symbol declared in include/mcstate_b200.h
SIGNATURES = {
    "mcb_version": (C.c_int, []),
    "mcb_last_error": (C.c_char_p, []),
    "mcb_device_count": (C.c_int, [i32p]),
    "mcb_device_info": (C.c_int, [C.c_int, C.c_char_p, C.c_size_t, i32p, i32p, i32p, szp]),
    "mcb_model_lookup": (C.c_int, [C.c_char_p, i32p, szp, szp, szp]),
    "mcb_model_par_name": (C.c_char_p, [C.c_int, C.c_size_t]),
    "mcb_model_par_default": (C.c_double, [C.c_int, C.c_size_t]),
    "mcb_model_state_name": (C.c_char_p, [C.c_int, C.c_size_t]),
    "mcb_model_data_name": (C.c_char_p, [C.c_int, C.c_size_t]),
    "mcb_rng_seed": (C.c_int, [C.c_uint64, u64p]),
    "mcb_rng_jump": (C.c_int, [u64p]),
    "mcb_rng_long_jump": (C.c_int, [u64p]),
    "mcb_rng_random_real": (C.c_int, [u64p, C.c_int, C.c_size_t, f64p]),
    "mcb_rng_jump_n": (C.c_int, [u64p, C.c_uint64]),
    "mcb_rng_distributed_state": (C.c_int, [u64p, C.c_size_t, C.c_size_t, u64p]),
    "mcb_alloc": (C.c_int, [C.c_int, f64p, C.c_size_t, C.c_size_t, C.c_uint64, u64p, C.c_size_t,
                            C.c_int, C.c_int, C.c_int, C.POINTER(handle_p)]),
    "mcb_alloc_ex": (C.c_int, [C.c_int, f64p, C.c_size_t, C.c_size_t, C.c_uint64, u64p, C.c_size_t,
                               C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(handle_p)]),
    "mcb_free": (C.c_int, [handle_p]),
    "mcb_set_stream": (C.c_int, [handle_p, C.c_void_p]),
    "mcb_synchronize": (C.c_int, [handle_p]),
    "mcb_shape": (C.c_int, [handle_p, szp, szp, szp, szp]),
    "mcb_time": (C.c_int, [handle_p, u64p]),
    "mcb_pars": (C.c_int, [handle_p, f64p]),
    "mcb_set_index": (C.c_int, [handle_p, szp, C.c_size_t]),
    "mcb_run": (C.c_int, [handle_p, C.c_uint64, f64p]),
    "mcb_simulate": (C.c_int, [handle_p, u64p, C.c_size_t, f64p]),
    "mcb_state": (C.c_int, [handle_p, szp, C.c_size_t, f64p]),
    "mcb_state_particles": (C.c_int, [handle_p, szp, C.c_size_t, f64p]),
    "mcb_update_state": (C.c_int, [handle_p, f64p, f64p, C.c_size_t, C.c_int, C.c_uint64, C.c_int]),
    "mcb_reorder": (C.c_int, [handle_p, i32p]),
    "mcb_rng_state": (C.c_int, [handle_p, C.c_int, u64p]),
    "mcb_set_rng_state": (C.c_int, [handle_p, u64p, C.c_size_t]),
    "mcb_set_data": (C.c_int, [handle_p, C.c_size_t, u64p, f64p, C.c_int]),
    "mcb_set_filter_stream": (C.c_int, [handle_p, u64p]),
    "mcb_set_stream_sharing": (C.c_int, [handle_p, C.c_size_t, C.c_size_t, C.c_char_p]),
    "mcb_copy_sets": (C.c_int, [handle_p, handle_p, C.c_char_p, C.c_int, C.c_int]),
    "mcb_compare_data": (C.c_int, [handle_p, f64p, i32p]),
    "mcb_resample": (C.c_int, [handle_p, f64p, i32p]),
    "mcb_filter_rows": (C.c_int, [handle_p, C.c_uint64, szp]),
    "mcb_filter": (C.c_int, [handle_p, C.c_uint64, C.c_int, u64p, C.c_size_t, f64p, C.c_size_t,
                             f64p, f64p, f64p]),
    "mcb_filter_history": (C.c_int, [handle_p, szp, C.c_size_t, f64p]),
    "mcb_set_persistent": (C.c_int, [handle_p, C.c_int]),
    "mcb_filter_enqueue": (C.c_int, [handle_p, C.c_uint64]),
    "mcb_filter_collect": (C.c_int, [handle_p, f64p]),
    "mcb_log_likelihood_device": (C.c_int, [handle_p, C.POINTER(C.c_void_p)]),
    "mcb_launch_count": (C.c_int, [handle_p, u64p]),
    "mcb_set_kernel_timing": (C.c_int, [handle_p, C.c_int]),
    "mcb_kernel_timing": (C.c_int, [handle_p, f64p, u64p]),
}


def build(force: bool = False) -> str:
    """Compile libmcstate_b200.so in-tree with nvcc for sm_100a (no GPU needed)."""
    src = os.path.join(_HERE, "csrc")
    newest = max(os.path.getmtime(os.path.join(src, f)) for f in os.listdir(src))
    hdr = os.path.join(os.path.dirname(_HERE), "include", "mcstate_b200.h")
    newest = max(newest, os.path.getmtime(hdr))
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < newest:
        subprocess.run(["make", "-C", src, "-s"] + (["-B"] if force else []), check=True)
    return LIB_PATH


_lib = None


def load(path: str):
    """dlopen one build of the engine and bind every entry point of the header."""
    L = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    return L


def lib():
    """Load the engine; raises if the shared library has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libmcstate_b200.so is missing: run `python -c 'import __graft_entry__ as g; "
                "g.build()'` (mcstate_b200 has no CPU fallback)")
        _lib = load(LIB_PATH)
    return _lib


# ---- user-supplied models (dust::dust(<file>) in the reference) ----------------------
USER_MODEL_DIR = os.path.join(_HERE, "_user_models")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
_USER_NVFLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
                 "-fmad=false", "-Xcompiler", "-fPIC,-Wall", "-ccbin", "/usr/bin/g++", "-shared"]
_user_libs = {}


def user_model_path(source: str) -> str:
    """Where the engine built around ``source`` lives: the name carries a digest of the
    model source, the engine sources and the flags, so a stale build is never loaded."""
    import hashlib

    src = os.path.join(_HERE, "csrc")
    h = hashlib.sha256()
    files = [source, os.path.join(os.path.dirname(_HERE), "include", "mcstate_b200.h")]
    files += [os.path.join(src, f) for f in sorted(os.listdir(src)) if f.endswith((".cu", ".cuh"))]
    for f in files:
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(_USER_NVFLAGS).encode())
    stem = os.path.splitext(os.path.basename(source))[0]
    return os.path.join(USER_MODEL_DIR, "libmcb_%s_%s.so" % (stem, h.hexdigest()[:12]))


def build_user_model(source: str, quiet: bool = True, force: bool = False) -> str:
    """Compile the engine around one user-supplied model header (sm_100a, in-tree; nvcc
    cross-compiles without a GPU).  There is no other way to run such a model: without
    nvcc, or if the source does not compile, this raises."""
    source = os.path.abspath(source)
    if not os.path.exists(source):
        raise FileNotFoundError("model source '%s' does not exist" % source)
    out = user_model_path(source)
    if force or not os.path.exists(out):
        if not os.path.exists(NVCC):
            raise RuntimeError("nvcc (%s) is needed to compile '%s': mcstate_b200 has no "
                               "interpreted or CPU path for user models" % (NVCC, source))
        os.makedirs(USER_MODEL_DIR, exist_ok=True)
        tmp = out + ".tmp%d" % os.getpid()
        cmd = [NVCC] + _USER_NVFLAGS + ['-DMCB_USER_MODEL="%s"' % source, "-o", tmp,
                                         os.path.join(_HERE, "csrc", "api.cu")]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            if os.path.exists(tmp):
                os.remove(tmp)
            raise RuntimeError("compiling '%s' failed:\n%s" % (source, r.stderr[-4000:]))
        if not quiet and r.stderr:
            print(r.stderr)
        os.replace(tmp, out)
        # builds of the same model from older sources can never be loaded again
        stem = os.path.basename(out).rsplit("_", 1)[0] + "_"
        for f in os.listdir(USER_MODEL_DIR):
            if f.startswith(stem) and f.endswith(".so") and f != os.path.basename(out):
                os.remove(os.path.join(USER_MODEL_DIR, f))
    return out


def user_lib(source: str, quiet: bool = True):
    """The loaded engine for one user model (built on first use, one dlopen per build)."""
    path = build_user_model(source, quiet)
    if path not in _user_libs:
        _user_libs[path] = load(path)
    return _user_libs[path]


def check(rc: int) -> None:
    if rc != OK:
        msg = lib().mcb_last_error()
        raise McbError(rc, msg.decode() if msg else "mcstate_b200 error %d" % rc)


def ptr(a, typ):
    return None if a is None else a.ctypes.data_as(typ)
