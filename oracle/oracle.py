"""ctypes loader for the CPU oracle -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py`` (cpu_baseline and
the ``--impl reference`` arm) may import this module.  The product package
``mcstate_b200`` never does.

Array conventions follow R: a state "matrix" ``[n_state, N]`` is
stored with ``n_state`` fastest, i.e. a C-contiguous numpy array of shape
``(N, n_state)``; helpers below return ``(n_state, N)`` views (Fortran order) so
tests read like the reference's.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmcstate_oracle.so")

PLUS, STARSTAR = 0, 1
SIR, SIRS, VOLATILITY, WALK, SEIR = 0, 1, 2, 3, 4
DUST_FP64, EXACT = 0, 1

_u64p = C.POINTER(C.c_uint64)
_f64p = C.POINTER(C.c_double)
_i32p = C.POINTER(C.c_int)
_szp = C.POINTER(C.c_size_t)


def build(force: bool = False) -> str:
    """Compile the oracle with its Makefile (gcc only; no GPU, no reference needed)."""
    if force or not os.path.exists(_LIB_PATH) or (
        os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "mcstate_oracle.c"))
    ):
        subprocess.run(["make", "-C", _HERE, "-s"] + (["-B"] if force else []), check=True)
    return _LIB_PATH


def _cpu_has_fma() -> bool:
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.startswith("flags"):
                    return " fma " in line + " "
    except OSError:
        pass
    return True


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()  # no-op unless the C source is newer than the library
    if not _cpu_has_fma():
        raise RuntimeError("oracle is built with -mfma but this CPU lacks FMA")
    L = C.CDLL(_LIB_PATH)
    sig = {
        "orc_splitmix64": (C.c_uint64, [_u64p]),
        "orc_seed": (None, [C.c_uint64, _u64p]),
        "orc_next": (C.c_uint64, [_u64p, C.c_int]),
        "orc_jump": (None, [_u64p]),
        "orc_long_jump": (None, [_u64p]),
        "orc_random_real": (C.c_double, [_u64p, C.c_int]),
        "orc_rng_streams": (None, [_u64p, C.c_size_t, C.c_size_t, _u64p]),
        "orc_rng_distributed_state": (None, [_u64p, C.c_size_t, C.c_size_t, _u64p]),
        "orc_exp": (C.c_double, [C.c_double]),
        "orc_log": (C.c_double, [C.c_double]),
        "orc_exponential": (C.c_double, [_u64p, C.c_int, C.c_double]),
        "orc_cos2pi": (C.c_double, [C.c_double]),
        "orc_normal": (C.c_double, [_u64p, C.c_int, C.c_double, C.c_double]),
        "orc_density_normal_log": (C.c_double, [C.c_double, C.c_double, C.c_double]),
        "orc_binomial": (C.c_double, [_u64p, C.c_int, C.c_double, C.c_double]),
        "orc_poisson": (C.c_double, [_u64p, C.c_int, C.c_double]),
        "orc_lfact": (C.c_double, [C.c_double]),
        "orc_density_poisson_log": (C.c_double, [C.c_double, C.c_double, C.c_double]),
        "orc_scale_log_weights_r": (C.c_double, [_f64p, C.c_size_t, _f64p]),
        "orc_scale_log_weights": (C.c_double, [_f64p, C.c_size_t, C.c_int]),
        "orc_weight_shift": (C.c_int, [C.c_size_t]),
        "orc_particle_resample_r": (None, [_f64p, C.c_size_t, C.c_double, _i32p]),
        "orc_resample_weight": (None, [_f64p, C.c_size_t, C.c_double, C.c_int, _i32p]),
        "orc_history_single": (None, [_f64p, _i32p, C.c_size_t, C.c_size_t, C.c_size_t, _f64p]),
        "orc_model_n_state": (C.c_size_t, [C.c_int]),
        "orc_model_n_par": (C.c_size_t, [C.c_int]),
        "orc_model_default_pars": (None, [C.c_int, _f64p]),
        "orc_alloc": (C.c_void_p, [C.c_int, _f64p, C.c_size_t, C.c_size_t, C.c_uint64, _u64p,
                                   C.c_size_t, C.c_int, C.c_int, C.c_int]),
        "orc_free": (None, [C.c_void_p]),
        "orc_set_resample_mode": (None, [C.c_void_p, C.c_int]),
        "orc_set_n_threads": (None, [C.c_void_p, C.c_int]),
        "orc_n_total": (C.c_size_t, [C.c_void_p]),
        "orc_time": (C.c_uint64, [C.c_void_p]),
        "orc_run": (None, [C.c_void_p, C.c_uint64]),
        "orc_state": (None, [C.c_void_p, _szp, C.c_size_t, _f64p]),
        "orc_update_state": (None, [C.c_void_p, _f64p, _f64p, C.c_size_t, C.c_int, C.c_uint64,
                                    C.c_int]),
        "orc_reorder": (None, [C.c_void_p, _i32p]),
        "orc_rng_state": (None, [C.c_void_p, _u64p]),
        "orc_set_rng_state": (None, [C.c_void_p, _u64p]),
        "orc_set_data": (None, [C.c_void_p, C.c_size_t, _u64p, _f64p, C.c_int]),
        "orc_compare_data": (C.c_int, [C.c_void_p, _f64p]),
        "orc_resample": (None, [C.c_void_p, _f64p, _i32p]),
        "orc_set_filter_stream": (None, [C.c_void_p, _u64p]),
        "orc_set_stream_sharing": (None, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_char_p]),
        "orc_filter": (C.c_size_t, [C.c_void_p, C.c_uint64, _f64p, _szp, C.c_size_t, _f64p,
                                    _u64p, C.c_size_t, _f64p, _f64p, C.c_size_t]),
        "orc_model_update": (None, [C.c_int, C.c_uint64, _f64p, _u64p, C.c_int, C.c_int, _f64p,
                                    _f64p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _p(a, typ):
    return None if a is None else a.ctypes.data_as(typ)


# --------------------------------------------------------------------------
# RNG helpers
# --------------------------------------------------------------------------
def seed_state(seed: int) -> np.ndarray:
    s = np.zeros(4, dtype=np.uint64)
    lib().orc_seed(C.c_uint64(seed), _p(s, _u64p))
    return s


def rng_streams(seed, n_streams: int) -> np.ndarray:
    """(n_streams, 4) uint64; ``seed`` is an int or an array of 4*k words."""
    base = seed_state(seed) if np.isscalar(seed) else np.ascontiguousarray(seed, dtype=np.uint64).ravel()
    out = np.zeros((n_streams, 4), dtype=np.uint64)
    lib().orc_rng_streams(_p(base, _u64p), base.size // 4, n_streams, _p(out, _u64p))
    return out


def rng_distributed_state(seed, n_streams: int, n_nodes: int) -> np.ndarray:
    base = seed_state(seed) if np.isscalar(seed) else np.ascontiguousarray(seed, dtype=np.uint64).ravel()[:4].copy()
    out = np.zeros((n_nodes, n_streams, 4), dtype=np.uint64)
    lib().orc_rng_distributed_state(_p(base, _u64p), n_streams, n_nodes, _p(out, _u64p))
    return out


def next_words(state: np.ndarray, n: int, scrambler: int = PLUS) -> np.ndarray:
    out = np.zeros(n, dtype=np.uint64)
    L = lib()
    for i in range(n):
        out[i] = L.orc_next(_p(state, _u64p), scrambler)
    return out


def scale_log_weights_r(logw):
    lw = np.ascontiguousarray(logw, dtype=np.float64)
    w = np.zeros_like(lw)
    avg = lib().orc_scale_log_weights_r(_p(lw, _f64p), lw.size, _p(w, _f64p))
    return w, avg


def scale_log_weights(logw, mode=EXACT):
    w = np.array(logw, dtype=np.float64)
    avg = lib().orc_scale_log_weights(_p(w, _f64p), w.size, mode)
    return w, avg


def particle_resample_r(weights, u0):
    w = np.ascontiguousarray(weights, dtype=np.float64)
    k = np.zeros(w.size, dtype=np.int32)
    lib().orc_particle_resample_r(_p(w, _f64p), w.size, u0, _p(k, _i32p))
    return k


def resample_weight(weights, u, mode=EXACT):
    w = np.ascontiguousarray(weights, dtype=np.float64)
    k = np.zeros(w.size, dtype=np.int32)
    lib().orc_resample_weight(_p(w, _f64p), w.size, u, mode, _p(k, _i32p))
    return k


def default_pars(model=SIR) -> np.ndarray:
    p = np.zeros(lib().orc_model_n_par(model), dtype=np.float64)
    lib().orc_model_default_pars(model, _p(p, _f64p))
    return p


class OracleModel:
    """The dust model-object protocol on the CPU oracle."""

    def __init__(self, pars=None, time=0, n_particles=1, n_pars=0, seed=1, model=SIR,
                 scrambler=PLUS, deterministic=False, n_threads=1, resample_mode=EXACT):
        L = lib()
        self.model = model
        self.n_state = L.orc_model_n_state(model)
        self.n_par = L.orc_model_n_par(model)
        self.n_particles = n_particles
        self.n_pars = n_pars
        self.n_sets = max(1, n_pars)
        self.n_total = n_particles * self.n_sets
        if pars is None:
            pars = np.tile(default_pars(model), (self.n_sets, 1))
        pars = np.ascontiguousarray(pars, dtype=np.float64).reshape(self.n_sets, self.n_par)
        base = seed_state(seed) if np.isscalar(seed) else np.ascontiguousarray(seed, dtype=np.uint64).ravel()
        self._h = L.orc_alloc(model, _p(pars, _f64p), n_particles, n_pars, time,
                              _p(base, _u64p), base.size // 4, scrambler, int(deterministic),
                              n_threads)
        L.orc_set_resample_mode(self._h, resample_mode)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_free(self._h)
            self._h = None

    def time(self):
        return lib().orc_time(self._h)

    def run(self, time_end):
        lib().orc_run(self._h, time_end)
        return self.state()

    def state(self, index=None):
        if index is None:
            out = np.zeros((self.n_total, self.n_state))
            lib().orc_state(self._h, None, 0, _p(out, _f64p))
        else:
            idx = np.ascontiguousarray(index, dtype=np.uintp)
            out = np.zeros((self.n_total, idx.size))
            lib().orc_state(self._h, _p(idx, _szp), idx.size, _p(out, _f64p))
        return out.T  # (n_state, n_total), R layout

    def update_state(self, pars=None, state=None, time=None, set_initial_state=True):
        p = None
        if pars is not None:
            p = np.ascontiguousarray(pars, dtype=np.float64).reshape(self.n_sets, self.n_par)
        s, slen = None, 0
        if state is not None:
            s = np.asarray(state, dtype=np.float64)
            s = np.ascontiguousarray(s.T if s.ndim == 2 else s).ravel()
            slen = s.size
        lib().orc_update_state(self._h, _p(p, _f64p), _p(s, _f64p), slen,
                               int(time is not None), 0 if time is None else time,
                               int(set_initial_state))

    def reorder(self, kappa0):
        k = np.ascontiguousarray(kappa0, dtype=np.int32)
        lib().orc_reorder(self._h, _p(k, _i32p))

    def rng_state(self):
        out = np.zeros((self.n_total + 1, 4), dtype=np.uint64)
        lib().orc_rng_state(self._h, _p(out, _u64p))
        return out

    def set_rng_state(self, s):
        s = np.ascontiguousarray(s, dtype=np.uint64)
        assert s.size == (self.n_total + 1) * 4
        lib().orc_set_rng_state(self._h, _p(s, _u64p))

    def set_data(self, time_end, values, shared=False):
        t = np.ascontiguousarray(time_end, dtype=np.uint64)
        v = np.ascontiguousarray(values, dtype=np.float64)
        lib().orc_set_data(self._h, t.size, _p(t, _u64p), _p(v, _f64p), int(shared))
        self._n_data = t.size
        self._data_time = t.copy()

    def set_filter_stream(self, state):
        s = np.ascontiguousarray(state, dtype=np.uint64)
        lib().orc_set_filter_stream(self._h, _p(s, _u64p))

    def set_stream_sharing(self, n_sets_global, set_offset, na_global):
        na = np.ascontiguousarray(na_global, dtype=np.uint8)
        lib().orc_set_stream_sharing(self._h, int(n_sets_global), int(set_offset), na.tobytes())

    def compare_data(self):
        w = np.zeros(self.n_total)
        ok = lib().orc_compare_data(self._h, _p(w, _f64p))
        return w if ok else None

    def resample(self, weights):
        w = np.ascontiguousarray(weights, dtype=np.float64)
        k = np.zeros(self.n_total, dtype=np.int32)
        lib().orc_resample(self._h, _p(w, _f64p), _p(k, _i32p))
        return k

    def set_n_threads(self, n):
        lib().orc_set_n_threads(self._h, n)

    def filter(self, time_end=None, save_trajectories=False, index=None, snapshot_times=None,
               min_log_likelihood=None):
        """Returns dict(log_likelihood, trajectories, snapshots, n_intervals).

        trajectories: (n_index, n_total, T+1); snapshots: (n_state, n_total, n_snap).
        """
        if time_end is None:
            time_end = int(self._data_time[-1])
        pending = int(np.sum((self._data_time > self.time()) & (self._data_time <= time_end)))
        ll = np.zeros(self.n_sets)
        idx = None if index is None else np.ascontiguousarray(index, dtype=np.uintp)
        ni = self.n_state if idx is None else idx.size
        traj = np.zeros((pending + 1, self.n_total, ni)) if save_trajectories else None
        snap_t = None if snapshot_times is None else np.ascontiguousarray(snapshot_times, dtype=np.uint64)
        n_snap = 0 if snap_t is None else snap_t.size
        snaps = np.full((n_snap, self.n_total, self.n_state), np.nan) if n_snap else None
        mll = None if min_log_likelihood is None else np.atleast_1d(
            np.asarray(min_log_likelihood, dtype=np.float64))
        done = lib().orc_filter(self._h, time_end, _p(ll, _f64p), _p(idx, _szp), ni,
                                _p(traj, _f64p), _p(snap_t, _u64p), n_snap, _p(snaps, _f64p),
                                _p(mll, _f64p), 0 if mll is None else mll.size)
        return {
            "log_likelihood": ll,
            "trajectories": None if traj is None else traj.transpose(2, 1, 0),
            "snapshots": None if snaps is None else snaps.transpose(2, 1, 0),
            "n_intervals": done,
        }
