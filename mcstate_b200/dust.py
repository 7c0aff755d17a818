"""The `dust_generator` model-object protocol on the B200 engine.

mcstate talks to its native engine only through this object:
``generator$new(...)`` then ``$run / $filter / $state / $reorder /
$update_state / $set_index / $set_data / $rng_state / ...``.  This module is the
Python twin of that R6 object: every method forwards to one C-ABI entry point
of ``libmcstate_b200.so`` (include/mcstate_b200.h), which runs CUDA kernels.
R is not available in the build image, so the host side is Python; the
equivalent R6/.Call binding is in INTEGRATION.md.

Arrays use R's dimension order: a state matrix is ``(n_state, n_particles)``
or ``(n_state, n_particles, n_pars)``; indices passed to ``set_index`` /
``state`` / ``reorder`` are 1-based, as R passes them.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import check, f64p, i32p, ptr, szp, u64p

_ALGORITHMS = {"xoshiro256plus": _lib.SCRAMBLER_PLUS, "xoshiro256starstar": _lib.SCRAMBLER_STARSTAR}


# --------------------------------------------------------------------------
# dust::dust_rng and friends (host-side RNG-state arithmetic)
# --------------------------------------------------------------------------
def _seed_words(seed) -> np.ndarray:
    """NULL / positive integer / raw vector of 32k bytes -> uint64 words
    (R/particle_filter.R:199-204)."""
    if seed is None:
        seed = int(np.random.randint(1, 2 ** 31 - 1))
    if isinstance(seed, (int, np.integer)):
        if seed < 0:
            raise ValueError("'seed' must be a positive integer")
        s = np.zeros(4, dtype=np.uint64)
        check(_lib.lib().mcb_rng_seed(C.c_uint64(int(seed)), ptr(s, u64p)))
        return s
    if isinstance(seed, (bytes, bytearray)):
        raw = np.frombuffer(bytes(seed), dtype=np.uint8)
    else:
        raw = np.asarray(seed)
        if raw.dtype == np.uint64:
            raw = np.ascontiguousarray(raw).view(np.uint8)
        raw = raw.astype(np.uint8, copy=False)
    if raw.size == 0 or raw.size % 32 != 0:
        raise ValueError("'seed' as a raw vector must have a length that is a multiple of 32")
    return np.frombuffer(raw.tobytes(), dtype=np.uint64).copy()


class dust_rng:
    """``dust::dust_rng$new(seed)``: one host-side stream
    (R/pmcmc_parallel.R:147-148 uses ``$long_jump()$random_real(1)``)."""

    def __init__(self, seed=None, algorithm="xoshiro256plus"):
        self._s = _seed_words(seed)[:4].copy()
        self._alg = _ALGORITHMS[algorithm]

    def jump(self, n=1):
        """``$jump()``; ``n`` jumps at once lands on stream n of this seed."""
        if n == 1:
            check(_lib.lib().mcb_rng_jump(ptr(self._s, u64p)))
        else:
            check(_lib.lib().mcb_rng_jump_n(ptr(self._s, u64p), int(n)))
        return self

    def long_jump(self):
        check(_lib.lib().mcb_rng_long_jump(ptr(self._s, u64p)))
        return self

    def random_real(self, n):
        out = np.zeros(n)
        check(_lib.lib().mcb_rng_random_real(ptr(self._s, u64p), self._alg, n, ptr(out, f64p)))
        return out

    def state(self) -> bytes:
        return self._s.tobytes()


class dust_rng_pointer:
    """``dust::dust_rng_pointer$new(seed, n_streams)$state()``
    (tests/testthat/test-particle-filter.R:1316): the raw state of ``n_streams`` streams,
    stream i = stream 0 jumped i times -- what a model with that seed holds."""

    def __init__(self, seed, n_streams=1, algorithm="xoshiro256plus"):
        first = dust_rng(seed, algorithm)
        out, r = [], first
        for i in range(int(n_streams)):
            out.append(bytes(r.state()))
            if i + 1 < int(n_streams):
                r = dust_rng(out[-1], algorithm).jump()
        self._state = b"".join(out)

    def state(self):
        return self._state


def dust_rng_distributed_state(seed=None, n_streams=1, n_nodes=1, algorithm="xoshiro256plus"):
    """``dust::dust_rng_distributed_state`` (R/pmcmc_parallel.R:133, R/smc2.R:68):
    a list of raw states, node k long-jumped k times from the seed."""
    base = _seed_words(seed)[:4].copy()
    out = np.zeros((n_nodes, n_streams, 4), dtype=np.uint64)
    check(_lib.lib().mcb_rng_distributed_state(ptr(base, u64p), n_streams, n_nodes, ptr(out, u64p)))
    return [out[k].tobytes() for k in range(n_nodes)]


def dust_data(data, name_time="time_end", multi=None):
    """``dust::dust_data`` (R/particle_filter_data.R:197-198): one entry per time,
    ``(time_end, [row, ...])`` with one row (a dict) per population."""
    cols = {k: np.asarray(v) for k, v in data.items()}
    n = len(cols[name_time])
    rows = [{k: (v[i].item() if hasattr(v[i], "item") else v[i]) for k, v in cols.items()}
            for i in range(n)]
    if multi is None:
        return [(int(r[name_time]), [r]) for r in rows]
    levels = sorted(set(cols[multi].tolist()))
    groups = [[r for r in rows if r[multi] == lv] for lv in levels]
    if len({len(g) for g in groups}) != 1:
        raise ValueError("All groups must have the same number of rows")
    out = []
    for i in range(len(groups[0])):
        t = {int(g[i][name_time]) for g in groups}
        if len(t) != 1:
            raise ValueError("All groups must have the same time")
        out.append((t.pop(), [g[i] for g in groups]))
    return out


# --------------------------------------------------------------------------
# the model object
# --------------------------------------------------------------------------
class dust_model:
    """One instantiated model: particle state, RNG streams and data in HBM."""

    def __init__(self, generator, pars, time, n_particles, n_threads=1, seed=None,
                 pars_multi=False, deterministic=False, gpu_config=None, seed_per_set=False):
        L = self._L = generator._library()   # the library that holds this generator's model
        self._ck = generator._check
        self._gen = generator
        self._n_threads = int(n_threads)
        self._pars_multi = bool(pars_multi)
        if pars_multi:
            if isinstance(pars, dict) or pars is None:
                raise ValueError("Expected an unnamed list of parameters for 'pars'")
            self._pars = [dict(p) for p in pars]
            n_pars = len(self._pars)
            if n_particles is None:  # R/if2.R:122: one particle per parameter set
                n_particles = 1
        else:
            self._pars = dict(pars or {})
            n_pars = 0
        if n_particles is None or int(n_particles) < 1:
            raise ValueError("'n_particles' must be at least 1")
        device_id = 0
        if gpu_config is not None:
            device_id = int(gpu_config["device_id"]) if isinstance(gpu_config, dict) else int(gpu_config)
        words = _seed_words(seed)
        vec = generator._pars_vector(self._pars, pars_multi)
        h = C.c_void_p()
        # seed_per_set: one 32-byte seed per parameter set; each set then owns the streams a
        # separate model object seeded with it would (independent filters in one handle)
        self._n_fs = max(1, n_pars) if seed_per_set else 1
        self._ck(L.mcb_alloc_ex(generator.model_id, ptr(vec, f64p), int(n_particles), n_pars,
                             int(time), ptr(words, u64p), words.size // 4,
                             _lib.SEED_PER_SET if seed_per_set else _lib.SEED_SHARED,
                             generator.scrambler, int(deterministic), device_id, C.byref(h)))
        self._h = h
        self._device = device_id
        self._n_state = generator.n_state
        self._n_particles = int(n_particles)
        self._n_pars = n_pars
        self._n_sets = max(1, n_pars)
        self._n_total = self._n_particles * self._n_sets
        self._n_index = self._n_state
        self._index_names = None
        self._data_times = None

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                self._L.mcb_free(h)
            except Exception:
                pass
            self._h = None

    # ---- shape helpers
    def _shape_out(self, buf, n_rows):
        """C buffer (n_total, n_rows) -> (n_rows, n_particles[, n_pars])."""
        if self._pars_multi:
            return buf.reshape(self._n_sets, self._n_particles, n_rows).transpose(2, 1, 0)
        return buf.reshape(self._n_particles, n_rows).T

    def _name_rows(self, arr):
        return arr

    # ---- protocol (each method cites its call site in mcstate)
    def n_state(self):
        return self._n_state

    def n_particles(self):
        return self._n_particles

    def n_particles_each(self):
        return self._n_particles

    def n_pars(self):
        """R/particle_filter_state.R:497: 0 unless pars_multi."""
        return self._n_pars

    def shape(self):
        """R/particle_filter_state.R:268: N or c(N, P)."""
        return (self._n_particles, self._n_pars) if self._pars_multi else (self._n_particles,)

    def time(self):
        t = C.c_uint64()
        self._ck(self._L.mcb_time(self._h, C.byref(t)))
        return int(t.value)

    def pars(self):
        return self._pars

    def info(self):
        """R/particle_filter_state.R:288: model-defined list (list of lists if multi)."""
        return [self._gen._info(p) for p in self._pars] if self._pars_multi else self._gen._info(self._pars)

    def n_threads(self):
        return self._n_threads

    def set_n_threads(self, n_threads):
        """Accepted for protocol compatibility; parallelism is the GPU's."""
        prev, self._n_threads = self._n_threads, int(n_threads)
        return prev

    def has_openmp(self):
        return False

    def uses_gpu(self, fake_gpu=False):
        """tests/testthat/test-particle-filter.R:1215."""
        return True

    def set_index(self, index):
        """R/particle_filter_state.R:261,263 (1-based; dict keeps row names)."""
        if index is None:
            self._ck(self._L.mcb_set_index(self._h, None, 0))
            self._n_index, self._index_names = self._n_state, None
            return
        if isinstance(index, dict):
            self._index_names = list(index.keys())
            index = list(index.values())
        else:
            self._index_names = None
        idx = np.ascontiguousarray(np.atleast_1d(index), dtype=np.int64) - 1
        if np.any(idx < 0) or np.any(idx >= self._n_state):
            raise ValueError("All elements of 'index' must lie in [1, %d]" % self._n_state)
        idx = idx.astype(np.uintp)
        self._ck(self._L.mcb_set_index(self._h, ptr(idx, szp), idx.size))
        self._n_index = int(idx.size)

    def run(self, time_end):
        """R/particle_filter_state.R:65."""
        out = np.zeros((self._n_total, self._n_index))
        self._ck(self._L.mcb_run(self._h, int(time_end), ptr(out, f64p)))
        return self._shape_out(out, self._n_index)

    def simulate(self, time_end):
        """R/predict.R:79: (n_index, n_particles[, n_pars], n_times)."""
        t = np.ascontiguousarray(np.atleast_1d(time_end), dtype=np.uint64)
        out = np.zeros((t.size, self._n_total, self._n_index))
        self._ck(self._L.mcb_simulate(self._h, ptr(t, u64p), t.size, ptr(out, f64p)))
        if self._pars_multi:
            return out.reshape(t.size, self._n_sets, self._n_particles, self._n_index).transpose(3, 2, 1, 0)
        return out.transpose(2, 1, 0)

    def state(self, index=None):
        """R/particle_filter_state.R:79,81,114,272."""
        if index is None:
            out = np.zeros((self._n_total, self._n_state))
            self._ck(self._L.mcb_state(self._h, None, 0, ptr(out, f64p)))
            return self._shape_out(out, self._n_state)
        if isinstance(index, dict):
            index = list(index.values())
        idx = np.ascontiguousarray(np.atleast_1d(index), dtype=np.int64) - 1
        if np.any(idx < 0) or np.any(idx >= self._n_state):
            raise ValueError("All elements of 'index' must lie in [1, %d]" % self._n_state)
        idx = idx.astype(np.uintp)
        out = np.zeros((self._n_total, idx.size))
        self._ck(self._L.mcb_state(self._h, ptr(idx, szp), idx.size, ptr(out, f64p)))
        return self._shape_out(out, idx.size)

    def state_particles(self, index_particle):
        """``model$state()[, index_particle]`` without moving the other columns:
        ``[n_state, n]`` (single parameter set; 1-based indices as in R)."""
        ip = np.atleast_1d(np.asarray(index_particle, dtype=np.int64))
        if ip.min() < 1 or ip.max() > self._n_total:
            raise IndexError("index_particle out of range")
        flat = np.ascontiguousarray(ip - 1, dtype=np.uintp)
        out = np.zeros((flat.size, self._n_state))
        self._ck(self._L.mcb_state_particles(self._h, ptr(flat, szp), flat.size, ptr(out, f64p)))
        return out.T

    def update_state(self, pars=None, state=None, time=None, set_initial_state=None):
        """R/particle_filter_state.R:248,253,512; R/smc2.R:141; R/if2.R:160-161."""
        vec = None
        if pars is not None:
            if self._pars_multi:
                if isinstance(pars, dict) or len(pars) != self._n_sets:
                    raise ValueError("'pars' must be an unnamed list of length %d" % self._n_sets)
                self._pars = [dict(p) for p in pars]
            else:
                self._pars = dict(pars)
            vec = self._gen._pars_vector(self._pars, self._pars_multi)
        st, st_len = None, 0
        if state is not None:
            s = np.asarray(state, dtype=np.float64)
            if s.ndim == 1:
                if s.size != self._n_state:
                    raise ValueError("Expected a vector of length %d for 'state'" % self._n_state)
                st = np.ascontiguousarray(s)
            else:
                want = (self._n_state, self._n_particles) + ((self._n_sets,) if self._pars_multi else ())
                if s.shape != want:
                    raise ValueError("Expected 'state' with dimensions %s" % (want,))
                st = np.ascontiguousarray(np.transpose(s)).ravel()  # (pars,) particle, state
            st_len = st.size
        if set_initial_state is None:
            set_initial_state = state is None
        self._ck(self._L.mcb_update_state(self._h, ptr(vec, f64p), ptr(st, f64p), st_len,
                                          int(time is not None), 0 if time is None else int(time),
                                          int(bool(set_initial_state))))

    def reorder(self, index):
        """R/particle_filter_state.R:100: 1-based, [N] or [N, P] within-block."""
        k = np.asarray(index)
        if self._pars_multi:
            if k.shape != (self._n_particles, self._n_sets):
                raise ValueError("Expected a matrix with %d rows and %d columns for 'index'"
                                 % (self._n_particles, self._n_sets))
            k = k.T
        elif k.size != self._n_particles:
            raise ValueError("Expected a vector of length %d for 'index'" % self._n_particles)
        k = np.ascontiguousarray(k, dtype=np.int32).ravel()
        self._ck(self._L.mcb_reorder(self._h, ptr(k, i32p)))

    def rng_state(self, first_only=False, last_only=False):
        """R/particle_filter.R:813; R/particle_filter_state.R:363,406,420: raw bytes."""
        out = np.zeros((self._n_total + self._n_fs) * 4, dtype=np.uint64)
        self._ck(self._L.mcb_rng_state(self._h, 0, ptr(out, u64p)))
        if first_only:
            return out[:4].tobytes()
        if last_only:  # the filter stream(s)
            return out[-4 * self._n_fs:].tobytes()
        return out.tobytes()

    def set_rng_state(self, rng_state):
        """R/particle_filter_state.R:369,420; R/particle_filter.R:877."""
        raw = np.frombuffer(bytes(rng_state), dtype=np.uint64).copy()
        if raw.size != (self._n_total + self._n_fs) * 4:
            raise ValueError("'rng_state' must be a raw vector of length %d (but was %d)"
                             % ((self._n_total + self._n_fs) * 32, raw.size * 8))
        self._ck(self._L.mcb_set_rng_state(self._h, ptr(raw, u64p), raw.size))

    def set_data(self, data, shared=False):
        """R/particle_filter_state.R:243-246: ``data`` = dust_data(...)."""
        n_sets_data = 1 if shared or not self._pars_multi else self._n_sets
        fields = self._gen.data_fields
        times = np.zeros(len(data), dtype=np.uint64)
        vals = np.full((len(data), n_sets_data, len(fields)), np.nan)
        for i, (t, rows) in enumerate(data):
            times[i] = t
            if len(rows) != n_sets_data:
                raise ValueError("Expected each element of data to have %d rows" % n_sets_data)
            for j, row in enumerate(rows):
                for f, name in enumerate(fields):
                    if name not in row:
                        raise KeyError("data is missing the field '%s'" % name)
                    x = row[name]
                    vals[i, j, f] = np.nan if x is None else float(x)
        self._ck(self._L.mcb_set_data(self._h, times.size, ptr(times, u64p), ptr(vals, f64p),
                                      int(bool(shared))))
        self._data_times = times

    # ---- sharding one multi-parameter model over several handles (include/mcstate_b200.h)
    def set_filter_stream(self, raw):
        """Install the (shared) filter stream: 32 raw bytes."""
        w = np.frombuffer(bytes(raw), dtype=np.uint64).copy()
        if w.size != 4:
            raise ValueError("the filter stream is 32 bytes")
        self._ck(self._L.mcb_set_filter_stream(self._h, ptr(w, u64p)))

    def set_stream_sharing(self, n_sets_global, set_offset, na_global):
        """This handle holds sets [set_offset, set_offset + n_pars) of ``n_sets_global``;
        ``na_global`` [n_data, n_sets_global] marks the data rows that are NA."""
        na = np.ascontiguousarray(na_global, dtype=np.uint8)
        if na.shape != (len(self._data_times), int(n_sets_global)):
            raise ValueError("'na_global' must be [n_data, n_sets_global]")
        self._ck(self._L.mcb_set_stream_sharing(self._h, int(n_sets_global), int(set_offset),
                                                na.tobytes()))

    def copy_sets_from(self, other, take, state=True, rng=True):
        """Take the particle state and/or RNG streams of the flagged parameter sets from
        ``other`` (same shape; smc2 replenishment, R/smc2.R:174-189)."""
        t = np.ascontiguousarray(take, dtype=np.uint8)
        if t.size != self._n_sets:
            raise ValueError("'take' must have one flag per parameter set")
        if other._L is not self._L:
            raise ValueError("copy_sets_from needs two objects of the same model")
        self._ck(self._L.mcb_copy_sets(self._h, other._h, t.tobytes(), int(state), int(rng)))

    def compare_data(self):
        """``$compare_data()``: log-weights for the datum at the current time, or None."""
        w = np.zeros(self._n_total)
        has = C.c_int(0)
        self._ck(self._L.mcb_compare_data(self._h, ptr(w, f64p), C.byref(has)))
        if not has.value:
            return None
        return w.reshape(self._n_sets, self._n_particles).T if self._pars_multi else w

    def resample(self, weights):
        """``$resample(weights)`` -> kappa (1-based, within block); reorders the state."""
        w = np.asarray(weights, dtype=np.float64)
        w = np.ascontiguousarray(w.T if w.ndim == 2 else w).ravel()
        k = np.zeros(self._n_total, dtype=np.int32)
        self._ck(self._L.mcb_resample(self._h, ptr(w, f64p), ptr(k, i32p)))
        return k.reshape(self._n_sets, self._n_particles).T if self._pars_multi else k

    def filter(self, time_end=None, save_trajectories=False, time_snapshot=None,
               min_log_likelihood=None):
        """R/particle_filter_state.R:151 -- the hot path, one native call.

        Returns ``dict(log_likelihood, trajectories, snapshots)``.
        """
        if self._data_times is None:
            raise _lib.McbError(_lib.ERR_STATE, "Data has not been set for this object")
        if time_end is None:
            time_end = int(self._data_times[-1])
        L = self._L
        n_rows = C.c_size_t()
        self._ck(L.mcb_filter_rows(self._h, int(time_end), C.byref(n_rows)))
        ll = np.zeros(self._n_sets)
        traj = None
        on_device = isinstance(save_trajectories, str)
        if on_device and save_trajectories != "device":
            raise ValueError("save_trajectories must be a logical or 'device'")
        if save_trajectories and not on_device:
            traj = np.zeros((n_rows.value + 1, self._n_total, self._n_index))
        self._hist_rows = n_rows.value if save_trajectories else None
        snap_t, snaps = None, None
        if time_snapshot is not None and len(time_snapshot) > 0:
            snap_t = np.ascontiguousarray(time_snapshot, dtype=np.uint64)
            snaps = np.zeros((snap_t.size, self._n_total, self._n_state))
        mll = None
        if min_log_likelihood is not None:
            mll = np.ascontiguousarray(np.atleast_1d(min_log_likelihood), dtype=np.float64)
        self._ck(L.mcb_filter(self._h, int(time_end), 2 if on_device else int(bool(save_trajectories)),
                           ptr(snap_t, u64p), 0 if snap_t is None else snap_t.size,
                           ptr(mll, f64p), 0 if mll is None else mll.size,
                           ptr(ll, f64p), ptr(traj, f64p), ptr(snaps, f64p)))
        if traj is not None:
            if self._pars_multi:
                traj = traj.reshape(-1, self._n_sets, self._n_particles, self._n_index).transpose(3, 2, 1, 0)
            else:
                traj = traj.transpose(2, 1, 0)
        if snaps is not None:
            if self._pars_multi:
                snaps = snaps.reshape(-1, self._n_sets, self._n_particles, self._n_state).transpose(3, 2, 1, 0)
            else:
                snaps = snaps.transpose(2, 1, 0)
        return {"log_likelihood": ll, "trajectories": traj, "snapshots": snaps}

    def filter_history(self, index_particle):
        """Trajectories of chosen final particles of the last
        ``filter(..., save_trajectories="device")`` (or True) call, traced on the device:
        ``[n_index, n, T+1]``; with several parameter sets ``index_particle`` is
        ``[n, n_pars]`` (1-based within each set) and the result ``[n_index, n, n_pars, T+1]``.
        The 1-based indices are R's (R/particle_filter.R:616-646 history_single /
        :653-714 history_multiple with an index)."""
        if getattr(self, "_hist_rows", None) is None:
            raise _lib.McbError(_lib.ERR_STATE, "filter was not run with save_trajectories")
        ip = np.asarray(index_particle, dtype=np.int64)
        if self._pars_multi:
            if ip.ndim < 2:
                ip = np.tile(np.atleast_1d(ip)[:, None], (1, self._n_sets))
            if ip.shape[1] != self._n_sets:
                raise ValueError("'index_particle' should have %d columns" % self._n_sets)
            if ip.min() < 1 or ip.max() > self._n_particles:
                raise IndexError("index_particle out of range")
            glob = (ip - 1) + np.arange(self._n_sets)[None, :] * self._n_particles
        else:
            ip = np.atleast_1d(ip)
            if ip.min() < 1 or ip.max() > self._n_particles:
                raise IndexError("index_particle out of range")
            glob = ip - 1
        flat = np.ascontiguousarray(glob.ravel(), dtype=np.uintp)
        out = np.zeros((self._hist_rows + 1, flat.size, self._n_index))
        self._ck(self._L.mcb_filter_history(self._h, ptr(flat, szp), flat.size, ptr(out, f64p)))
        if self._pars_multi:
            return out.reshape(-1, glob.shape[0], self._n_sets, self._n_index).transpose(3, 1, 2, 0)
        return out.transpose(2, 1, 0)

    def set_persistent(self, enable=True):
        """Small single-set filters run as one cooperative kernel per call
        (csrc/persistent.cuh); ``False`` keeps this model on the per-interval path."""
        self._ck(self._L.mcb_set_persistent(self._h, int(bool(enable))))

    # ---- asynchronous variant used to keep several GPUs / handles busy
    def filter_enqueue(self, time_end=None):
        if time_end is None:
            time_end = int(self._data_times[-1])
        self._ck(self._L.mcb_filter_enqueue(self._h, int(time_end)))

    def filter_collect(self):
        ll = np.zeros(self._n_sets)
        self._ck(self._L.mcb_filter_collect(self._h, ptr(ll, f64p)))
        return ll

    def log_likelihood_device(self):
        """The log-likelihood vector [n_pars] an enqueued filter writes, as a torch tensor
        over the handle's own device memory (no copy): what a sharded caller hands to its
        all-gather, in stream order after ``filter_enqueue`` (mcb_log_likelihood_device)."""
        import torch

        p = C.c_void_p()
        self._ck(self._L.mcb_log_likelihood_device(self._h, C.byref(p)))

        class _View:   # the CUDA array interface: torch wraps the memory, nothing is copied
            __cuda_array_interface__ = {"shape": (self._n_sets,), "typestr": "<f8",
                                        "data": (int(p.value), False), "version": 3}
        return torch.as_tensor(_View(), device=torch.device("cuda", self._device))

    def set_stream(self, cuda_stream):
        """Launch on this ``cudaStream_t`` (an integer handle, e.g. ``torch.cuda.Stream().
        cuda_stream``); ``None`` or 0 goes back to the stream the handle owns.  0 is also the
        handle of CUDA's default stream -- which cannot be used here (no graph capture on
        it): give the handle a stream of its own and order other work on that."""
        self._ck(self._L.mcb_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    def synchronize(self):
        self._ck(self._L.mcb_synchronize(self._h))

    def set_kernel_timing(self, enable=True):
        self._ck(self._L.mcb_set_kernel_timing(self._h, int(bool(enable))))

    def kernel_timing(self):
        """(ms[4], launches[4]) for run+compare, weights scan, ancestors, the final gather."""
        ms = np.zeros(4)
        n = np.zeros(4, dtype=np.uint64)
        self._ck(self._L.mcb_kernel_timing(self._h, ptr(ms, f64p), ptr(n, u64p)))
        return ms, n

    def launch_count(self):
        n = C.c_uint64()
        self._ck(self._L.mcb_launch_count(self._h, C.byref(n)))
        return int(n.value)


class dust_generator:
    """What ``dust::dust_example(name)`` returns: the generator with its static
    capability queries (R/particle_filter.R:590-593, 991, 241, 1029, 212)."""

    name = "dust_generator"

    def __init__(self, model_name, algorithm="xoshiro256plus", library=None, source=None):
        # `library`: a build of the engine that holds a user-supplied model (dust() below,
        # `source` = the model's header); the stock library holds the four dust_example models
        L = self._L = _lib.lib() if library is None else library
        self._source = source
        check = self._check
        mid, ns, npar, nd = C.c_int(), C.c_size_t(), C.c_size_t(), C.c_size_t()
        check(L.mcb_model_lookup(model_name.encode(), C.byref(mid), C.byref(ns), C.byref(npar),
                                 C.byref(nd)))
        if algorithm not in _ALGORITHMS:
            raise ValueError("Unknown algorithm '%s'" % algorithm)
        self.model_name = model_name
        self.model_id = mid.value
        self.n_state = ns.value
        self.n_par = npar.value
        self.algorithm = algorithm
        self.scrambler = _ALGORITHMS[algorithm]
        self.par_names = [L.mcb_model_par_name(mid.value, i).decode() for i in range(npar.value)]
        self.par_defaults = [L.mcb_model_par_default(mid.value, i) for i in range(npar.value)]
        self.state_names = [L.mcb_model_state_name(mid.value, i).decode() for i in range(ns.value)]
        self.data_fields = [L.mcb_model_data_name(mid.value, i).decode() for i in range(nd.value)]
        self.public_methods = self  # R: model$public_methods$has_compare()

    # A generator travels between processes (pmcmc_chains_prepare / _run, the samples'
    # `predict` element) without its library handle; the receiving side binds it again
    # -- dust::dust_repair_environment (R/pmcmc_chains.R:78).
    def __getstate__(self):
        d = dict(self.__dict__)
        d["_L"] = None
        d["public_methods"] = None
        return d

    def __setstate__(self, d):
        self.__dict__.update(d)
        self.public_methods = self
        try:
            dust_repair_environment(self)
        except Exception:
            # no engine here (e.g. results collected on a machine without the library):
            # the description survives, the library is bound when something needs it
            self._L = None

    def _library(self):
        if self._L is None:
            dust_repair_environment(self)
        return self._L

    def _check(self, rc):
        if rc != _lib.OK:
            msg = self._library().mcb_last_error()
            raise _lib.McbError(rc, msg.decode() if msg else "mcstate_b200 error %d" % rc)

    # -- static capability queries
    def has_compare(self):
        return len(self.data_fields) > 0      # "walk" has no compiled compare

    def has_gpu_support(self, fake_gpu=False):
        return True

    def has_openmp(self):
        return False

    def time_type(self):
        return "discrete"

    def rng_algorithm(self):
        return self.algorithm

    def gpu_info(self):
        L, check = self._library(), self._check
        n = C.c_int(0)
        rc = L.mcb_device_count(C.byref(n))
        devices = []
        if rc == _lib.OK:
            for d in range(n.value):
                name = C.create_string_buffer(256)
                major, minor, sm, mem = C.c_int(), C.c_int(), C.c_int(), C.c_size_t()
                check(L.mcb_device_info(d, name, 256, C.byref(major), C.byref(minor), C.byref(sm),
                                        C.byref(mem)))
                devices.append({"id": d, "name": name.value.decode(), "memory": mem.value,
                                "version": major.value * 10 + minor.value, "n_sm": sm.value})
        return {"has_cuda": n.value > 0, "devices": devices}

    # -- instantiate
    def new(self, pars, time, n_particles, n_threads=1, seed=None, pars_multi=False,
            deterministic=False, gpu_config=None, seed_per_set=False, **ignored):
        """R/particle_filter_state.R:237-241."""
        return dust_model(self, pars, time, n_particles, n_threads=n_threads, seed=seed,
                          pars_multi=pars_multi, deterministic=deterministic,
                          gpu_config=gpu_config, seed_per_set=seed_per_set)

    # -- helpers
    def _pars_vector(self, pars, multi):
        sets = pars if multi else [pars]
        vec = np.zeros((len(sets), self.n_par))
        for j, p in enumerate(sets):
            for i, name in enumerate(self.par_names):
                x = p.get(name, self.par_defaults[i])  # extra keys ignored (R/particle_filter.R:283-284)
                vec[j, i] = float(x)
        return np.ascontiguousarray(vec)

    def _info(self, pars):
        vals = {n: float(pars.get(n, d)) for n, d in zip(self.par_names, self.par_defaults)}
        return {"vars": list(self.state_names), "pars": vals,
                "index": {n: i + 1 for i, n in enumerate(self.state_names)}}


def dust(path, quiet=True, algorithm="xoshiro256plus"):
    """``dust::dust(path)``: compile a user-supplied model and return its generator
    (tests/testthat/test-pmcmc-parallel.R:197; with a compiled compare as in
    tests/testthat/helper-mcstate.R:478-492 + compare_variable.cpp).

    ``path`` is a CUDA header defining ``mcb::UserModel`` and ``MCB_USER_MODEL_DESC`` by
    the contract of csrc/models.cuh (``mcstate_b200/examples/seir.cuh`` is one).  The
    engine -- every kernel of the filter, templated on the model -- is compiled around
    it for sm_100a into its own shared library with the same C ABI; the generator
    returned behaves like the ``dust_example`` ones everywhere (particle_filter, pmcmc,
    smc2, sharding).  No GPU is needed to build; none means no way to run."""
    L = _lib.user_lib(path, quiet)
    # a user build holds exactly one model (id 0), named after its source file
    name = os.path.splitext(os.path.basename(path))[0]
    if L.mcb_model_lookup(name.encode(), None, None, None, None) != _lib.OK:
        raise ValueError("'%s': MCB_USER_MODEL_DESC must name the model after its file "
                         "('%s')" % (path, name))
    return dust_generator(name, algorithm, library=L, source=os.path.abspath(path))


def dust_repair_environment(generator, quiet=True):
    """``dust::dust_repair_environment(generator)`` (R/pmcmc_chains.R:78): make a generator
    that was serialised in another process usable here -- bind the engine library again
    (for a user-supplied model: its own build, compiled here if this machine has not got
    it yet).  Harmless on a generator that is already bound, and on test doubles."""
    if not isinstance(generator, dust_generator) or generator._L is not None:
        return generator
    src = generator._source
    if src is not None and not os.path.exists(src):
        # a model shipped inside the package, serialised from a checkout somewhere else
        pkg = os.path.dirname(os.path.abspath(__file__))
        tail = src.replace(os.sep, "/").rsplit("/mcstate_b200/", 1)
        if len(tail) == 2 and os.path.exists(os.path.join(pkg, tail[1])):
            src = generator._source = os.path.join(pkg, tail[1])
    generator._L = _lib.lib() if src is None else _lib.user_lib(src, quiet)
    mid = C.c_int()
    generator._check(generator._L.mcb_model_lookup(generator.model_name.encode(), C.byref(mid),
                                                   None, None, None))
    generator.model_id = mid.value
    return generator


def dust_openmp_threads(n=None, action="error"):
    """``dust::dust_openmp_threads(n, action)`` (R/particle_filter.R:505): a thread count
    checked against what the host offers.  The engine's parallelism is the GPU's and
    ignores ``n_threads``; the count still reaches host-side callers (pmcmc_predict)."""
    limit = os.cpu_count() or 1
    if n is None:
        return limit
    n = int(n)
    if n < 1:
        raise ValueError("'n' must be at least 1")
    if n > limit:
        if action == "fix":
            return limit
        msg = "Requested number of threads '%d' exceeds a limit of '%d'" % (n, limit)
        if action == "error":
            raise ValueError(msg)
        if action in ("message", "warning"):
            import warnings

            warnings.warn(msg)
            return limit
        raise ValueError("'action' must be one of 'error', 'fix', 'message', 'warning'")
    return n


def dust_example(name, algorithm="xoshiro256plus"):
    """``dust::dust_example("sir")`` (tests/testthat/helper-mcstate.R:3)."""
    return dust_generator(name, algorithm)
