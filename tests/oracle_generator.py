"""A dust-protocol generator backed by the CPU oracle -- a TEST DOUBLE.

Lets the host-side logic of mcstate_b200.particle_filter run in the CPU-only test
suite (the reference tests its MCMC machinery the same way, with fake filters:
tests/testthat/helper-mcstate.R:222-386), and gives the GPU tests a second
implementation of the whole protocol to compare against.  Never imported by the
product.
"""
import numpy as np

from oracle import oracle as O

PAR_NAMES = {O.SIR: ["beta", "gamma", "I0", "exp_noise", "S0", "R0", "freq"],
             O.SIRS: ["alpha", "beta", "gamma", "I0", "exp_noise", "S0", "R0", "freq"],
             O.VOLATILITY: ["alpha", "sigma", "gamma", "tau", "x0"],
             O.WALK: ["sd"],
             O.SEIR: ["beta", "sigma", "gamma", "I0", "exp_noise", "S0", "E0", "freq", "import"]}
STATE_NAMES = {O.SIR: ["S", "I", "R", "cases_cumul", "cases_inc"],
               O.SIRS: ["S", "I", "R", "cases_inc"],
               O.VOLATILITY: ["x"],
               O.WALK: ["x"],
               O.SEIR: ["S", "E", "I", "R", "cases_cumul", "cases_inc"]}
DATA_FIELD = {O.SIR: "incidence", O.SIRS: "incidence", O.VOLATILITY: "observed", O.WALK: None,
              O.SEIR: "incidence"}


def _seed_words(seed):
    if seed is None:
        seed = int(np.random.randint(1, 2 ** 31 - 1))
    if isinstance(seed, (int, np.integer)):
        return O.seed_state(int(seed))
    return np.frombuffer(bytes(seed), dtype=np.uint64).copy()


class OracleDustModel:
    def __init__(self, gen, pars, time, n_particles, n_threads, seed, pars_multi, deterministic):
        self._gen = gen
        self._multi = bool(pars_multi)
        self._pars = [dict(p) for p in pars] if pars_multi else dict(pars or {})
        self._np = 1 if n_particles is None else int(n_particles)  # R/if2.R:122
        self._n_pars = len(pars) if pars_multi else 0
        self._n_sets = max(1, self._n_pars)
        self._n_state = len(STATE_NAMES[gen.model_id])
        self._index = None
        self._o = O.OracleModel(pars=gen._pars_vector(self._pars, self._multi), time=int(time),
                                n_particles=self._np, n_pars=self._n_pars,
                                seed=_seed_words(seed), model=gen.model_id,
                                scrambler=gen.scrambler, deterministic=deterministic,
                                n_threads=max(1, int(n_threads)))
        self._n_threads = max(1, int(n_threads))

    def _shape_out(self, a):  # (rows, n_total) -> (rows, N[, P])
        if self._multi:
            return a.reshape(a.shape[0], self._n_sets, self._np).transpose(0, 2, 1)
        return a

    def n_state(self): return self._n_state
    def n_particles(self): return self._np
    def n_pars(self): return self._n_pars
    def shape(self): return (self._np, self._n_pars) if self._multi else (self._np,)
    def time(self): return int(self._o.time())
    def pars(self): return self._pars
    def n_threads(self): return self._n_threads

    def set_n_threads(self, n):
        prev, self._n_threads = self._n_threads, int(n)
        self._o.set_n_threads(int(n))
        return prev
    def uses_gpu(self, fake_gpu=False): return False

    def info(self):
        return [self._gen._info(p) for p in self._pars] if self._multi else self._gen._info(self._pars)

    def set_index(self, index):
        if index is None:
            self._index = None
            return
        if isinstance(index, dict):
            index = list(index.values())
        self._index = np.atleast_1d(index).astype(np.int64) - 1

    def run(self, time_end):
        self._o.run(int(time_end))
        return self._shape_out(self._o.state(self._index))

    def simulate(self, time_end):
        out = [self.run(t) for t in np.atleast_1d(time_end)]
        return np.stack(out, axis=-1)

    def state(self, index=None):
        if index is None:
            return self._shape_out(self._o.state())
        if isinstance(index, dict):
            index = list(index.values())
        return self._shape_out(self._o.state(np.atleast_1d(index).astype(np.int64) - 1))

    def update_state(self, pars=None, state=None, time=None, set_initial_state=None):
        vec = None
        if pars is not None:
            self._pars = [dict(p) for p in pars] if self._multi else dict(pars)
            vec = self._gen._pars_vector(self._pars, self._multi)
        st = None
        if state is not None:
            s = np.asarray(state, dtype=np.float64)
            st = s if s.ndim == 1 else np.ascontiguousarray(np.transpose(s)).reshape(-1, self._n_state).T
        if set_initial_state is None:
            set_initial_state = state is None
        self._o.update_state(pars=vec, state=st, time=time, set_initial_state=set_initial_state)

    def reorder(self, index):
        k = np.asarray(index)
        if self._multi:
            k = (k.T - 1) + (np.arange(self._n_sets) * self._np)[:, None]
        else:
            k = k - 1
        self._o.reorder(np.ascontiguousarray(k).ravel())

    def rng_state(self, first_only=False, last_only=False):
        s = self._o.rng_state()
        if first_only:
            return s[0].tobytes()
        if last_only:
            return s[-1].tobytes()
        return s.tobytes()

    def set_rng_state(self, raw):
        self._o.set_rng_state(np.frombuffer(bytes(raw), dtype=np.uint64).copy())

    def set_filter_stream(self, raw):
        self._o.set_filter_stream(np.frombuffer(bytes(raw), dtype=np.uint64).copy())

    def set_stream_sharing(self, n_sets_global, set_offset, na_global):
        self._o.set_stream_sharing(n_sets_global, set_offset, na_global)

    def set_data(self, data, shared=False):
        n_sets = 1 if shared or not self._multi else self._n_sets
        t = np.array([d[0] for d in data], dtype=np.uint64)
        f = DATA_FIELD[self._gen.model_id]
        v = np.array([[np.nan if r[f] is None else float(r[f]) for r in d[1]] for d in data])
        assert v.shape == (len(data), n_sets)
        self._o.set_data(t, v, shared=bool(shared))
        self._times = t

    def filter(self, time_end=None, save_trajectories=False, time_snapshot=None,
               min_log_likelihood=None):
        if time_end is None:
            time_end = int(self._times[-1])
        snap = None if time_snapshot is None or len(time_snapshot) == 0 else time_snapshot
        r = self._o.filter(int(time_end), save_trajectories, self._index, snap, min_log_likelihood)
        traj, snaps = r["trajectories"], r["snapshots"]
        if self._multi:
            if traj is not None:
                traj = traj.reshape(traj.shape[0], self._n_sets, self._np, -1).transpose(0, 2, 1, 3)
            if snaps is not None:
                snaps = snaps.reshape(snaps.shape[0], self._n_sets, self._np, -1).transpose(0, 2, 1, 3)
        return {"log_likelihood": r["log_likelihood"], "trajectories": traj, "snapshots": snaps}


class OraclePerSetModel:
    """``pars_multi`` with per-set seeds (MCB_SEED_PER_SET) = independent filters: here,
    literally one single-set oracle model per parameter set.  Covers what smc2 uses."""

    def __init__(self, gen, pars, time, n_particles, n_threads, seed, deterministic):
        raw = bytes(seed)
        assert len(raw) == 32 * len(pars)
        self._gen, self._np, self._n_sets = gen, int(n_particles), len(pars)
        self._m = [OracleDustModel(gen, p, time, n_particles, n_threads, raw[32 * k:32 * k + 32],
                                   False, deterministic) for k, p in enumerate(pars)]

    def n_particles(self): return self._np
    def n_pars(self): return self._n_sets
    def shape(self): return (self._np, self._n_sets)
    def time(self): return self._m[0].time()

    def set_data(self, data, shared=False):
        assert shared
        for m in self._m:
            m.set_data(data, False)

    def filter(self, time_end=None, **kw):
        return {"log_likelihood": np.array([m.filter(time_end)["log_likelihood"][0]
                                            for m in self._m]),
                "trajectories": None, "snapshots": None}

    def state(self, index=None):
        return np.stack([m.state(index) for m in self._m], axis=-1)

    def update_state(self, pars=None, state=None, time=None, set_initial_state=None):
        assert state is None
        for k, m in enumerate(self._m):
            m.update_state(pars=None if pars is None else pars[k], time=time,
                           set_initial_state=set_initial_state)

    def rng_state(self, first_only=False, last_only=False):
        # the engine's order: every particle stream, then the filter streams
        parts = [np.frombuffer(m.rng_state(), dtype=np.uint64).reshape(-1, 4) for m in self._m]
        out = np.concatenate([p[:-1] for p in parts] + [p[-1:] for p in parts])
        return out.tobytes()

    def set_rng_state(self, raw):
        a = np.frombuffer(bytes(raw), dtype=np.uint64).reshape(-1, 4)
        nt = self._np * self._n_sets
        for k, m in enumerate(self._m):
            m.set_rng_state(np.concatenate([a[k * self._np:(k + 1) * self._np],
                                            a[nt + k:nt + k + 1]]).tobytes())

    def copy_sets_from(self, other, take, state=True, rng=True):
        for k, t in enumerate(take):
            if not t:
                continue
            if state:
                self._m[k].update_state(state=other._m[k].state(), set_initial_state=False)
            if rng:
                self._m[k].set_rng_state(other._m[k].rng_state())


class OracleGenerator:
    name = "dust_generator"

    def __init__(self, model_name="sir", algorithm="xoshiro256plus"):
        self.model_id = {"sir": O.SIR, "sirs": O.SIRS, "volatility": O.VOLATILITY,
                         "walk": O.WALK, "seir": O.SEIR}[model_name]
        self.scrambler = {"xoshiro256plus": O.PLUS, "xoshiro256starstar": O.STARSTAR}[algorithm]
        self.par_names = PAR_NAMES[self.model_id]
        self.par_defaults = list(O.default_pars(self.model_id))
        self.state_names = STATE_NAMES[self.model_id]
        self.public_methods = self

    def has_compare(self): return DATA_FIELD[self.model_id] is not None
    def has_gpu_support(self, fake_gpu=False): return False
    def time_type(self): return "discrete"

    def new(self, pars, time, n_particles, n_threads=1, seed=None, pars_multi=False,
            deterministic=False, gpu_config=None, seed_per_set=False, **kw):
        if seed_per_set:
            assert pars_multi
            return OraclePerSetModel(self, pars, time, n_particles, n_threads, seed, deterministic)
        return OracleDustModel(self, pars, time, n_particles, n_threads, seed, pars_multi,
                               deterministic)

    def _pars_vector(self, pars, multi):
        sets = pars if multi else [pars]
        return np.array([[float(p.get(n, d)) for n, d in zip(self.par_names, self.par_defaults)]
                         for p in sets])

    def _info(self, pars):
        vals = {n: float(pars.get(n, d)) for n, d in zip(self.par_names, self.par_defaults)}
        return {"vars": list(self.state_names), "pars": vals,
                "index": {n: i + 1 for i, n in enumerate(self.state_names)}}
