"""``particle_filter`` / ``particle_filter_state``: the filter facade mcstate users call.

Python twin of R/particle_filter.R and R/particle_filter_state.R for the hot path:
``particle_filter(data, model, n_particles, compare).run(pars)``.  With
``compare = None`` the whole time series is handed to the model object in one call
(``step_compiled``, R/particle_filter_state.R:138-162) -- on the B200 engine that is
one CUDA-graph launch.  With a host ``compare`` callable the per-interval loop of
``step_r`` (R/particle_filter_state.R:35-136) runs here, crossing to the device twice
per interval exactly as the reference crosses into dust; it works against the same
model object but is never the fast path.

Any object following the dust generator protocol can be the ``model`` (the boundary
is the protocol).  Error messages are the reference's.
"""
from __future__ import annotations

import numpy as np

from ._rng import host_rng
from .particle_filter_data import particle_filter_data, particle_filter_data_split


# ------------------------------------------------------------------ helpers
def scale_log_weights(log_weights):
    """R/particle_filter.R:570-587."""
    lw = np.array(log_weights, dtype=np.float64)
    lw[np.isnan(lw)] = -np.inf
    m = lw.max()
    if not np.isfinite(m):
        return {"weights": np.full(lw.shape, np.nan), "average": -np.inf}
    w = np.exp(lw - m)
    return {"weights": w, "average": float(np.log(np.mean(w)) + m)}


def particle_resample(weights, rng=None):
    """R/particle_filter.R:515-523: systematic resampling with the HOST generator
    (R's runif; here numpy's), 1-based indices; a matrix is resampled column-wise."""
    w = np.asarray(weights, dtype=np.float64)
    if w.ndim == 2:
        return np.stack([particle_resample(w[:, j], rng) for j in range(w.shape[1])], axis=1)
    n = w.size
    draw = (rng.uniform if rng is not None else np.random.uniform)(0.0, 1.0 / n)
    u = draw + np.arange(n) * (1.0 / n)
    cw = np.cumsum(w / w.sum())
    return (np.searchsorted(cw, u, side="right") + 1).astype(np.int64)


def is_dust_generator(x):
    """R/particle_filter.R:590-593."""
    return getattr(x, "name", None) == "dust_generator" and hasattr(x, "new")


def particle_filter_early_exit(log_likelihood, min_log_likelihood):
    """R/particle_filter_state.R:463-474."""
    ll = np.atleast_1d(np.asarray(log_likelihood, dtype=np.float64))
    mn = np.atleast_1d(np.asarray(min_log_likelihood, dtype=np.float64))
    if np.any(ll == -np.inf):
        return True
    if ll.size == 1:
        return bool(ll[0] < mn[0])
    if mn.size == 1:
        return bool(ll.sum() < mn[0])
    return bool(np.all(ll < mn))


def check_n_parameters(n_parameters, data):
    """R/particle_filter.R:1000-1024."""
    has_multiple_data = data.is_nested
    n_data = len(data.populations) if has_multiple_data else 1
    if n_parameters is None:
        has_multiple_parameters = has_multiple_data
        n_parameters = n_data
    else:
        if int(n_parameters) != n_parameters or n_parameters < 1:
            raise ValueError("'n_parameters' must be at least 1")
        if has_multiple_data and n_parameters != n_data:
            raise ValueError("To match the number of populations in your data, "
                             "n_parameters must be %d (if not NULL)" % n_data)
        has_multiple_parameters = True
        n_parameters = int(n_parameters)
    return {"has_multiple_parameters": has_multiple_parameters,
            "has_multiple_data": has_multiple_data, "n_parameters": n_parameters,
            "n_data": n_data}


def check_save_restart(save_restart, data):
    """R/particle_filter.R:596-613: model times -> step times."""
    if save_restart is None:
        return np.zeros(0, dtype=np.int64)
    sr = np.atleast_1d(np.asarray(save_restart))
    if np.any(np.diff(sr) <= 0):
        raise ValueError("'save_restart' must be strictly increasing")
    time_end = data.model_times[:, 1]
    idx = [int(np.where(time_end == s)[0][0]) if np.any(time_end == s) else -1 for s in sr]
    bad = [str(s) for s, i in zip(sr.tolist(), idx) if i < 0]
    if bad:
        raise ValueError("'save_restart' contains times not in '%s': %s" % (data.time, ", ".join(bad)))
    return data.times[idx, 1].astype(np.int64)


def check_time_step(curr, time_index, times, name):
    """R/particle_filter_state.R:428-444."""
    n_times = times.shape[0]
    if curr >= n_times:
        raise ValueError("%s has reached the end of the data" % name)
    if time_index > n_times:
        raise ValueError("time_index %d is beyond the length of the data (max %d)"
                         % (time_index, n_times))
    if time_index <= curr:
        raise ValueError("%s has already run step index %d (to model step %d)"
                         % (name, time_index, times[time_index - 1, 1]))


class _PendingHistory(dict):
    """``history`` of a filter state that has not stepped yet, on the compiled path:
    ``value`` / ``order`` are allocated on first access."""

    def __init__(self, allocate, index):
        super().__init__(index=index)
        self._allocate = allocate

    def __getitem__(self, key):
        if key in ("value", "order") and not dict.__contains__(self, key):
            value, order = self._allocate()
            dict.__setitem__(self, "value", value)
            dict.__setitem__(self, "order", order)
        return dict.__getitem__(self, key)


class _DeviceHistory:
    """Trajectories of a compiled-filter run that stayed in device memory: lineages
    are traced there on request (model.filter_history), so asking for one sampled
    particle -- all that pMCMC keeps, R/pmcmc_state.R:32-51 -- moves
    n_index x (T+1) doubles instead of the whole [n_index, N, T+1] array.  Valid
    until the model's next filter call, like the reference's last_history."""

    def __init__(self, model, n_particles, multi):
        self._model, self._np, self._multi = model, int(n_particles), bool(multi)

    def get(self, index_particle=None):
        if index_particle is None:
            index_particle = np.arange(1, self._np + 1)
        return self._model.filter_history(index_particle)


def history_single(history_value, history_order, index_particle=None):
    """R/particle_filter.R:616-646 (1-based indices)."""
    hv = np.asarray(history_value)
    if history_order is None:
        return hv if index_particle is None else hv[:, np.atleast_1d(index_particle) - 1, :]
    ip = (np.arange(1, hv.shape[1] + 1) if index_particle is None
          else np.atleast_1d(index_particle))
    nt = history_order.shape[1]
    out = np.empty((hv.shape[0], ip.size, nt))
    for t in range(nt - 1, -1, -1):
        ip = history_order[ip - 1, t]
        out[:, :, t] = hv[:, ip - 1, t]
    return out


def history_multiple(history_value, history_order, index_particle=None):
    """R/particle_filter.R:653-714."""
    hv = np.asarray(history_value)
    ny, npart, npop, nt = hv.shape
    if history_order is None:
        if index_particle is None:
            return hv
        ip = np.asarray(index_particle)
        if ip.ndim < 2:
            return hv[:, np.atleast_1d(ip) - 1, :, :]
        if ip.shape[1] != npop:
            raise ValueError("'index_particle' should have %d columns" % npop)
        out = np.empty((ny, ip.shape[0], npop, nt))
        for i in range(npop):
            out[:, :, i, :] = hv[:, ip[:, i] - 1, i, :]
        return out
    if index_particle is None:
        ip = np.tile(np.arange(1, npart + 1)[:, None], (1, npop))
    else:
        ip = np.asarray(index_particle)
        if ip.ndim == 2:
            if ip.shape[1] != npop:
                raise ValueError("'index_particle' should have %d columns" % npop)
        else:
            ip = np.tile(np.atleast_1d(ip)[:, None], (1, npop))
    out = np.empty((ny, ip.shape[0], npop, nt))
    for t in range(nt - 1, -1, -1):
        nxt = np.empty_like(ip)
        for j in range(npop):
            nxt[:, j] = history_order[ip[:, j] - 1, j, t]
            out[:, :, j, t] = hv[:, nxt[:, j] - 1, j, t]
        ip = nxt
    return out


def _restart_subset(restart_state, index_particle):
    rs = np.asarray(restart_state)
    if index_particle is None:
        return rs
    ip = np.asarray(index_particle)
    if rs.ndim == 3:
        return rs[:, np.atleast_1d(ip) - 1, :]
    if ip.ndim < 2:
        return rs[:, np.atleast_1d(ip) - 1, :, :]
    if ip.shape[1] != rs.shape[2]:
        raise ValueError("'index_particle' should have %d columns" % rs.shape[2])
    out = np.empty((rs.shape[0], ip.shape[0], rs.shape[2], rs.shape[3]))
    for i in range(rs.shape[2]):
        out[:, :, i, :] = rs[:, ip[:, i] - 1, i, :]
    return out


# ------------------------------------------- single / multiple parameter adapters
def _pfs_compare_single(state, compare, data, pars):
    """R/particle_filter_state.R:566-572."""
    lw = compare(state, data, pars)
    return None if lw is None else scale_log_weights(lw)


def _pfs_compare_multiple(has_multiple_data, state, compare, data, pars):
    """R/particle_filter_state.R:577-593."""
    n_layer = state.shape[2]
    lws = [compare(state[:, :, i], data[i] if has_multiple_data else data, pars[i])
           for i in range(n_layer)]
    if all(lw is None or len(lw) == 0 for lw in lws):
        return None
    ws = [scale_log_weights(lw) for lw in lws]
    return {"average": np.array([w["average"] for w in ws]),
            "weights": np.stack([w["weights"] for w in ws], axis=1)}


# ---------------------------------------------------------------------------
# multistage parameters (R/particle_filter_multistage.R): the model's parameters --
# and, through transform_state, its state -- change at given times of the run
# ---------------------------------------------------------------------------
def transform_state_identity(x, *args):
    return x


class multistage_epoch:
    """``multistage_epoch(start, pars = NULL, transform_state = NULL)``
    (R/particle_filter_multistage.R:113-126).  ``start`` is in the data's time units."""

    def __init__(self, start, pars=None, transform_state=None):
        if transform_state is None:
            transform_state = transform_state_identity
        elif not callable(transform_state):
            raise TypeError("'transform_state' must be a function")
        self.start, self.pars, self.transform_state = start, pars, transform_state


class multistage_parameters(list):
    """``multistage_parameters(pars, epochs)`` (R/particle_filter_multistage.R:60-93): a
    list of stages ``dict(pars, start, transform_state)``, the first being the base."""

    def __init__(self, pars, epochs):
        if not all(isinstance(e, multistage_epoch) for e in epochs):
            raise TypeError("Expected all elements of 'epochs' to be 'multistage_epoch' objects")
        sets_pars = [e.pars is None for e in epochs]
        if any(sets_pars) and not all(sets_pars):
            raise ValueError("If changing parameters, all epochs must contain 'pars'")
        stages = [{"pars": pars, "start": None, "transform_state": None}]
        stages += [{"pars": e.pars, "start": e.start, "transform_state": e.transform_state}
                   for e in epochs]
        super().__init__(stages)


def filter_check_times(pars, data, save_restart):
    """R/particle_filter_multistage.R:134-188: which stages overlap the data, the (1-based)
    row index each one ends at, and which restart times fall into which stage."""
    times = data.model_times
    stages = [dict(st) for st in pars]
    time_pars = np.array([st["start"] for st in stages[1:]], dtype=float)
    start = np.concatenate([[-np.inf], time_pars])
    end = np.concatenate([time_pars, [np.inf]])
    drop = (end <= times[0, 0]) | (start > times[-1, 1])
    index = [i + 1 for i in range(len(stages)) if not drop[i]]
    stages = [st for st, d in zip(stages, drop) if not d]
    time_pars = np.array([st["start"] for st in stages[1:]], dtype=float)
    ends = list(times[:, 1])
    time_index = [ends.index(t) + 1 if t in ends else None for t in time_pars] + [len(ends)]
    bad = [str(index[i]) for i, t in enumerate(time_index) if t is None]
    if bad:
        raise ValueError("Could not map epoch to filter time: error for stage %s" % ", ".join(bad))
    sr = [] if save_restart is None else list(np.atleast_1d(save_restart))
    err = [t for t in sr if t in set(time_pars.tolist())]
    if err:
        raise ValueError("save_restart cannot include epoch change: error for %s"
                         % ", ".join("%g" % e for e in err))
    stage_of = np.searchsorted(np.concatenate([[0.0], time_pars]), sr, side="right")
    for i, st in enumerate(stages):
        st["time_index"] = time_index[i]
        if save_restart is not None:
            st["restart_index"] = [j for j in range(len(sr)) if stage_of[j] == i + 1]
    return stages


class _JoinedDeviceHistory:
    """The stages' device-resident histories (one model object per stage), joined on
    request (join_histories)."""

    def __init__(self, parts, stages, names):
        self._parts, self._stages, self._names = parts, stages, names

    def get(self, index_particle=None):
        hs = [{"value": p["value"].get(index_particle), "order": None, "index": p["index"]}
              for p in self._parts]
        return join_histories(hs, self._stages)["value"]


def join_histories(history, stages):
    """R/particle_filter_multistage.R:191-243.  A stage's arrays either span the whole
    series (the R loop: only its own interval is filled) or hold just the rows of its own
    filter call (the compiled filter: slice 0 is the state the call started from)."""
    time_index = [st["time_index"] for st in stages]
    if not isinstance(history[0]["index"], dict):
        raise ValueError("Named index required")
    nms = []
    for h in history:
        nms += [n for n in h["index"] if n not in nms]
    if any(isinstance(h["value"], _DeviceHistory) for h in history):
        return {"value": _JoinedDeviceHistory(history, stages, nms), "order": None,
                "index": {n: None for n in nms}}
    has_order = history[0]["order"] is not None
    n_time = time_index[-1] + 1
    first = history[0]["value"]
    value = np.full((len(nms),) + first.shape[1:-1] + (n_time,), np.nan)
    order = None
    if has_order:
        order = np.zeros(history[0]["order"].shape[:-1] + (n_time,), dtype=np.int64)
    end = [t + 1 for t in time_index]                 # 1-based, inclusive
    start = [1] + [e + 1 for e in end[:-1]]
    for i, h in enumerate(history):
        j = [nms.index(n) for n in h["index"]]
        k = np.arange(start[i] - 1, end[i])           # 0-based slices of the joined array
        v = np.asarray(h["value"])
        if v.shape[-1] == n_time or has_order:        # full-length arrays: same slices
            value[j, ..., k[0]:k[-1] + 1] = v[..., k[0]:k[-1] + 1]
        else:                                         # this call's rows only
            local = v if i == 0 else v[..., 1:]
            value[j, ..., k[0]:k[-1] + 1] = local[..., :k.size]
        if has_order:
            order[..., k[0]:k[-1] + 1] = h["order"][..., k[0]:k[-1] + 1]
    return {"value": value, "order": order, "index": {n: None for n in nms}}


def join_restart_state(restart, stages):
    """R/particle_filter_multistage.R:253-266."""
    parts = []
    for r, st in zip(restart, stages):
        idx = st.get("restart_index", [])
        if r is not None and len(idx):
            parts.append(np.asarray(r)[..., idx])
    if len({p.shape[0] for p in parts}) > 1:
        raise ValueError("Restart state varies in size over the simulation:\n" + "\n".join(
            "  - %d: %d rows" % (i + 1, p.shape[0]) for i, p in enumerate(parts)))
    return np.concatenate(parts, axis=-1) if parts else None


def particle_filter_update_state(transform, model_old, model_new):
    """R/particle_filter_state.R:495-513."""
    info_old, info_new = model_old.info(), model_new.info()
    n_pars = model_old.n_pars()
    if n_pars == 0:
        state = transform(model_old.state(), info_old, info_new)
    else:
        old = model_old.state()
        state = np.stack([transform(old[:, :, i], info_old[i], info_new[i])
                          for i in range(n_pars)], axis=-1)
    model_new.update_state(state=np.asarray(state, dtype=float), time=model_old.time())


class particle_filter_state:
    """R/particle_filter_state.R: one run of the filter, steppable."""

    def __init__(self, pars, generator, model, data, data_split, times, n_particles,
                 has_multiple_parameters, n_threads, initial, index, compare,
                 constant_log_likelihood, gpu_config, seed, min_log_likelihood,
                 save_history, save_restart):
        has_multiple_data = data.is_nested
        if model is None:
            # first run: create the model object and upload the observations ONCE
            # (R/particle_filter_state.R:237-246)
            model = generator.new(pars=pars, time=int(times[0, 0]), n_particles=n_particles,
                                  n_threads=n_threads, seed=seed, gpu_config=gpu_config,
                                  pars_multi=has_multiple_parameters)
            if compare is None:
                data_is_shared = has_multiple_parameters and not has_multiple_data
                model.set_data(data_split, data_is_shared)
        else:
            # later runs reuse it: new parameters, initial state, RNG continues (:248)
            model.update_state(pars=pars, time=int(times[0, 0]))

        if initial is not None:
            if has_multiple_parameters:
                infos = model.info()
                states = [initial(infos[i], n_particles, pars[i]) for i in range(len(pars))]
                if not all(s is None for s in states):   # pfs_initial_multiple (:528-546)
                    if any(isinstance(s, dict) for s in states):
                        raise ValueError("Setting 'time' from initial no longer supported")
                    lens = [np.asarray(s).shape for s in states]
                    if len(set(lens)) != 1:
                        raise ValueError("initial() produced unequal state lengths %s"
                                         % ", ".join(str(int(np.prod(l))) for l in lens))
                    st = np.stack([np.asarray(s, dtype=float) for s in states], axis=-1)
                    if st.ndim == 2:  # vectors: recycle over particles
                        st = np.repeat(st[:, None, :], n_particles, axis=1)
                    model.update_state(state=st)
            else:
                st = initial(model.info(), n_particles, pars)
                if isinstance(st, dict):
                    raise ValueError("Setting 'time' from initial no longer supported")
                model.update_state(state=st)

        if index is None:
            index_data = None
        else:
            if has_multiple_parameters:
                all_idx = [index(i) for i in model.info()]
                if any(_idx_key(a) != _idx_key(all_idx[0]) for a in all_idx[1:]):
                    raise ValueError("index must be identical across populations")
                index_data = all_idx[0]
            else:
                index_data = index(model.info())
            model.set_index(index_data["run"] if compare is not None else index_data["state"])

        shape = model.shape()
        self.history = None
        if save_history:
            length = times.shape[0] + 1
            index_state = None if index_data is None else index_data["state"]

            def allocate():
                # R/particle_filter_state.R:270-285
                state = model.state(index_state)
                value = np.full(state.shape + (length,), np.nan)
                value[..., 0] = state
                order = np.empty(tuple(shape) + (length,), dtype=np.int64)
                order[...] = np.arange(1, n_particles + 1).reshape((n_particles,) + (1,) * len(shape))
                return value, order

            if compare is None:
                # the compiled filter returns its own history and never writes these
                # arrays ([n_index, N, T+1]: 160 MB at 2^16 particles, per run): build
                # them only if somebody looks
                self.history = _PendingHistory(allocate, index_state)
            else:
                value, order = allocate()
                self.history = {"value": value, "order": order, "index": index_state}

        self._save_restart_time = check_save_restart(save_restart, data)
        self.restart_state = None
        if self._save_restart_time.size > 0:
            self.restart_state = np.full((model.n_state(),) + tuple(shape) +
                                         (self._save_restart_time.size,), np.nan)

        self._generator = generator
        self._pars = pars
        self._data = data
        self._data_split = data_split
        self._times = times
        self._n_particles = n_particles
        self._n_threads = n_threads
        self._has_multiple_parameters = has_multiple_parameters
        self._has_multiple_data = has_multiple_data
        self._initial = initial
        self._index = index
        self._compare = compare
        self._constant_log_likelihood = constant_log_likelihood
        self._gpu = gpu_config is not None
        self._gpu_config = gpu_config
        self._min_log_likelihood = min_log_likelihood
        self._save_restart = save_restart

        self.model = model
        self.current_time_index = 0
        self.log_likelihood_step = None
        # R/particle_filter_state.R:313-314, 477-493
        if constant_log_likelihood is None:
            self.log_likelihood = np.zeros(len(pars)) if has_multiple_parameters else 0.0
        elif has_multiple_parameters:
            self.log_likelihood = np.array([constant_log_likelihood(p) for p in pars], dtype=float)
        else:
            self.log_likelihood = float(constant_log_likelihood(pars))

    # ---- public
    def run(self):
        return self.step(self._times.shape[0])

    def step(self, time_index, partial=False):
        if self._compare is None:
            return self._step_compiled(time_index, partial)
        return self._step_r(time_index, partial)

    # ---- R/particle_filter_state.R:138-162
    def _step_compiled(self, time_index, partial=False):
        if partial:
            raise ValueError("'partial' not supported with compiled compare")
        check_time_step(self.current_time_index, time_index, self._times, "Particle filter")
        time = int(self._times[time_index - 1, 1])
        save_history = self.history is not None
        # snapshots the model can take in THIS call: the restart times in (now, time]
        # (a stage of a multistage run covers part of the series)
        t_now = int(self._times[self.current_time_index - 1, 1]) if self.current_time_index else \
            int(self._times[0, 0])
        all_snap = self._save_restart_time
        here = np.where((all_snap > t_now) & (all_snap <= time))[0]
        snap = all_snap[here] if here.size else None
        # the one native call for all the intervals; min_log_likelihood is NOT forwarded,
        # as in the reference (:151)
        on_device = save_history and hasattr(self.model, "filter_history")
        res = self.model.filter(time, "device" if on_device else save_history, snap)
        self.log_likelihood_step = np.nan
        ll = res["log_likelihood"]
        self.log_likelihood = self.log_likelihood + (ll if self._has_multiple_parameters else float(ll[0]))
        self.current_time_index = time_index
        if save_history:
            value = res["trajectories"]
            if on_device:
                value = _DeviceHistory(self.model, self._n_particles,
                                       self._has_multiple_parameters)
            self.history = {"value": value, "order": None, "index": self.history["index"]}
        if all_snap.size and res["snapshots"] is not None:
            if here.size == all_snap.size:
                self.restart_state = res["snapshots"]
            else:
                if self.restart_state is None:
                    self.restart_state = np.full(res["snapshots"].shape[:-1] + (all_snap.size,),
                                                 np.nan)
                self.restart_state[..., here] = res["snapshots"]
        elif not all_snap.size:
            self.restart_state = None
        return self.log_likelihood

    # ---- R/particle_filter_state.R:35-136
    def _step_r(self, time_index, partial=False):
        curr = self.current_time_index
        check_time_step(curr, time_index, self._times, "Particle filter")
        model = self.model
        multi = self._has_multiple_parameters
        save_history = self.history is not None
        save_restart = self.restart_state is not None
        log_likelihood = self.log_likelihood
        log_likelihood_step = None
        n_particles = self._n_particles
        for t in range(curr + 1, time_index + 1):
            time_end = int(self._times[t - 1, 1])
            state = model.run(time_end)
            if save_history:
                self.history["value"][..., t] = model.state(self.history["index"])
            if multi:
                weights = _pfs_compare_multiple(self._has_multiple_data, state, self._compare,
                                                self._data_split[t - 1], self._pars)
            else:
                weights = _pfs_compare_single(state, self._compare, self._data_split[t - 1],
                                              self._pars)
            if weights is None:
                log_likelihood_step = np.nan
                kappa = np.arange(1, n_particles + 1)
                if multi:
                    kappa = np.tile(kappa[:, None], (1, len(self._pars)))
            else:
                log_likelihood_step = weights["average"]
                log_likelihood = log_likelihood + log_likelihood_step
                if particle_filter_early_exit(log_likelihood, self._min_log_likelihood):
                    log_likelihood = (np.full_like(np.asarray(log_likelihood, dtype=float), -np.inf)
                                      if multi else -np.inf)
                    break
                kappa = particle_resample(weights["weights"])
                model.reorder(kappa)
            if save_history:
                self.history["order"][..., t] = kappa
            if save_restart:
                hit = np.where(self._save_restart_time == time_end)[0]
                if hit.size:
                    self.restart_state[..., hit[0]] = model.state()
        self.log_likelihood_step = log_likelihood_step
        self.log_likelihood = log_likelihood
        self.current_time_index = time_index
        return log_likelihood_step if partial else log_likelihood

    def fork_multistage(self, model, pars, transform_state):
        """R/particle_filter_state.R:360-389: a new filter state for the next stage --
        its own model object (created on the first run, reused afterwards) with the new
        parameters, this one's RNG state, and this one's model state passed through
        ``transform_state``; position in the data and likelihood carry over.  The
        reference refuses this for its GPU back end (:361, the RNG state could not be
        handed over); here it is a device copy."""
        seed = self.model.rng_state()
        save_history = self.history is not None
        if model is not None:
            model.set_rng_state(seed)
        if pars is None:
            pars = self.model.pars()
        ret = particle_filter_state(
            pars, self._generator, model, self._data, self._data_split, self._times,
            self._n_particles, self._has_multiple_parameters, self._n_threads, None,
            self._index, self._compare, None, self._gpu_config, seed,
            self._min_log_likelihood, save_history, self._save_restart)
        particle_filter_update_state(transform_state, self.model, ret.model)
        ret.current_time_index = self.current_time_index
        ret.log_likelihood = self.log_likelihood
        ret.log_likelihood_step = self.log_likelihood_step
        return ret

    def fork_smc2(self, pars):
        """R/particle_filter_state.R:402-424: re-run from the start with new parameters
        on the parent's RNG state, then hand the advanced RNG back to the parent.  The
        reference forbids this on its GPU back end (:403); RNG state get/set is cheap
        here, so it is allowed."""
        seed = self.model.rng_state()
        ret = particle_filter_state(
            pars, self._generator, None, self._data, self._data_split, self._times,
            self._n_particles, self._has_multiple_parameters, self._n_threads, self._initial,
            self._index, self._compare, self._constant_log_likelihood, None, seed,
            self._min_log_likelihood, self.history is not None, self._save_restart)
        ret.step(self.current_time_index)
        self.model.set_rng_state(ret.model.rng_state())
        return ret


def _idx_key(d):
    return {k: (list(v.items()) if isinstance(v, dict) else list(np.atleast_1d(v)))
            for k, v in d.items()}


class particle_filter:
    """R/particle_filter.R: ``particle_filter$new(data, model, n_particles, compare, ...)``."""

    def __init__(self, data, model, n_particles, compare, index=None, initial=None,
                 constant_log_likelihood=None, n_threads=1, seed=None, n_parameters=None,
                 gpu_config=None, stochastic_schedule=None, ode_control=None):
        if not is_dust_generator(model):
            raise TypeError("'model' must be a dust_generator")
        for name, fn in (("index", index), ("initial", initial),
                         ("constant_log_likelihood", constant_log_likelihood)):
            if fn is not None and not callable(fn):
                raise TypeError("'%s' must be function if not NULL" % name)
        if not isinstance(data, particle_filter_data):
            raise TypeError("'data' must be a particle_filter_data")
        if compare is None:
            if not model.public_methods.has_compare():
                raise ValueError("Your model does not have a built-in 'compare' function")
        elif not callable(compare):
            raise TypeError("'compare' must be a function")
        if gpu_config is not None and not model.public_methods.has_gpu_support(True):
            raise ValueError("'gpu_config' provided, but 'model' does not have GPU support")
        if model.public_methods.time_type() != "discrete":
            raise ValueError("'model' is continuous but 'data' is of type "
                             "'particle_filter_data_discrete'")
        if stochastic_schedule is not None:
            raise ValueError("'stochastic_schedule' provided but 'model' does not support this")
        if ode_control is not None:
            raise ValueError("'ode_control' provided but 'model' does not support this")
        self.model = model
        self._data = data
        info = check_n_parameters(n_parameters, data)
        self.has_multiple_parameters = info["has_multiple_parameters"]
        self.has_multiple_data = info["has_multiple_data"]
        self.n_parameters = info["n_parameters"]
        self.n_data = info["n_data"]
        self._times = data.times
        self._data_split = particle_filter_data_split(data, compare is None)
        self._compare = compare
        self._gpu_config = gpu_config
        self._index = index
        self._initial = initial
        self._constant_log_likelihood = constant_log_likelihood
        if int(n_particles) != n_particles or n_particles < 1:
            raise ValueError("'n_particles' must be at least 1")
        self.n_particles = int(n_particles)
        if int(n_threads) != n_threads or n_threads < 1:
            raise ValueError("'n_threads' must be at least 1")
        self._n_threads = int(n_threads)
        self._seed = seed
        self._last_model = None
        self._last_history = None
        self._last_state = None
        self._last_restart_state = None
        self._last_stages = None

    def run(self, pars=None, save_history=False, save_restart=None, min_log_likelihood=None):
        """R/particle_filter.R:313-317 -> filter_run -> filter_run_simple (:819-849)."""
        pars = {} if pars is None else pars
        if not isinstance(save_history, (bool, np.bool_)):
            raise TypeError("'save_history' must be a logical")
        multistage = isinstance(pars, multistage_parameters)
        if self.has_multiple_parameters and not multistage:
            if isinstance(pars, dict):
                raise ValueError("Expected an unnamed list of parameters for 'pars'")
            if len(pars) != self.n_parameters:
                raise ValueError("'pars' must have length %d" % self.n_parameters)
        # the number of stages is fixed by the first run: one model object per stage is
        # kept between runs (particle_filter_check_multistage_pars, R/particle_filter.R:971-986)
        n_stages = len(pars) if multistage else 1
        if self._last_stages is not None and self._last_stages != n_stages:
            if self._last_stages == 1:
                raise ValueError("Expected single-stage parameters (but given one with %d stages)"
                                 % n_stages)
            raise ValueError("Expected multistage_pars with %d stages (but given one with %d)"
                             % (self._last_stages, n_stages))
        self._last_stages = n_stages
        if multistage:
            return self._run_multistage(pars, save_history, save_restart, min_log_likelihood)
        obj = self.run_begin(pars, save_history, save_restart, min_log_likelihood)
        obj.run()
        self._last_history = obj.history
        self._last_model = [obj.model]
        self._last_state = obj.model.state
        self._last_restart_state = obj.restart_state
        return obj.log_likelihood

    def _run_multistage(self, pars, save_history, save_restart, min_log_likelihood):
        """filter_run_multistage (R/particle_filter.R:852-898)."""
        if self.has_multiple_parameters:
            raise NotImplementedError("multistage parameters with a nested filter")
        stages = filter_check_times(pars, self._data, save_restart)
        models = list(self._last_model) if self._last_model else []
        if len(models) != len(stages):
            models = [models[0] if models else None] + [None] * (len(stages) - 1)
        history, restart = [], []
        obj = None
        for i, st in enumerate(stages):
            if i == 0:
                self._last_model = None if models[0] is None else [models[0]]
                obj = self.run_begin(st["pars"], save_history, save_restart, min_log_likelihood)
            else:
                obj = obj.fork_multistage(models[i], st["pars"], st["transform_state"])
            obj.step(st["time_index"])
            models[i] = obj.model
            history.append(obj.history)
            restart.append(obj.restart_state)
        # the final RNG state goes into the first model, completing the cycle: the next
        # run starts from the first model (:877-879)
        if len(models) > 1:
            models[0].set_rng_state(models[-1].rng_state())
        self._last_model = models
        self._last_state = models[-1].state
        self._last_history = join_histories(history, stages) if save_history else None
        self._last_restart_state = (join_restart_state(restart, stages)
                                    if save_restart is not None else None)
        return obj.log_likelihood

    def run_begin(self, pars=None, save_history=False, save_restart=None,
                  min_log_likelihood=None):
        """R/particle_filter.R:338-351."""
        pars = {} if pars is None else pars
        if min_log_likelihood is None:
            min_log_likelihood = -np.inf
        last = None if self._last_model is None else self._last_model[0]
        return particle_filter_state(
            pars, self.model, last, self._data, self._data_split, self._times,
            self.n_particles, self.has_multiple_parameters, self._n_threads, self._initial,
            self._index, self._compare, self._constant_log_likelihood, self._gpu_config,
            self._seed, min_log_likelihood, bool(save_history), save_restart)

    def state(self, index_state=None):
        """R/particle_filter.R:360-366."""
        if self._last_state is None:
            raise RuntimeError("Model has not yet been run")
        return self._last_state(index_state)

    def state_particles(self, index_particle):
        """``filter$state()[, index_particle]`` ([n_state, n]): only those columns leave the
        device when the model object can select them (mcb_state_particles)."""
        if self._last_model is None:
            raise RuntimeError("Model has not yet been run")
        model = self._last_model[-1]
        if hasattr(model, "state_particles") and not self.has_multiple_parameters:
            return model.state_particles(index_particle)
        return np.asarray(self.state())[:, np.atleast_1d(index_particle) - 1]

    def history(self, index_particle=None):
        """R/particle_filter.R:377-397."""
        if self._last_model is None:
            raise RuntimeError("Model has not yet been run")
        if self._last_history is None:
            raise RuntimeError("Can't get history as model was run with save_history = FALSE")
        value, order = self._last_history["value"], self._last_history["order"]
        if isinstance(value, (_DeviceHistory, _JoinedDeviceHistory)):
            return value.get(index_particle)
        if value.ndim == 4:
            return history_multiple(value, order, index_particle)
        return history_single(value, order, index_particle)

    def history_names(self):
        """``rownames(filter$history())`` (tests/testthat/test-particle-filter.R:367-392):
        the names of a named ``index$state``, else None -- numpy arrays carry no row
        names, so they are asked for here."""
        if self._last_model is None:
            raise RuntimeError("Model has not yet been run")
        if self._last_history is None:
            raise RuntimeError("Can't get history as model was run with save_history = FALSE")
        idx = self._last_history["index"]
        return list(idx) if isinstance(idx, dict) else None

    def restart_state(self, index_particle=None, save_restart=None, restart_match=False):
        """R/particle_filter.R:444-468 (the compiled path has no order array, so
        ``restart_match`` falls back to plain subsetting exactly as there, :718)."""
        if self._last_model is None:
            raise RuntimeError("Model has not yet been run")
        if self._last_restart_state is None:
            raise RuntimeError("Can't get history as model was run with save_restart = NULL")
        order = None if self._last_history is None else self._last_history["order"]
        rs = self._last_restart_state
        if index_particle is None or order is None or not restart_match:
            return _restart_subset(rs, index_particle)
        times_restart = check_save_restart(save_restart, self._data)
        idx_save = [int(np.where(self._times[:, 0] == s)[0][0]) + 1 for s in times_restart]
        ip = np.atleast_1d(index_particle)
        if rs.ndim == 3:
            nt = order.shape[1]
            idx = np.empty((ip.size, nt), dtype=np.int64)
            for t in range(nt - 1, -1, -1):
                ip = idx[:, t] = order[ip - 1, t]
            out = np.empty((rs.shape[0], idx.shape[0], len(idx_save)))
            for i, s in enumerate(idx_save):
                out[:, :, i] = rs[:, idx[:, s] - 1, i]
            return out
        npop, nt = rs.shape[2], order.shape[2]
        ipm = ip if ip.ndim == 2 else np.tile(ip[:, None], (1, npop))
        idx = np.empty((ipm.shape[0], npop, nt), dtype=np.int64)
        for t in range(nt - 1, -1, -1):
            for j in range(npop):
                idx[:, j, t] = order[ipm[:, j] - 1, j, t]
            ipm = idx[:, :, t]
        out = np.empty((rs.shape[0], idx.shape[0], npop, len(idx_save)))
        for i in range(npop):
            for j, s in enumerate(idx_save):
                out[:, :, i, j] = rs[:, idx[:, i, s] - 1, i, j]
        return out

    def inputs(self):
        """R/particle_filter.R:479-496: everything needed to rebuild the filter; the
        seed is the live first-stream RNG state (:811-816)."""
        seed = self._seed
        if self._last_model is not None and hasattr(self._last_model[-1], "rng_state"):
            seed = self._last_model[-1].rng_state(first_only=True)
        return {"data": self._data, "model": self.model, "n_particles": self.n_particles,
                "index": self._index, "initial": self._initial, "compare": self._compare,
                "constant_log_likelihood": self._constant_log_likelihood,
                "gpu_config": self._gpu_config, "n_threads": self._n_threads,
                "n_parameters": self.n_parameters if self.has_multiple_parameters else None,
                "stochastic_schedule": None, "ode_control": None, "seed": seed}

    def set_n_threads(self, n_threads):
        """R/particle_filter.R:508-510."""
        prev, self._n_threads = self._n_threads, n_threads
        for m in self._last_model or []:
            if m is not None:
                m.set_n_threads(n_threads)
        return prev


def particle_filter_initial(state, rng=None):
    """``particle_filter_initial(state)`` (R/particle_filter_initial.R:18-25): an
    ``initial`` function for a filter restarted from saved state -- ``state`` is a matrix
    ``[n_state, n_samples]`` (``pmcmc(...)["restart"][:, :, k]``); each call samples
    ``n_particles`` of its columns with replacement (``sample.int``; here a
    ``numpy.random.Generator``, a fresh one unless given)."""
    state = np.asarray(state)
    if state.ndim != 2:
        raise TypeError("'state' must be a matrix")
    state = state.copy()
    gen = host_rng() if rng is None else rng

    def initial(info, n_particles, pars):
        return state[:, gen.integers(0, state.shape[1], int(n_particles))]

    return initial


def particle_filter_from_inputs(inputs, seed=None):
    """R/particle_filter.R:528-567."""
    return particle_filter(
        inputs["data"], inputs["model"], inputs["n_particles"], inputs["compare"],
        index=inputs["index"], initial=inputs["initial"],
        constant_log_likelihood=inputs["constant_log_likelihood"],
        n_threads=inputs["n_threads"], seed=inputs["seed"] if seed is None else seed,
        n_parameters=inputs["n_parameters"], gpu_config=inputs["gpu_config"])
