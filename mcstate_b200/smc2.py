"""SMC^2 driving the B200 filter (BASELINE config C5).

The reference's ``smc2()`` keeps one ``particle_filter_state`` per parameter particle
and advances them one data row at a time (R/smc2.R:58-114), re-running whole filters
from t = 0 when it replenishes (``fork_smc2``, R/particle_filter_state.R:402-424).  It
forbids all of that on its GPU back end (:403).  Here the parameter particles of a rank
are ONE batched model handle -- ``pars_multi`` with per-set seeds, i.e. exactly the
streams the reference's separate model objects would hold (R/smc2.R:68-71) -- so every
data row is one incremental ``model$filter(time_end)`` call over all of them, and a
replenishment is one more batched filter run plus a device-side per-set copy
(``mcb_copy_sets``).

Sharding (one rank per GPU): parameter particles [lo, hi) live on rank r; the per-row
log-likelihoods are all-gathered (8 bytes x n_parameter_sets) and every rank then makes
the same host-side decisions from the same host RNG.  The reference's resampling step
permutes parameters and probabilities but NOT the filters (SURVEY.md Q1,
R/smc2.R:120-149), so no particle state ever moves between GPUs; that quirk and the
one of passing log-weights to ``particle_resample`` (:122) are reproduced as they are.
"""
from __future__ import annotations

import math

import numpy as np

from ._rng import host_rng
from .dust import dust_rng_distributed_state
from .particle_filter import particle_resample, scale_log_weights
from .particle_filter_data import particle_filter_data_split
from .pmcmc import reflect_proposal_rows, rmvnorm_generator
from .sharding import block_range


class smc2_parameter:
    """``smc2_parameter(name, sample, prior, min, max, discrete, integer)``
    (R/smc2_parameters.R): ``sample(n, rng)`` draws n values, ``prior(x)`` is the log density."""

    def __init__(self, name, sample, prior, min=-math.inf, max=math.inf, discrete=False,
                 integer=False):
        if not callable(sample) or not callable(prior):
            raise TypeError("'sample' and 'prior' must be functions")
        if discrete:   # R/smc2_parameters.R: deprecated spelling of `integer`
            import warnings

            warnings.warn("'discrete' is deprecated.\nUse 'integer' instead.", DeprecationWarning,
                          stacklevel=2)
        if max <= min:
            raise ValueError("'max' must be > 'min' (%g)" % min)
        self.name, self.sample, self.prior = name, sample, prior
        self.min, self.max = float(min), float(max)
        self.integer = bool(discrete or integer)


class smc2_parameters:
    """``smc2_parameters$new(parameters, transform)`` (R/smc2_parameters.R)."""

    def __init__(self, parameters, transform=None):
        if len(parameters) == 0:
            raise ValueError("At least one parameter is required")
        self._p = list(parameters)
        self._names = [p.name for p in self._p]
        if len(set(self._names)) != len(self._names):
            raise ValueError("Duplicate parameter names")
        self._min = np.array([p.min for p in self._p])
        self._max = np.array([p.max for p in self._p])
        self._integer = np.array([p.integer for p in self._p])
        self._transform = transform if transform is not None else (lambda d: d)

    def names(self):
        return list(self._names)

    def summary(self):
        """R/smc2_parameters.R: one record per parameter."""
        return [{"name": p.name, "min": p.min, "max": p.max, "discrete": p.integer,
                 "integer": p.integer} for p in self._p]

    def sample(self, n, rng):
        return np.stack([np.asarray(p.sample(n, rng), dtype=float) for p in self._p], axis=1)

    def prior(self, theta):
        """Sum of the parameters' log densities per row.  A ``prior`` that takes a whole
        column at once (numpy in, numpy out) is called once per parameter; a scalar-only
        one is called per value, as the reference does."""
        theta = np.atleast_2d(theta)
        total = np.zeros(theta.shape[0])
        for j, p in enumerate(self._p):
            col = None
            try:
                col = np.asarray(p.prior(theta[:, j]), dtype=float)
            except Exception:
                col = None
            if col is None or col.shape != (theta.shape[0],):
                col = np.array([p.prior(x) for x in theta[:, j]], dtype=float)
            total = total + col
        return total

    def propose(self, theta, vcv, rng):
        # R/smc2_parameters.R: one multivariate-normal step per row, reflected at the bounds;
        # all rows at once (with 1024 parameter particles a Python loop over the rows costs
        # more than the GPU work of a replenishment)
        out = rmvnorm_generator(vcv).rows(theta, rng)
        out[:, self._integer] = np.round(out[:, self._integer])
        return reflect_proposal_rows(out, self._min, self._max)

    def model(self, theta):
        return [self._transform(dict(zip(self._names, (float(x) for x in row))))
                for row in np.atleast_2d(theta)]


class smc2_control:
    """R/smc2_control.R:30-40."""

    def __init__(self, n_parameter_sets, degeneracy_threshold=0.5, covariance_scaling=0.5,
                 progress=False, save_trajectories=False, seed=1):
        if n_parameter_sets < 2:
            raise ValueError("'n_parameter_sets' must be at least 2")
        if save_trajectories:
            raise ValueError("save_trajectories is not supported by the batched smc2 engine")
        self.n_parameter_sets = int(n_parameter_sets)
        self.degeneracy_threshold = float(degeneracy_threshold)
        self.covariance_scaling = float(covariance_scaling)
        self.progress = bool(progress)
        self.seed = seed  # host RNG (R: the session's RNG state)


def cov_wt(x, w):
    """``stats::cov.wt(x, wt = w)$cov`` (R/smc2.R:116-118): weights normalised to sum 1,
    unbiased correction 1 / (1 - sum w^2)."""
    w = np.asarray(w, dtype=float)
    w = w / w.sum()
    xc = x - (w[:, None] * x).sum(axis=0)
    return (xc * w[:, None]).T @ xc / (1.0 - np.sum(w ** 2))


class smc2_engine:
    """R/smc2.R:38-220 on batched, optionally sharded, filters."""

    def __init__(self, pars, filter, control, group=None):
        inputs = filter.inputs()
        if inputs["compare"] is not None:
            raise ValueError("the batched smc2 engine needs the compiled compare (compare = NULL)")
        if filter.has_multiple_parameters:
            raise ValueError("smc2 takes a single-parameter filter")
        self._pars, self._control = pars, control
        self._rng = np.random.default_rng(control.seed)
        self._group = group
        self.rank, self.world = 0, 1
        self._dist = None
        try:
            import torch.distributed as dist

            if dist.is_available() and dist.is_initialized():
                self._dist = dist
                self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        except ImportError:
            pass
        n = control.n_parameter_sets
        self._lo, self._hi = block_range(n, self.rank, self.world)
        self._counts = [hi - lo for lo, hi in (block_range(n, r, self.world)
                                               for r in range(self.world))]
        if min(self._counts) < 1:
            raise ValueError("more ranks than parameter sets")

        # R/smc2.R:63-66: sample, prior; every rank draws the same values
        theta = pars.sample(n, self._rng)
        self.pars = theta
        self.log_prior = pars.prior(theta)
        self.log_likelihood = np.zeros(n)
        self.log_posterior = self.log_prior + self.log_likelihood
        self.log_weights = np.zeros(n)
        self.weights = np.ones(n)

        # R/smc2.R:68-78: filter i runs on node i of the distributed seed
        algorithm = getattr(inputs["model"], "algorithm", "xoshiro256plus")
        seeds = dust_rng_distributed_state(inputs["seed"], 1, n, algorithm)
        self._seed_local = b"".join(seeds[self._lo:self._hi])
        self._inputs = inputs
        self._data = inputs["data"]
        self._times = self._data.times
        self._data_split = particle_filter_data_split(self._data, True)
        self.n_particles = inputs["n_particles"]
        self.n_steps = self._times.shape[0]
        # the parameters each local filter really runs with: resample() permutes self.pars
        # but not the filters (Q1), so the two drift apart until a proposal is accepted
        self._model_theta = theta[self._lo:self._hi].copy()
        self._model = self._new_model(pars.model(self._model_theta))
        # the proposal's model handle, allocated once -- here, with the engine, not at the
        # first replenishment (a few hundred MB of device memory: its allocation belongs to
        # setting the engine up, not to a data row) -- and reused by every replenishment
        self._child = self._new_model(pars.model(self._model_theta))
        self._gather_obj = None
        self._on_torch_stream = False
        self._filter_facade = filter

        self.step_current = 0
        self.phase_seconds = {}
        self.acceptance_rate = np.full(self.n_steps, np.nan)
        self.ess = np.full(self.n_steps, np.nan)
        self.n_filter_rows = 0  # model rows x parameter sets run on this rank

    # ---- model handles ------------------------------------------------------------
    def _new_model(self, pars_model, rng_state=None):
        i = self._inputs
        m = i["model"].new(pars=pars_model, time=int(self._times[0, 0]),
                           n_particles=self.n_particles, n_threads=i["n_threads"],
                           seed=self._seed_local, gpu_config=i["gpu_config"], pars_multi=True,
                           seed_per_set=True)
        m.set_data(self._data_split, True)  # one data stream shared by every parameter set
        if i["initial"] is not None:
            raise ValueError("'initial' is not supported by the batched smc2 engine")
        if rng_state is not None:
            m.set_rng_state(rng_state)
        return m

    def _fork(self, pars_model):
        """The proposal's filter: a model at the start time with the proposed parameters and
        the parent's generator state (fork_smc2, R/particle_filter_state.R:400-424).  One
        handle serves every replenishment -- it is reset (F2) and takes the streams by a
        device copy, instead of an allocation, a host round trip of the RNG state and a
        free per proposal."""
        self._child.update_state(pars=pars_model, time=int(self._times[0, 0]),
                                 set_initial_state=True)
        self._child.copy_sets_from(self._model, np.ones(self._hi - self._lo, dtype=bool),
                                   state=False, rng=True)
        return self._child

    def _gatherer(self):
        if self._gather_obj is None:
            from .sharding import likelihood_gather

            self._gather_obj = likelihood_gather(self._counts, self._group)
        return self._gather_obj

    def _gather(self, local):
        if self._dist is None or self.world == 1:
            return np.asarray(local, dtype=float)
        return self._gatherer().gather_host(local)

    def _step_likelihood(self, time_end):
        """One data row on every local filter and the step likelihoods of ALL parameter
        particles (R/smc2.R:91).  Sharded over GPUs the row is queued, gathered and copied
        out on one stream and waited for once."""
        m = self._model
        if self._dist is None or self.world == 1 or not hasattr(m, "log_likelihood_device"):
            return self._gather(m.filter(time_end)["log_likelihood"])
        g = self._gatherer()
        if not g.nccl:
            return g.gather_host(m.filter(time_end)["log_likelihood"])
        import torch

        if not self._on_torch_stream:
            # a stream of this engine's own that torch knows about: the handle launches on
            # it, the copy into the gather buffer, the collective's hand-over and the copy
            # out are ordered on it.  (Not torch's default stream: its handle is 0, which
            # set_stream reads as "the handle's own stream", and graphs cannot be captured
            # on it.)
            self._stream = torch.cuda.Stream()
            m.set_stream(self._stream.cuda_stream)
            self._on_torch_stream = True
        with torch.cuda.stream(self._stream):
            m.filter_enqueue(time_end)
            return g.gather_device(m.log_likelihood_device(), wait=m.filter_collect)

    # ---- R/smc2.R:89-108 ------------------------------------------------------------
    def _lap(self, key, t0):
        """Wall time per phase (filter + gather, host arithmetic, the pieces of a
        replenishment), kept for profiles/tools/c5_phases.py; a perf_counter call each."""
        import time

        now = time.perf_counter()
        self.phase_seconds[key] = self.phase_seconds.get(key, 0.0) + (now - t0)
        return now

    def step(self):
        import time

        t0 = time.perf_counter()
        t = self.step_current + 1
        time_end = int(self._times[t - 1, 1])
        step_ll = self._step_likelihood(time_end)
        t0 = self._lap("row: filter + gather", t0)
        self.n_filter_rows += self._hi - self._lo
        self.log_likelihood = self.log_likelihood + step_ll
        self.log_posterior = self.log_posterior + step_ll
        self.log_weights = self.log_weights + step_ll
        self.weights = scale_log_weights(self.log_weights)["weights"]
        with np.errstate(invalid="ignore", divide="ignore"):
            ess = self.weights.sum() ** 2 / np.sum(self.weights ** 2)
        self.step_current = t
        self.ess[t - 1] = ess
        t0 = self._lap("row: host arithmetic", t0)
        if ess < self._control.degeneracy_threshold * self._control.n_parameter_sets:
            kernel = cov_wt(self.pars, self.weights) * self._control.covariance_scaling
            self.resample()
            t0 = self._lap("replenish: kernel + resample", t0)
            proposed = self.propose(kernel)
            t0 = self._lap("replenish: propose (batched re-run + gather)", t0)
            self.update(proposed)
            self._lap("replenish: update", t0)

    def run(self):
        for _ in range(self.step_current, self.n_steps):
            self.step()

    # ---- R/smc2.R:120-149 (Q1: log-weights in, filters stay where they are) -----------
    def resample(self):
        kappa = particle_resample(self.log_weights, self._rng) - 1
        self.pars = self.pars[kappa]
        self.log_prior = self.log_prior[kappa]
        self.log_likelihood = self.log_likelihood[kappa]
        self.log_posterior = self.log_posterior[kappa]
        self.log_weights = self.log_weights[kappa]

    # ---- R/smc2.R:151-172 + fork_smc2 ---------------------------------------------------
    def propose(self, kernel):
        pars = self._pars.propose(self.pars, kernel, self._rng)
        log_prior = self._pars.prior(pars)
        finite = np.isfinite(log_prior)
        lo, hi = self._lo, self._hi
        # sets with an impossible proposal are not forked (:164-169); the batched child
        # still needs valid parameters in those slots, so they keep the current ones
        run_pars = np.where(finite[lo:hi, None], pars[lo:hi], self._model_theta)
        child = self._fork(self._pars.model(run_pars))
        time_end = int(self._times[self.step_current - 1, 1])
        ll_child = child.filter(time_end)["log_likelihood"]
        self.n_filter_rows += (hi - lo) * self.step_current
        # the parent's generator takes the child's, for the sets that were forked (:420)
        self._model.copy_sets_from(child, finite[lo:hi], state=False, rng=True)
        ll = self._gather(np.where(finite[lo:hi], ll_child, -np.inf))
        return {"pars": pars, "log_prior": log_prior, "log_likelihood": ll,
                "log_posterior": log_prior + ll, "model": child, "finite": finite}

    # ---- R/smc2.R:174-189 -----------------------------------------------------------
    def update(self, proposed):
        with np.errstate(over="ignore", invalid="ignore"):
            pr = np.exp(proposed["log_posterior"] - self.log_posterior)
        accept = self._rng.random(pr.size) < pr
        if accept.any():
            self.pars[accept] = proposed["pars"][accept]
            self.log_prior[accept] = proposed["log_prior"][accept]
            self.log_likelihood[accept] = proposed["log_likelihood"][accept]
            self.log_posterior[accept] = proposed["log_posterior"][accept]
            lo, hi = self._lo, self._hi
            take = accept[lo:hi]
            if take.any():
                # filter[accept] <- proposed$filter[accept]: parameters and particle state
                self._model_theta[take] = proposed["pars"][lo:hi][take]
                self._model.update_state(pars=self._pars.model(self._model_theta),
                                         set_initial_state=False)
                self._model.copy_sets_from(proposed["model"], take, state=True, rng=False)
            self.log_weights[:] = 0.0
        self.acceptance_rate[self.step_current - 1] = accept.mean()

    # ---- R/smc2.R:191-220 -----------------------------------------------------------
    def results(self):
        weight = scale_log_weights(self.log_posterior)["weights"]
        acc = self.acceptance_rate
        effort = sum(1 if np.isnan(a) else i + 2 for i, a in enumerate(acc)) / self.n_steps \
            * self.pars.shape[0]
        return {
            "pars": self.pars.copy(), "pars_names": self._pars.names(),
            "probabilities": np.stack([self.log_prior, self.log_likelihood, self.log_posterior,
                                       weight / weight.sum()], axis=1),
            "statistics": {"ess": self.ess.copy(), "acceptance_rate": acc.copy(),
                           "n_particles": self.n_particles,
                           "n_parameter_sets": self.pars.shape[0], "n_steps": self.n_steps,
                           "effort": effort},
        }


def smc2(pars, filter, control, group=None):
    """``smc2(pars, filter, control)`` (R/smc2.R:29-36)."""
    obj = smc2_engine(pars, filter, control, group)
    obj.run()
    return obj.results()


def smc2_predict(result, n=None, rng=None):
    """``predict.smc2_result`` (R/smc2.R:236-240)."""
    rng = host_rng() if rng is None else rng
    w = result["probabilities"][:, 3]
    i = rng.choice(w.size, size=w.size if n is None else n, replace=True, p=w)
    return result["pars"][i]
