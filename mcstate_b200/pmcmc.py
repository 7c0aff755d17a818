"""pMCMC driving the B200 filter: the caller side of the hot path.

The reference's ``pmcmc()`` is host logic around ``filter$run(pars)``
(R/pmcmc_state.R:101-106): propose -> prior -> uniform -> filter -> accept
(:108-138).  This module restates that control flow so the B200 filter can be
driven exactly as mcstate drives its own (BASELINE config C3), and adds the
B200 analogue of ``pmcmc_parallel`` (R/pmcmc_parallel.R:8-20): chains are
sharded one per GPU / rank with ``torch.distributed``, seeded by long jumps so a
result does not depend on where a chain ran (R/pmcmc_parallel.R:125-151), and
the per-chain samples are gathered at the end (``pmcmc_combine``,
R/pmcmc_tools.R:104-156).  There is no data-path collective: a chain's filter
never leaves its GPU.

Only what the path needs is here: simple (non-nested) parameters with a
multivariate-normal proposal and reflecting bounds.  Nested parameter
containers, adaptive proposals and the result-thinning helpers are out of scope
(SURVEY.md section 2, rows 9, 10, 13).  R's global RNG (``runif``, ``rnorm``,
``sample.int``) becomes a ``numpy.random.Generator`` seeded where the reference
calls ``set.seed`` (R/pmcmc.R:75-81).
"""
from __future__ import annotations

import math
import os

import numpy as np

from ._rng import host_rng
from .dust import dust_rng, dust_rng_distributed_state
from .particle_filter import particle_filter_from_inputs


# ---------------------------------------------------------------------------
# parameters (R/pmcmc_parameters.R)
# ---------------------------------------------------------------------------
class pmcmc_parameter:
    """``pmcmc_parameter(name, initial, min, max, discrete, integer, prior)``
    (R/pmcmc_parameters.R:56-108)."""

    def __init__(self, name, initial, min=-math.inf, max=math.inf, discrete=False,
                 integer=False, prior=None, mean=None):
        if discrete:
            import warnings

            warnings.warn("'discrete' is deprecated.\nUse 'integer' instead.", DeprecationWarning,
                          stacklevel=2)
        if initial < min:
            raise ValueError("'initial' must be >= 'min' (%g)" % min)
        if initial > max:
            raise ValueError("'initial' must be <= 'max' (%g)" % max)
        self.name, self.initial, self.min, self.max = name, float(initial), float(min), float(max)
        self.mean = self.initial if mean is None else float(mean)
        self.integer = bool(discrete or integer)
        self.prior = prior if prior is not None else (lambda p: 0.0)
        try:
            value = self.prior(self.initial)
        except Exception as e:
            raise ValueError("Prior function for '%s' failed to evaluate initial value: %s"
                             % (name, e)) from None
        if not math.isfinite(value):
            raise ValueError("Prior function for '%s' returned a non-finite value on initial "
                             "value" % name)


def reflect_proposal(x, x_min, x_max):
    """R/pmcmc_parameters.R:382-404: reflect a proposal at its bounds."""
    x = np.array(x, dtype=float)
    x_min, x_max = np.asarray(x_min, dtype=float), np.asarray(x_max, dtype=float)
    out = x < x_min
    out |= x > x_max
    if out.any():
        fmin, fmax = np.isfinite(x_min), np.isfinite(x_max)
        both, lo, hi = out & fmin & fmax, out & fmin & ~fmax, out & ~fmin & fmax
        xr = x_max[both] - x_min[both]
        x[both] = np.abs(np.mod(x[both] + xr - x_min[both], 2 * xr) - xr) + x_min[both]
        x[lo] = 2 * x_min[lo] - x[lo]
        x[hi] = 2 * x_max[hi] - x[hi]
    return x


def rmvnorm_generator(vcv):
    """R/utils.R:37-75: draws from N(mean, scale * vcv) by eigendecomposition, skipping
    parameters whose row of ``vcv`` is all zero."""
    vcv = np.asarray(vcv, dtype=float)
    if vcv.ndim != 2 or vcv.shape[0] != vcv.shape[1] or not np.allclose(vcv, vcv.T, atol=1.5e-8):
        raise ValueError("vcv must be symmetric")
    include = ~np.all(vcv == 0, axis=1)
    sub = vcv[np.ix_(include, include)]
    values, vectors = np.linalg.eigh(sub) if sub.size else (np.zeros(0), np.zeros((0, 0)))
    if values.size and not np.all(values >= -math.sqrt(np.finfo(float).eps) * abs(values[-1])):
        raise ValueError("vcv must be positive definite")
    res = (vectors * np.sqrt(np.maximum(values, 0))) @ vectors.T

    def draw(mean, rng, scale=1.0):
        x = np.zeros(include.size)
        x[include] = rng.standard_normal(res.shape[0]) @ (res * math.sqrt(scale))
        return np.asarray(mean, dtype=float) + x

    def draw_rows(means, rng, scale=1.0):
        """One draw per row of ``means`` in one go: the normal deviates in the order of
        ``len(means)`` successive ``draw`` calls (so the generator ends in the same state),
        the linear map term by term."""
        means = np.atleast_2d(np.asarray(means, dtype=float))
        z = rng.standard_normal((means.shape[0], res.shape[0]))
        r = res * math.sqrt(scale)
        x = np.zeros(means.shape)
        acc = np.zeros((means.shape[0], res.shape[0]))
        for j in range(res.shape[0]):
            acc += z[:, j:j + 1] * r[j]
        x[:, include] = acc
        return means + x

    draw.rows = draw_rows
    return draw


def reflect_proposal_rows(x, x_min, x_max):
    """``reflect_proposal`` for a matrix of proposals, one per row (same formulas)."""
    x = np.array(x, dtype=float)
    lo_b = np.broadcast_to(np.asarray(x_min, dtype=float), x.shape)
    hi_b = np.broadcast_to(np.asarray(x_max, dtype=float), x.shape)
    out = (x < lo_b) | (x > hi_b)
    if out.any():
        fmin, fmax = np.isfinite(lo_b), np.isfinite(hi_b)
        both, lo, hi = out & fmin & fmax, out & fmin & ~fmax, out & ~fmin & fmax
        xr = hi_b[both] - lo_b[both]
        x[both] = np.abs(np.mod(x[both] + xr - lo_b[both], 2 * xr) - xr) + lo_b[both]
        x[lo] = 2 * lo_b[lo] - x[lo]
        x[hi] = 2 * hi_b[hi] - x[hi]
    return x


class pmcmc_parameters:
    """``pmcmc_parameters$new(parameters, proposal, transform)`` (R/pmcmc_parameters.R)."""

    def __init__(self, parameters, proposal, transform=None):
        given_names = None
        if isinstance(parameters, dict):          # R: a named list
            given_names, parameters = list(parameters), list(parameters.values())
        if not isinstance(parameters, (list, tuple)):
            raise TypeError("'parameters' must be a list")
        if len(parameters) == 0:
            raise ValueError("At least one parameter is required")
        if not all(isinstance(p, pmcmc_parameter) for p in parameters):
            raise TypeError("Expected all elements of '...' to be 'pmcmc_parameter' objects")
        self._p = list(parameters)
        names = [p.name for p in self._p]
        dup = [n for k, n in enumerate(names) if n in names[:k]]
        if dup:
            raise ValueError("Duplicate parameter names: %s"
                             % ", ".join("'%s'" % n for n in dict.fromkeys(dup)))
        if given_names is not None and given_names != names:
            raise ValueError("'parameters' is named, but the names do not match parameters")
        if hasattr(proposal, "index") and hasattr(proposal, "columns"):   # pandas.DataFrame
            if list(proposal.index) != names or list(proposal.columns) != names:
                raise ValueError("Expected dimension names of 'proposal' to match parameters")
        proposal = np.asarray(proposal, dtype=float)
        if proposal.shape != (len(names), len(names)):
            raise ValueError("Expected a square proposal matrix with %d rows and columns"
                             % len(names))
        self._names = names
        self._vcv = proposal.copy()
        self._proposal = rmvnorm_generator(proposal)
        self._min = np.array([p.min for p in self._p])
        self._max = np.array([p.max for p in self._p])
        self._integer = np.array([p.integer for p in self._p])
        self._transform = transform if transform is not None else (lambda d: d)

    def names(self):
        return list(self._names)

    def initial(self):
        return np.array([p.initial for p in self._p])

    def prior(self, theta):
        """By position, not by name (a dict's values are taken in its own order)."""
        if isinstance(theta, dict):
            theta = list(theta.values())
        return float(sum(p.prior(t) for p, t in zip(self._p, theta)))

    def propose(self, theta, rng, scale=1.0):
        theta = self._proposal(theta, rng, scale)
        theta[self._integer] = np.round(theta[self._integer])
        return reflect_proposal(theta, self._min, self._max)

    def model(self, theta):
        """theta -> what the filter's ``run`` takes (a dict of named values, transformed)."""
        return self._transform(dict(zip(self._names, (float(t) for t in theta))))

    def mean(self):
        return np.array([p.mean for p in self._p])

    def vcv(self):
        return self._vcv.copy()

    def summary(self):
        """R/pmcmc_parameters.R:297-303: one record per parameter."""
        return [{"name": p.name, "min": p.min, "max": p.max, "discrete": p.integer,
                 "integer": p.integer} for p in self._p]

    def fix(self, fixed):
        """``pars$fix(c(name = value, ...))`` (R/pmcmc_parameters.R:355-377): a parameter set
        over the remaining parameters; the fixed values re-enter in ``model()``, ahead of
        the original transform."""
        if not isinstance(fixed, dict):
            raise ValueError("'fixed' must be named")
        unknown = [n for n in fixed if n not in self._names]
        if unknown:
            raise ValueError("Fixed parameters not found in model: %s"
                             % ", ".join("'%s'" % n for n in unknown))
        if len(fixed) == len(self._names):
            raise ValueError("Cannot fix all parameters")
        vary = [i for i, n in enumerate(self._names) if n not in fixed]
        names, base_transform = list(self._names), self._transform
        values = {n: float(v) for n, v in fixed.items()}

        def transform(d):
            return base_transform({n: values[n] if n in values else d[n] for n in names})

        return pmcmc_parameters([self._p[i] for i in vary], self._vcv[np.ix_(vary, vary)],
                                transform)


# ---------------------------------------------------------------------------
# control (R/pmcmc_control.R:187-290, the fields this path reads)
# ---------------------------------------------------------------------------
def pmcmc_filter_on_generation(n_steps, n_burnin=None, n_steps_retain=None):
    """R/pmcmc_control.R:295-332: which steps a chain keeps -- every ``n_steps_every``-th
    one, counted back from the LAST step, ``n_steps_retain`` of them; ``n_burnin`` is then
    back-calculated (it can only grow)."""
    n_burnin = 0 if n_burnin is None else n_burnin
    if int(n_burnin) != n_burnin or n_burnin < 0:
        raise ValueError("'n_burnin' must be a non-negative integer")
    if n_burnin >= n_steps:
        raise ValueError("'n_burnin' cannot be greater than or equal to 'n_steps'")
    possible = int(n_steps - n_burnin)
    n_steps_retain = possible if n_steps_retain is None else n_steps_retain
    if int(n_steps_retain) != n_steps_retain or n_steps_retain < 1:
        raise ValueError("'n_steps_retain' must be at least 1")
    n_steps_retain = int(n_steps_retain)
    if n_steps_retain > possible:
        raise ValueError("'n_steps_retain' is too large, max possible is %d but given %d"
                         % (possible, n_steps_retain))
    every = possible // n_steps_retain
    if every == 1 and (possible - n_steps_retain) / possible > 0.05:
        raise ValueError("'n_steps_retain' is too large to skip any samples, and would result "
                         "in just increasing 'n_burnin' by more than 5% of your post-burnin "
                         "samples. Please adjust 'n_steps' 'n_burnin' or 'n_steps_retain' to "
                         "make your intentions clearer")
    return {"n_burnin": int(n_steps - every * (n_steps_retain - 1) - 1),
            "n_steps_retain": n_steps_retain, "n_steps_every": every}


def pmcmc_check_control(control):
    """R/pmcmc_control.R:335-343."""
    if control.n_steps != control.n_burnin + (control.n_steps_retain - 1) * control.n_steps_every + 1:
        raise ValueError("Corrupt pmcmc_control (n_steps/n_steps_retain/n_burnin), "
                         "perhaps you modified it after creation?")


class pmcmc_control:
    """``pmcmc_control(...)`` (R/pmcmc_control.R:187-290, the fields this path reads).
    ``n_workers > 1`` shards the chains over the ranks of the process group (one GPU each);
    ``n_threads_total`` is validated as the reference does when given, but not demanded
    with workers: the filter's parallelism is the GPU's, not host threads."""

    def __init__(self, n_steps, n_chains=1, n_burnin=None, n_steps_retain=None, save_state=True,
                 save_restart=None, save_trajectories=False, rerun_every=math.inf,
                 rerun_random=False, filter_early_exit=False, use_parallel_seed=False,
                 n_workers=1, progress=False, n_threads_total=None, path=None):
        if n_steps < 1 or n_chains < 1 or n_workers < 1:
            raise ValueError("'n_steps', 'n_chains' and 'n_workers' must be at least 1")
        if n_threads_total is not None:
            if int(n_threads_total) != n_threads_total or n_threads_total < 1:
                raise ValueError("'n_threads_total' must be at least 1")
            if n_threads_total < n_workers:
                raise ValueError("'n_threads_total' (%d) is less than 'n_workers' (%d)"
                                 % (n_threads_total, n_workers))
            if n_threads_total % n_workers != 0:
                raise ValueError("'n_threads_total' (%d) is not a multiple of 'n_workers' (%d)"
                                 % (n_threads_total, n_workers))
        if n_chains < n_workers:
            raise ValueError("'n_chains' (%d) is less than 'n_workers' (%d)"
                             % (n_chains, n_workers))
        if path is not None and n_workers == 1:
            import warnings

            warnings.warn("'path' given when n_workers = 1 has no effect and is ignored")
        self.n_steps, self.n_chains = int(n_steps), int(n_chains)
        self.n_threads_total = None if n_threads_total is None else int(n_threads_total)
        self.path = path
        # keep n_steps_retain steps, every n_steps_every-th, ending at the last one
        # (R/pmcmc_control.R:295-332)
        gen = pmcmc_filter_on_generation(self.n_steps, n_burnin, n_steps_retain)
        self.n_burnin, self.n_steps_retain = gen["n_burnin"], gen["n_steps_retain"]
        self.n_steps_every = gen["n_steps_every"]
        self.save_state, self.save_trajectories = bool(save_state), bool(save_trajectories)
        if save_restart is not None:
            save_restart = [save_restart] if np.isscalar(save_restart) else list(save_restart)
            if any(b <= a for a, b in zip(save_restart, save_restart[1:])):
                raise ValueError("'save_restart' must be strictly increasing")
        self.save_restart = save_restart
        self.rerun_every, self.rerun_random = rerun_every, bool(rerun_random)
        self.filter_early_exit = bool(filter_early_exit)
        # more than one worker always uses the parallel seeds (R/pmcmc_control.R:260-262)
        self.use_parallel_seed = bool(use_parallel_seed) or n_workers > 1
        self.n_workers, self.progress = int(n_workers), bool(progress)


# ---------------------------------------------------------------------------
# seeds (R/pmcmc_parallel.R:125-151)
# ---------------------------------------------------------------------------
def make_seeds(n, seed, algorithm="xoshiro256plus"):
    """One ``dict(dust=<raw 32 k bytes>, r=<int>)`` per chain.  Chain k's filter streams
    are the base state long-jumped k times (chain 1 = the input, prefix-stable in ``n``:
    tests/testthat/test-pmcmc-parallel.R:40-60); the host-RNG seeds are
    ``ceiling(random_real * 2^24)`` from one more long-jumped copy."""
    n_streams = len(bytes(seed)) // 32 if isinstance(seed, (bytes, bytearray)) else 1
    seed_dust = dust_rng_distributed_state(seed, n_streams, n, algorithm)
    rng = dust_rng(seed_dust[0], algorithm).long_jump()
    seed_r = np.ceil(np.asarray(rng.random_real(n)) * 2 ** 24).astype(np.int64)
    return [{"dust": seed_dust[i], "r": int(seed_r[i])} for i in range(n)]


# ---------------------------------------------------------------------------
# one chain (R/pmcmc_state.R)
# ---------------------------------------------------------------------------
def _make_rerun(every, random, rng):
    if not math.isfinite(every):
        return lambda i: False
    if random:
        return lambda i: rng.random() < 1.0 / every
    return lambda i: i % every == 0


class pmcmc_state:
    """R/pmcmc_state.R: the Metropolis-Hastings state machine of one chain."""

    def __init__(self, pars, initial, filter, control, rng):
        if getattr(filter, "has_multiple_parameters", False):
            if not getattr(filter, "has_multiple_data", False):    # R/pmcmc_state.R:220-223
                raise ValueError("Can't use a filter with multiple parameter sets but not "
                                 "multiple data")
            # R/pmcmc_state.R:225-227: these are simple (non-nested) parameters
            raise ValueError("'pars' and 'filter' disagree on nestedness (nested filters need "
                             "pmcmc_parameters_nested: out of scope here)")
        self._filter, self._pars, self._control, self._rng = filter, pars, control, rng
        self.n_filter_runs = 0
        self._curr_pars = np.array(initial, dtype=float)
        self._curr_lprior = pars.prior(self._curr_pars)
        self._curr_llik = self._run_filter(self._curr_pars)
        self._curr_lpost = self._curr_lprior + self._curr_llik
        self._curr_state = self._curr_traj = self._curr_restart = None
        self._update_particle_history()
        self._h_pars, self._h_prob = [], []
        self._h_state, self._h_traj, self._h_restart = [], [], []

    def _run_filter(self, theta, min_log_likelihood=-math.inf):
        # R/pmcmc_state.R:101-106 -- the hot path, once per step
        self.n_filter_runs += 1
        c = self._control
        return float(self._filter.run(self._pars.model(theta), c.save_trajectories,
                                      c.save_restart, min_log_likelihood))

    def _update_particle_history(self):
        # R/pmcmc_state.R:32-51: ONE particle's lineage leaves the device
        c = self._control
        i = int(self._rng.integers(1, self._filter.n_particles + 1))
        if c.save_trajectories:
            self._curr_traj = np.asarray(self._filter.history(i))[:, 0]
        if c.save_state:
            if hasattr(self._filter, "state_particles"):
                self._curr_state = np.asarray(self._filter.state_particles(i))[:, 0]
            else:
                self._curr_state = np.asarray(self._filter.state())[:, i - 1]
        if c.save_restart is not None and len(c.save_restart) > 0:
            self._curr_restart = np.asarray(self._filter.restart_state(i, c.save_restart))[:, 0]

    def _min_log_likelihood(self, prop_lprior, u):
        # u < exp(prop_llik + prop_lprior - curr_lpost)   (R/pmcmc_state.R:85-99)
        if not self._control.filter_early_exit:
            return -math.inf
        return self._curr_lpost - prop_lprior + math.log(u)

    def _update_simple(self):
        # R/pmcmc_state.R:108-138
        prop_pars = self._pars.propose(self._curr_pars, self._rng)
        prop_lprior = self._pars.prior(prop_pars)
        u = self._rng.random()
        prop_llik = self._run_filter(prop_pars, self._min_log_likelihood(prop_lprior, u))
        prop_lpost = prop_lprior + prop_llik
        d = prop_lpost - self._curr_lpost
        accept_prob = 1.0 if d >= 0 else (math.exp(d) if d == d else 0.0)
        if u < accept_prob:
            self._curr_pars, self._curr_lprior = prop_pars, prop_lprior
            self._curr_llik, self._curr_lpost = prop_llik, prop_lpost
            self._update_particle_history()

    def _record(self, i):
        # R/pmcmc_state.R:53-84
        c = self._control
        self._h_pars.append(self._curr_pars.copy())
        self._h_prob.append((self._curr_lprior, self._curr_llik, self._curr_lpost))
        j = i - c.n_burnin - 1
        if j >= 0 and j % c.n_steps_every == 0:
            if c.save_trajectories:
                self._h_traj.append(self._curr_traj)
            if c.save_state:
                self._h_state.append(self._curr_state)
            if self._curr_restart is not None:
                self._h_restart.append(self._curr_restart)

    def run(self):
        # R/pmcmc_state.R:284-308
        c = self._control
        pmcmc_check_control(c)
        rerun = _make_rerun(c.rerun_every, c.rerun_random, self._rng)
        for i in range(1, c.n_steps + 1):
            if rerun(i):
                self._curr_llik = self._run_filter(self._curr_pars)
                self._curr_lpost = self._curr_lprior + self._curr_llik
                self._update_particle_history()
            self._update_simple()
            self._record(i)

    def finish(self):
        # R/pmcmc_state.R:310-428: keep every n_steps_every-th step after the burn-in
        c = self._control
        keep = [i for i in range(c.n_steps) if i - c.n_burnin >= 0
                and (i - c.n_burnin) % c.n_steps_every == 0]
        out = {
            "pars": np.array(self._h_pars)[keep],
            "probabilities": np.array(self._h_prob)[keep],
            # every step, kept or not (the reference's results$pars_full: R/pmcmc_utils.R:100-121)
            "pars_full": np.array(self._h_pars),
            "probabilities_full": np.array(self._h_prob),
            "pars_names": self._pars.names(),
            "iteration": np.array(keep) + 1,
            "chain": None,
            "state": np.stack(self._h_state, axis=1) if self._h_state else None,
            "trajectories": np.stack(self._h_traj, axis=1) if self._h_traj else None,
            "restart": np.stack(self._h_restart, axis=1) if self._h_restart else None,
            "n_filter_runs": self.n_filter_runs,
            "predict": None,
        }
        if c.save_state or c.save_trajectories:
            # R/pmcmc_state.R:336-353: what pmcmc_predict needs to carry the samples forward
            data = self._filter.inputs()["data"]
            time = int(data.times[-1, 1])
            last_history = getattr(self._filter, "_last_history", None)
            out["predict"] = {"is_continuous": False, "transform": self._pars.model,
                              "index": None if last_history is None else last_history["index"],
                              "time": time, "rate": data.rate, "model_time": time / data.rate,
                              "filter": self._filter.inputs()}
            if c.save_trajectories:
                # the model steps of the saved trajectories (R/pmcmc_state.R:386-396)
                out["trajectories_time"] = np.concatenate([data.times[:1, 0], data.times[:, 1]])
                # rownames(results$trajectories$state): the names of a named index$state
                # (tests/testthat/test-pmcmc.R:239-267), None for an unnamed one
                idx = out["predict"]["index"]
                out["trajectories_names"] = list(idx) if isinstance(idx, dict) else None
        return out


def pmcmc_run_chain(pars, initial, filter, control, rng, n_threads=None):
    """R/pmcmc.R:93-102: the thread count given, else the control's share per worker, is
    set on the filter before the chain starts."""
    if hasattr(filter, "set_n_threads"):
        if n_threads is not None:
            filter.set_n_threads(n_threads)
        elif getattr(control, "n_threads_total", None) is not None:
            filter.set_n_threads(control.n_threads_total // control.n_workers)
    obj = pmcmc_state(pars, initial, filter, control, rng)
    obj.run()
    return obj.finish()


def pmcmc_combine(*args, samples=None):
    """``pmcmc_combine(..., samples = list(...))`` (R/pmcmc_tools.R:104-190): stack chains,
    remembering which sample came from which; the chains must agree in length, iterations,
    parameters and in what they saved."""
    if samples is None:
        samples = list(args[0]) if len(args) == 1 and isinstance(args[0], (list, tuple)) \
            else list(args)
    elif isinstance(samples, dict):
        samples = list(samples.values())        # names are ignored
    elif not isinstance(samples, (list, tuple)):
        raise TypeError("'samples' must be a list")
    samples = list(samples)
    if any(not (isinstance(x, dict) and "pars" in x and "iteration" in x) for x in samples):
        raise TypeError("Elements of '...' must be all be 'mcstate_pmcmc' objects")
    if len(samples) < 2:
        raise ValueError("At least 2 samples objects must be provided")
    if any(x.get("chain") is not None for x in samples):
        raise ValueError("Chains have already been combined")
    if len({np.asarray(x["pars"]).shape[0] for x in samples}) != 1:
        raise ValueError("All chains must have the same length")
    if any(x["pars_names"] != samples[0]["pars_names"]
           or np.asarray(x["pars"]).shape[1] != np.asarray(samples[0]["pars"]).shape[1]
           for x in samples):
        raise ValueError("All parameters must have the same dimension names")
    if any(not np.array_equal(x["iteration"], samples[0]["iteration"]) for x in samples):
        raise ValueError("All chains must have the same iterations")
    for key in ("state", "trajectories", "restart"):
        have = [x.get(key) is not None for x in samples]
        if any(have) and not all(have):
            raise ValueError("If '%s' is present for any samples, it must be present for all"
                             % key)
    if any(not np.array_equal(x.get("trajectories_time"), samples[0].get("trajectories_time"))
           for x in samples):
        raise ValueError("trajectories data is inconsistent")
    out = {"pars_names": samples[0]["pars_names"]}
    for key in ("pars", "probabilities", "iteration"):
        out[key] = np.concatenate([s[key] for s in samples], axis=0)
    out["chain"] = np.concatenate([np.full(np.asarray(s["pars"]).shape[0], k + 1)
                                   for k, s in enumerate(samples)])
    for key in ("pars_full", "probabilities_full"):
        parts = [s.get(key) for s in samples]
        out[key] = None if any(p is None for p in parts) else np.concatenate(parts, axis=0)
    for key in ("state", "trajectories", "restart"):
        parts = [s.get(key) for s in samples]
        out[key] = None if any(p is None for p in parts) else np.concatenate(parts, axis=1)
    out["n_filter_runs"] = int(sum(s.get("n_filter_runs", 0) for s in samples))
    # the last chain's, as R/pmcmc_tools.R:148-152
    out["predict"] = samples[-1].get("predict")
    if samples[-1].get("trajectories_time") is not None:
        out["trajectories_time"] = samples[-1]["trajectories_time"]
        out["trajectories_names"] = samples[-1].get("trajectories_names")
    return out


def pmcmc_filter(object, i):
    """R/pmcmc_tools.R:64-90: keep the samples ``i`` (a boolean mask or 0-based indices)."""
    i = np.asarray(i)
    out = dict(object)
    if out.get("chain") is not None:
        out["chain"] = np.asarray(out["chain"])[i]
    out["iteration"] = np.asarray(out["iteration"])[i]
    out["pars"] = np.asarray(out["pars"])[i]
    out["probabilities"] = np.asarray(out["probabilities"])[i]
    for key in ("state", "trajectories", "restart"):     # samples along the second axis
        if out.get(key) is not None:
            out[key] = np.asarray(out[key])[:, i]
    # the full chain does not survive a filter (R/pmcmc_tools.R:85-86)
    out["pars_full"] = out["probabilities_full"] = None
    return out


def pmcmc_thin(object, burnin=None, thin=None):
    """``pmcmc_thin(object, burnin, thin)`` (R/pmcmc_tools.R:23-45): drop every sample
    whose iteration is at most ``burnin``, then keep every ``thin``-th of the rest (per
    chain: the iteration numbers decide, not the positions)."""
    it = np.asarray(object["iteration"])
    keep = np.ones(it.size, dtype=bool)
    if burnin is not None:
        if int(burnin) != burnin or burnin < 0:
            raise ValueError("'burnin' must be a positive integer")
        burnin_max = int(it.max())
        if burnin >= burnin_max:
            raise ValueError("'burnin' must be less than %d for your results" % burnin_max)
        keep &= it > burnin
    if thin is not None:
        if int(thin) != thin or thin < 1:
            raise ValueError("'thin' must be a positive integer")
        offset = it[keep].min()
        keep &= (it - offset) % int(thin) == 0
    return pmcmc_filter(object, keep)


def pmcmc_sample(object, n_sample, burnin=None, rng=None):
    """``pmcmc_sample(object, n_sample, burnin)`` (R/pmcmc_tools.R:53-61): ``n_sample``
    samples drawn with replacement from what is left after ``burnin``."""
    object = pmcmc_thin(object, burnin)
    gen = host_rng() if rng is None else rng
    return pmcmc_filter(object, gen.integers(0, np.asarray(object["pars"]).shape[0], int(n_sample)))


def pmcmc_predict(object, times, prepend_trajectories=False, n_threads=None, seed=None):
    """``pmcmc_predict`` (R/predict.R:36-89): run the model forward from every retained
    sample's final state with that sample's parameters -- one multi-parameter model
    object, one particle per sample (``model$new(pars, times[1], NULL, pars_multi =
    TRUE)``, ``update_state(state = state)``, ``simulate(times)``).  On the engine this
    is one launch of the run kernel per requested time over all samples, the saved
    rows collected in HBM and copied out once.

    Returns ``dict(time, rate, state [n_index, n_samples, n_times], predicted)``
    (``mcstate_trajectories_discrete``, R/trajectories.R:1-8)."""
    pred = object.get("predict")
    if pred is None:
        raise ValueError("mcmc was run with return_state = FALSE, can't predict")
    times = [int(t) for t in times]
    if len(times) < 2:
        raise ValueError("At least two times required for predict")
    if times[0] != pred["time"]:
        raise ValueError("Expected times[1] to be %d" % pred["time"])
    if prepend_trajectories and object.get("trajectories") is None:
        raise ValueError("mcmc was run with return_trajectories = FALSE, can't prepend trajectories")
    state = object.get("state")
    if state is None:
        raise ValueError("mcmc was run with return_state = FALSE, can't predict")
    inputs = pred["filter"]
    pars = [pred["transform"](theta) for theta in np.asarray(object["pars"])]
    nt = inputs["n_threads"] if n_threads is None else n_threads
    mod = inputs["model"].new(pars, times[0], None, n_threads=nt, seed=seed, pars_multi=True)
    # state is [n_state, n_samples]; the model's layout is [n_state, 1 particle, n_pars]
    mod.update_state(state=np.asarray(state)[:, None, :])
    if pred["index"] is not None:
        mod.set_index(pred["index"])
    y = np.asarray(mod.simulate(times))           # [n_index, 1, n_samples, n_times]
    y = y.reshape(y.shape[0], -1, y.shape[-1])
    # rownames(y$state): names are copied from the index (tests/testthat/test-predict.R:129-141)
    names = list(pred["index"]) if isinstance(pred["index"], dict) else None
    res = {"time": np.array(times), "rate": pred["rate"], "state": y,
           "predicted": np.ones(len(times), dtype=bool), "names": names}
    if prepend_trajectories:
        a_time = np.asarray(object["trajectories_time"])
        a_state = np.asarray(object["trajectories"])
        if a_time[-1] != res["time"][0] or a_state.shape[:2] != y.shape[:2]:
            raise ValueError("trajectories and predictions do not line up")
        # bind_mcstate_trajectories_discrete (R/trajectories.R:28-45)
        res = {"time": np.concatenate([a_time, res["time"][1:]]), "rate": pred["rate"],
               "state": np.concatenate([a_state, y[:, :, 1:]], axis=2),
               "predicted": np.concatenate([np.zeros(a_time.size, dtype=bool),
                                            res["predicted"][1:]]), "names": names}
    return res


def pmcmc_check_initial(initial, pars, n_chains):
    """``pmcmc_check_initial`` (R/pmcmc.R:112-154): None -> the parameters' own starting
    point; a vector (or a dict, R's named vector) of ``n_pars`` values is used for every
    chain; a matrix is ``[n_pars, n_chains]``, one COLUMN per chain as in R.  Returns the
    per-chain starting points ``[n_chains, n_pars]`` (R: a list with one vector per chain)."""
    nms = pars.names()
    n_pars = len(nms)
    if initial is None:
        initial = pars.initial()
    if isinstance(initial, dict):
        if list(initial) != nms:
            raise ValueError("Expected names of 'initial' to match parameters (%s)"
                             % ", ".join("'%s'" % n for n in nms))
        initial = [initial[n] for n in nms]
    initial = np.asarray(initial, dtype=float)
    if initial.ndim == 2:
        if initial.shape != (n_pars, n_chains):
            raise ValueError("Expected 'initial' to be a matrix with dimensions %d x %d"
                             % (n_pars, n_chains))
        initial = initial.T.copy()
    else:
        if initial.shape != (n_pars,):
            raise ValueError("Expected 'initial' to be a vector with length %d" % n_pars)
        initial = np.tile(initial[None, :], (n_chains, 1))
    bad = [k + 1 for k, row in enumerate(initial) if not math.isfinite(pars.prior(row))]
    if bad:
        raise ValueError("Starting point does not have finite prior probability (chain %s)"
                         % ", ".join(map(str, bad)))
    return initial


_check_initial = pmcmc_check_initial


def _chain_from_inputs(inputs, seed, pars, initial_k, control, device=None, n_threads=None):
    """One chain from a filter's inputs: a fresh filter on the chain's own streams and a
    host RNG seeded as ``set.seed(seed$r)`` (R/pmcmc.R:75-81, R/pmcmc_chains.R:86-93)."""
    inputs = dict(inputs)
    if device is not None and inputs.get("gpu_config") is not None:
        inputs["gpu_config"] = device
    f = particle_filter_from_inputs(inputs, seed["dust"])
    return pmcmc_run_chain(pars, initial_k, f, control, np.random.default_rng(seed["r"]),
                           n_threads)


def _run_one(k, pars, initial, filter, control, seeds, device):
    """Chain k (0-based) on ``device``."""
    if seeds is None:
        return pmcmc_run_chain(pars, initial[k], filter, control, host_rng())
    return _chain_from_inputs(filter.inputs(), seeds[k], pars, initial[k], control, device)


def pmcmc(pars, filter, initial=None, control=None):
    """``pmcmc(pars, filter, initial, control)`` (R/pmcmc.R:47-58).

    ``control.n_workers == 1``: chains run one after another on this process's device
    (``pmcmc_multiple_series``, R/pmcmc.R:61-90).  ``n_workers > 1``: the chains are
    sharded over the ranks of the initialised ``torch.distributed`` group, rank r taking
    chains r, r + world, ... on its own GPU (``pmcmc_sharded``)."""
    if control is None:
        raise TypeError("'control' must be a pmcmc_control")
    pmcmc_check_control(control)
    if control.n_workers > 1:
        return pmcmc_sharded(pars, filter, initial, control)
    initial = _check_initial(initial, pars, control.n_chains)
    seeds = None
    if control.use_parallel_seed:
        seeds = make_seeds(control.n_chains, filter.inputs()["seed"],
                           getattr(filter.model, "algorithm", "xoshiro256plus"))
    samples = [_run_one(k, pars, initial, filter, control, seeds, None)
               for k in range(control.n_chains)]
    return samples[0] if control.n_chains == 1 else pmcmc_combine(samples=samples)


def pmcmc_parallel_threads(n_threads_total, n_workers, n_chains):
    """R/pmcmc_parallel.R:199-209: threads per chain when ``n_workers`` run at a time --
    an even split while every worker is busy, the spare threads shared by the last
    ``n_chains % n_workers`` chains.  (The engine's parallelism is the GPU's; the counts
    still size host-side work such as ``pmcmc_predict``.)"""
    out = [n_threads_total / n_workers] * (n_chains // n_workers * n_workers)
    spare = n_chains % n_workers
    if spare:
        each = n_threads_total // spare
        out += [each] * (spare - 1) + [each + n_threads_total % spare]
    return [int(x) if float(x).is_integer() else x for x in out]


def chain_owner(chain, world):
    """Which rank runs chain ``chain`` (0-based): round robin, like the reference hands
    the next chain to the next free worker (R/pmcmc_parallel.R:71-102)."""
    return chain % world


def pmcmc_sharded(pars, filter, initial=None, control=None, group=None, device=None):
    """The B200 ``pmcmc_parallel``: one chain per GPU at a time, no collective while the
    chains run, one gather of the per-chain samples at the end.  Every rank returns the
    combined result, identical to the serial run with ``use_parallel_seed``
    (tests/testthat/test-pmcmc-parallel.R:3-25)."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError("pmcmc with n_workers > 1 needs an initialised torch.distributed "
                           "process group (one rank per GPU)")
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    initial = _check_initial(initial, pars, control.n_chains)
    seeds = make_seeds(control.n_chains, filter.inputs()["seed"],
                       getattr(filter.model, "algorithm", "xoshiro256plus"))
    mine = {k: _run_one(k, pars, initial, filter, control, seeds, device)
            for k in range(control.n_chains) if chain_owner(k, world) == rank}
    # `predict` holds callables and the model generator: send only what differs by chain
    # (the filter's live seed), every rank rebuilds the rest from its own pars / filter
    for smp in mine.values():
        if smp.get("predict") is not None:
            smp["predict"] = {"seed": smp["predict"]["filter"]["seed"],
                              **{k: smp["predict"][k] for k in
                                 ("is_continuous", "index", "time", "rate", "model_time")}}
    gathered = [None] * world
    dist.all_gather_object(gathered, mine, group=group)
    merged = {}
    for part in gathered:
        merged.update(part)
    samples = [merged[k] for k in range(control.n_chains)]
    for smp in samples:
        if smp.get("predict") is not None:
            seed = smp["predict"].pop("seed")
            smp["predict"].update(transform=pars.model, filter=dict(filter.inputs(), seed=seed))
    return samples[0] if control.n_chains == 1 else pmcmc_combine(samples=samples)


# ---------------------------------------------------------------------------
# chains run by hand, one process (and GPU) each, through files (R/pmcmc_chains.R)
# ---------------------------------------------------------------------------
def _serializer():
    """R's saveRDS keeps closures (priors, index and compare functions); the counterpart
    here is cloudpickle, with plain pickle (module-level functions only) as the fallback."""
    try:
        import cloudpickle

        return cloudpickle
    except ImportError:  # pragma: no cover
        import pickle

        return pickle


def _save(obj, path):
    tmp = path + ".tmp%d" % os.getpid()
    with open(tmp, "wb") as fh:
        _serializer().dump(obj, fh)
    os.replace(tmp, path)


def _load(path):
    import pickle

    with open(path, "rb") as fh:
        return pickle.load(fh)


def pmcmc_chains_path(path, chain_id=None):
    """R/pmcmc_chains.R:130-136 (``.pkl`` where the reference writes ``.rds``)."""
    if not isinstance(path, (str, os.PathLike)):
        raise TypeError("'path' must be a scalar character")
    path = os.fspath(path)
    ids = [] if chain_id is None else np.atleast_1d(chain_id).tolist()
    results = [os.path.join(path, "results_%d.pkl" % k) for k in ids]
    return {"root": path, "inputs": os.path.join(path, "inputs.pkl"),
            "control": os.path.join(path, "control.pkl"),
            "results": results[0] if np.isscalar(chain_id) and chain_id is not None else results}


def pmcmc_chains_prepare(path, pars, filter, control, initial=None):
    """``pmcmc_chains_prepare`` (R/pmcmc_chains.R:34-61): write everything a chain needs
    -- parameters, starting points, the filter's inputs, the control and each chain's
    seeds -- under ``path``; any process that can read it (another GPU of the box, another
    box) then runs chain k with ``pmcmc_chains_run(k, path)``.  Returns ``path``."""
    p = pmcmc_chains_path(path)
    if not isinstance(pars, pmcmc_parameters):
        raise TypeError("'pars' must be a pmcmc_parameters")
    if not hasattr(filter, "inputs"):
        raise TypeError("'filter' must be a particle_filter")
    if not isinstance(control, pmcmc_control):
        raise TypeError("'control' must be a pmcmc_control")
    if control.n_workers != 1:
        raise ValueError("'n_workers' must be 1")
    if not control.use_parallel_seed:
        raise ValueError("'use_parallel_seed' must be TRUE")
    initial = _check_initial(initial, pars, control.n_chains)
    inputs = filter.inputs()
    seed = make_seeds(control.n_chains, inputs["seed"],
                      getattr(filter.model, "algorithm", "xoshiro256plus"))
    os.makedirs(p["root"], exist_ok=True)
    _save({"class": "pmcmc_inputs", "pars": pars, "initial": initial, "filter": inputs,
           "control": control, "seed": seed}, p["inputs"])
    _save(control, p["control"])
    return p["root"]


def pmcmc_chains_run(chain_id, path, n_threads=None, device=None):
    """``pmcmc_chains_run`` (R/pmcmc_chains.R:71-98): run chain ``chain_id`` (1-based, as
    in R) from the prepared inputs and write its samples; returns the results file.
    ``device``: the GPU to run on when the filter was built with a ``gpu_config`` (the
    prepared inputs name the device of the process that prepared them)."""
    if isinstance(chain_id, bool) or int(chain_id) != chain_id or chain_id < 1:
        raise ValueError("'chain_id' must be at least 1")
    chain_id = int(chain_id)
    p = pmcmc_chains_path(path, chain_id)
    inputs = _load(p["inputs"])
    if not (isinstance(inputs, dict) and inputs.get("class") == "pmcmc_inputs"):
        raise TypeError("'inputs' must be a pmcmc_inputs")
    from .dust import dust_repair_environment

    dust_repair_environment(inputs["filter"]["model"])
    control = inputs["control"]
    if chain_id > control.n_chains:
        raise ValueError("'chain_id' must be an integer in 1..%d" % control.n_chains)
    k = chain_id - 1
    samples = _chain_from_inputs(inputs["filter"], inputs["seed"][k], inputs["pars"],
                                 inputs["initial"][k], control, device, n_threads)
    _save(samples, p["results"])
    return p["results"]


def pmcmc_chains_collect(path):
    """``pmcmc_chains_collect`` (R/pmcmc_chains.R:101-117): the combined samples, equal to
    ``pmcmc()`` run with the same control in one process."""
    control = _load(pmcmc_chains_path(path)["control"])
    p = pmcmc_chains_path(path, list(range(1, control.n_chains + 1)))
    missing = [k + 1 for k, f in enumerate(p["results"]) if not os.path.exists(f)]
    if missing:
        raise RuntimeError("Results missing for chains %s" % ", ".join(map(str, missing)))
    samples = [_load(f) for f in p["results"]]
    for smp in samples:   # the generator inside `predict` was serialised with the samples
        if smp.get("predict") is not None:
            from .dust import dust_repair_environment

            dust_repair_environment(smp["predict"]["filter"]["model"])
    return samples[0] if control.n_chains == 1 else pmcmc_combine(samples=samples)


def pmcmc_chains_cleanup(path):
    """``pmcmc_chains_cleanup`` (R/pmcmc_chains.R:120-127): remove the files written, and
    the directory if nothing else is in it."""
    control = _load(pmcmc_chains_path(path)["control"])
    p = pmcmc_chains_path(path, list(range(1, control.n_chains + 1)))
    for f in [p["inputs"], p["control"]] + p["results"]:
        if os.path.exists(f):
            os.remove(f)
    if os.path.isdir(p["root"]) and not os.listdir(p["root"]):
        os.rmdir(p["root"])
