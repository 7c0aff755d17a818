"""The configurations of BASELINE.json that shard over GPUs -- C3, C4, C5 -- timed under
the same launch as the headline (`bench.py --gpus N`, one rank per GPU, NCCL), so the
driver's clock sees them at N = 1, 2, 4, 8.  Each goes through the caller the reference's
user would use and includes that caller's collective:

  c3  SIR pmcmc, N chains x 2^16 particles x 1000 Metropolis-Hastings steps, one chain per
      GPU through pmcmc_sharded (R/pmcmc_parallel.R:8-20,125-151): state and trajectories
      kept, one gather of the samples at the end
  c4  nested SIR filter, max(2, N) populations x 2^16 particles through
      nested_filter_sharded (R/particle_filter.R:1000-1024): every evaluation ends with the
      all-gather of the populations' log-likelihoods
  c5  smc2 on SIR, 128 N parameter particles x 2^14 state particles x 100 data rows through
      smc2_engine (R/smc2.R:89-108,120-189): one all-gather of step likelihoods per data
      row, replenishment by batched re-run

Work per GPU is fixed as N grows (weak scaling).  Times are wall-clock between barriers with
a device synchronisation on both sides, the maximum over ranks.
"""
from __future__ import annotations

import math
import os
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
RATE = 4


def _sir_frame():
    d = np.loadtxt(os.path.join(ROOT, "tests", "golden", "sir_incidence.csv"), delimiter=",",
                   skiprows=1)
    return d[:, 1].astype(int), d[:, 0]


def _nested_data(n_pop):
    """inst/nested_sir_incidence.csv (populations A, B); further populations are synthetic
    copies of A's series, shifted, so that every GPU has one."""
    from mcstate_b200.particle_filter_data import particle_filter_data

    rows = [line.strip().split(",")
            for line in open(os.path.join(ROOT, "tests", "golden", "nested_sir_incidence.csv"))][1:]
    cases = np.array([float(r[0]) for r in rows])
    day = np.array([int(r[1]) for r in rows])
    pop = np.array([r[2].strip('"') for r in rows])
    cols = {"day": list(day[pop == "A"]) + list(day[pop == "B"]),
            "incidence": list(cases[pop == "A"]) + list(cases[pop == "B"]),
            "population": ["A"] * 100 + ["B"] * 100}
    for k in range(2, n_pop):
        cols["day"] += list(day[pop == "A"])
        cols["incidence"] += list(np.roll(cases[pop == "A"], 3 * k) + (k % 2))
        cols["population"] += [chr(ord("A") + k)] * 100
    cols["incidence"] = np.array(cols["incidence"], dtype=float)
    return particle_filter_data(cols, "day", RATE, 0, population="population")


class _Timer:
    """barrier + synchronize on both sides; clocks sampled in between."""

    def __init__(self, dist, torch, sampler_cls, local):
        self.dist, self.torch, self.sampler = dist, torch, sampler_cls(local)

    def __enter__(self):
        self.dist.barrier()
        self.torch.cuda.synchronize()
        self.sampler.__enter__()
        self.t0 = time.perf_counter()
        return self

    def __exit__(self, *a):
        self.torch.cuda.synchronize()
        self.dist.barrier()
        self.seconds = time.perf_counter() - self.t0
        self.sampler.__exit__(None, None, None)

    def clocks(self):
        return self.sampler.summary()


def run_configs(local, sampler_cls, c3_steps=1000, c4_evals=30):
    """-> {"c3": {...}, "c4": {...}, "c5": {...}} on rank 0 (None elsewhere).  Needs an
    initialised process group (bench.py makes a one-rank group at N = 1)."""
    import torch
    import torch.distributed as dist

    from mcstate_b200 import dust
    from mcstate_b200.particle_filter import particle_filter
    from mcstate_b200.particle_filter_data import particle_filter_data
    from mcstate_b200.pmcmc import (pmcmc, pmcmc_control, pmcmc_parameter, pmcmc_parameters)
    from mcstate_b200.sharding import nested_filter_sharded
    from mcstate_b200.smc2 import smc2_control, smc2_engine, smc2_parameter, smc2_parameters

    rank, world = dist.get_rank(), dist.get_world_size()
    gen = dust.dust_example("sir")
    day, inc = _sir_frame()
    data = particle_filter_data({"day": day, "incidence": inc}, "day", RATE, 0)
    out = {}

    def mx(x):   # the slowest rank's time
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    # ---- C3: tests/testthat/helper-mcstate.R:44-52 parameters, one chain per GPU
    n3 = 1 << 16
    flat = lambda p: math.log(1e-10)  # noqa: E731
    pars3 = pmcmc_parameters([pmcmc_parameter("beta", 0.2, min=0, max=1, prior=flat),
                              pmcmc_parameter("gamma", 0.1, min=0, max=1, prior=flat)],
                             [[0.00057, 0.00034], [0.00034, 0.00026]])
    index = lambda info: {"run": [5], "state": [1, 2, 3]}  # noqa: E731
    f3 = particle_filter(data, gen, n3, None, index=index, seed=1, gpu_config=local)

    def ctl(n):
        return pmcmc_control(n, n_chains=world, n_workers=world, use_parallel_seed=True,
                             save_state=True, save_trajectories=True, n_steps_retain=min(n, 100))

    pmcmc(pars3, f3, control=ctl(5))          # warm-up: model object, kernels, graphs
    with _Timer(dist, torch, sampler_cls, local) as t:
        res = pmcmc(pars3, f3, control=ctl(c3_steps))
    secs = mx(t.seconds)
    runs = int(res["n_filter_runs"])
    out["c3"] = {"workload": "SIR pmcmc, %d chains x 2^16 particles x %d steps, one chain per GPU "
                             "(pmcmc_sharded; state + trajectories kept)" % (world, c3_steps),
                 "seconds": secs, "filter_evals": runs, "filter_evals_per_s": runs / secs,
                 "particle_steps_per_s": runs * n3 * int(day[-1]) * RATE / secs,
                 "ms_per_mcmc_step_per_chain": 1e3 * secs / c3_steps,
                 "posterior_mean": [float(x) for x in np.asarray(res["pars"]).mean(axis=0)],
                 "clocks": t.clocks()}
    del f3

    # ---- C4: populations sharded, log-likelihoods all-gathered per evaluation
    n_pop, n4 = max(2, world), 1 << 16
    f4 = nested_filter_sharded(_nested_data(n_pop), gen, n4, seed=7, gpu_config=local)
    p4 = [{"beta": 0.2 + 0.02 * k, "gamma": 0.1 + 0.005 * k} for k in range(n_pop)]
    for _ in range(3):
        ll = f4.run(p4)
    with _Timer(dist, torch, sampler_cls, local) as t:
        for _ in range(c4_evals):
            ll = f4.run(p4)
    secs = mx(t.seconds) / c4_evals
    out["c4"] = {"workload": "nested SIR filter, %d populations x 2^16 particles over %d GPUs "
                             "(nested_filter_sharded), all-gather of the log-likelihoods included"
                             % (n_pop, world),
                 "ms_per_evaluation": 1e3 * secs, "evaluations": c4_evals,
                 "particle_steps_per_s": n_pop * n4 * int(day[-1]) * RATE / secs,
                 "log_likelihood": [float(x) for x in ll], "clocks": t.clocks()}
    del f4

    # ---- C5: parameter particles sharded, one all-gather per data row
    n5 = 1 << 14
    f5 = particle_filter(data, gen, n5, None, seed=1, gpu_config=local)
    unif = lambda k, rng: rng.uniform(0.05, 0.5, k)  # noqa: E731
    dunif = lambda v: np.where((v >= 0.05) & (v <= 0.5), 0.0, -math.inf)  # noqa: E731
    pars5 = smc2_parameters([smc2_parameter("beta", unif, dunif, min=0.05, max=0.5),
                             smc2_parameter("gamma", unif, dunif, min=0.05, max=0.5)])
    warm = smc2_engine(pars5, f5, smc2_control(128 * world, seed=3))
    for _ in range(3):
        warm.step()                            # kernels and the per-row launch sequences
    del warm
    eng = smc2_engine(pars5, f5, smc2_control(128 * world, seed=4))
    with _Timer(dist, torch, sampler_cls, local) as t:
        eng.run()
    secs = mx(t.seconds)
    res = eng.results()
    w = res["probabilities"][:, 3]
    out["c5"] = {"workload": "smc2 on SIR, %d parameter particles x 2^14 state particles over %d "
                             "GPUs, 100 data rows (smc2_engine)" % (128 * world, world),
                 "seconds": secs, "filter_rows_per_gpu": int(eng.n_filter_rows),
                 "particle_steps_per_s_per_gpu": eng.n_filter_rows * n5 * RATE / secs,
                 "replenishments": int(np.sum(~np.isnan(eng.acceptance_rate))),
                 "posterior_mean": [float(x) for x in (w[:, None] * res["pars"]).sum(axis=0)],
                 "clocks": t.clocks()}
    return out if rank == 0 else None
