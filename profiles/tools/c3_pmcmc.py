"""BASELINE config C3: SIR pmcmc, one chain of 2^16 particles x 1000 steps per GPU
(pmcmc_parallel equivalent: mcstate_b200.pmcmc.pmcmc_sharded).  Run under torchrun, one
rank per GPU; rank 0 prints one JSON line.

torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 profiles/tools/c3_pmcmc.py
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from mcstate_b200 import dust  # noqa: E402
from mcstate_b200.particle_filter import particle_filter  # noqa: E402
from mcstate_b200.particle_filter_data import particle_filter_data  # noqa: E402
from mcstate_b200.pmcmc import pmcmc, pmcmc_control  # noqa: E402
from test_pmcmc import example_pars  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
d = np.loadtxt(os.path.join(ROOT, "tests", "golden", "sir_incidence.csv"), delimiter=",", skiprows=1)
data = particle_filter_data({"day": d[:, 1].astype(int), "incidence": d[:, 0]}, "day", 4, 0)
index = lambda info: {"run": [5], "state": [1, 2, 3]}  # noqa: E731
n_steps = int(os.environ.get("C3_STEPS", "1000"))
f = particle_filter(data, dust.dust_example("sir"), 1 << 16, None, index=index, seed=1, gpu_config=local)
ctl = lambda n: pmcmc_control(n, n_chains=world, n_workers=max(world, 2), save_state=True,  # noqa: E731
                              save_trajectories=True, n_steps_retain=min(n, 100))
pmcmc(example_pars(), f, control=ctl(5))          # warm-up: model objects, kernels
dist.barrier(); torch.cuda.synchronize()
t0 = time.perf_counter()
res = pmcmc(example_pars(), f, control=ctl(n_steps))
torch.cuda.synchronize(); dist.barrier()
secs = time.perf_counter() - t0
if rank == 0:
    runs = int(res["n_filter_runs"])
    print(json.dumps({"config": "C3: SIR pmcmc, %d chains x 2^16 particles x %d steps, one chain per GPU"
                                % (world, n_steps),
                      "n_gpus": world, "seconds": secs, "filter_runs": runs,
                      "filter_evals_per_s": runs / secs,
                      "particle_steps_per_s": runs * (1 << 16) * 400 / secs,
                      "ms_per_mcmc_step_per_chain": 1e3 * secs / n_steps,
                      "posterior_mean": [float(x) for x in np.asarray(res["pars"]).mean(axis=0)]}))
dist.destroy_process_group()
