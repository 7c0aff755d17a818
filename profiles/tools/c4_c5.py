"""BASELINE configs C4 and C5 on the GPUs of one box (run under torchrun, one rank per GPU).

C4  nested SIR filter on inst/nested_sir_incidence.csv (populations A, B; synthetic copies
    with shifted parameters when there are more GPUs than populations), 2^16 particles per
    population, populations sharded over the ranks (mcstate_b200.sharding).
C5  smc2 on SIR: 128 parameter particles x 2^14 state particles PER GPU (1024 over 8 GPUs),
    parameter particles sharded over the ranks (mcstate_b200.smc2).

torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 profiles/tools/c4_c5.py
"""
import json
import math
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
from mcstate_b200.sharding import nested_filter_sharded  # noqa: E402
from mcstate_b200.smc2 import smc2_control, smc2_engine, smc2_parameter, smc2_parameters  # noqa: E402
from test_sharding import nested_data, pars_for  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
gen = dust.dust_example("sir")

# ---- C4
n_pop, n = max(2, world), 1 << 16
f = nested_filter_sharded(nested_data(n_pop), gen, n, seed=7, gpu_config=local)
p = pars_for(n_pop)
f.run(p)
dist.barrier(); torch.cuda.synchronize()
t0 = time.perf_counter()
K = 50
for _ in range(K):
    ll = f.run(p)
torch.cuda.synchronize(); dist.barrier()
c4 = (time.perf_counter() - t0) / K
if rank == 0:
    print(json.dumps({"config": "C4: nested SIR filter, %d populations x 2^16 particles over %d GPUs"
                                % (n_pop, world), "ms_per_evaluation": 1e3 * c4,
                      "particle_steps_per_s": n_pop * n * 400 / c4,
                      "log_likelihood": [float(x) for x in ll]}))

# ---- C5
d = np.loadtxt(os.path.join(ROOT, "tests", "golden", "sir_incidence.csv"), delimiter=",", skiprows=1)
data = particle_filter_data({"day": d[:, 1].astype(int), "incidence": d[:, 0]}, "day", 4, 0)
flt = particle_filter(data, gen, 1 << 14, None, seed=1, gpu_config=local)
unif = lambda k, rng: rng.uniform(0.05, 0.5, k)  # noqa: E731
dunif = lambda v: 0.0 if 0.05 <= v <= 0.5 else -math.inf  # noqa: E731
pars = smc2_parameters([smc2_parameter("beta", unif, dunif, min=0.05, max=0.5),
                        smc2_parameter("gamma", unif, dunif, min=0.05, max=0.5)])
eng = smc2_engine(pars, flt, smc2_control(128 * world, seed=4))
dist.barrier(); torch.cuda.synchronize()
t0 = time.perf_counter()
eng.run()
torch.cuda.synchronize(); dist.barrier()
c5 = time.perf_counter() - t0
res = eng.results()
if rank == 0:
    w = res["probabilities"][:, 3]
    print(json.dumps({"config": "C5: smc2 on SIR, %d parameter particles x 2^14 state particles over %d GPUs, "
                                "100 data rows" % (128 * world, world), "seconds": c5,
                      "filter_rows_per_rank": int(eng.n_filter_rows),
                      "particle_steps_per_s_per_gpu": eng.n_filter_rows * (1 << 14) * 4 / c5,
                      "replenishments": int(np.sum(~np.isnan(eng.acceptance_rate))),
                      "posterior_mean": [float(x) for x in (w[:, None] * res["pars"]).sum(axis=0)]}))
dist.destroy_process_group()
