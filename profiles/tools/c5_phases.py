"""Where config C5 (smc2, 128 parameter particles x 2^14 state particles per GPU) spends its
wall time, per phase and rank.  torchrun --nproc-per-node N profiles/tools/c5_phases.py"""
import math, os, sys, time
import numpy as np
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mcstate_b200 import dust
from mcstate_b200.particle_filter import particle_filter
from mcstate_b200.particle_filter_data import particle_filter_data
from mcstate_b200.smc2 import smc2_control, smc2_engine, smc2_parameter, smc2_parameters
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
d = np.loadtxt(os.path.join(ROOT, "tests", "golden", "sir_incidence.csv"), delimiter=",", skiprows=1)
data = particle_filter_data({"day": d[:, 1].astype(int), "incidence": d[:, 0]}, "day", 4, 0)
f5 = particle_filter(data, dust.dust_example("sir"), 1 << 14, None, seed=1, gpu_config=local)
unif = lambda k, rng: rng.uniform(0.05, 0.5, k)  # noqa: E731
dunif = lambda v: np.where((v >= 0.05) & (v <= 0.5), 0.0, -math.inf)  # noqa: E731
pars5 = smc2_parameters([smc2_parameter("beta", unif, dunif, min=0.05, max=0.5),
                         smc2_parameter("gamma", unif, dunif, min=0.05, max=0.5)])
warm = smc2_engine(pars5, f5, smc2_control(128 * world, seed=3))
for _ in range(3):
    warm.step()
del warm
for rep in range(2):
    eng = smc2_engine(pars5, f5, smc2_control(128 * world, seed=4))
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    eng.run()
    torch.cuda.synchronize(); dist.barrier()
    secs = time.perf_counter() - t0
    for r in range(world):
        if r == rank and (rank in (0, world - 1)):
            print("rep %d rank %d: %.3f s, rows/gpu %d, %s" % (rep, rank, secs, eng.n_filter_rows,
                  {k: round(v, 4) for k, v in eng.phase_seconds.items()}), flush=True)
        dist.barrier()
dist.destroy_process_group()
