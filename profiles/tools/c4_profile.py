"""Where an evaluation of config C4 (nested filter, one population of 2^16 particles per GPU)
spends its host time: cProfile of rank 0.  torchrun --nproc-per-node N profiles/tools/c4_profile.py"""
import cProfile, io, os, pstats, sys, time
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench_configs
from mcstate_b200 import dust
from mcstate_b200.sharding import nested_filter_sharded
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
n_pop = max(2, world)
f = nested_filter_sharded(bench_configs._nested_data(n_pop), dust.dust_example("sir"), 1 << 16, seed=7,
                          gpu_config=local)
p = [{"beta": 0.2 + 0.02 * k, "gamma": 0.1 + 0.005 * k} for k in range(n_pop)]
for _ in range(5):
    f.run(p)
dist.barrier(); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(100):
    f.run(p)
torch.cuda.synchronize(); dist.barrier()
if rank == 0:
    print("%.3f ms per evaluation" % ((time.perf_counter() - t0) * 10))
pr = cProfile.Profile(); pr.enable()
for _ in range(100):
    f.run(p)
pr.disable()
if rank == 0:
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(32)
    print(s.getvalue()[:5000])
dist.destroy_process_group()
