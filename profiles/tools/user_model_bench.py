"""A user-supplied model (mcstate_b200/examples/seir.cuh, compiled by dust.dust) at config
C2's size: 2^20 particles x 100 data intervals x 4 steps.  The model has three binomial
draws per step; the example header also carries the optional lean path (memo tables for its two
fixed-probability draws, certified fast walk for the infection draw), so this is what a user model
costs once it has one (the general form alone: profiles/user_model_r01_v41.txt, 21.1 ms);
the stock SIR model is timed beside it through the same calls."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from mcstate_b200 import dust
n = 1 << 20
d = np.loadtxt(os.path.join(ROOT, "tests", "golden", "sir_incidence.csv"), delimiter=",", skiprows=1)
t, x = d[:, 1].astype(np.uint64) * 4, d[:, 0]
if len(sys.argv) > 1:   # a tuning build of the SEIR engine (A/B runs): python user_model_bench.py <lib.so>
    from mcstate_b200 import _lib
    models = ((os.path.basename(sys.argv[1]), dust.dust_generator("seir", library=_lib.load(sys.argv[1]))),)
else:
    models = (("seir (user model)", dust.dust(os.path.join(ROOT, "mcstate_b200", "examples", "seir.cuh"))),
              ("sir (stock)", dust.dust_example("sir")))
for name, g in models:
    m = g.new({}, 0, n, seed=1)
    m.set_data(dust.dust_data({"time_end": t, "incidence": x}), False)
    for _ in range(3):
        m.update_state(pars={}, time=0); m.filter()
    torch.cuda.synchronize()
    K = 10
    t0 = time.perf_counter()
    for _ in range(K):
        m.update_state(pars={}, time=0); ll = m.filter()["log_likelihood"]
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / K
    m.set_kernel_timing(True)
    m.update_state(pars={}, time=0); m.filter()
    kt = m.kernel_timing()
    print("%-18s 2^20 x 100 intervals x 4 steps: %.3f ms/eval, %.1f us/interval, %.2f G particle-steps/s, ll %.6f, kernels %s"
          % (name, dt * 1e3, dt * 1e4, n * 400 / dt / 1e9, ll[0], kt))
