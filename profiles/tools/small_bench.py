import sys, time, os, numpy as np
sys.path.insert(0, '.')
import torch
from mcstate_b200 import dust
d = np.loadtxt("tests/golden/sir_incidence.csv", delimiter=",", skiprows=1)
t, x = d[:, 1].astype(np.uint64) * 4, d[:, 0]
g = dust.dust_example("sir")
for n in [1 << 12, 1 << 14, 1 << 15, 1 << 16, 75000, 100000, 1 << 17, 150000]:
    for persistent in (True, False):
        m = g.new({}, 0, n, seed=1)
        m.set_persistent(persistent)
        m.set_data(dust.dust_data({"time_end": t, "incidence": x}), False)
        def once():
            m.update_state(pars={}, time=0)
            return m.filter()["log_likelihood"]
        for _ in range(3): once()
        torch.cuda.synchronize(); t0 = time.perf_counter(); K = 20
        for _ in range(K): ll = once()
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / K
        print("n %7d persistent=%d: %.3f ms/eval  %.1f us/interval  %.2f G particle-steps/s  ll %.6f"
              % (n, persistent, dt * 1e3, dt * 1e4, n * 400 / dt / 1e9, ll[0]))
