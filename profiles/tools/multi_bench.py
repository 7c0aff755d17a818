import sys, time, numpy as np
sys.path.insert(0, '.')
import torch
from mcstate_b200 import dust
d = np.loadtxt("tests/golden/sir_incidence.csv", delimiter=",", skiprows=1)
t, x = d[:, 1].astype(np.uint64) * 4, d[:, 0]
g = dust.dust_example("sir")
for n_sets, n in [(128, 1 << 14), (8, 1 << 16), (1, 1 << 16), (1024, 1 << 10)]:
    pars = [{"beta": 0.2 + 0.0001 * k} for k in range(n_sets)]
    m = g.new(pars, 0, n, seed=1, pars_multi=True) if n_sets > 1 else g.new({}, 0, n, seed=1)
    m.set_data(dust.dust_data({"time_end": t, "incidence": x}), n_sets > 1)
    def once():
        m.update_state(pars=pars if n_sets > 1 else {}, time=0)
        return m.filter()["log_likelihood"]
    for _ in range(3): once()
    torch.cuda.synchronize(); t0 = time.perf_counter(); K = 10
    for _ in range(K): ll = once()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / K
    print("sets %4d x %6d particles: %.3f ms/eval  %.1f us/interval  %.2f G particle-steps/s  ll0 %.4f"
          % (n_sets, n, dt * 1e3, dt * 1e4, n_sets * n * 400 / dt / 1e9, ll[0]))
