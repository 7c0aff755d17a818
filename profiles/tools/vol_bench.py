import sys, time, numpy as np
sys.path.insert(0, '.')
import torch
from mcstate_b200 import dust
n = 1 << 20
g = dust.dust_example("volatility")
rs = np.random.RandomState(1)
T = 100
t = np.arange(1, T + 1, dtype=np.uint64); obs = rs.normal(size=T) * 2
m = g.new({}, 0, n, seed=1)
m.set_data(dust.dust_data({"time_end": t, "observed": obs}), False)
for _ in range(3):
    m.update_state(pars={}, time=0); m.filter()
torch.cuda.synchronize()
t0 = time.perf_counter()
K = 10
for _ in range(K):
    m.update_state(pars={}, time=0); ll = m.filter()["log_likelihood"]
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / K
# per particle-interval: state 8 in + 8 out, rng 32 + 32, logw 8 w + 8 r, cw 8 w + 8 probe
print("volatility 2^20 x %d intervals: %.3f ms/eval, %.1f us/interval, ll %.6f, K1+K2 algorithmic 104 B/particle -> %.0f GB/s"
      % (T, dt * 1e3, dt * 1e6 / T, ll[0], 104 * n * T / dt / 1e9))
