"""the weights / resampling / gather stages on four weight distributions.

In this engine the resampling search and the state gather of interval t are fused into the
load of K1 at interval t+1, so the kernels timed are K1 (search + gather + update + compare)
and K2 (exp, fixed point, scan, log-mean-exp), on the volatility model (one real state: the
model itself is cheap, what is left is the pipeline) at 2^20 particles, 100 intervals.  The
compare is N(observed | gamma x, tau), so the weight distribution is set by (gamma, tau):

  uniform     gamma = 0           every weight equal: kappa is the identity (best-case gather)
  mild        gamma = 1, tau = 1  the model's own parameters (tests/testthat/helper-mcstate.R:99)
  heavy       tau = 0.05          log-weights spread over hundreds of units: few ancestors
  one-hot     tau = 1e-4          every fixed-point weight but the largest rounds to zero:
                                  kappa is constant (broadcast gather)

ESS and distinct ancestors are those of the first interval (host arithmetic on compare_data()).
"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mcstate_b200 import dust
n, T = 1 << 20, 100
g = dust.dust_example("volatility")
rs = np.random.RandomState(1)
t = np.arange(1, T + 1, dtype=np.uint64)
obs = rs.normal(size=T) * 2
data = dust.dust_data({"time_end": t, "observed": obs})
for name, pars in (("uniform", {"gamma": 0.0}), ("mild", {}), ("heavy", {"tau": 0.05}),
                   ("one-hot", {"tau": 1e-4})):
    p = g.new(pars, 0, n, seed=1)
    p.set_data(data, False)
    p.run(1)
    lw = p.compare_data()
    w = np.exp(lw - lw.max())
    ess = w.sum() ** 2 / np.sum(w ** 2)
    kappa = p.resample(w)
    m = g.new(pars, 0, n, seed=1)
    m.set_data(data, False)
    for _ in range(3):
        m.update_state(pars=pars, time=0); m.filter()
    m.set_kernel_timing(True)
    ms = np.zeros(4); cnt = np.zeros(4)
    for _ in range(5):
        m.update_state(pars=pars, time=0); ll = m.filter()["log_likelihood"][0]
        a, b = m.kernel_timing(); ms += a; cnt += b
    k1, k2 = 1e3 * ms[0] / cnt[0], 1e3 * ms[1] / cnt[1]
    # K1 moves state 8+8, rng 32+32, logw 8, one cw probe 8 = 96 B; K2 logw 8 + cw 8
    print("%-8s ESS %10.1f  distinct ancestors %8d  K1 %6.1f us (%5.0f GB/s of 96 B/particle)  "
          "K2 %5.1f us (%4.0f GB/s of 16 B/particle)  ll %.4f"
          % (name, ess, np.unique(kappa).size, k1, 96 * n / k1 / 1e3, k2, 16 * n / k2 / 1e3, ll))
