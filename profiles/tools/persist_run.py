import sys, numpy as np
sys.path.insert(0, '.')
from mcstate_b200 import dust
d = np.loadtxt("tests/golden/sir_incidence.csv", delimiter=",", skiprows=1)
t, x = d[:, 1].astype(np.uint64) * 4, d[:, 0]
g = dust.dust_example("sir")
m = g.new({}, 0, 1 << 16, seed=1)
m.set_data(dust.dust_data({"time_end": t, "incidence": x}), False)
for _ in range(3):
    m.update_state(pars={}, time=0); print(m.filter()["log_likelihood"])
