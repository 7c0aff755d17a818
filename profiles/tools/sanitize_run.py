"""Small runs of every kernel family for `compute-sanitizer --tool memcheck` (and racecheck):
all stock models and the user-supplied SEIR build, both execution paths of the filter,
single and batched parameter sets, NA data rows, trajectories, snapshots, early exit, the
per-set copies SMC^2 uses, simulate and the on-device lineage trace.

compute-sanitizer --tool memcheck python profiles/tools/sanitize_run.py
(this pool refuses compute-sanitizer since round 1 -- "closed on this pool" -- so the committed
output, profiles/allpaths_r01_v42.txt, is a plain run: every path executes, and the
single-launch and two-launch paths print the same likelihoods)
"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mcstate_b200 import dust
d = np.loadtxt(os.path.join(ROOT, "tests", "golden", "sir_incidence.csv"), delimiter=",", skiprows=1)
t, x = d[:30, 1].astype(np.uint64) * 4, d[:30, 0].copy()
x[[3, 11]] = np.nan
seir = dust.dust(os.path.join(ROOT, "mcstate_b200", "examples", "seir.cuh"))
for name, pars in (("sir", {}), ("sirs", {"alpha": 0.05}), ("volatility", {}), ("seir", {"beta": 0.5})):
    g = seir if name == "seir" else dust.dust_example(name)
    for persistent in (True, False):
        for n, multi in ((3000, False), (700, True)):
            if multi:
                m = g.new([pars, pars, pars], 0, n, seed=1, pars_multi=True)
            else:
                m = g.new(pars, 0, n, seed=1)
            m.set_persistent(persistent)
            col = "observed" if name == "volatility" else "incidence"
            m.set_data(dust.dust_data({"time_end": t, col: x}), multi)
            m.set_index([1])
            r = m.filter(save_trajectories=True, time_snapshot=None if persistent else [t[5], t[20]])
            m.update_state(pars=[pars] * 3 if multi else pars, time=0)
            r2 = m.filter(min_log_likelihood=-1e3)
            m.update_state(pars=[pars] * 3 if multi else pars, time=0)
            m.filter(save_trajectories="device")
            if not multi:
                m.filter_history([1, n])
                m.state_particles([2])
            m.simulate([int(t[-1]) + 4, int(t[-1]) + 8])
            print(name, persistent, n, multi, r["log_likelihood"][:1], r2["log_likelihood"][:1])
    # per-set seeds + copies between two handles (smc2 replenishment)
    a = g.new([pars] * 4, 0, 500, seed=np.arange(16, dtype=np.uint64).tobytes(), pars_multi=True,
              seed_per_set=True)
    b = g.new([pars] * 4, 0, 500, seed=np.arange(16, 32, dtype=np.uint64).tobytes(), pars_multi=True,
              seed_per_set=True)
    col = "observed" if name == "volatility" else "incidence"
    for m in (a, b):
        m.set_data(dust.dust_data({"time_end": t, col: x}), True)
    a.filter(int(t[10])); b.filter(int(t[10]))
    a.copy_sets_from(b, [True, False, True, False], state=True, rng=True)
    print(name, "copy_sets", a.filter()["log_likelihood"][:2])
print("done")
