#!/usr/bin/env python
"""Regression vectors of the CPU oracle -> tests/golden/oracle_vectors.json.

The GPU tests compare the engine with the oracle on the fly; that cannot notice the two
drifting TOGETHER (both are ours).  These vectors freeze what the oracle computed when
they were made -- RNG words, draws, the reproducible math functions (as bit patterns)
and whole-filter log-likelihoods -- so a later change to either side that alters a bit
shows up.  They pin the oracle against itself over time, not against dust (PARITY_NOTES.md).

usage: python tests/golden/make_oracle_vectors.py        (rewrites the JSON)
"""
import json
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as O  # noqa: E402


def bits(x):
    return "%016x" % struct.unpack("<Q", struct.pack("<d", float(x)))[0]


def vectors():
    O.build()
    L = O.lib()
    v = {}
    v["seed_1_streams_0_1_2"] = [[int(w) for w in s] for s in O.rng_streams(1, 3)]
    v["seed_42_long_jump_nodes"] = [[int(w) for w in n[0]] for n in O.rng_distributed_state(42, 1, 3)]
    for scr, name in ((O.PLUS, "plus"), (O.STARSTAR, "starstar")):
        st = O.seed_state(7).copy()
        v["next_words_seed7_" + name] = [int(w) for w in O.next_words(st, 5, scr)]
    st = O.seed_state(42).copy()
    p = st.ctypes.data_as(O._u64p)
    v["random_real_seed42"] = [bits(L.orc_random_real(p, O.PLUS)) for _ in range(4)]
    v["binomial_seed42"] = {}
    for n, pr in ((10, 0.3), (150, 0.0247), (1000, 0.05), (5000, 0.9), (37, 0.5)):
        st = O.seed_state(42).copy()
        p = st.ctypes.data_as(O._u64p)
        v["binomial_seed42"]["%d,%r" % (n, pr)] = [L.orc_binomial(p, O.PLUS, float(n), pr)
                                                  for _ in range(8)]
    st = O.seed_state(42).copy()
    p = st.ctypes.data_as(O._u64p)
    v["poisson_seed42"] = [L.orc_poisson(p, O.PLUS, lam) for lam in (0.5, 3.0, 9.9, 25.0, 400.0)]
    st = O.seed_state(42).copy()
    p = st.ctypes.data_as(O._u64p)
    v["normal_seed42"] = [bits(L.orc_normal(p, O.PLUS, 0.0, 1.0)) for _ in range(4)]
    xs = [-700.5, -37.25, -1.0, -1e-9, 0.0, 0.3, 1.0, 12.75, 700.0]
    v["exp"] = {repr(x): bits(L.orc_exp(x)) for x in xs}
    xs = [1e-300, 1e-6, 0.1, 0.5, 0.999999, 1.0, 1.5, 10.0, 1e6, 1e300]
    v["log"] = {repr(x): bits(L.orc_log(x)) for x in xs}
    v["lfact"] = {repr(k): bits(L.orc_lfact(k)) for k in (0.0, 1.0, 15.0, 16.0, 17.0, 250.0, 1e6)}
    v["poisson_ptrs_seed7"] = {}
    for lam in (10.0, 37.5, 1e4, 1e6):
        st = O.seed_state(7).copy()
        p = st.ctypes.data_as(O._u64p)
        v["poisson_ptrs_seed7"][repr(lam)] = [L.orc_poisson(p, O.PLUS, lam) for _ in range(6)]
    xs = [0.0, 1e-9, 0.1, 0.125, 0.2, 0.3, 0.49, 0.6, 0.77, 0.999]
    v["cos2pi"] = {repr(x): bits(L.orc_cos2pi(x)) for x in xs}

    d = np.loadtxt(os.path.join(HERE, "sir_incidence.csv"), delimiter=",", skiprows=1)
    t, x = d[:, 1].astype(np.uint64) * 4, d[:, 0]
    fl = {}
    for name, model, n, seed in (("sir_n100_seed1", O.SIR, 100, 1), ("sir_n4096_seed1", O.SIR, 4096, 1),
                                 ("sirs_n64_seed3", O.SIRS, 64, 3)):
        m = O.OracleModel(n_particles=n, seed=seed, model=model)
        m.set_data(t, x)
        ll = m.filter()["log_likelihood"][0]
        fl[name] = {"log_likelihood": bits(ll), "value": float(ll),
                    "state_sum": bits(float(m.state().sum())),
                    "rng_word_0": int(m.rng_state()[0, 0])}
    obs = np.sin(np.arange(1, 41) * 0.37) * 2.0
    m = O.OracleModel(n_particles=50, seed=5, model=O.VOLATILITY)
    m.set_data(np.arange(1, 41, dtype=np.uint64), obs)
    ll = m.filter()["log_likelihood"][0]
    fl["volatility_n50_seed5"] = {"log_likelihood": bits(ll), "value": float(ll),
                                  "state_sum": bits(float(m.state().sum())),
                                  "rng_word_0": int(m.rng_state()[0, 0])}
    v["filters"] = fl
    return v


if __name__ == "__main__":
    out = os.path.join(HERE, "oracle_vectors.json")
    with open(out, "w") as fh:
        json.dump(vectors(), fh, indent=1, sort_keys=True)
    print("wrote", out)
