"""BASELINE.json's configurations at their full sizes on one B200 (what one GPU of the
8-GPU box runs), checked through properties that do not need the CPU oracle to finish
the same work -- plus a short oracle-checked prefix where that is cheap.

  C3  SIR pmcmc, one chain of 2^16 particles x 1000 steps per GPU
  C4  nested SIR filter, populations as parameter sets in one handle
  C5  smc2, 1024 parameter particles x 2^14 state particles over 8 GPUs = 128 per GPU
"""
import math
import time

import numpy as np
import pytest

from mcstate_b200.particle_filter import particle_filter
from mcstate_b200.particle_filter_data import particle_filter_data
from mcstate_b200.pmcmc import pmcmc, pmcmc_control
from oracle_generator import OracleGenerator
from test_pmcmc import example_pars

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dust():
    from mcstate_b200 import dust as d

    if not d.dust_example("sir").gpu_info()["has_cuda"]:
        pytest.fail("no CUDA device visible: the engine has no CPU fallback")
    return d


def _filter(gen, sir_data, n_particles, seed=1, gpu=0):
    data = particle_filter_data({"day": sir_data["day"], "incidence": sir_data["incidence"]},
                                "day", 4, 0)
    index = lambda info: {"run": [5], "state": [1, 2, 3]}  # noqa: E731
    return particle_filter(data, gen, n_particles, None, index=index, seed=seed, gpu_config=gpu)


def test_c3_chain_of_2_16_particles(dust, sir_data):
    """One chain of config C3 at full size: 2^16 particles x 1000 MCMC steps.  The first
    steps are replayed on the CPU oracle (same seeds => identical chain), the whole run
    is checked for a sane acceptance rate and for the known posterior region."""
    n = 1 << 16
    np.random.seed(3)
    short = pmcmc_control(3, save_state=False, use_parallel_seed=True)
    got = pmcmc(example_pars(), _filter(dust.dust_example("sir"), sir_data, n), control=short)
    np.random.seed(3)
    want = pmcmc(example_pars(), _filter(OracleGenerator("sir"), sir_data, n, gpu=None),
                 control=short)
    assert np.array_equal(got["pars"], want["pars"])
    assert np.array_equal(got["probabilities"], want["probabilities"])

    np.random.seed(3)
    t0 = time.perf_counter()
    full = pmcmc(example_pars(), _filter(dust.dust_example("sir"), sir_data, n),
                 control=pmcmc_control(1000, save_state=True, save_trajectories=True,
                                       n_steps_retain=100, use_parallel_seed=True))
    secs = time.perf_counter() - t0
    ll = full["probabilities"][:, 1]
    assert np.isfinite(ll).all()
    moved = np.mean(np.any(np.diff(full["pars"], axis=0) != 0, axis=1))
    assert 0.02 < moved <= 1.0            # retained every 10th step: most pairs differ
    assert full["trajectories"].shape == (3, 100, 101) and full["state"].shape == (5, 100)
    # the data were simulated with beta = 0.2, gamma = 0.1
    assert abs(full["pars"][50:, 0].mean() - 0.2) < 0.05
    assert abs(full["pars"][50:, 1].mean() - 0.1) < 0.05
    assert full["n_filter_runs"] >= 1001
    # 1001+ filter evaluations of 2^16 particles x 400 steps: seconds, not minutes
    assert secs < 60, secs


def test_c4_nested_populations_at_full_size(dust, sir_data):
    """Populations as parameter sets of one model (R/particle_filter.R:1000-1024): 8
    populations x 2^17 particles = 2^20 per handle.  The same model split into two
    shards of 4 populations, as two GPUs of config C4 hold it (each takes its particles'
    streams and replays the shared filter stream: mcb_set_stream_sharing), returns
    bit-identical log-likelihoods and states."""
    n, P = 1 << 17, 8
    t = sir_data["day"].astype(np.uint64) * 4
    x = sir_data["incidence"]
    g = dust.dust_example("sir")
    sets = [{"beta": 0.18 + 0.01 * p} for p in range(P)]
    data = dust.dust_data({"time_end": t, "incidence": x})
    whole = g.new(sets, 0, n, seed=1, pars_multi=True)
    whole.set_data(data, True)
    words = np.frombuffer(whole.rng_state(), dtype=np.uint64).reshape(-1, 4).copy()
    ll = whole.filter()["log_likelihood"]
    assert ll.shape == (P,) and np.isfinite(ll).all()
    s = whole.state()
    assert s.shape == (5, n, P)
    assert np.all(s[0] + s[1] + s[2] == 1010.0)      # closed population, every particle
    na = np.zeros((t.size, P), dtype=bool)
    for lo, hi in ((0, 4), (4, 8)):
        shard = g.new(sets[lo:hi], 0, n, seed=words[lo * n:hi * n].tobytes(), pars_multi=True)
        shard.set_data(data, True)
        shard.set_filter_stream(words[-1].tobytes())
        shard.set_stream_sharing(P, lo, na)
        assert np.array_equal(shard.filter()["log_likelihood"], ll[lo:hi])
        assert np.array_equal(shard.state(), s[:, :, lo:hi])


def test_c5_smc2_block_of_128_parameter_particles(dust, sir_data):
    """One GPU's share of config C5: 128 parameter particles x 2^14 state particles, one
    batched handle with per-set seeds (mcstate_b200/smc2.py), over the first 25 days."""
    from mcstate_b200.smc2 import smc2_control, smc2_engine, smc2_parameter, smc2_parameters

    data = particle_filter_data({"day": sir_data["day"][:25],
                                 "incidence": sir_data["incidence"][:25]}, "day", 4, 0)
    f = particle_filter(data, dust.dust_example("sir"), 1 << 14, None, seed=1, gpu_config=0)
    unif = lambda k, rng: rng.uniform(0.05, 0.5, k)  # noqa: E731
    dunif = lambda v: 0.0 if 0.05 <= v <= 0.5 else -math.inf  # noqa: E731
    p = smc2_parameters([smc2_parameter("beta", unif, dunif, min=0.05, max=0.5),
                         smc2_parameter("gamma", unif, dunif, min=0.05, max=0.5)])
    eng = smc2_engine(p, f, smc2_control(128, seed=4))
    t0 = time.perf_counter()
    eng.run()
    secs = time.perf_counter() - t0
    res = eng.results()
    assert res["pars"].shape == (128, 2)
    assert np.isclose(res["probabilities"][:, 3].sum(), 1.0)
    assert np.isfinite(res["probabilities"][:, 1]).all()
    assert res["statistics"]["n_particles"] == 1 << 14
    assert np.nanmin(res["statistics"]["ess"]) > 0
    assert secs < 120, secs
