"""SMC^2 on the batched filter (mcstate_b200.smc2) -- tests/testthat/test-smc2.R:1-50 and
the per-set seeding the engine relies on (R/smc2.R:68-78)."""
import math
import os
import pickle
import subprocess
import sys
import textwrap

import numpy as np
import pytest
from conftest import free_port

from mcstate_b200.particle_filter import particle_filter
from mcstate_b200.particle_filter_data import particle_filter_data
from mcstate_b200.smc2 import (cov_wt, smc2, smc2_control, smc2_engine, smc2_parameter,
                               smc2_parameters, smc2_predict)
from oracle_generator import OracleGenerator

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BACKENDS = [pytest.param("oracle", id="oracle"),
            pytest.param("gpu", id="b200", marks=pytest.mark.gpu)]


def make_generator(kind):
    if kind == "oracle":
        return OracleGenerator("sir")
    from mcstate_b200 import dust

    g = dust.dust_example("sir")
    if not g.gpu_info()["has_cuda"]:
        raise RuntimeError("no CUDA device visible: the engine has no CPU fallback")
    return g


def load_sir():
    d = np.loadtxt(os.path.join(ROOT, "tests", "golden", "sir_incidence.csv"), delimiter=",",
                   skiprows=1)
    return {"day": d[:, 1].astype(int), "incidence": d[:, 0]}


def _device():
    import torch

    return int(os.environ.get("LOCAL_RANK", "0")) % max(1, torch.cuda.device_count())


def example(kind, n_particles=42, n_days=None):
    sir = load_sir()
    data = particle_filter_data({"day": sir["day"][:n_days], "incidence": sir["incidence"][:n_days]},
                                "day", 4, 0)
    f = particle_filter(data, make_generator(kind), n_particles, None, seed=1,
                        gpu_config=_device() if kind == "gpu" else None)
    unif = lambda n, rng: rng.uniform(0, 1, n)  # noqa: E731
    dunif = lambda x: 0.0 if 0 <= x <= 1 else -math.inf  # noqa: E731
    p = smc2_parameters([smc2_parameter("beta", unif, dunif, min=0, max=1),
                         smc2_parameter("gamma", unif, dunif, min=0, max=1)])
    return f, p


def test_cov_wt_matches_definition():
    rng = np.random.default_rng(3)
    x, w = rng.normal(size=(50, 2)), rng.uniform(size=50)
    wn = w / w.sum()
    mu = (wn[:, None] * x).sum(0)
    want = sum(wi * np.outer(xi - mu, xi - mu) for wi, xi in zip(wn, x)) / (1 - np.sum(wn ** 2))
    assert np.allclose(cov_wt(x, w), want)
    assert np.allclose(cov_wt(x, np.ones(50)), np.cov(x.T))  # equal weights: the usual one


@pytest.mark.parametrize("kind", BACKENDS)
def test_can_run_smc2(kind):
    """tests/testthat/test-smc2.R:3-50: shapes, names, the effort band."""
    f, p = example(kind)
    res = smc2(p, f, smc2_control(20, seed=2))
    assert res["pars"].shape == (20, 2) and res["pars_names"] == ["beta", "gamma"]
    assert res["probabilities"].shape == (20, 4)
    assert np.isclose(res["probabilities"][:, 3].sum(), 1.0)
    st = res["statistics"]
    assert set(st) == {"ess", "acceptance_rate", "n_particles", "n_parameter_sets", "n_steps",
                       "effort"}
    assert st["n_particles"] == 42 and st["n_parameter_sets"] == 20 and st["n_steps"] == 100
    assert 150 < st["effort"] < 500
    assert smc2_predict(res, rng=np.random.default_rng(1)).shape == (20, 2)
    assert smc2_predict(res, 100, np.random.default_rng(1)).shape == (100, 2)
    assert smc2_predict(res, 1, np.random.default_rng(1)).shape == (1, 2)
    # replenishment moved the particles to where the data can be explained at all
    assert res["probabilities"][:, 1].max() > -330


@pytest.mark.parametrize("kind", BACKENDS)
def test_engine_is_deterministic_and_counts_work(kind):
    f, p = example(kind, n_days=30)
    a = smc2_engine(p, f, smc2_control(8, seed=5)); a.run()
    f2, p2 = example(kind, n_days=30)
    b = smc2_engine(p2, f2, smc2_control(8, seed=5)); b.run()
    assert np.array_equal(a.pars, b.pars)
    assert np.array_equal(a.log_likelihood, b.log_likelihood)
    n_repl = int(np.sum(~np.isnan(a.acceptance_rate)))
    assert n_repl >= 1
    # rows run: one per step per set, plus t rows per set at each replenishment at step t
    want = 8 * 30 + 8 * sum(t + 1 for t in range(30) if not np.isnan(a.acceptance_rate[t]))
    assert a.n_filter_rows == want


@pytest.mark.gpu
def test_batched_per_set_model_equals_independent_oracle_filters():
    """MCB_SEED_PER_SET: a pars_multi handle with one seed per set holds exactly the
    streams of separate model objects (R/smc2.R:68-78), so its incremental filter gives,
    bit for bit, what independent single-set oracle filters give."""
    from mcstate_b200 import dust
    from mcstate_b200.dust import dust_rng_distributed_state

    sir = load_sir()
    t = (sir["day"] * 4).astype(np.uint64)
    data = dust.dust_data({"time_end": t, "incidence": sir["incidence"]})
    pars = [{"beta": 0.15 + 0.03 * k, "gamma": 0.08 + 0.01 * k} for k in range(5)]
    seeds = dust_rng_distributed_state(11, 1, 5)
    m = dust.dust_example("sir").new(pars, 0, 1000, seed=b"".join(seeds), pars_multi=True,
                                     seed_per_set=True, gpu_config=0)
    m.set_data(data, True)
    o = OracleGenerator("sir").new(pars, 0, 1000, seed=b"".join(seeds), pars_multi=True,
                                   seed_per_set=True)
    o.set_data(data, True)
    for te in (40, 200, 400):
        assert np.array_equal(m.filter(te)["log_likelihood"], o.filter(te)["log_likelihood"])
    assert m.rng_state() == o.rng_state()
    assert np.array_equal(m.state(), o.state())
    # per-set copies between two handles
    m2 = dust.dust_example("sir").new(pars, 0, 1000, seed=b"".join(seeds), pars_multi=True,
                                      seed_per_set=True, gpu_config=0)
    take = np.array([1, 0, 1, 0, 0], dtype=bool)
    before = m2.state().copy()
    m2.copy_sets_from(m, take, state=True, rng=True)
    s2, s = m2.state(), m.state()
    assert np.array_equal(s2[:, :, take], s[:, :, take])
    assert np.array_equal(s2[:, :, ~take], before[:, :, ~take])


_WORKER = textwrap.dedent("""
    import os, sys, pickle
    import numpy as np
    import torch.distributed as dist
    sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
    from test_smc2 import example
    from mcstate_b200.smc2 import smc2, smc2_control
    kind = sys.argv[1]
    backend = "gloo"
    if kind == "gpu":   # one GPU per rank and NCCL when the box has the GPUs, else gloo
        import torch
        backend = os.environ.get("MCB_TEST_BACKEND") or (
            "nccl" if torch.cuda.device_count() >= int(os.environ["WORLD_SIZE"]) else "gloo")
    if backend == "nccl":
        import torch
        torch.cuda.set_device(int(os.environ["LOCAL_RANK"]) % torch.cuda.device_count())
    import datetime
    dist.init_process_group(backend, timeout=datetime.timedelta(seconds=120))
    f, p = example(kind, n_days=30)
    res = smc2(p, f, smc2_control(7, seed=5))
    if dist.get_rank() == 0:
        with open(sys.argv[2], "wb") as fh:
            pickle.dump(res, fh)
    dist.barrier()
    dist.destroy_process_group()
""")


@pytest.mark.parametrize("kind", BACKENDS)
def test_sharded_smc2_equals_single_rank(tmp_path, kind):
    """7 parameter particles over 2 ranks (4 + 3): per-set seeds make the result
    independent of the placement -- identical to the single-process run."""
    f, p = example(kind, n_days=30)
    want = smc2(p, f, smc2_control(7, seed=5))
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    out = tmp_path / "res.pkl"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()),
           str(script), kind, str(out)]
    r = subprocess.run(cmd, env=dict(os.environ, MASTER_ADDR="127.0.0.1"), capture_output=True,
                       text=True, timeout=400)
    assert r.returncode == 0, r.stderr[-3000:]
    with open(out, "rb") as fh:
        got = pickle.load(fh)
    assert np.array_equal(want["pars"], got["pars"])
    assert np.array_equal(want["probabilities"], got["probabilities"])
    assert np.array_equal(want["statistics"]["ess"], got["statistics"]["ess"])


def test_smc2_parameters():
    """tests/testthat/test-smc2-parameters.R:6-112."""
    import warnings

    unif = lambda lo, hi: (lambda n, rng: rng.uniform(lo, hi, n))  # noqa: E731
    dunif = lambda lo, hi: (lambda x: -math.log(hi - lo) if lo <= x <= hi else -math.inf)  # noqa: E731
    p = smc2_parameter("a", unif(0, 10), dunif(0, 10), min=0, max=10, integer=True)
    assert (p.name, p.min, p.max, p.integer) == ("a", 0.0, 10.0, True)
    assert p.prior(1) == pytest.approx(math.log(0.1))
    assert np.array_equal(p.sample(20, np.random.default_rng(1)),
                          np.random.default_rng(1).uniform(0, 10, 20))
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        q = smc2_parameter("a", unif(0, 10), dunif(0, 10), min=0, max=10, discrete=True)
    assert q.integer and "'discrete' is deprecated" in str(w[0].message)
    with pytest.raises(ValueError, match=r"'max' must be > 'min' \(20\)"):
        smc2_parameter("a", unif(0, 10), dunif(0, 10), min=20, max=10)
    ps = smc2_parameters([smc2_parameter("beta", unif(0, 10), dunif(0, 10), min=0, max=10),
                          smc2_parameter("gamma", unif(1, 2), dunif(1, 2), min=1, max=2)])
    assert ps.names() == ["beta", "gamma"]
    assert ps.summary() == [
        {"name": "beta", "min": 0.0, "max": 10.0, "discrete": False, "integer": False},
        {"name": "gamma", "min": 1.0, "max": 2.0, "discrete": False, "integer": False}]
    theta = ps.sample(10, np.random.default_rng(2))
    assert theta.shape == (10, 2) and np.all((theta[:, 1] >= 1) & (theta[:, 1] <= 2))
    assert np.allclose(ps.prior(theta), math.log(0.1))
    ps = smc2_parameters([smc2_parameter("beta", unif(0, 10), dunif(0, 10), min=0, max=10),
                          smc2_parameter("gamma", lambda n, rng: rng.exponential(1.0, n),
                                         lambda x: -x)])
    theta = ps.sample(100, np.random.default_rng(3))
    new = ps.propose(theta, np.eye(2) * 100, np.random.default_rng(4))
    assert new.shape == (100, 2) and np.all((new[:, 0] > 0) & (new[:, 0] < 10))
    assert ps.model(theta[:2])[1] == {"beta": theta[1, 0], "gamma": theta[1, 1]}
    with pytest.raises(ValueError, match="Duplicate parameter names"):
        smc2_parameters([p, p])
