"""pMCMC driving the filter (mcstate_b200.pmcmc) against the relations the reference
asserts for pmcmc / pmcmc_parallel (file:line per test).

CPU runs use the oracle-backed generator (a test double, as the reference's own MCMC
tests use fake filters: tests/testthat/helper-mcstate.R:222-386); `-m gpu` runs the
same on the B200 engine.  The sharded path is covered with world_size-2 gloo processes.
"""
import math
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest
from conftest import free_port

from mcstate_b200.particle_filter import particle_filter
from mcstate_b200.particle_filter_data import particle_filter_data
from mcstate_b200.pmcmc import (chain_owner, make_seeds, pmcmc, pmcmc_control, pmcmc_parameter,
                                pmcmc_parameters, pmcmc_predict, reflect_proposal,
                                rmvnorm_generator)
from oracle_generator import OracleGenerator

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BACKENDS = [pytest.param("oracle", id="oracle"),
            pytest.param("gpu", id="b200", marks=pytest.mark.gpu)]


def _generator(kind):
    if kind == "oracle":
        return OracleGenerator("sir")
    from mcstate_b200 import dust

    g = dust.dust_example("sir")
    if not g.gpu_info()["has_cuda"]:
        pytest.fail("no CUDA device visible: the engine has no CPU fallback")
    return g


def example_pars():
    """tests/testthat/helper-mcstate.R:44-52."""
    kernel = [[0.00057, 0.00034], [0.00034, 0.00026]]
    flat = lambda p: math.log(1e-10)  # noqa: E731
    return pmcmc_parameters([pmcmc_parameter("beta", 0.2, min=0, max=1, prior=flat),
                             pmcmc_parameter("gamma", 0.1, min=0, max=1, prior=flat)], kernel)


def _device():
    """The GPU this process should use: its local rank under torchrun, else 0."""
    import torch

    return int(os.environ.get("LOCAL_RANK", "0")) % max(1, torch.cuda.device_count())


def example_filter(kind, sir_data, n_particles=64, seed=1):
    data = particle_filter_data({"day": sir_data["day"], "incidence": sir_data["incidence"]},
                                "day", 4, 0)
    index = lambda info: {"run": [5], "state": [1, 2, 3]}  # noqa: E731
    gpu = _device() if kind == "gpu" else None
    return particle_filter(data, _generator(kind), n_particles, None, index=index, seed=seed,
                           gpu_config=gpu)


# ---------------------------------------------------------------- pure host pieces
def test_reflect_proposal_known_answers():
    """tests/testthat/test-pmcmc.R:40-75."""
    inf = math.inf
    assert list(reflect_proposal([0.4, 1.2, 2.6], [0] * 3, [inf] * 3)) == [0.4, 1.2, 2.6]
    assert np.allclose(reflect_proposal([-0.4, 1.2, 2.6], [0] * 3, [inf] * 3), [0.4, 1.2, 2.6])
    assert np.allclose(reflect_proposal([0.4, 1.4, 2.6], [-inf] * 3, [1] * 3), [0.4, 0.6, -0.6])
    x = np.array([0.5, 1.25, 2.5, -0.25, -1.75])
    assert np.allclose(reflect_proposal(x, [0] * 5, [1] * 5), [0.5, 0.75, 0.5, 0.25, 0.25])


def test_rmvnorm_generator_moments_and_zero_rows():
    vcv = np.array([[4.0, 1.0, 0.0], [1.0, 2.0, 0.0], [0.0, 0.0, 0.0]])
    draw = rmvnorm_generator(vcv)
    rng = np.random.default_rng(1)
    x = np.array([draw([1.0, 2.0, 3.0], rng) for _ in range(20000)])
    assert np.all(x[:, 2] == 3.0)  # the all-zero row never moves (R/utils.R:49-53)
    assert np.allclose(x.mean(axis=0), [1, 2, 3], atol=0.05)
    assert np.allclose(np.cov(x[:, :2].T), vcv[:2, :2], atol=0.1)
    with pytest.raises(ValueError, match="symmetric"):
        rmvnorm_generator([[1.0, 0.5], [0.0, 1.0]])


def test_control_validation_and_retention():
    """tests/testthat/test-pmcmc-control.R:4-120."""
    from mcstate_b200.pmcmc import pmcmc_check_control, pmcmc_filter_on_generation

    with pytest.raises(ValueError, match=r"'n_chains' \(2\) is less than 'n_workers' \(5\)"):
        pmcmc_control(100, n_chains=2, n_workers=5, n_threads_total=5)
    with pytest.raises(ValueError, match=r"'n_threads_total' \(3\) is less than 'n_workers' \(5\)"):
        pmcmc_control(100, n_chains=5, n_workers=5, n_threads_total=3)
    with pytest.raises(ValueError,
                       match=r"'n_threads_total' \(8\) is not a multiple of 'n_workers' \(5\)"):
        pmcmc_control(100, n_chains=5, n_workers=5, n_threads_total=8)
    assert pmcmc_control(100).save_restart is None
    assert pmcmc_control(100, save_restart=1).save_restart == [1]
    assert pmcmc_control(100, save_restart=[10, 20]).save_restart == [10, 20]
    with pytest.raises(ValueError, match="'save_restart' must be strictly increasing"):
        pmcmc_control(100, save_restart=[20, 10])
    # filter on generation: the kept steps end at the last one
    assert pmcmc_filter_on_generation(100, None, None) == \
        {"n_burnin": 0, "n_steps_retain": 100, "n_steps_every": 1}
    dat = pmcmc_filter_on_generation(100, 40, 20)
    assert dat == {"n_burnin": 42, "n_steps_retain": 20, "n_steps_every": 3}
    steps = [dat["n_burnin"] + 1 + k * dat["n_steps_every"] for k in range(dat["n_steps_retain"])]
    assert steps == list(range(43, 101, 3))
    for args in ((10, 100, 5), (100, 100, 5)):
        with pytest.raises(ValueError,
                           match="'n_burnin' cannot be greater than or equal to 'n_steps'"):
            pmcmc_filter_on_generation(*args)
    with pytest.raises(ValueError,
                       match="'n_steps_retain' is too large, max possible is 90 but given 500"):
        pmcmc_filter_on_generation(100, 10, 500)
    with pytest.raises(ValueError, match="'n_steps_retain' is too large to skip any samples,"):
        pmcmc_filter_on_generation(100, 10, 75)
    c = pmcmc_control(100, n_burnin=10, n_steps_retain=30)
    assert (c.n_steps_every, c.n_burnin, c.n_steps_retain) == (3, 12, 30)
    c = pmcmc_control(100, n_steps_retain=15, n_burnin=5)
    pmcmc_check_control(c)
    c.n_steps = 30
    with pytest.raises(ValueError, match=r"Corrupt pmcmc_control \(n_steps/n_steps_retain/n_burnin\)"):
        pmcmc_check_control(c)
    with pytest.warns(UserWarning,
                      match="'path' given when n_workers = 1 has no effect and is ignored"):
        pmcmc_control(10, path="location")
    assert pmcmc_control(10, n_chains=2, n_workers=2).use_parallel_seed


def test_retained_steps_end_at_the_last_one(sir_data):
    """tests/testthat/test-pmcmc.R:609-645 ("Can filter pmcmc on creation"): a run that
    retains 7 of 30 steps after a burn-in equals the full run at those iterations."""
    full = pmcmc(example_pars(), example_filter("oracle", sir_data, 10),
                 control=pmcmc_control(30, save_state=True, save_trajectories=True,
                                       use_parallel_seed=True))
    part = pmcmc(example_pars(), example_filter("oracle", sir_data, 10),
                 control=pmcmc_control(30, save_state=True, save_trajectories=True,
                                       use_parallel_seed=True, n_burnin=5, n_steps_retain=7))
    assert list(part["iteration"]) == list(range(12, 31, 3)) and part["iteration"][-1] == 30
    i = part["iteration"] - 1
    assert np.array_equal(part["pars"], full["pars"][i])
    assert np.array_equal(part["probabilities"], full["probabilities"][i])
    assert np.array_equal(part["state"], full["state"][:, i])
    assert np.array_equal(part["trajectories"], full["trajectories"][:, i])


def test_make_seeds_deterministic_prefix_stable():
    """tests/testthat/test-pmcmc-parallel.R:40-60."""
    s2, s5 = make_seeds(2, 1), make_seeds(5, 1)
    assert all(len(s["dust"]) == 32 for s in s5)
    assert [s["dust"] for s in s5[:2]] == [s["dust"] for s in s2]  # prefix stable
    assert [s["r"] for s in s5[:2]] == [s["r"] for s in s2]
    from mcstate_b200.dust import dust_rng

    assert s5[0]["dust"] == dust_rng(1).state()  # chain 1 runs on the input seed
    assert s5[1]["dust"] == dust_rng(1).long_jump().state()
    assert all(1 <= s["r"] <= 2 ** 24 for s in s5)
    raw = make_seeds(3, s5[1]["dust"] + s5[2]["dust"])  # a raw 64-byte seed: two streams
    assert all(len(s["dust"]) == 64 for s in raw)
    assert [chain_owner(k, 2) for k in range(5)] == [0, 1, 0, 1, 0]


# ---------------------------------------------------------------- chains on a filter
@pytest.mark.parametrize("kind", BACKENDS)
def test_pmcmc_runs_and_is_reproducible(kind, sir_data):
    """tests/testthat/test-pmcmc.R:80-108: shapes, and same seeds => same chain."""
    control = pmcmc_control(30, n_chains=1, save_state=True, save_trajectories=True,
                            use_parallel_seed=True)
    a = pmcmc(example_pars(), example_filter(kind, sir_data), control=control)
    b = pmcmc(example_pars(), example_filter(kind, sir_data), control=control)
    assert a["pars"].shape == (30, 2) and a["probabilities"].shape == (30, 3)
    assert a["state"].shape == (5, 30)
    assert a["trajectories"].shape == (3, 30, 101)
    assert np.array_equal(a["pars"], b["pars"])
    assert np.array_equal(a["probabilities"], b["probabilities"])
    assert np.array_equal(a["trajectories"], b["trajectories"])
    assert np.allclose(a["probabilities"][:, 0] + a["probabilities"][:, 1],
                       a["probabilities"][:, 2])
    assert a["n_filter_runs"] == 31  # one at initialisation, one per step
    # lineages: S never rises, R never falls (tests/testthat/test-particle-filter.R:270-293)
    assert np.all(np.diff(a["trajectories"][0], axis=1) <= 0)
    assert np.all(np.diff(a["trajectories"][2], axis=1) >= 0)


@pytest.mark.parametrize("kind", BACKENDS)
def test_chain_independent_of_what_is_saved(kind, sir_data):
    """tests/testthat/test-pmcmc.R:110-128: saving state/trajectories changes nothing.
    (The particle index is drawn from the host RNG either way.)"""
    pars = example_pars()
    c1 = pmcmc_control(20, save_state=False, save_trajectories=False, use_parallel_seed=True)
    c2 = pmcmc_control(20, save_state=True, save_trajectories=True, use_parallel_seed=True)
    a = pmcmc(pars, example_filter(kind, sir_data), control=c1)
    b = pmcmc(pars, example_filter(kind, sir_data), control=c2)
    assert a["state"] is None and a["trajectories"] is None
    assert np.array_equal(a["pars"], b["pars"])
    assert np.array_equal(a["probabilities"], b["probabilities"])


@pytest.mark.parametrize("kind", BACKENDS)
def test_early_exit_does_not_change_the_chain(kind, sir_data):
    """filter_early_exit only skips work for proposals that would be rejected anyway
    (R/pmcmc_state.R:85-99); on the compiled path the reference never forwards the
    bound (R/particle_filter_state.R:151), so the chain is identical."""
    pars = example_pars()
    a = pmcmc(pars, example_filter(kind, sir_data),
              control=pmcmc_control(20, use_parallel_seed=True))
    b = pmcmc(pars, example_filter(kind, sir_data),
              control=pmcmc_control(20, use_parallel_seed=True, filter_early_exit=True))
    assert np.array_equal(a["pars"], b["pars"])


@pytest.mark.parametrize("kind", BACKENDS)
def test_multiple_chains_and_rerun(kind, sir_data):
    control = pmcmc_control(12, n_chains=3, use_parallel_seed=True, rerun_every=4)
    res = pmcmc(example_pars(), example_filter(kind, sir_data), control=control)
    assert res["pars"].shape == (36, 2)
    assert list(np.unique(res["chain"])) == [1, 2, 3]
    assert res["n_filter_runs"] == 3 * (1 + 12 + 3)  # + a rerun at steps 4, 8, 12
    assert not np.array_equal(res["pars"][:12], res["pars"][12:24])  # chains differ
    # chain 1 of a 3-chain run == the single-chain run (same seeds: prefix stability)
    one = pmcmc(example_pars(), example_filter(kind, sir_data),
                control=pmcmc_control(12, n_chains=1, use_parallel_seed=True, rerun_every=4))
    assert np.array_equal(one["pars"], res["pars"][:12])


def test_posterior_moves_towards_the_data(sir_data):
    """A sanity check of the sampler itself: starting from poor parameters the chain
    climbs (tests/testthat/test-pmcmc.R uses the same style of check on its fixtures)."""
    f = example_filter("oracle", sir_data, n_particles=128)
    res = pmcmc(example_pars(), f, initial=[0.35, 0.25],
                control=pmcmc_control(150, use_parallel_seed=True))
    assert res["probabilities"][-30:, 1].mean() > res["probabilities"][0, 1] + 5


# ---------------------------------------------------------------- sharded == serial
_WORKER = textwrap.dedent("""
    import os, sys, pickle
    import numpy as np
    import torch.distributed as dist
    sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
    from test_pmcmc import example_filter, example_pars
    from mcstate_b200.pmcmc import pmcmc, pmcmc_control
    kind, n_chains = sys.argv[1], int(sys.argv[2])
    backend = {backend!r}
    if kind == "gpu":   # one GPU per rank and NCCL when the box has the GPUs, else as asked
        import torch
        backend = os.environ.get("MCB_TEST_BACKEND") or (
            "nccl" if torch.cuda.device_count() >= int(os.environ["WORLD_SIZE"]) else {backend!r})
    if backend == "nccl":
        import torch
        torch.cuda.set_device(int(os.environ["LOCAL_RANK"]) % torch.cuda.device_count())
    import datetime
    dist.init_process_group(backend, timeout=datetime.timedelta(seconds=120))
    rank, world = dist.get_rank(), dist.get_world_size()
    d = np.loadtxt(os.path.join({root!r}, "tests", "golden", "sir_incidence.csv"), delimiter=",",
                   skiprows=1)
    sir = {{"day": d[:, 1].astype(int), "incidence": d[:, 0]}}
    control = pmcmc_control(10, n_chains=n_chains, n_workers=world, save_trajectories=True)
    res = pmcmc(example_pars(), example_filter(kind, sir), control=control)
    if rank == 0:
        assert res["predict"]["filter"]["model"] is not None
        res["predict"] = {{k: res["predict"][k] for k in ("time", "rate", "index")}}
        with open(sys.argv[3], "wb") as fh:
            pickle.dump(res, fh)
    dist.barrier()
    dist.destroy_process_group()
""")


def _run_sharded(tmp_path, kind, n_chains, world, backend):
    import pickle

    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT, backend=backend))
    out = tmp_path / "res.pkl"
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
           "--nproc-per-node=%d" % world, "--master-addr", "127.0.0.1", "--master-port",
           str(free_port()), str(script), kind, str(n_chains), str(out)]
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-3000:]
    with open(out, "rb") as fh:
        return pickle.load(fh)


def test_sharded_chains_equal_serial_gloo(tmp_path, sir_data):
    """tests/testthat/test-pmcmc-parallel.R:3-25: results do not depend on the number of
    workers.  Two gloo ranks, three chains (rank 0 runs chains 1 and 3)."""
    serial = pmcmc(example_pars(), example_filter("oracle", sir_data),
                   control=pmcmc_control(10, n_chains=3, use_parallel_seed=True,
                                         save_trajectories=True))
    sharded = _run_sharded(tmp_path, "oracle", 3, 2, "gloo")
    for key in ("pars", "probabilities", "chain", "iteration", "state", "trajectories"):
        assert np.array_equal(serial[key], sharded[key]), key


@pytest.mark.gpu
def test_sharded_chains_equal_serial_b200(tmp_path, sir_data):
    """The same on the B200 engine: two ranks, one chain at a time per GPU (both on device 0
    when the box has one GPU; MCB_TEST_BACKEND=nccl uses one GPU per rank and NCCL)."""
    serial = pmcmc(example_pars(), example_filter("gpu", sir_data),
                   control=pmcmc_control(10, n_chains=3, use_parallel_seed=True,
                                         save_trajectories=True))
    sharded = _run_sharded(tmp_path, "gpu", 3, 2, "gloo")
    for key in ("pars", "probabilities", "chain", "iteration", "state", "trajectories"):
        assert np.array_equal(serial[key], sharded[key]), key


# ---------------------------------------------------------------- pmcmc_predict (8f N3)
@pytest.mark.parametrize("kind", BACKENDS)
def test_predict_from_a_mcmc_run(kind, sir_data):
    """tests/testthat/test-predict.R:3-29, 43-56, 59-101."""
    p = example_filter(kind, sir_data, n_particles=100)
    control = pmcmc_control(30, save_state=True, save_trajectories=True)
    results = pmcmc(example_pars(), p, control=control)
    last = int(sir_data["day"][-1]) * 4
    assert results["predict"]["time"] == last and results["predict"]["rate"] == 4
    steps = [last + 4 * k for k in range(26)]
    y = pmcmc_predict(results, steps, seed=1)
    assert list(y["time"]) == steps and y["rate"] == 4 and y["predicted"].all()
    assert y["state"].shape == (3, 30, 26)
    assert np.array_equal(y["state"][:, :, 0], results["state"][:3, :])   # starts where the run ended
    assert np.all(np.diff(y["state"][0], axis=1) <= 0)                     # S falls
    assert np.all(np.diff(y["state"][2], axis=1) >= 0)                     # R rises
    # same seed, same prediction (test-predict.R:32-40)
    assert np.array_equal(pmcmc_predict(results, steps, seed=1)["state"], y["state"])
    # prepend the filter's trajectories (test-predict.R:43-56)
    y1 = pmcmc_predict(results, steps, prepend_trajectories=True, seed=1)
    assert list(y1["time"]) == list(range(0, last + 4 * 25 + 1, 4))
    assert list(y1["predicted"]) == [False] * 101 + [True] * 25
    assert np.array_equal(y1["state"][:, :, 100:], y["state"])
    assert np.array_equal(y1["state"][:, :, :101], results["trajectories"])
    with pytest.raises(ValueError, match="At least two times required"):
        pmcmc_predict(results, [last])
    with pytest.raises(ValueError, match="Expected times\\[1\\] to be %d" % last):
        pmcmc_predict(results, [last + 4, last + 8])
    no_traj = pmcmc(example_pars(), p, control=pmcmc_control(5, save_state=True,
                                                             save_trajectories=False))
    with pytest.raises(ValueError, match="can't prepend trajectories"):
        pmcmc_predict(no_traj, steps, prepend_trajectories=True)
    no_state = pmcmc(example_pars(), p, control=pmcmc_control(5, save_state=False))
    with pytest.raises(ValueError, match="return_state = FALSE, can't predict"):
        pmcmc_predict(no_state, steps)


@pytest.mark.gpu
def test_predict_engine_equals_oracle(sir_data):
    """The same retained samples carried forward by the engine and by the CPU oracle:
    identical on identical seeds."""
    res = {}
    for kind in ("oracle", "gpu"):
        p = example_filter(kind, sir_data, n_particles=64, seed=3)
        np.random.seed(0)
        r = pmcmc(example_pars(), p, control=pmcmc_control(12, save_state=True,
                                                           use_parallel_seed=True))
        last = r["predict"]["time"]
        res[kind] = (r, pmcmc_predict(r, [last + 2 * k for k in range(15)], seed=7))
    assert np.array_equal(res["oracle"][0]["pars"], res["gpu"][0]["pars"])
    assert np.array_equal(res["oracle"][0]["state"], res["gpu"][0]["state"])
    assert np.array_equal(res["oracle"][1]["state"], res["gpu"][1]["state"])


# ---------------------------------------------------------------- chains run by hand
from mcstate_b200.pmcmc import (pmcmc_chains_cleanup, pmcmc_chains_collect,  # noqa: E402
                                pmcmc_chains_prepare, pmcmc_chains_run)


def _same_samples(a, b):
    assert a["pars_names"] == b["pars_names"]
    for key in ("pars", "probabilities", "iteration", "chain", "state", "trajectories"):
        if a.get(key) is None:
            assert b.get(key) is None, key
        else:
            assert np.array_equal(a[key], b[key]), key
    assert a["n_filter_runs"] == b["n_filter_runs"]


@pytest.mark.parametrize("kind", BACKENDS)
def test_split_chains_manually(kind, sir_data, tmp_path):
    """tests/testthat/test-pmcmc-chains.R:1-28: prepare / run each chain / collect equals
    pmcmc() with the same control; cleanup removes what was written."""
    control = pmcmc_control(10, n_chains=4, use_parallel_seed=True, save_trajectories=True)
    res1 = pmcmc(example_pars(), example_filter(kind, sir_data, 30), control=control)
    path = str(tmp_path / "chains")
    assert pmcmc_chains_prepare(path, example_pars(), example_filter(kind, sir_data, 30),
                                control) == path
    device = _device() if kind == "gpu" else None
    results = [pmcmc_chains_run(k, path, device=device) for k in (3, 1, 4, 2)]  # any order
    assert sorted(os.listdir(path)) == sorted(
        ["control.pkl", "inputs.pkl"] + ["results_%d.pkl" % k for k in range(1, 5)])
    assert results[1] == os.path.join(path, "results_1.pkl")
    res2 = pmcmc_chains_collect(path)
    _same_samples(res1, res2)
    assert res2["predict"]["filter"]["model"].name == "dust_generator"
    pmcmc_chains_cleanup(path)
    assert not os.path.exists(path)


def test_split_chains_in_other_processes(sir_data, tmp_path):
    """What the files are for: each chain in its own interpreter, nothing shared but the
    directory (the reference: one callr / cluster job per chain)."""
    import subprocess
    import sys

    control = pmcmc_control(6, n_chains=2, use_parallel_seed=True)
    res1 = pmcmc(example_pars(), example_filter("oracle", sir_data, 20), control=control)
    path = str(tmp_path / "chains")
    pmcmc_chains_prepare(path, example_pars(), example_filter("oracle", sir_data, 20), control)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for k in (1, 2):
        code = ("import sys; sys.path[:0] = [%r, %r]; "
                "from mcstate_b200.pmcmc import pmcmc_chains_run; "
                "print(pmcmc_chains_run(%d, %r))" % (root, os.path.join(root, "tests"), k, path))
        out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True,
                             timeout=300)
        assert out.returncode == 0, out.stderr[-2000:]
        assert out.stdout.strip().endswith("results_%d.pkl" % k)
    _same_samples(res1, pmcmc_chains_collect(path))


def test_split_chains_errors(sir_data, tmp_path):
    """tests/testthat/test-pmcmc-chains.R:62-124."""
    f = example_filter("oracle", sir_data, 5)
    with pytest.raises(ValueError, match="'n_workers' must be 1"):
        pmcmc_chains_prepare(str(tmp_path / "a"), example_pars(), f,
                             pmcmc_control(10, n_chains=4, n_workers=2, use_parallel_seed=True))
    with pytest.raises(ValueError, match="'use_parallel_seed' must be TRUE"):
        pmcmc_chains_prepare(str(tmp_path / "b"), example_pars(), f,
                             pmcmc_control(10, n_chains=4, use_parallel_seed=False))
    control = pmcmc_control(3, n_chains=4, use_parallel_seed=True)
    path = pmcmc_chains_prepare(str(tmp_path / "c"), example_pars(), f, control)
    with pytest.raises(RuntimeError, match="Results missing for chains 1, 2, 3, 4"):
        pmcmc_chains_collect(path)
    pmcmc_chains_run(2, path)
    pmcmc_chains_run(3, path)
    with pytest.raises(RuntimeError, match="Results missing for chains 1, 4"):
        pmcmc_chains_collect(path)
    with pytest.raises(ValueError, match="'chain_id' must be at least 1"):
        pmcmc_chains_run(0, path)
    with pytest.raises(ValueError, match=r"'chain_id' must be an integer in 1\.\.4"):
        pmcmc_chains_run(5, path)


def test_generator_survives_serialisation():
    """dust::dust_repair_environment (R/pmcmc_chains.R:78): a generator read back in has
    no library handle until it is repaired; stock and user-supplied models."""
    import pickle

    import cloudpickle

    from mcstate_b200 import dust

    for g in (dust.dust_example("sirs", "xoshiro256starstar"),
              dust.dust(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                     "mcstate_b200", "examples", "seir.cuh"))):
        h = pickle.loads(cloudpickle.dumps(g))
        assert h._L is g._L and h.public_methods is h
        assert (h.model_name, h.model_id, h.scrambler, h.par_names, h.state_names) == \
            (g.model_name, g.model_id, g.scrambler, g.par_names, g.state_names)
        h._L = None
        assert dust.dust_repair_environment(h)._L is g._L
    assert dust.dust_repair_environment("not a generator") == "not a generator"
    assert dust.dust_openmp_threads(1) == 1
    assert dust.dust_openmp_threads(10 ** 6, "fix") == dust.dust_openmp_threads()
    with pytest.raises(ValueError, match="exceeds a limit"):
        dust.dust_openmp_threads(10 ** 6)


# ---------------------------------------------------------------- thin / sample / restart
from mcstate_b200.particle_filter import particle_filter_initial  # noqa: E402
from mcstate_b200.pmcmc import pmcmc_sample, pmcmc_thin  # noqa: E402


@pytest.fixture(scope="module")
def sir_pmcmc(sir_data):
    """tests/testthat/helper-mcstate.R example_sir_pmcmc: 30 steps, state, trajectories and
    restart state saved."""
    control = pmcmc_control(30, save_state=True, save_trajectories=True, save_restart=[40],
                            use_parallel_seed=True)
    return pmcmc(example_pars(), example_filter("oracle", sir_data, 20), control=control)


def test_pmcmc_thin(sir_pmcmc):
    """tests/testthat/test-pmcmc-tools.R:3-68."""
    res = sir_pmcmc
    same = pmcmc_thin(res)
    for k in ("pars", "probabilities", "state", "trajectories", "restart", "iteration"):
        assert np.array_equal(same[k], res[k])
    for args, i in (((9, None), np.arange(9, 30)), ((None, 4), np.arange(0, 30, 4)),
                    ((9, 4), np.arange(9, 30, 4))):
        out = pmcmc_thin(res, *args)
        assert np.array_equal(out["pars"], res["pars"][i])
        assert np.array_equal(out["probabilities"], res["probabilities"][i])
        assert np.array_equal(out["state"], res["state"][:, i])
        assert np.array_equal(out["trajectories"], res["trajectories"][:, i])
        assert np.array_equal(out["restart"], res["restart"][:, i])
        assert np.array_equal(out["iteration"], i + 1)
    for b in (30, 100):
        with pytest.raises(ValueError, match="'burnin' must be less than 30 for your results"):
            pmcmc_thin(res, b)
    bare = dict(res, state=None, trajectories=None)
    out = pmcmc_thin(bare, 9, 4)
    assert out["state"] is None and out["trajectories"] is None
    assert np.array_equal(out["pars"], res["pars"][np.arange(9, 30, 4)])


def test_pmcmc_thin_and_sample_combined_chains(sir_data):
    """tests/testthat/test-pmcmc-tools.R:135-160, 276-299."""
    control = pmcmc_control(30, n_chains=3, save_state=True, use_parallel_seed=True)
    res = pmcmc(example_pars(), example_filter("oracle", sir_data, 10), control=control)
    out = pmcmc_thin(res, 10)
    assert np.array_equal(out["chain"], np.repeat([1, 2, 3], 20))
    assert np.array_equal(out["iteration"], np.tile(np.arange(11, 31), 3))
    out = pmcmc_thin(res, thin=4)
    assert np.array_equal(out["iteration"], np.tile(np.arange(1, 31, 4), 3))
    assert np.array_equal(out["state"], res["state"][:, np.isin(res["iteration"], np.arange(1, 31, 4))])
    sub = pmcmc_sample(res, 50, burnin=9, rng=np.random.default_rng(1))
    assert sub["pars"].shape == (50, 2) and sub["state"].shape == (5, 50)
    assert np.all(sub["iteration"] >= 10) and set(sub["chain"]) == {1, 2, 3}
    one = pmcmc_sample(pmcmc_thin(res, 25), 50, rng=np.random.default_rng(2))
    assert len(set(zip(one["chain"], one["iteration"]))) < 50        # with replacement


def test_particle_filter_initial():
    """tests/testthat/test-pmcmc.R:598-606."""
    m = np.arange(70).reshape(7, 10).T          # R: matrix(1:70 - 1, 10)
    f = particle_filter_initial(m, np.random.default_rng(1))
    res = f(None, 3, None)
    assert res.shape == (10, 3)
    assert np.array_equal(res % 10, np.tile(np.arange(10)[:, None], (1, 3)))
    assert np.all(np.diff(res // 10, axis=0) == 0)
    m[:] = -1                                    # the closure holds its own copy
    assert f(None, 2, None).min() >= 0
    with pytest.raises(TypeError, match="'state' must be a matrix"):
        particle_filter_initial(np.zeros(3))


@pytest.mark.parametrize("kind", BACKENDS)
def test_restart_pmcmc_from_saved_state(kind, sir_data):
    """tests/testthat/test-pmcmc.R:478-508: a second pmcmc on the later part of the data,
    started from the restart state the first one saved."""
    control1 = pmcmc_control(20, save_state=True, save_trajectories=True, save_restart=[40],
                             use_parallel_seed=True)
    res1 = pmcmc(example_pars(), example_filter(kind, sir_data, 50), control=control1)
    assert res1["restart"].shape == (5, 20, 1)
    s = res1["restart"][:, :, 0]
    late = sir_data["day"] > 40
    data2 = particle_filter_data({"day": sir_data["day"][late],
                                  "incidence": sir_data["incidence"][late]}, "day", 4, 40)
    gpu = _device() if kind == "gpu" else None
    p2 = particle_filter(data2, _generator(kind), 50, None,
                         index=lambda info: {"run": [5], "state": [1, 2, 3]}, seed=1,
                         initial=particle_filter_initial(s, np.random.default_rng(4)),
                         gpu_config=gpu)
    control2 = pmcmc_control(20, save_state=True, save_trajectories=True, use_parallel_seed=True)
    res2 = pmcmc(example_pars(), p2, control=control2)
    assert res2["trajectories"].shape == (3, 20, 61)
    assert np.array_equal(res2["trajectories_time"], np.arange(40, 101) * 4)
    # every starting particle is one of the saved columns
    start = res2["trajectories"][:, :, 0]
    assert all(any(np.array_equal(start[:, j], s[:3, k]) for k in range(s.shape[1]))
               for j in range(start.shape[1]))


def test_thread_allocation():
    """tests/testthat/test-pmcmc-parallel.R:180-190 (known answers)."""
    from mcstate_b200.pmcmc import pmcmc_parallel_threads as f

    assert f(1, 1, 1) == [1]
    assert f(20, 2, 2) == [10, 10]
    assert f(20, 2, 4) == [10] * 4
    assert f(20, 2, 5) == [10] * 4 + [20]
    assert f(6, 3, 7) == [2] * 6 + [6]
    assert f(6, 3, 8) == [2] * 6 + [3, 3]
    assert f(21, 3, 8) == [7] * 6 + [10, 11]


def test_fix_parameters():
    """tests/testthat/test-pmcmc-parameters.R:256-308."""
    rs = np.random.RandomState(1)
    names = list("abcde")
    vcv = np.cov(rs.uniform(size=(5, 5)), rowvar=False)
    initial, rate = rs.uniform(size=5), rs.uniform(size=5)
    plist = [pmcmc_parameter(n, i, prior=(lambda r: lambda p: math.log(r))(r))
             for n, i, r in zip(names, initial, rate)]
    p = pmcmc_parameters(plist, vcv)
    assert np.array_equal(p.vcv(), vcv) and np.array_equal(p.mean(), initial)
    assert [r["name"] for r in p.summary()] == names
    p2 = p.fix({"b": 0.5, "d": 0.2})
    assert p2.names() == ["a", "c", "e"]
    assert np.array_equal(p2.initial(), initial[[0, 2, 4]])
    assert np.allclose(p2.propose(p2.initial(), np.random.default_rng(1), 0.0), p2.initial())
    gen = np.random.default_rng(2)
    draws = np.array([p2.propose(initial[[0, 2, 4]], gen) for _ in range(20000)])
    assert np.allclose(np.cov(draws, rowvar=False), vcv[np.ix_([0, 2, 4], [0, 2, 4])], atol=1e-2)
    want = dict(zip(names, initial))
    want.update(b=0.5, d=0.2)
    assert p2.model(p2.initial()) == want
    assert math.isclose(p2.prior(p2.initial()), sum(math.log(rate[i]) for i in (0, 2, 4)))
    # a transform given to the original still runs, after the fixed values are put back
    p3 = pmcmc_parameters(plist, vcv, transform=lambda d: dict(d, total=sum(d.values())))
    got = p3.fix({"a": 1.0}).model(initial[1:])
    assert math.isclose(got["total"], 1.0 + initial[1:].sum()) and got["a"] == 1.0
    with pytest.raises(ValueError, match="'fixed' must be named"):
        p.fix([1, 2])
    with pytest.raises(ValueError, match="Fixed parameters not found in model: 'f'"):
        p.fix({"a": 1, "b": 2, "f": 1})
    with pytest.raises(ValueError, match="Cannot fix all parameters"):
        p.fix(dict.fromkeys(names, 1.0))


def test_pmcmc_with_fixed_parameter(sir_data):
    """tests/testthat/test-pmcmc.R:573-596: gamma fixed, beta sampled."""
    pars2 = example_pars().fix({"gamma": 0.1})
    control = pmcmc_control(10, save_state=True, save_trajectories=True, use_parallel_seed=True)
    res = pmcmc(pars2, example_filter("oracle", sir_data, 40), control=control)
    assert res["pars"].shape == (10, 1) and res["pars_names"] == ["beta"]
    assert res["state"].shape == (5, 10) and res["trajectories"].shape == (3, 10, 101)
    assert res["predict"]["transform"](res["pars"][-1])["gamma"] == 0.1


# ---------------------------------------------------------------- the MH machinery on known targets
class _ToyFilter:
    """tests/testthat/helper-mcstate.R:222-249, 312-335: a "filter" whose run() is a known
    log density -- the reference tests its MCMC machinery this way."""
    n_particles = 10
    has_multiple_parameters = False
    has_multiple_data = False

    def __init__(self, target):
        self._target = target

    def run(self, pars, *args):
        return self._target(pars)

    def state(self):
        return np.ones((2, 10))


def _unit_square_pars(proposal=None):
    flat = lambda p: 0.0 if 0.0 <= p <= 1.0 else -math.inf  # noqa: E731  dunif(p, log = TRUE)
    return pmcmc_parameters([pmcmc_parameter("a", 0.5, min=0, max=1, prior=flat),
                             pmcmc_parameter("b", 0.5, min=0, max=1, prior=flat)],
                            np.eye(2) * 0.1 if proposal is None else proposal)


def test_mcmc_uniform_on_unit_square():
    """tests/testthat/test-pmcmc.R:7-20: a flat target accepts every (reflected) proposal
    and the chain's means sit at the centre."""
    from mcstate_b200.pmcmc import pmcmc_run_chain

    control = pmcmc_control(1000, save_state=False, save_trajectories=False)
    res = pmcmc_run_chain(_unit_square_pars(), [0.5, 0.5], _ToyFilter(lambda p: 1.0), control,
                          np.random.default_rng(1))
    assert res["pars"].shape == (1000, 2)
    assert np.all(np.any(np.diff(res["pars"], axis=0) != 0, axis=1))      # acceptance rate 1
    assert np.all((res["pars"] >= 0) & (res["pars"] <= 1))
    assert abs(res["pars"][:, 0].mean() - 0.5) < 0.05 and abs(res["pars"][:, 1].mean() - 0.5) < 0.05


def test_mcmc_multivariate_gaussian():
    """tests/testthat/test-pmcmc.R:23-37: a standard bivariate normal target is recovered
    (Kolmogorov-Smirnov on thinned draws, uncorrelated components)."""
    from scipy import stats

    from mcstate_b200.pmcmc import pmcmc_run_chain

    pars = pmcmc_parameters([pmcmc_parameter("a", 0, min=-100, max=100),
                             pmcmc_parameter("b", 0, min=-100, max=100)], np.eye(2))
    target = lambda p: float(-0.5 * (p["a"] ** 2 + p["b"] ** 2) - math.log(2 * math.pi))  # noqa: E731
    control = pmcmc_control(4000, save_state=False, save_trajectories=False)
    res = pmcmc_run_chain(pars, [0.0, 0.0], _ToyFilter(target), control, np.random.default_rng(1))
    thin = res["pars"][::40]
    assert stats.kstest(thin[:, 0], "norm").pvalue > 0.05
    assert stats.kstest(thin[:, 1], "norm").pvalue > 0.05
    assert abs(np.cov(res["pars"], rowvar=False)[0, 1]) < 0.1
    assert np.allclose(res["probabilities"][:, 0] + res["probabilities"][:, 1],
                       res["probabilities"][:, 2])


def test_initial_conditions():
    """tests/testthat/test-pmcmc.R:270-337."""
    from mcstate_b200.pmcmc import pmcmc_check_initial

    sir, uni = example_pars(), _unit_square_pars()
    assert np.array_equal(pmcmc_check_initial(None, sir, 1), [[0.2, 0.1]])
    assert np.array_equal(pmcmc_check_initial(None, sir, 5), np.tile([0.2, 0.1], (5, 1)))
    for given in ([0.1, 0.2], {"a": 0.1, "b": 0.2}):
        assert np.array_equal(pmcmc_check_initial(given, uni, 1), [[0.1, 0.2]])
        assert np.array_equal(pmcmc_check_initial(given, uni, 5), np.tile([0.1, 0.2], (5, 1)))
    with pytest.raises(ValueError, match="Expected 'initial' to be a vector with length 2"):
        pmcmc_check_initial([0.1, 0.2, 0.4], uni, 1)
    with pytest.raises(ValueError,
                       match=r"Expected names of 'initial' to match parameters \('a', 'b'\)"):
        pmcmc_check_initial({"x": 0.1, "y": 0.2}, uni, 1)
    with pytest.raises(ValueError, match="Starting point does not have finite prior probability"):
        pmcmc_check_initial([-0.1, 0.2], uni, 1)
    # a matrix has one COLUMN per chain, as in R
    assert np.array_equal(pmcmc_check_initial(np.array([[0.1], [0.2]]), uni, 1), [[0.1, 0.2]])
    m = np.random.default_rng(1).uniform(size=(2, 5))
    assert np.array_equal(pmcmc_check_initial(m, uni, 5), m.T)
    for shape in ((3, 5), (2, 6)):
        with pytest.raises(ValueError,
                           match="Expected 'initial' to be a matrix with dimensions 2 x 5"):
            pmcmc_check_initial(np.full(shape, 0.5), uni, 5)
    m[[1, 0, 1], [1, 3, 4]] *= -1
    with pytest.raises(ValueError, match=r"Starting point does not have finite prior "
                                         r"probability \(chain 2, 4, 5\)"):
        pmcmc_check_initial(m, uni, 5)


def test_pmcmc_from_a_matrix_of_starting_points(sir_data):
    """tests/testthat/test-pmcmc.R:340-362."""
    initial = np.array([[0.2], [0.1]]) + np.random.default_rng(2).uniform(0, 0.001, (2, 3))
    control = pmcmc_control(2, n_chains=3, use_parallel_seed=True)
    res = pmcmc(example_pars(), example_filter("oracle", sir_data, 10), initial=initial,
                control=control)
    first = res["pars"][res["iteration"] == 1]
    assert first.shape == (3, 2) and len({tuple(r) for r in first}) == 3


def test_rerun_schedule():
    """tests/testthat/test-pmcmc-support.R:3-25."""
    from mcstate_b200.pmcmc import _make_rerun

    gen = np.random.default_rng(1)
    fixed, rand = _make_rerun(3, False, gen), _make_rerun(3, True, gen)
    i = range(11)
    x = np.array([[fixed(k) for k in i] for _ in range(1000)])
    y = np.array([[rand(k) for k in i] for _ in range(1000)])
    assert np.array_equal(x, np.tile(np.resize([True, False, False], 11), (1000, 1)))
    assert not np.array_equal(x, y) and 0.32 < y.mean() < 0.345
    for random in (False, True):
        never = _make_rerun(math.inf, random, gen)
        assert not any(never(k) for k in i)


def test_names_on_pmcmc_output_and_predictions(sir_data):
    """tests/testthat/test-pmcmc.R:239-267, tests/testthat/test-predict.R:129-141: a named
    ``index$state`` names the rows of the saved trajectories and of what is predicted from
    them (numpy arrays carry no row names: they come beside the arrays)."""
    data = particle_filter_data({"day": sir_data["day"], "incidence": sir_data["incidence"]},
                                "day", 4, 0)
    control = pmcmc_control(5, n_chains=2, save_state=True, save_trajectories=True,
                            use_parallel_seed=True)
    out = {}
    for key, state in (("plain", [1, 2, 3]), ("named", {"a": 1, "b": 2, "c": 3})):
        f = particle_filter(data, OracleGenerator("sir"), 10, None, seed=1,
                            index=lambda info, state=state: {"run": [4], "state": state})
        out[key] = pmcmc(example_pars(), f, control=control)
    assert out["plain"]["trajectories_names"] is None
    assert out["named"]["trajectories_names"] == ["a", "b", "c"]
    assert np.array_equal(out["plain"]["trajectories"], out["named"]["trajectories"])
    steps = [400, 404, 408]
    assert pmcmc_predict(out["plain"], steps, seed=1)["names"] is None
    for prepend in (False, True):
        y = pmcmc_predict(out["named"], steps, prepend_trajectories=prepend, seed=1)
        assert y["names"] == ["a", "b", "c"] and y["state"].shape[0] == 3


def test_thread_count_and_filter_kind_checks(sir_data):
    """tests/testthat/test-pmcmc.R:413-428 (thread count via the control), :693-704 (a
    replicated, non-nested filter is refused)."""
    f = example_filter("oracle", sir_data, 30)
    res = pmcmc(example_pars(), f, control=pmcmc_control(5, n_threads_total=2))
    assert res["predict"]["filter"]["n_threads"] == 2
    f.set_n_threads(1)
    res = pmcmc(example_pars(), f, control=pmcmc_control(5))
    assert res["predict"]["filter"]["n_threads"] == 1
    data = particle_filter_data({"day": sir_data["day"], "incidence": sir_data["incidence"]},
                                "day", 4, 0)
    rep = particle_filter(data, OracleGenerator("sir"), 42, None, seed=1, n_parameters=3)
    with pytest.raises(ValueError, match="Can't use a filter with multiple parameter sets but "
                                         "not multiple data"):
        pmcmc(example_pars(), rep, control=pmcmc_control(2))


def test_rerun_every_reruns_the_filter(sir_data):
    """tests/testthat/test-pmcmc.R:365-410: with rerun_every = 2 no likelihood value
    survives more than two steps, and run() is called n_steps + 1 + n_steps // 2 times."""
    pars = pmcmc_parameters([pmcmc_parameter("beta", 0.2, min=0, max=1, prior=lambda p: -23.0),
                             pmcmc_parameter("gamma", 0.1, min=0, max=1, prior=lambda p: -23.0)],
                            np.eye(2) * 1e-4 * 50)

    def longest_run(x):
        best = cur = 1
        for a, b in zip(x, x[1:]):
            cur = cur + 1 if a == b else 1
            best = max(best, cur)
        return best

    for every, check in ((2, lambda n: n <= 2), (math.inf, lambda n: n > 5)):
        control = pmcmc_control(30, save_state=True, save_trajectories=True, rerun_every=every,
                                use_parallel_seed=True)
        res = pmcmc(pars, example_filter("oracle", sir_data, 100), control=control)
        assert check(longest_run(list(res["probabilities"][:, 1]))), every
    calls = []
    toy = _ToyFilter(lambda p: calls.append(1) or 1.0)
    from mcstate_b200.pmcmc import pmcmc_run_chain

    pmcmc_run_chain(_unit_square_pars(), [0.5, 0.5], toy,
                    pmcmc_control(10, rerun_every=2, save_state=False), np.random.default_rng(1))
    assert len(calls) == 16


def test_pmcmc_parameter_objects():
    """tests/testthat/test-pmcmc-parameters.R:3-52, 160-255, 310-330."""
    import warnings

    p = pmcmc_parameter("a", 1, 0, 10)
    assert (p.name, p.initial, p.min, p.max, p.integer, p.prior(1)) == ("a", 1.0, 0.0, 10.0, False, 0.0)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        assert pmcmc_parameter("a", 1, 0, 10, discrete=True).integer
    assert "'discrete' is deprecated" in str(w[0].message)
    f = lambda p: math.log(1 / p)  # noqa: E731
    assert pmcmc_parameter("a", 1, prior=f).prior is f
    for kw in ({}, {"min": 0}, {"max": 0}):
        pmcmc_parameter("a", 0, **kw)
    with pytest.raises(ValueError, match=r"'initial' must be >= 'min' \(1\)"):
        pmcmc_parameter("a", 0, min=1)
    with pytest.raises(ValueError, match=r"'initial' must be <= 'max' \(-1\)"):
        pmcmc_parameter("a", 0, max=-1)
    pmcmc_parameter("a", 10, prior=math.log)
    with pytest.raises(ValueError, match="Prior function for 'a' returned a non-finite value on "
                                         "initial value"):
        pmcmc_parameter("a", 0, prior=lambda p: -math.inf)

    def boom(p):
        raise RuntimeError("an error")

    with pytest.raises(ValueError, match="Prior function for 'a' failed to evaluate initial "
                                         "value: an error"):
        pmcmc_parameter("a", 0, prior=boom)
    with pytest.raises(TypeError, match="'parameters' must be a list"):
        pmcmc_parameters(None, np.eye(2))
    with pytest.raises(ValueError, match="At least one parameter is required"):
        pmcmc_parameters([], np.eye(2))
    with pytest.raises(TypeError, match="to be 'pmcmc_parameter' objects"):
        pmcmc_parameters([1, 2], np.eye(2))
    pars = [pmcmc_parameter("beta", 0.2), pmcmc_parameter("gamma", 0.1)]
    for bad in (np.eye(1), np.eye(3), np.zeros((3, 2))):
        with pytest.raises(ValueError,
                           match="Expected a square proposal matrix with 2 rows and columns"):
            pmcmc_parameters(pars, bad)
    with pytest.raises(ValueError, match="Duplicate parameter names: 'a'"):
        pmcmc_parameters([pmcmc_parameter("a", 0.2), pmcmc_parameter("a", 0.1)], np.ones((2, 2)))
    six = [pmcmc_parameter(n, 0.1) for n in "aaabcc"]
    with pytest.raises(ValueError, match="Duplicate parameter names: 'a', 'c'"):
        pmcmc_parameters(six, np.ones((6, 6)))
    ab = [pmcmc_parameter("a", 0.2), pmcmc_parameter("b", 0.1)]
    pmcmc_parameters({"a": ab[0], "b": ab[1]}, np.ones((2, 2)))
    with pytest.raises(ValueError,
                       match="'parameters' is named, but the names do not match parameters"):
        pmcmc_parameters({"b": ab[0], "a": ab[1]}, np.ones((2, 2)))


def test_pmcmc_parameters_propose_transform_prior():
    """tests/testthat/test-pmcmc-parameters.R:55-157."""
    from mcstate_b200.pmcmc import rmvnorm_generator

    kernel = np.eye(2) * 1e-4
    p = example_pars().__class__(
        [pmcmc_parameter("beta", 0.2, min=0, max=1, prior=lambda p: math.log(1e-10)),
         pmcmc_parameter("gamma", 0.1, min=0, max=1, prior=lambda p: math.log(1e-10))], kernel)
    p0 = p.initial()
    assert np.array_equal(p0, [0.2, 0.1])
    p1 = p.propose(p0, np.random.default_rng(1))
    p2 = p.propose(p0, np.random.default_rng(1), 10)
    assert np.array_equal(p1, rmvnorm_generator(kernel)(p0, np.random.default_rng(1)))
    assert np.allclose(p2 - p0, (p1 - p0) * math.sqrt(10), rtol=1e-12, atol=0)
    assert p.model(p0) == {"beta": 0.2, "gamma": 0.1}
    assert p.model(p0 + 0.5) == pytest.approx({"beta": 0.7, "gamma": 0.6})
    transform = lambda d: {"alpha": d["beta"] + d["gamma"], "gamma": [d["gamma"]] * 4,  # noqa: E731
                           "beta": d["beta"]}
    dexp = lambda x: -x  # noqa: E731
    dnorm = lambda x: -0.5 * x * x - 0.5 * math.log(2 * math.pi)  # noqa: E731
    q = pmcmc_parameters([pmcmc_parameter("beta", 0.2, prior=dexp),
                          pmcmc_parameter("gamma", 0.1, prior=dnorm)], np.eye(2), transform)
    got = q.model(q.initial())
    assert got["alpha"] == pytest.approx(0.3) and got["gamma"] == [0.1] * 4 and got["beta"] == 0.2
    assert q.prior(q.initial()) == pytest.approx(dexp(0.2) + dnorm(0.1))
    assert q.prior([0.2, 0.1]) == pytest.approx(dexp(0.2) + dnorm(0.1))       # by position
    assert q.prior({"beta": 1, "gamma": 0}) == pytest.approx(dexp(1) + dnorm(0))
    assert [r["name"] for r in p.summary()] == ["beta", "gamma"]
    assert all(r["min"] == 0 and r["max"] == 1 and not r["integer"] for r in p.summary())


def test_rmvnorm_generator_reference_cases():
    """tests/testthat/test-utils.R:26-72."""
    from mcstate_b200.pmcmc import rmvnorm_generator

    for bad in (np.array([[1.0, 3.0], [2.0, 4.0]]), np.ones((2, 6))):
        with pytest.raises(ValueError, match="vcv must be symmetric"):
            rmvnorm_generator(bad)
    with pytest.raises(ValueError, match="vcv must be positive definite"):
        rmvnorm_generator(np.array([[1.0, 2.0], [2.0, 1.0]]))
    n = 10
    vcv = np.cov(np.random.default_rng(1).normal(size=(n, n)), rowvar=False)
    i = [1, 4]
    vcv[i, :] = 0
    vcv[:, i] = 0
    f = rmvnorm_generator(vcv)
    gen = np.random.default_rng(2)
    res = np.array([f(np.zeros(n), gen) for _ in range(100)])
    assert np.all(res[:, i] == 0) and np.all(res[:, [0, 2, 3]] != 0)
    vcv = np.array([[4.0, 2.0], [2.0, 3.0]])
    theta = np.random.default_rng(3).uniform(size=2)
    y1 = rmvnorm_generator(vcv)(theta, np.random.default_rng(1))
    y2 = rmvnorm_generator(vcv)(theta, np.random.default_rng(1), 2)
    y3 = rmvnorm_generator(vcv * 2)(theta, np.random.default_rng(1))
    assert np.allclose((y1 - theta) * math.sqrt(2), y2 - theta, rtol=1e-12, atol=0)
    assert np.allclose(y2, y3, rtol=1e-12, atol=1e-15)


def test_pmcmc_combine(sir_data):
    """tests/testthat/test-pmcmc-tools.R:71-272."""
    from mcstate_b200.pmcmc import pmcmc_combine

    control = pmcmc_control(10, save_state=True, save_trajectories=True, save_restart=[40],
                            use_parallel_seed=True)
    results = [pmcmc(example_pars(), example_filter("oracle", sir_data, 10, seed=k), control=control)
               for k in (1, 2, 3)]
    res = pmcmc_combine(*results)
    assert res["pars"].shape == (30, 2) and res["probabilities"].shape == (30, 3)
    assert res["state"].shape == (5, 30) and res["trajectories"].shape == (3, 30, 101)
    assert res["restart"].shape == (5, 30, 1)
    assert list(res["chain"]) == [1] * 10 + [2] * 10 + [3] * 10
    i = slice(10, 20)
    assert np.array_equal(res["pars"][i], results[1]["pars"])
    assert np.array_equal(res["state"][:, i], results[1]["state"])
    assert np.array_equal(res["trajectories"][:, i], results[1]["trajectories"])
    assert np.array_equal(res["restart"][:, i], results[1]["restart"])
    for other in (pmcmc_combine(samples=results), pmcmc_combine(results),
                  pmcmc_combine(samples={"a": results[0], "b": results[1], "c": results[2]})):
        assert all(np.array_equal(res[k], other[k]) for k in ("pars", "chain", "state", "iteration"))
    bare = [dict(r, state=None, trajectories=None, predict=None) for r in results]
    comb = pmcmc_combine(samples=bare)
    assert comb["state"] is None and comb["trajectories"] is None and comb["predict"] is None
    assert np.array_equal(comb["pars"], res["pars"])
    for bad, msg in (((), "At least 2 samples objects must be provided"),
                     ((results[0],), "At least 2 samples objects must be provided")):
        with pytest.raises(ValueError, match=msg):
            pmcmc_combine(*bad)
    with pytest.raises(TypeError, match="'samples' must be a list"):
        pmcmc_combine(samples=3)
    with pytest.raises(ValueError, match="Chains have already been combined"):
        pmcmc_combine(results[0], pmcmc_combine(results[1], results[2]))
    with pytest.raises(ValueError, match="All chains must have the same length"):
        pmcmc_combine(results[0], pmcmc_thin(results[1], burnin=2))
    wide = dict(results[0], pars=np.hstack([results[0]["pars"], np.ones((10, 1))]),
                pars_names=["beta", "gamma", "zeta"])
    with pytest.raises(ValueError, match="All parameters must have the same dimension names"):
        pmcmc_combine(wide, results[1])
    with pytest.raises(ValueError, match="All chains must have the same iterations"):
        pmcmc_combine(dict(results[0], iteration=results[0]["iteration"] + 1), results[1])
    for key in ("state", "trajectories", "restart"):
        with pytest.raises(ValueError, match="If '%s' is present for any samples, it must be "
                                             "present for all" % key):
            pmcmc_combine(dict(results[0], **{key: None}), results[1])
    with pytest.raises(ValueError, match="trajectories data is inconsistent"):
        pmcmc_combine(dict(results[0], trajectories_time=results[0]["trajectories_time"] + 1),
                      results[1])
    for bad in ((results[0], None),):
        with pytest.raises(TypeError, match="Elements of"):
            pmcmc_combine(*bad)
    with pytest.raises(TypeError, match="Elements of"):
        pmcmc_combine(samples=[results[0], None])


def test_filter_on_creation_after_combining(sir_data):
    """tests/testthat/test-pmcmc.R:648-690: the retained samples of a 2-chain run equal the
    full run filtered at those iterations; the full parameter chain stays available until
    the samples are filtered."""
    from mcstate_b200.pmcmc import pmcmc_filter

    def run(**kw):
        control = pmcmc_control(30, n_chains=2, save_state=True, save_trajectories=True,
                                use_parallel_seed=True, **kw)
        return pmcmc(example_pars(), example_filter("oracle", sir_data, 10), control=control), control

    r1, c1 = run()
    r2, c2 = run(n_burnin=3, n_steps_retain=7)
    assert r2["pars"].shape == (14, 2)
    assert np.array_equal(r1["pars_full"], r1["pars"])
    assert np.array_equal(r1["probabilities_full"], r1["probabilities"])
    assert np.array_equal(r2["pars_full"], r1["pars"])
    assert np.array_equal(r2["probabilities_full"], r1["probabilities"])
    assert list(r1["iteration"]) == list(range(1, 31)) * 2
    i = np.array([c2.n_burnin + 1 + k * c2.n_steps_every for k in range(c2.n_steps_retain)])
    assert list(r2["iteration"]) == list(i) * 2
    cmp = pmcmc_filter(r1, np.concatenate([i, i + c1.n_steps]) - 1)
    for key in ("pars", "probabilities", "state", "trajectories", "chain", "iteration"):
        assert np.array_equal(r2[key], cmp[key]), key
    assert cmp["pars_full"] is None
    r3 = pmcmc_thin(r2)
    assert np.array_equal(r3["pars"], r2["pars"]) and r3["pars_full"] is None


def test_generator_unpickles_without_the_engine(monkeypatch):
    """Results that carry a generator (``predict``) can be read where the engine library is
    not available; the library is bound when something first needs it."""
    import pickle

    import cloudpickle

    from mcstate_b200 import _lib, dust

    g = dust.dust_example("sir")
    blob = cloudpickle.dumps(g)

    def missing():
        raise RuntimeError("libmcstate_b200.so is missing")

    monkeypatch.setattr(_lib, "lib", missing)
    h = pickle.loads(blob)
    assert h._L is None and h.par_names == g.par_names and h.has_compare()
    with pytest.raises(RuntimeError, match="is missing"):
        h.gpu_info()
    monkeypatch.undo()
    assert isinstance(h.gpu_info(), dict) and h._L is g._L


def test_pmcmc_result_layout(sir_data):
    """tests/testthat/test-pmcmc.R:88-166: what a chain returns (R's nested list
    ``trajectories$state/time/rate`` is flat here: ``trajectories``, ``trajectories_time``,
    ``predict["rate"]``)."""
    f = example_filter("oracle", sir_data, 100)
    res = pmcmc(example_pars(), f, control=pmcmc_control(30, save_state=True, save_trajectories=True,
                                                        use_parallel_seed=True))
    assert res["chain"] is None and list(res["iteration"]) == list(range(1, 31))
    assert res["pars"].shape == (30, 2) and res["pars_names"] == ["beta", "gamma"]
    assert res["probabilities"].shape == (30, 3)       # log prior, likelihood, posterior
    assert np.any(np.diff(res["pars"], axis=0) != 0)   # it mixed
    assert res["state"].shape == (5, 30) and res["trajectories"].shape == (3, 30, 101)
    assert np.array_equal(res["trajectories"][:, :, -1], res["state"][:3])
    assert np.array_equal(res["trajectories_time"], np.arange(0, 401, 4))
    pred = res["predict"]
    assert set(pred) == {"is_continuous", "transform", "index", "rate", "model_time", "time",
                         "filter"}
    assert pred["index"] == [1, 2, 3] and pred["rate"] == 4
    assert pred["time"] == 400 and pred["model_time"] == 100 and not pred["is_continuous"]
    assert pred["transform"]([0.2, 0.1]) == {"beta": 0.2, "gamma": 0.1}
    assert pred["filter"]["n_particles"] == 100 and len(pred["filter"]["seed"]) == 32


def test_default_run_is_reproducible_with_set_seed(sir_data):
    """The reference's default pmcmc draws proposals from R's global RNG, so set.seed()
    makes a run reproducible (R/pmcmc.R:61-90); here mcstate_b200.set_seed() does."""
    import mcstate_b200

    def once():
        mcstate_b200.set_seed(42)
        f = example_filter("oracle", sir_data, n_particles=32, seed=1)
        return pmcmc(example_pars(), f, control=pmcmc_control(15, save_state=False))["pars"]

    a, b = once(), once()
    assert np.array_equal(a, b)
    mcstate_b200.set_seed(43)
    f = example_filter("oracle", sir_data, n_particles=32, seed=1)
    c = pmcmc(example_pars(), f, control=pmcmc_control(15, save_state=False))["pars"]
    assert not np.array_equal(a, c)
