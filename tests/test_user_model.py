"""User-supplied compiled models: ``mcstate_b200.dust.dust(path)`` -- the engine's
``dust::dust(file)`` (tests/testthat/test-pmcmc-parallel.R:197; a model with its own
compiled compare as in tests/testthat/helper-mcstate.R:478-492 + compare_variable.cpp).

The example is the SEIR model (mcstate_b200/examples/seir.cuh); its CPU twin is
oracle/mcstate_oracle.c: seir_update.  CPU tests: the model compiles into its own
library that exports the whole C ABI and describes itself; the oracle twin behaves.
GPU tests: engine == oracle bit for bit, on both execution paths of the filter, both
scramblers, batched parameter sets, and through particle_filter / pmcmc.
"""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SEIR = os.path.join(ROOT, "mcstate_b200", "examples", "seir.cuh")
SEIR_ID = 4  # oracle.SEIR


def _seir_data(sir_data, n=None, rate=4):
    return sir_data["day"][:n] * rate, sir_data["incidence"][:n]


# ---------------------------------------------------------------- CPU (no device needed)
def test_user_model_library_builds_and_exports_the_abi():
    from mcstate_b200 import _lib

    path = _lib.build_user_model(SEIR)
    assert os.path.dirname(path) == _lib.USER_MODEL_DIR and os.path.exists(path)
    assert _lib.build_user_model(SEIR) == path          # cached: same digest, same file
    L = C.CDLL(path)
    for name in _lib.SIGNATURES:
        assert hasattr(L, name), name
    L = _lib.load(path)
    mid, ns, npar, nd = C.c_int(-1), C.c_size_t(), C.c_size_t(), C.c_size_t()
    assert L.mcb_model_lookup(b"seir", C.byref(mid), C.byref(ns), C.byref(npar), C.byref(nd)) == 0
    assert (mid.value, ns.value, npar.value, nd.value) == (0, 6, 9, 1)
    assert [L.mcb_model_par_name(0, i).decode() for i in range(8)] == \
        ["beta", "sigma", "gamma", "I0", "exp_noise", "S0", "E0", "freq"]
    assert [L.mcb_model_state_name(0, i).decode() for i in range(6)] == \
        ["S", "E", "I", "R", "cases_cumul", "cases_inc"]
    assert L.mcb_model_data_name(0, 0) == b"incidence"
    # a user build holds its model only; the stock library does not hold it
    assert L.mcb_model_lookup(b"sir", None, None, None, None) != 0
    assert _lib.lib().mcb_model_lookup(b"seir", None, None, None, None) != 0


def test_user_model_generator_describes_itself():
    from mcstate_b200 import dust

    g = dust.dust(SEIR)
    assert g.name == "dust_generator" and g.model_name == "seir"
    assert g.has_compare() and g.time_type() == "discrete"
    assert g.state_names == ["S", "E", "I", "R", "cases_cumul", "cases_inc"]
    assert g.par_defaults == [0.3, 0.25, 0.1, 10.0, 1e6, 1000.0, 0.0, 4.0, 0.0]
    assert g.par_names[-1] == "import"   # Poisson arrivals (samplers.cuh: draw_poisson)
    assert g._info({"beta": 0.4})["pars"]["beta"] == 0.4


def test_user_model_errors(tmp_path):
    from mcstate_b200 import dust

    with pytest.raises(FileNotFoundError):
        dust.dust(str(tmp_path / "missing.cuh"))
    bad = tmp_path / "broken.cuh"
    bad.write_text("namespace mcb { struct UserModel { this does not compile }; }\n")
    with pytest.raises(RuntimeError, match="compiling .* failed"):
        dust.dust(str(bad))
    # compiles, but the descriptor names another model than the file
    other = tmp_path / "renamed.cuh"
    other.write_text(open(SEIR).read())
    with pytest.raises(ValueError, match="must name the model after its file"):
        dust.dust(str(other))
    from mcstate_b200 import _lib

    os.remove(_lib.user_model_path(str(other)))   # keep the in-tree build directory tidy


def test_oracle_seir_twin(orc, sir_data):
    n = 500
    o = orc.OracleModel(n_particles=n, seed=3, model=SEIR_ID, n_threads=4)
    y0 = o.state()
    assert np.array_equal(y0[:, 0], [1000.0, 0.0, 10.0, 0.0, 0.0, 0.0])
    y = o.run(200)                                       # [n_state, n]
    assert np.all(y[:4].sum(axis=0) == 1010.0)           # closed population
    assert np.all(y[0] <= 1000.0) and np.all(y[4] >= y[5])
    assert np.all(y[4] == 1000.0 - y[0] - y[1])          # cumulative E->I = who left S and E
    # deterministic mode consumes no draws
    d = orc.OracleModel(n_particles=2, seed=3, model=SEIR_ID, deterministic=True)
    d.run(40)
    assert np.array_equal(d.rng_state(), orc.rng_streams(3, 3))
    t, x = _seir_data(sir_data)
    o = orc.OracleModel(n_particles=n, seed=3, model=SEIR_ID, n_threads=4)
    o.set_data(t, x)
    ll = o.filter()["log_likelihood"]
    assert np.isfinite(ll).all()


# ---------------------------------------------------------------- GPU: engine == oracle
@pytest.fixture(scope="module")
def seir():
    from mcstate_b200 import dust

    g = dust.dust(SEIR)
    if not g.gpu_info()["has_cuda"]:
        pytest.fail("no CUDA device visible: the engine has no CPU fallback")
    return g


def _words(m):
    return np.frombuffer(m.rng_state(), dtype=np.uint64).reshape(-1, 4)


@pytest.mark.gpu
@pytest.mark.parametrize("alg,scr", [("xoshiro256plus", 0), ("xoshiro256starstar", 1)])
def test_user_model_run_bit_exact(orc, alg, scr):
    from mcstate_b200 import dust

    g = dust.dust(SEIR, algorithm=alg)
    n = 5000
    for pars in ({}, {"beta": 0.5, "sigma": 0.4, "I0": 30.0, "E0": 12.0},
                 {"S0": 1e6, "I0": 5000.0, "beta": 0.6}):           # the last one reaches BTRS
        m = g.new(pars, 0, n, seed=11)
        o = orc.OracleModel(pars=g._pars_vector(pars, False), n_particles=n, seed=11,
                            model=SEIR_ID, scrambler=scr, n_threads=8)
        assert np.array_equal(m.state(), o.state())
        for t_end in (1, 4, 7, 200):
            assert np.array_equal(m.run(t_end), o.run(t_end)), (pars, t_end)
            assert np.array_equal(_words(m), o.rng_state())


@pytest.mark.gpu
@pytest.mark.parametrize("alg,scr", [("xoshiro256plus", 0), ("xoshiro256starstar", 1)])
def test_poisson_draws_bit_exact(orc, alg, scr):
    """The device Poisson sampler (samplers.cuh: draw_poisson -- Knuth's product below 10,
    PTRS from 10 up, log(k!) from det_lfact) against orc_poisson over a sweep of lambda,
    through the SEIR model's importation term: with an empty population every binomial
    returns 0 without a draw, so after one step E holds Poisson(import) from each
    particle's own stream (BASELINE north_star item 1)."""
    from mcstate_b200 import dust

    g = dust.dust(SEIR, algorithm=alg)
    n = 20000
    L = orc.lib()
    for lam in (0.0, 1e-3, 0.5, 3.0, 9.999, 10.0, 10.5, 37.5, 1e3, 1e6):
        pars = {"S0": 0.0, "I0": 0.0, "E0": 0.0, "freq": 1.0, "import": lam}
        m = g.new(pars, 0, n, seed=13)
        y = m.run(1)
        streams = orc.rng_streams(13, n + 1)[:n].copy()
        want = np.array([L.orc_poisson(streams[i].ctypes.data_as(orc._u64p), scr, lam)
                         for i in range(n)])
        assert np.array_equal(y[1], want), lam
        assert np.array_equal(_words(m)[:n], streams), lam
        if lam >= 0.5:
            assert abs(want.mean() - lam) < 5 * np.sqrt(lam / n)
    # the importation term inside a running epidemic, filter and all, against the twin
    t, x = np.arange(1, 31, dtype=np.uint64) * 4, np.arange(30) % 7 + 1.0
    for pars in ({"import": 2.0}, {"import": 60.0, "beta": 0.1}):
        m = g.new(pars, 0, 3000, seed=2)
        m.set_data(dust.dust_data({"time_end": t, "incidence": x}), False)
        o = orc.OracleModel(pars=g._pars_vector(pars, False), n_particles=3000, seed=2,
                            model=SEIR_ID, scrambler=scr, n_threads=8)
        o.set_data(t, x)
        assert np.array_equal(m.filter()["log_likelihood"], o.filter()["log_likelihood"])
        assert np.array_equal(m.state(), o.state())
        assert np.array_equal(_words(m), o.rng_state())


@pytest.mark.gpu
@pytest.mark.parametrize("single_launch", [False, True])
@pytest.mark.parametrize("n", [100, 4096, 70001])
def test_user_model_filter_bit_exact(seir, orc, sir_data, n, single_launch):
    from mcstate_b200 import dust

    t, x = _seir_data(sir_data)
    pars = {"beta": 0.45, "sigma": 0.5, "gamma": 0.12}
    m = seir.new(pars, 0, n, seed=5)
    m.set_persistent(single_launch)
    m.set_data(dust.dust_data({"time_end": t, "incidence": x}), False)
    o = orc.OracleModel(pars=seir._pars_vector(pars, False), n_particles=n, seed=5,
                        model=SEIR_ID, n_threads=8)
    o.set_data(t, x)
    got = m.filter(save_trajectories=True)
    want = o.filter(save_trajectories=True)
    assert np.array_equal(got["log_likelihood"], want["log_likelihood"])
    assert np.array_equal(m.state(), o.state())
    assert np.array_equal(_words(m), o.rng_state())
    assert np.array_equal(got["trajectories"], want["trajectories"])


@pytest.mark.gpu
def test_user_model_parameter_sets_and_deterministic(seir, orc, sir_data):
    from mcstate_b200 import dust

    t, x = _seir_data(sir_data, 40)
    sets = [{"beta": b, "sigma": s} for b, s in ((0.3, 0.25), (0.5, 0.5), (0.8, 0.2))]
    n = 3000
    m = seir.new(sets, 0, n, seed=2, pars_multi=True)
    m.set_data(dust.dust_data({"time_end": t, "incidence": x}), True)
    o = orc.OracleModel(pars=seir._pars_vector(sets, True), n_particles=n, n_pars=3, seed=2,
                        model=SEIR_ID, n_threads=8)
    o.set_data(t, x, shared=True)
    assert np.array_equal(m.filter()["log_likelihood"], o.filter()["log_likelihood"])
    assert np.array_equal(_words(m), o.rng_state())
    d = seir.new({}, 0, 4, seed=1, deterministic=True)
    od = orc.OracleModel(n_particles=4, seed=1, model=SEIR_ID, deterministic=True)
    assert np.array_equal(d.run(100), od.run(100))


@pytest.mark.gpu
def test_user_model_through_particle_filter_and_pmcmc(seir, sir_data):
    """The generator drops into the host mirror unchanged: particle_filter$run, history,
    and a short pmcmc -- against the same calls on the oracle-backed generator."""
    from oracle_generator import OracleGenerator

    from mcstate_b200.particle_filter import particle_filter
    from mcstate_b200.particle_filter_data import particle_filter_data
    from mcstate_b200.pmcmc import pmcmc, pmcmc_control, pmcmc_parameter, pmcmc_parameters

    data = particle_filter_data({"day": sir_data["day"], "incidence": sir_data["incidence"]},
                                "day", 4, 0)
    res = []
    for gen in (seir, OracleGenerator("seir")):
        f = particle_filter(data, gen, 2000, None, seed=9,
                            index=lambda info: {"run": [info["index"]["cases_inc"]],
                                                "state": [1, 2, 3, 4]})
        ll = [f.run({"beta": 0.4}, save_history=True)]
        h = f.history()
        ll.append(f.run({"beta": 0.35, "sigma": 0.3}))
        pars = pmcmc_parameters([pmcmc_parameter("beta", 0.4, min=0.05, max=2.0),
                                 pmcmc_parameter("sigma", 0.3, min=0.05, max=2.0)],
                                np.diag([0.002, 0.002]))
        # use_parallel_seed: the chain's filter and host RNG are seeded from the filter's seed
        out = pmcmc(pars, f, control=pmcmc_control(15, save_state=True, progress=False,
                                                   use_parallel_seed=True))
        res.append((ll, h, out["pars"], out["probabilities"], out["state"]))
    a, b = res
    assert a[0] == b[0]
    assert np.array_equal(a[1], b[1])
    for k in (2, 3, 4):
        assert np.array_equal(a[k], b[k]), k


@pytest.mark.gpu
def test_user_model_chains_in_other_processes(seir, sir_data, tmp_path):
    """tests/testthat/test-pmcmc-parallel.R:193-210 ("Can use workers on non-package
    models": the worker must be able to load the compiled model): chains of a
    user-supplied model prepared here and run in fresh interpreters, which bind the
    model's own library again from the serialised generator."""
    import subprocess
    import sys

    from mcstate_b200.particle_filter import particle_filter
    from mcstate_b200.particle_filter_data import particle_filter_data
    from mcstate_b200.pmcmc import (pmcmc, pmcmc_chains_collect, pmcmc_chains_prepare,
                                    pmcmc_control, pmcmc_parameter, pmcmc_parameters)

    data = particle_filter_data({"day": sir_data["day"], "incidence": sir_data["incidence"]},
                                "day", 4, 0)

    def make():
        return particle_filter(data, seir, 500, None, seed=4, gpu_config=0,
                               index=lambda info: {"run": [info["index"]["cases_inc"]],
                                                   "state": [1, 2, 3, 4]})

    pars = pmcmc_parameters([pmcmc_parameter("beta", 0.4, min=0.05, max=2.0),
                             pmcmc_parameter("sigma", 0.3, min=0.05, max=2.0)],
                            np.diag([0.002, 0.002]))
    control = pmcmc_control(8, n_chains=2, save_state=True, use_parallel_seed=True)
    want = pmcmc(pars, make(), control=control)
    path = pmcmc_chains_prepare(str(tmp_path / "chains"), pars, make(), control)
    for k in (1, 2):
        code = ("import sys; sys.path.insert(0, %r); "
                "from mcstate_b200.pmcmc import pmcmc_chains_run; "
                "print(pmcmc_chains_run(%d, %r))" % (ROOT, k, path))
        out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True,
                             timeout=300)
        assert out.returncode == 0, out.stderr[-2000:]
    got = pmcmc_chains_collect(path)
    for key in ("pars", "probabilities", "state", "chain", "iteration"):
        assert np.array_equal(got[key], want[key]), key
