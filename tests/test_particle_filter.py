"""The filter facade (mcstate_b200.particle_filter) against the relations the
reference's test-suite asserts (file:line cited per test).

Every test runs twice: on the CPU-oracle generator (host logic; runs without a GPU)
and on the B200 engine (`-m gpu`).  A final group checks the two back ends against
each other, as the reference does for its CPU/GPU pair
(tests/testthat/test-particle-filter.R:1480-1504).
"""
import math

import numpy as np
import pytest

from mcstate_b200.particle_filter import (particle_filter, particle_filter_from_inputs,
                                          particle_resample, scale_log_weights)
from mcstate_b200.particle_filter_data import particle_filter_data
from oracle_generator import OracleGenerator


def _gpu_generator(name="sir"):
    from mcstate_b200 import dust

    g = dust.dust_example(name)
    if not g.gpu_info()["has_cuda"]:
        pytest.fail("no CUDA device visible: the engine has no CPU fallback")
    return g


BACKENDS = [pytest.param("oracle", id="oracle"),
            pytest.param("gpu", id="b200", marks=pytest.mark.gpu)]


@pytest.fixture(params=BACKENDS)
def gen(request):
    return OracleGenerator("sir") if request.param == "oracle" else _gpu_generator("sir")


def example_sir(sir_data, n_days=None, initial_time=0, offset=0):
    """tests/testthat/helper-mcstate.R:1-61 with the shipped incidence series."""
    day = sir_data["day"][:n_days] + offset
    inc = sir_data["incidence"][:n_days]
    data = particle_filter_data({"day": day, "incidence": inc}, "day", 4, initial_time + offset)

    def compare(state, observed, pars=None):
        if observed["incidence"] is None or np.isnan(observed["incidence"]):
            return None
        exp_noise = 1e6
        if pars and "compare" in pars and "exp_noise" in pars["compare"]:
            exp_noise = pars["compare"]["exp_noise"]
        modelled = state[0, :]
        lam = modelled + np.random.exponential(1.0 / exp_noise, size=modelled.size)
        x = observed["incidence"]
        with np.errstate(divide="ignore", invalid="ignore"):
            return x * np.log(lam) - lam - math.lgamma(x + 1)

    def index(info):
        return {"run": [5], "state": [1, 2, 3]}

    return {"data": data, "compare": compare, "index": index, "day": day, "incidence": inc}


# ------------------------------------------------------------- pure helpers (no model)
def test_scale_log_weights_known_answers():
    """tests/testthat/test-particle-filter.R:174-187."""
    inf, nan, e = math.inf, math.nan, math.e
    r = scale_log_weights([-inf, -inf])
    assert np.isnan(r["weights"]).all() and r["average"] == -inf
    r = scale_log_weights([-inf, 1])
    assert list(r["weights"]) == [0, 1] and r["average"] == pytest.approx(math.log(e / 2))
    r = scale_log_weights([-inf, 1, 1])
    assert list(r["weights"]) == [0, 1, 1] and r["average"] == pytest.approx(math.log(e * 2 / 3))
    r = scale_log_weights([nan, nan])
    assert np.isnan(r["weights"]).all() and r["average"] == -inf
    r = scale_log_weights([nan, 1])
    assert list(r["weights"]) == [0, 1] and r["average"] == pytest.approx(math.log(e / 2))


def test_particle_resample_is_systematic():
    np.random.seed(1)
    w = np.random.uniform(size=50)
    k = particle_resample(w)
    assert k.min() >= 1 and k.max() <= 50 and np.all(np.diff(k) >= 0)
    counts = np.bincount(k - 1, minlength=50)
    expect = 50 * w / w.sum()
    assert np.all(np.abs(counts - expect) < 1)
    km = particle_resample(np.random.uniform(size=(20, 3)))
    assert km.shape == (20, 3)


def test_constructor_validation(sir_data):
    """tests/testthat/test-particle-filter.R:140-171 (messages are the reference's)."""
    dat = example_sir(sir_data)
    g = OracleGenerator()
    with pytest.raises(TypeError, match="'model' must be a dust_generator"):
        particle_filter(dat["data"], None, 10, dat["compare"])
    with pytest.raises(TypeError, match="'index' must be function if not NULL"):
        particle_filter(dat["data"], g, 10, dat["compare"], index=[1, 3, 5])
    with pytest.raises(TypeError, match="'initial' must be function if not NULL"):
        particle_filter(dat["data"], g, 10, dat["compare"], initial=[1, 3, 5])
    with pytest.raises(ValueError, match="'n_particles' must be at least 1"):
        particle_filter(dat["data"], g, 0, dat["compare"])
    with pytest.raises(ValueError, match="'n_threads' must be at least 1"):
        particle_filter(dat["data"], g, 1, dat["compare"], n_threads=-1)
    with pytest.raises(TypeError, match="'data' must be a particle_filter_data"):
        particle_filter({"day": [1]}, g, 1, dat["compare"])
    with pytest.raises(ValueError, match="'gpu_config' provided, but 'model' does not have GPU"):
        particle_filter(dat["data"], g, 1, None, gpu_config=0)
    p = particle_filter(dat["data"], g, 10, None)
    with pytest.raises(RuntimeError, match="Model has not yet been run"):
        p.state()
    with pytest.raises(RuntimeError, match="Model has not yet been run"):
        p.history()


@pytest.mark.parametrize("kind", BACKENDS)
def test_model_without_a_compiled_compare(kind, sir_data):
    """tests/testthat/test-particle-filter.R:579-586: the walk model has no built-in
    compare, so a filter without a compare function refuses it; with one, it runs through
    the R loop (model$run / $state / $reorder) like any other model."""
    dat = example_sir(sir_data, n_days=30)
    g = OracleGenerator("walk") if kind == "oracle" else _gpu_generator("walk")
    assert not g.public_methods.has_compare()
    with pytest.raises(ValueError, match="Your model does not have a built-in 'compare' function"):
        particle_filter(dat["data"], g, 100, None)

    def compare(state, observed, pars=None):
        # the walk's position against the (scaled) incidence series: any smooth density will do
        return -0.5 * (state[0, :] - observed["incidence"] / 10.0) ** 2

    res = []
    for _ in range(2):
        np.random.seed(1)       # the R loop resamples with the host RNG
        p = particle_filter(dat["data"], g, 200, compare, seed=3,
                            gpu_config=0 if kind == "gpu" else None)
        res.append((p.run({"sd": 0.5}, save_history=True), p.history(), p.state()))
    assert np.isfinite(res[0][0]) and res[0][0] == res[1][0]
    assert np.array_equal(res[0][1], res[1][1]) and res[0][1].shape == (1, 200, 31)
    assert np.array_equal(res[0][2], res[1][2])


@pytest.mark.gpu
def test_walk_model_engine_equals_oracle(sir_data):
    from mcstate_b200 import dust

    m = dust.dust_example("walk", "xoshiro256starstar").new({"sd": 2.5}, 0, 3000, seed=4)
    from oracle import oracle as O

    o = O.OracleModel(pars=np.array([[2.5]]), n_particles=3000, seed=4, model=O.WALK,
                      scrambler=O.STARSTAR)
    for t_end in (1, 7, 40):
        assert np.array_equal(m.run(t_end), o.run(t_end))
    assert np.array_equal(np.frombuffer(m.rng_state(), dtype=np.uint64).reshape(-1, 4),
                          o.rng_state())
    with pytest.raises(Exception, match="does not have a built-in 'compare'"):
        m.set_data(dust.dust_data({"time_end": [4, 8], "x": [0.0, 1.0]}), False)


# ------------------------------------------------------------------ relations
def test_same_seed_same_result_and_rng_continues(gen, sir_data):
    """tests/testthat/test-particle-filter.R:22-34."""
    dat = example_sir(sir_data)
    p1 = particle_filter(dat["data"], gen, 42, None, index=dat["index"], seed=1)
    p2 = particle_filter(dat["data"], gen, 42, None, index=dat["index"], seed=1)
    a = p1.run(save_history=True)
    b = p2.run(save_history=True)
    assert a == b and np.isfinite(a)
    assert np.array_equal(p1.history(), p2.history())
    assert np.array_equal(p1.state(), p2.state())
    a2 = p1.run()
    assert a2 != a  # the model object persists, so its RNG streams continue
    assert p2.run() == a2


def test_index_does_not_change_likelihood(gen, sir_data):
    """tests/testthat/test-particle-filter.R:37-57."""
    dat = example_sir(sir_data)
    a = particle_filter(dat["data"], gen, 42, None, index=dat["index"], seed=1).run()
    b = particle_filter(dat["data"], gen, 42, None, seed=1).run()
    assert a == b


def test_shifted_start_date_gives_identical_likelihood(gen, sir_data):
    """tests/testthat/test-particle-filter.R:111-139."""
    d0 = example_sir(sir_data, n_days=40)
    d1 = example_sir(sir_data, n_days=40, offset=100)
    a = particle_filter(d0["data"], gen, 64, None, seed=1).run()
    b = particle_filter(d1["data"], gen, 64, None, seed=1).run()
    assert a == b


def test_worse_parameters_give_worse_likelihood(gen, sir_data):
    """tests/testthat/test-particle-filter.R:60-68."""
    dat = example_sir(sir_data)
    p = particle_filter(dat["data"], gen, 100, None, index=dat["index"], seed=1)
    assert p.run() > p.run({"gamma": 1.0, "beta": 1.0})


def test_incremental_steps_equal_one_shot(gen, sir_data):
    """tests/testthat/test-particle-filter.R:604-711."""
    dat = example_sir(sir_data, n_days=40)
    p1 = particle_filter(dat["data"], gen, 42, None, index=dat["index"], seed=1)
    whole = p1.run()
    p2 = particle_filter(dat["data"], gen, 42, None, index=dat["index"], seed=1)
    obj = p2.run_begin()
    assert obj.current_time_index == 0
    obj.step(10)
    assert obj.current_time_index == 10
    obj.step(11)
    got = obj.step(40)
    assert got == pytest.approx(whole, rel=1e-12)
    assert np.array_equal(obj.model.state(), p1.state())
    with pytest.raises(ValueError, match="Particle filter has reached the end of the data"):
        obj.step(41)
    obj = p2.run_begin()
    with pytest.raises(ValueError, match="time_index 41 is beyond the length of the data"):
        obj.step(41)
    obj.step(10)
    with pytest.raises(ValueError, match=r"already run step index 5 \(to model step 20\)"):
        obj.step(5)
    with pytest.raises(ValueError, match="'partial' not supported with compiled compare"):
        obj.step(12, partial=True)


def test_incremental_steps_r_compare(gen, sir_data):
    """tests/testthat/test-particle-filter.R:604-632 (R compare path)."""
    dat = example_sir(sir_data, n_days=30)
    np.random.seed(1)
    whole = particle_filter(dat["data"], gen, 42, dat["compare"], index=dat["index"], seed=1).run()
    np.random.seed(1)
    obj = particle_filter(dat["data"], gen, 42, dat["compare"], index=dat["index"],
                          seed=1).run_begin()
    parts = [obj.step(10, partial=True), obj.step(30)]
    assert parts[1] == pytest.approx(whole, rel=1e-12)


def test_fork_smc2_equals_fresh_filter_on_parent_rng(gen, sir_data):
    """tests/testthat/test-particle-filter.R:714-743."""
    dat = example_sir(sir_data, n_days=30)
    pars = {"beta": 0.25, "gamma": 0.12}
    p = particle_filter(dat["data"], gen, 42, None, index=dat["index"], seed=1)
    obj = p.run_begin()
    obj.step(10)
    rng_before = obj.model.rng_state()
    child = obj.fork_smc2(pars)
    assert child.current_time_index == 10
    # a fresh filter seeded with the parent's pre-fork RNG state is the same thing
    fresh = particle_filter(dat["data"], gen, 42, None, index=dat["index"],
                            seed=rng_before).run_begin(pars)
    fresh.step(10)
    assert child.log_likelihood == fresh.log_likelihood
    assert np.array_equal(child.model.state(), fresh.model.state())
    # and the parent now continues from the child's RNG
    assert obj.model.rng_state() == child.model.rng_state() != rng_before


def test_constant_log_likelihood_is_additive(gen, sir_data):
    """tests/testthat/test-particle-filter.R:1356-1377."""
    dat = example_sir(sir_data, n_days=30)
    base = particle_filter(dat["data"], gen, 42, None, seed=1).run()
    p = particle_filter(dat["data"], gen, 42, None, seed=1,
                        constant_log_likelihood=lambda pars: -12.5)
    assert p.run() == base - 12.5


def test_restart_snapshot_equals_truncated_run(gen, sir_data):
    """tests/testthat/test-particle-filter.R:417-477."""
    dat = example_sir(sir_data, n_days=40)
    p = particle_filter(dat["data"], gen, 42, None, index=dat["index"], seed=1)
    p.run(save_restart=[10, 25])
    rs = p.restart_state()
    assert rs.shape == (5, 42, 2)
    for k, d in enumerate((10, 25)):
        short = example_sir(sir_data, n_days=d)
        q = particle_filter(short["data"], gen, 42, None, index=dat["index"], seed=1)
        q.run()
        assert np.array_equal(q.state(), rs[:, :, k])
    assert np.array_equal(p.restart_state([3, 1])[:, 0, :], rs[:, 2, :])
    with pytest.raises(ValueError, match="'save_restart' contains times not in 'day': 1000"):
        p.run(save_restart=[10, 1000])
    p2 = particle_filter(dat["data"], gen, 42, None, seed=1)
    p2.run()
    with pytest.raises(RuntimeError, match="save_restart = NULL"):
        p2.restart_state()


def test_history_shapes_and_lineages(gen, sir_data):
    """tests/testthat/test-particle-filter.R:270-293, 536-554: S never rises, R never
    falls along an ancestor-traced lineage."""
    dat = example_sir(sir_data)
    for compare in (None, dat["compare"]):
        np.random.seed(1)
        p = particle_filter(dat["data"], gen, 42, compare, index=dat["index"], seed=1)
        p.run(save_history=True)
        h = p.history()
        assert h.shape == (3, 42, 101)
        assert np.all(np.diff(h[0], axis=1) <= 0) and np.all(np.diff(h[2], axis=1) >= 0)
        assert np.array_equal(p.history([5, 9]), h[:, [4, 8], :])
        assert np.array_equal(h[:, :, 0], np.tile([[1000.0], [10.0], [0.0]], (1, 42)))
    p = particle_filter(dat["data"], gen, 42, None, seed=1)
    p.run()
    with pytest.raises(RuntimeError, match="save_history = FALSE"):
        p.history()


def test_initial_function_sets_state(gen, sir_data):
    """tests/testthat/test-particle-filter.R:215-250."""
    dat = example_sir(sir_data)

    def initial(info, n_particles, pars):
        return [1000, pars["I0"], 0, 0, 0]

    p = particle_filter(dat["data"], gen, 50, None, index=dat["index"], initial=initial, seed=1)
    p.run({"I0": 200}, save_history=True)
    assert np.array_equal(p.history()[:, :, 0], np.tile([[1000.0], [200.0], [0.0]], (1, 50)))

    def initial_matrix(info, n_particles, pars):
        return np.tile(np.array([[1000, 5, 0, 0, 0.0]]).T, (1, n_particles)) + \
            np.arange(n_particles) * np.array([[0, 1, 0, 0, 0.0]]).T

    p = particle_filter(dat["data"], gen, 50, None, index=dat["index"], initial=initial_matrix,
                        seed=1)
    # (a traced history shows ancestors, which coalesce: look at the state itself)
    obj = p.run_begin(save_history=True)
    assert np.array_equal(obj.model.state()[1], 5.0 + np.arange(50))
    assert np.array_equal(obj.history["value"][1, :, 0], 5.0 + np.arange(50))


def test_null_compare_means_no_resampling(gen, sir_data):
    """tests/testthat/test-particle-filter.R:208-215."""
    dat = example_sir(sir_data)
    p = particle_filter(dat["data"], gen, 42, lambda *a: None, index=dat["index"], seed=1)
    assert p.run() == 0


def test_impossible_likelihood_stops_filter(gen, sir_data):
    """tests/testthat/test-particle-filter.R:71-93."""
    dat = example_sir(sir_data)

    def compare(state, observed, pars):
        ret = dat["compare"](state, observed, pars)
        if observed["incidence"] > 15:
            ret[:] = -np.inf
        return ret

    np.random.seed(1)
    p = particle_filter(dat["data"], gen, 42, compare, index=dat["index"], seed=1)
    assert p.run(save_history=True) == -np.inf
    first = int(np.where(dat["incidence"] > 15)[0][0]) + 1  # 1-based data row
    h = p.history()
    assert not np.isnan(h[:, :, :first + 1]).any()
    assert np.isnan(h[:, :, first + 1:]).all()


def test_early_exit(gen, sir_data):
    """tests/testthat/test-particle-filter.R:1220-1282 (R compare path: the compiled
    path never receives min_log_likelihood, R/particle_filter_state.R:151)."""
    dat = example_sir(sir_data)
    np.random.seed(1)
    p = particle_filter(dat["data"], gen, 42, dat["compare"], index=dat["index"], seed=1)
    ll = p.run()
    np.random.seed(1)
    p = particle_filter(dat["data"], gen, 42, dat["compare"], index=dat["index"], seed=1)
    assert p.run(min_log_likelihood=ll - 10) == pytest.approx(ll)
    np.random.seed(1)
    p = particle_filter(dat["data"], gen, 42, dat["compare"], index=dat["index"], seed=1)
    assert p.run(min_log_likelihood=ll + 50) == -np.inf
    # compiled path ignores it, as the reference does
    q = particle_filter(dat["data"], gen, 42, None, seed=1)
    assert np.isfinite(q.run(min_log_likelihood=0.0))


def test_missing_data_rows_equal_dropped_rows(gen, sir_data):
    """tests/testthat/test-particle-filter.R:1698-1776 (both compare paths)."""
    inc = sir_data["incidence"].copy()
    inc[::2] = np.nan
    keep = ~np.isnan(inc)
    d1 = particle_filter_data({"day": sir_data["day"], "incidence": inc}, "day", 4, 0)
    d2 = particle_filter_data({"day": sir_data["day"][keep], "incidence": inc[keep]}, "day", 4, 0)
    dat = example_sir(sir_data)
    for compare in (None, dat["compare"]):
        np.random.seed(1)
        a = particle_filter(d1, gen, 42, compare, index=dat["index"], seed=1).run()
        np.random.seed(1)
        b = particle_filter(d2, gen, 42, compare, index=dat["index"], seed=1).run()
        assert a == b and np.isfinite(a)


def test_compiled_and_r_compare_agree_in_distribution(gen, sir_data):
    """tests/testthat/test-particle-filter.R:514-533: mean of 50 runs within 1 %."""
    dat = example_sir(sir_data)
    np.random.seed(1)
    p1 = particle_filter(dat["data"], gen, 100, dat["compare"], index=dat["index"], seed=1)
    p2 = particle_filter(dat["data"], gen, 100, None, index=dat["index"], seed=1)
    a = np.array([p1.run() for _ in range(50)])
    b = np.array([p2.run() for _ in range(50)])
    assert abs(a.mean() - b.mean()) < 0.01 * abs(a.mean())


def test_inputs_round_trip(gen, sir_data):
    """tests/testthat/test-particle-filter.R:1660-1695; seed = live RNG state."""
    dat = example_sir(sir_data, n_days=20)
    p = particle_filter(dat["data"], gen, 42, None, index=dat["index"], seed=100)
    assert p.inputs()["seed"] == 100
    p.run()
    inputs = p.inputs()
    assert isinstance(inputs["seed"], bytes) and len(inputs["seed"]) == 32
    assert inputs["n_particles"] == 42 and inputs["n_parameters"] is None
    p2 = particle_filter_from_inputs(inputs)
    assert p2.inputs()["seed"] == inputs["seed"]
    assert np.isfinite(p2.run())
    assert p.set_n_threads(2) == 1


# ------------------------------------------------------ nested / multi-parameter
def nested_data(sir_data, n_days=30):
    day = sir_data["day"][:n_days]
    a = sir_data["incidence"][:n_days]
    b = np.roll(sir_data["incidence"], 2)[:n_days]
    return particle_filter_data({"day": np.concatenate([day, day]),
                                 "incidence": np.concatenate([a, b]),
                                 "population": np.repeat(["a", "b"], n_days)},
                                "day", 4, 0, population="population"), day, a, b


def test_nested_filter_returns_one_likelihood_per_population(gen, sir_data):
    """tests/testthat/test-particle-filter.R:746-807, 1285-1353."""
    data, day, a, b = nested_data(sir_data)
    pars = [{"beta": 0.2}, {"beta": 0.25}]
    p = particle_filter(data, gen, 42, None, seed=1)
    assert p.has_multiple_parameters and p.has_multiple_data and p.n_parameters == 2
    ll = p.run(pars, save_history=True)
    assert ll.shape == (2,) and np.isfinite(ll).all()
    assert p.history().shape == (5, 42, 2, 31)
    assert p.state().shape == (5, 42, 2)
    with pytest.raises(ValueError, match="Expected an unnamed list of parameters for 'pars'"):
        p.run({"beta": 0.2})
    with pytest.raises(ValueError, match="'pars' must have length 2"):
        p.run([{"beta": 0.2}])
    with pytest.raises(ValueError, match="n_parameters must be 2"):
        particle_filter(data, gen, 42, None, n_parameters=3)
    # population 1 alone, on the first n streams + the shared filter stream, sees the
    # same first-interval likelihood (streams 1..n -> population 1: :1316-1318)
    single = particle_filter(
        particle_filter_data({"day": day, "incidence": a}, "day", 4, 0), gen, 42, None, seed=1)
    obj = single.run_begin(pars[0])
    nested = particle_filter(data, gen, 42, None, seed=1).run_begin(pars)
    assert obj.step(1) == nested.step(1)[0]


def test_replicated_parameters_equal_nested_replicated_data(gen, sir_data):
    """tests/testthat/test-particle-filter.R:1405-1466."""
    n_days = 25
    day = sir_data["day"][:n_days]
    inc = sir_data["incidence"][:n_days]
    shared = particle_filter_data({"day": day, "incidence": inc}, "day", 4, 0)
    nested = particle_filter_data({"day": np.tile(day, 3), "incidence": np.tile(inc, 3),
                                   "pop": np.repeat(["a", "b", "c"], n_days)},
                                  "day", 4, 0, population="pop")
    pars = [{"beta": 0.2}, {"beta": 0.22}, {"beta": 0.18}]
    for compare_of in (lambda d: None, lambda d: d["compare"]):
        dat = example_sir(sir_data, n_days=n_days)
        np.random.seed(1)
        p1 = particle_filter(shared, gen, 42, compare_of(dat), index=dat["index"], seed=1,
                             n_parameters=3)
        assert p1.has_multiple_parameters and not p1.has_multiple_data
        a = p1.run(pars)
        np.random.seed(1)
        p2 = particle_filter(nested, gen, 42, compare_of(dat), index=dat["index"], seed=1)
        b = p2.run(pars)
        assert np.array_equal(a, b) and a.shape == (3,)


def test_nested_r_compare_with_history(gen, sir_data):
    """tests/testthat/test-particle-filter.R:773-807, 868-899."""
    data, day, a, b = nested_data(sir_data)
    dat = example_sir(sir_data, n_days=30)
    np.random.seed(1)
    p = particle_filter(data, gen, 42, dat["compare"], index=dat["index"], seed=1)
    ll = p.run([{"beta": 0.2}, {"beta": 0.25}], save_history=True)
    assert ll.shape == (2,) and np.isfinite(ll).all()
    h = p.history()
    assert h.shape == (3, 42, 2, 31)
    assert np.all(np.diff(h[0], axis=2) <= 0) and np.all(np.diff(h[2], axis=2) >= 0)
    assert np.array_equal(p.history([2, 7]), h[:, [1, 6], :, :])
    ip = np.array([[1, 5], [2, 6]])
    sub = p.history(ip)
    assert np.array_equal(sub[:, :, 0, :], h[:, [0, 1], 0, :])
    assert np.array_equal(sub[:, :, 1, :], h[:, [4, 5], 1, :])


# ---------------------------------------------- the two back ends against each other
@pytest.mark.gpu
@pytest.mark.parametrize("n_particles", [42, 1000])
def test_b200_and_oracle_backends_identical(sir_data, n_particles):
    """tests/testthat/test-particle-filter.R:1480-1504: identical log-likelihoods,
    histories, states and restart states from the two implementations."""
    dat = example_sir(sir_data)
    res = []
    for g in (OracleGenerator("sir"), _gpu_generator("sir")):
        p = particle_filter(dat["data"], g, n_particles, None, index=dat["index"], seed=1)
        lls = [p.run({"beta": 0.2 + 0.01 * k}, save_history=True, save_restart=[20, 50])
               for k in range(3)]
        res.append((lls, p.history(), p.state(), p.restart_state(), p.inputs()["seed"]))
    for a, b in zip(res[0], res[1]):
        if isinstance(a, np.ndarray):
            assert np.array_equal(a, b)
        else:
            assert a == b


@pytest.mark.gpu
def test_b200_and_oracle_backends_identical_nested(sir_data):
    data, day, a, b = nested_data(sir_data)
    pars = [{"beta": 0.2}, {"beta": 0.25}]
    out = []
    for g in (OracleGenerator("sir"), _gpu_generator("sir")):
        p = particle_filter(data, g, 300, None, seed=3)
        out.append((p.run(pars, save_history=True), p.history(), p.state()))
    for x, y in zip(out[0], out[1]):
        assert np.array_equal(x, y)
