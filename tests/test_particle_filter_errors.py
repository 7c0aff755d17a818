"""Extraction and error behaviour of the filter facade, ported from the blocks of
tests/testthat/test-particle-filter.R that tests/test_particle_filter.py does not cover
(file:line per test; messages are the reference's).  Oracle-backed generator on the CPU,
the B200 engine under ``-m gpu``."""
import numpy as np
import pytest

from mcstate_b200.particle_filter import particle_filter
from test_particle_filter import BACKENDS, OracleGenerator, _gpu_generator, example_sir


@pytest.fixture(params=BACKENDS)
def gen(request):
    return OracleGenerator("sir") if request.param == "oracle" else _gpu_generator("sir")


def test_nothing_to_extract_before_a_run(gen, sir_data):
    """tests/testthat/test-particle-filter.R:296-310."""
    dat = example_sir(sir_data)
    p = particle_filter(dat["data"], gen, 42, dat["compare"], index=dat["index"])
    for call in (p.state, p.history, p.restart_state):
        with pytest.raises(RuntimeError, match="Model has not yet been run"):
            call()


@pytest.mark.parametrize("compiled", [False, True])
def test_filter_state_on_extraction(gen, sir_data, compiled):
    """tests/testthat/test-particle-filter.R:313-325."""
    dat = example_sir(sir_data)
    p = particle_filter(dat["data"], gen, 42, None if compiled else dat["compare"],
                        index=dat["index"], seed=1)
    assert isinstance(p.run(), float)
    state = p.state()
    assert state.shape == (5, 42)
    assert np.array_equal(p.state([1]), state[:1])
    assert np.array_equal(p.state([2, 3]), state[1:3])


@pytest.mark.parametrize("compiled", [False, True])
def test_extract_one_restart_state(gen, sir_data, compiled):
    """tests/testthat/test-particle-filter.R:452-477: the restart state at day 40 is the
    final state of a filter run on the first 40 days; one particle of it can be asked for."""
    n, end = 42, 40
    dat = example_sir(sir_data)
    np.random.seed(1)
    p = particle_filter(dat["data"], gen, n, None if compiled else dat["compare"],
                        index=dat["index"], seed=100)
    p.run(save_restart=[end])
    s, s1 = p.restart_state(), p.restart_state(1)
    assert s.shape == (5, n, 1) and s1.shape == (5, 1, 1)
    dat40 = example_sir(sir_data, n_days=end)
    np.random.seed(1)
    q = particle_filter(dat40["data"], gen, n, None if compiled else dat40["compare"],
                        index=dat40["index"], seed=100)
    q.run()
    cmp = q.state()
    assert np.array_equal(s[:, :, 0], cmp)
    assert np.array_equal(s1[:, 0, 0], cmp[:, 0])


def test_restart_state_needs_saving(gen, sir_data):
    """tests/testthat/test-particle-filter.R:480-491."""
    dat = example_sir(sir_data)
    p = particle_filter(dat["data"], gen, 42, dat["compare"], index=dat["index"], seed=100)
    p.run()
    with pytest.raises(RuntimeError,
                       match="Can't get history as model was run with save_restart = NULL"):
        p.restart_state()


def test_restart_dates_must_exist(gen, sir_data):
    """tests/testthat/test-particle-filter.R:494-511."""
    dat = example_sir(sir_data, n_days=20)
    p = particle_filter(dat["data"], gen, 42, dat["compare"], index=dat["index"], seed=100)
    for ask, bad in (([30], "30"), ([30, 50], "30, 50"), ([10, 30, 50], "30, 50")):
        with pytest.raises(ValueError,
                           match="'save_restart' contains times not in 'day': %s" % bad):
            p.run(save_restart=ask)


def test_no_partial_likelihood_with_compiled_compare(gen, sir_data):
    """tests/testthat/test-particle-filter.R:589-601."""
    dat = example_sir(sir_data)
    p = particle_filter(dat["data"], gen, 100, None, index=dat["index"])
    obj = p.run_begin()
    with pytest.raises(ValueError, match="'partial' not supported with compiled compare"):
        obj.step(10, True)


def test_history_needs_saving_and_index_particle(gen, sir_data):
    """tests/testthat/test-particle-filter.R:270-293 (tail), 536-554: history of chosen
    particles equals the matching columns of the whole history."""
    dat = example_sir(sir_data)
    p = particle_filter(dat["data"], gen, 30, None, index=dat["index"], seed=3)
    p.run()
    with pytest.raises(RuntimeError,
                       match="Can't get history as model was run with save_history = FALSE"):
        p.history()
    p.run(save_history=True)
    h = p.history()
    assert h.shape == (3, 30, 101)
    assert np.array_equal(p.history([4, 9]), h[:, [3, 8], :])
    assert np.array_equal(p.history(1), h[:, :1, :])


def test_gpu_config_requires_gpu_support(sir_data):
    """tests/testthat/test-particle-filter.R:1186-1194: a generator without GPU support
    refuses ``gpu_config`` (the oracle-backed generator is such a model)."""
    dat = example_sir(sir_data)
    with pytest.raises(ValueError, match="'gpu_config' provided, but 'model' does not have GPU support"):
        particle_filter(dat["data"], OracleGenerator("sir"), 10, None, index=dat["index"],
                        gpu_config=0)


def test_nested_initial_errors(gen, sir_data):
    """tests/testthat/test-particle-filter.R:918-932, 1153-1167."""
    from test_particle_filter import nested_data

    data = nested_data(sir_data)[0]
    pars = [{"beta": 0.2, "gamma": 0.1}, {"beta": 0.3, "gamma": 0.1}]
    p = particle_filter(data, gen, 42, None, initial=lambda info, n, pars: {"time": 0})
    with pytest.raises(ValueError, match="Setting 'time' from initial no longer supported"):
        p.run(pars)
    unequal = [dict(pars[0], extra=[]), dict(pars[1], extra=[10.0])]
    p = particle_filter(data, gen, 42, None,
                        initial=lambda info, n, pars: [1000.0] + list(pars["extra"]) + [0.0] * 3)
    with pytest.raises(ValueError, match="unequal state"):
        p.run(unequal)


def test_names_on_history(gen, sir_data):
    """tests/testthat/test-particle-filter.R:367-392: a named ``index$state`` names the
    rows of the history; an unnamed one leaves them unnamed."""
    dat = example_sir(sir_data)
    named = particle_filter(dat["data"], gen, 42, dat["compare"], seed=100,
                            index=lambda info: {"run": [4], "state": {"S": 1, "I": 2, "R": 3}})
    named.run(save_history=True)
    h = named.history()
    assert h.shape == (3, 42, 101)
    assert named.history_names() == ["S", "I", "R"]
    plain = particle_filter(dat["data"], gen, 42, dat["compare"], seed=100,
                            index=lambda info: {"run": [4], "state": [1, 2, 3]})
    plain.run(save_history=True)
    assert plain.history_names() is None


def test_change_thread_count(gen, sir_data):
    """tests/testthat/test-particle-filter.R:395-414."""
    dat = example_sir(sir_data)
    p = particle_filter(dat["data"], gen, 42, dat["compare"], index=dat["index"], seed=100,
                        n_threads=1)
    assert p.set_n_threads(2) == 1
    p.run()
    assert p._last_model[0].n_threads() == 2
    assert p.set_n_threads(1) == 2
    assert p._last_model[0].n_threads() == 1
    p.run()
    assert p._last_model[0].n_threads() == 1


@pytest.mark.parametrize("compiled", [False, True])
def test_stepping_errors(gen, sir_data, compiled):
    """tests/testthat/test-particle-filter.R:661-711."""
    dat = example_sir(sir_data)
    p = particle_filter(dat["data"], gen, 42, None if compiled else dat["compare"],
                        index=dat["index"], seed=1)
    obj = p.run_begin()
    obj.run()
    for call in (obj.run, lambda: obj.step(100), lambda: obj.step(101)):
        with pytest.raises(ValueError, match="Particle filter has reached the end of the data"):
            call()
    obj = p.run_begin()
    obj.step(10)
    for k, t in ((10, 40), (5, 20)):
        with pytest.raises(ValueError, match=r"Particle filter has already run step index %d "
                                             r"\(to model step %d\)" % (k, t)):
            obj.step(k)
    obj = p.run_begin()
    with pytest.raises(ValueError,
                       match=r"time_index 101 is beyond the length of the data \(max 100\)"):
        obj.step(101)


def _nested(sir_data, gen, compare="example", **kw):
    from test_particle_filter import nested_data

    data = nested_data(sir_data, n_days=100)[0]
    dat = example_sir(sir_data)
    cmp = dat["compare"] if compare == "example" else compare
    return particle_filter(data, gen, 42, cmp, index=kw.pop("index", dat["index"]), **kw), dat


NESTED_PARS = [{"beta": 0.2, "gamma": 0.1}, {"beta": 0.3, "gamma": 0.1}]


def test_nested_impossible_likelihood(gen, sir_data):
    """tests/testthat/test-particle-filter.R:1007-1034: the filter stops at the first
    interval one population finds impossible; the history after it stays NA."""
    dat = example_sir(sir_data)

    def compare(state, observed, pars):
        ret = dat["compare"](state, observed, pars)
        if observed["incidence"] > 15:
            ret[:] = -np.inf
        return ret

    p, _ = _nested(sir_data, gen, compare)
    np.random.seed(1)
    res = p.run(NESTED_PARS, save_history=True)
    assert -np.inf in res
    first = int(np.argmax(sir_data["incidence"] > 15))     # 0-based data row of population a
    h = p.history()
    assert h.shape == (3, 42, 2, 101)
    assert not np.isnan(h[:, :, 0, :first + 2]).any()
    assert np.isnan(h[:, :, 0, first + 2:]).all()


def test_nested_null_compare_and_initial(gen, sir_data):
    """tests/testthat/test-particle-filter.R:1036-1058."""
    p, _ = _nested(sir_data, gen, lambda *a: None)
    res = p.run(NESTED_PARS, save_history=True)
    assert np.array_equal(res, [0.0, 0.0])
    assert p.history().shape == (3, 42, 2, 101)
    p, _ = _nested(sir_data, gen, initial=lambda *a: None, seed=100)
    np.random.seed(1)
    assert np.isfinite(p.run(NESTED_PARS)).all()


def test_nested_index_must_agree(gen, sir_data):
    """tests/testthat/test-particle-filter.R:1078-1093."""
    p, _ = _nested(sir_data, gen, seed=100,
                   index=lambda info: {"run": [int(round(info["pars"]["beta"] * 10))]})
    with pytest.raises(ValueError, match="index must be identical across populations"):
        p.run(NESTED_PARS, save_history=True)


def test_nested_early_exit(gen, sir_data):
    """tests/testthat/test-particle-filter.R:1251-1282: a scalar threshold applies to the
    sum over populations, a vector one population by population (all must fall below)."""
    def runs(min_ll=None):
        np.random.seed(1)
        p, _ = _nested(sir_data, gen, seed=1)
        return np.array([p.run(NESTED_PARS, min_log_likelihood=min_ll) for _ in range(10)]).T

    ll1 = runs()
    tot = ll1.sum(axis=0)
    ll2 = runs(float(tot.mean()))
    assert np.isinf(ll2).any()
    fin = np.isfinite(ll2).all(axis=0)
    assert ll2[:, fin].sum(axis=0).min() >= tot.mean()
    vec = ll1[:, int(np.argmin(np.abs(tot - tot.mean())))]
    ll3 = runs(vec)
    alive = (ll3 > -np.inf).any(axis=0)
    assert not alive.all()
    assert np.all(ll3[:, ~alive] == -np.inf)
    assert np.all((ll3[:, alive] >= vec[:, None]).any(axis=0))
