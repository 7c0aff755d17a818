"""Multistage parameters: the relations of
tests/testthat/test-particle-filter-multistage.R on the Python mirror, with the R-loop
filter (a compare function) and the compiled filter, on the CPU-oracle generator and on
the B200 engine.  The reference refuses `fork_multistage` on its GPU back end
(R/particle_filter_state.R:361); handing the RNG state to the next stage's model is a
device copy here, so the engine runs it.
"""
import numpy as np
import pytest

from mcstate_b200.particle_filter import (filter_check_times, multistage_epoch,
                                          multistage_parameters, particle_filter)
from mcstate_b200.particle_filter_data import particle_filter_data
from oracle_generator import OracleGenerator
from test_particle_filter import example_sir

BACKENDS = [pytest.param("oracle", id="oracle"),
            pytest.param("gpu", id="b200", marks=pytest.mark.gpu)]
PARS = {"beta": 0.2, "gamma": 0.1}


def _gen(kind):
    if kind == "oracle":
        return OracleGenerator("sir")
    from mcstate_b200 import dust

    g = dust.dust_example("sir")
    if not g.gpu_info()["has_cuda"]:
        pytest.fail("no CUDA device visible: the engine has no CPU fallback")
    return g


def named_index(info):
    return {"run": [5], "state": {"S": 1, "I": 2, "R": 3}}


def _filter(kind, dat, compare, n=42):
    return particle_filter(dat["data"], _gen(kind), n, compare, index=named_index, seed=1,
                           gpu_config=0 if kind == "gpu" else None)


def _rng(model):
    return bytes(model.rng_state())


# ---- constructors (test-particle-filter-multistage.R:1-80) --------------------------
def test_epoch_and_parameters_objects():
    e = multistage_epoch(10)
    assert e.start == 10 and e.pars is None
    assert e.transform_state(np.ones(3), None, None).tolist() == [1, 1, 1]
    with pytest.raises(TypeError, match="'transform_state' must be a function"):
        multistage_epoch(10, transform_state=1)
    p = multistage_parameters(PARS, [])
    assert len(p) == 1 and p[0]["pars"] == PARS and p[0]["start"] is None
    p = multistage_parameters(PARS, [multistage_epoch(10, {"beta": 0.3}),
                                     multistage_epoch(20, {"beta": 0.1})])
    assert [st["start"] for st in p] == [None, 10, 20]
    assert [st["pars"]["beta"] for st in p] == [0.2, 0.3, 0.1]
    with pytest.raises(ValueError, match="all epochs must contain 'pars'"):
        multistage_parameters(PARS, [multistage_epoch(10, {"beta": 0.3}), multistage_epoch(20)])
    with pytest.raises(TypeError, match="multistage_epoch"):
        multistage_parameters(PARS, [None])


def test_filter_check_times(sir_data):
    """test-particle-filter-multistage.R:401-441, 444-484."""
    dat = example_sir(sir_data)
    p = multistage_parameters(PARS, [multistage_epoch(10), multistage_epoch(25)])
    st = filter_check_times(p, dat["data"], [5, 15, 30])
    assert [s["time_index"] for s in st] == [10, 25, 100]
    assert [s["restart_index"] for s in st] == [[0], [1], [2]]
    with pytest.raises(ValueError, match="save_restart cannot include epoch change: error for 10"):
        filter_check_times(p, dat["data"], [5, 10, 15])
    with pytest.raises(ValueError, match="error for 10, 25"):
        filter_check_times(p, dat["data"], [5, 10, 15, 20, 25, 30])
    # data that start after the first epochs: those stages drop out
    late = particle_filter_data({"day": sir_data["day"][35:], "incidence": sir_data["incidence"][35:]},
                                "day", 4, 35)
    st = filter_check_times(p, late, None)
    assert len(st) == 1 and st[0]["time_index"] == 65
    with pytest.raises(ValueError, match="Could not map epoch to filter time"):
        filter_check_times(multistage_parameters(PARS, [multistage_epoch(10.5)]), dat["data"], None)


# ---- the filter ---------------------------------------------------------------------
@pytest.mark.parametrize("kind", BACKENDS)
def test_trivial_and_effectless_multistage_equal_single_stage(kind, sir_data):
    """test-particle-filter-multistage.R:81-139, R-loop filter: identical likelihood, RNG
    state, history and restart state."""
    dat = example_sir(sir_data)
    restart = [5, 15, 25]
    np.random.seed(1)
    f1 = _filter(kind, dat, dat["compare"])
    ll1 = f1.run(PARS, save_history=True, save_restart=restart)
    for epochs in ([], [multistage_epoch(10), multistage_epoch(20)]):
        np.random.seed(1)
        f2 = _filter(kind, dat, dat["compare"])
        ll2 = f2.run(multistage_parameters(PARS, epochs), save_history=True, save_restart=restart)
        assert ll1 == ll2
        assert _rng(f1._last_model[0]) == _rng(f2._last_model[-1])
        assert _rng(f2._last_model[0]) == _rng(f2._last_model[-1])   # the cycle is closed
        assert np.array_equal(f1.history(), f2.history())
        assert np.array_equal(f1.restart_state(), f2.restart_state())
        assert np.array_equal(f1.state(), f2.state())


@pytest.mark.parametrize("kind", BACKENDS)
def test_effectless_multistage_compiled_filter(kind, sir_data):
    """The compiled filter: a stage is one native call on its own model object.  Same draws
    and states as the single call; the likelihood is the same sum associated by stage."""
    dat = example_sir(sir_data)
    restart = [5, 15, 25]
    f1 = _filter(kind, dat, None, n=500)
    ll1 = f1.run(PARS, save_history=True, save_restart=restart)
    f2 = _filter(kind, dat, None, n=500)
    pars = multistage_parameters(PARS, [multistage_epoch(10), multistage_epoch(20)])
    ll2 = f2.run(pars, save_history=True, save_restart=restart)
    assert ll2 == pytest.approx(ll1, rel=1e-13)
    assert np.array_equal(f1.state(), f2.state())
    assert _rng(f1._last_model[0]) == _rng(f2._last_model[-1])
    assert np.array_equal(f1.restart_state(), f2.restart_state())
    h = f2.history()
    assert h.shape == (3, 500, 101) and not np.isnan(h).any()
    assert np.array_equal(f2.history([3, 77]), h[:, [2, 76], :])
    # the last stage's lineages are those of the single call over the same rows
    assert np.array_equal(h[:, :, 21:], f1.history()[:, :, 21:])
    # a second run re-uses the three model objects and continues their streams
    models = list(f2._last_model)
    f1.run(PARS)
    f2.run(pars)
    assert all(a is b for a, b in zip(models, f2._last_model))
    assert np.array_equal(f1.state(), f2.state())


@pytest.mark.parametrize("kind", BACKENDS)
@pytest.mark.parametrize("compiled", [False, True], ids=["r_compare", "compiled"])
def test_transform_state_and_changing_parameters(kind, compiled, sir_data):
    """test-particle-filter-multistage.R:142-180: 20 new infecteds at each epoch show in the
    final state and in the restart states."""
    dat = example_sir(sir_data)

    def add_infecteds(y, info_old, info_new):
        y = np.array(y, dtype=float)
        y[1] += 20
        return y

    epochs = [multistage_epoch(10, transform_state=add_infecteds),
              multistage_epoch(20, transform_state=add_infecteds)]
    compare = None if compiled else dat["compare"]
    np.random.seed(1)
    f1 = _filter(kind, dat, compare)
    ll1 = f1.run(PARS)
    np.random.seed(1)
    f2 = _filter(kind, dat, compare)
    ll2 = f2.run(multistage_parameters(PARS, epochs), save_restart=[5, 15, 25])
    assert ll1 != ll2
    assert np.all(f1.state([1, 2, 3]).sum(axis=0) == 1010)
    assert np.all(f2.state([1, 2, 3]).sum(axis=0) == 1050)
    r = f2.restart_state()
    assert r.shape == (5, 42, 3)
    assert np.array_equal(r[:3].sum(axis=0), np.tile([1010.0, 1030.0, 1050.0], (42, 1)))
    # parameters that change at the epochs
    np.random.seed(1)
    f3 = _filter(kind, dat, compare)
    ll3 = f3.run(multistage_parameters(PARS, [multistage_epoch(30, {"beta": 0.05, "gamma": 0.1})]))
    assert ll3 < ll1 - 5            # the epidemic is throttled from day 30: a poor fit
    assert f3._last_model[1].pars()["beta"] == 0.05


@pytest.mark.parametrize("kind", BACKENDS)
def test_multistage_early_exit_and_partway(kind, sir_data):
    """test-particle-filter-multistage.R:359-398, 444-484."""
    dat = example_sir(sir_data)
    pars = multistage_parameters(PARS, [multistage_epoch(10), multistage_epoch(25)])
    np.random.seed(1)
    f = _filter(kind, dat, dat["compare"])
    assert f.run(pars, save_restart=[40], min_log_likelihood=-20) == -np.inf
    assert np.isnan(f.restart_state()).all() and f.restart_state().shape == (5, 42, 1)
    late = particle_filter_data({"day": sir_data["day"][35:], "incidence": sir_data["incidence"][35:]},
                                "day", 4, 35)
    g = particle_filter(late, _gen(kind), 42, None, index=named_index, seed=1,
                        gpu_config=0 if kind == "gpu" else None)
    ll = g.run(pars, save_history=True)
    assert np.isfinite(ll)
    h = g.history()
    assert h.shape == (3, 42, 66) and not np.isnan(h).any()


@pytest.mark.gpu
def test_multistage_engine_equals_oracle(sir_data):
    dat = example_sir(sir_data)

    def swap(y, *info):
        y = np.array(y, dtype=float)
        y[0] -= 5
        y[1] += 5
        return y

    pars = multistage_parameters(PARS, [multistage_epoch(12, {"beta": 0.25, "gamma": 0.1}, swap),
                                        multistage_epoch(40, {"beta": 0.15, "gamma": 0.12})])
    out = []
    for kind in ("oracle", "gpu"):
        f = _filter(kind, dat, None, n=300)
        ll = f.run(pars, save_history=True, save_restart=[5, 20, 50])
        out.append((ll, f.state(), f.history(), f.restart_state(), _rng(f._last_model[-1])))
    for a, b in zip(*out):
        assert np.array_equal(a, b) if not isinstance(a, (bytes, float)) else a == b
