"""Nested (multi-population) relations of SURVEY.md Appendix C that the other modules do not
hold: items 4 (time shift, nested), 6 (fork_smc2, nested) and 9 (constant log-likelihood with
history).  Host-mirror logic on the oracle-backed generator (the engine's nested filter is
compared with the oracle bit for bit in tests/test_engine_gpu.py and test_sharding.py)."""
import numpy as np

from mcstate_b200.particle_filter import particle_filter
from mcstate_b200.particle_filter_data import particle_filter_data
from test_particle_filter import OracleGenerator, example_sir

PARS = [{"I0": 10, "beta": 0.2, "gamma": 0.1}, {"I0": 10, "beta": 0.3, "gamma": 0.1}]


def _nested(sir_data, n_days=40, offset=0):
    day = sir_data["day"][:n_days] + offset
    a = sir_data["incidence"][:n_days]
    b = np.roll(sir_data["incidence"], 2)[:n_days]
    return particle_filter_data({"day": np.concatenate([day, day]),
                                 "incidence": np.concatenate([a, b]),
                                 "population": np.repeat(["a", "b"], n_days)},
                                "day", 4, offset, population="population")


def test_nested_shifted_start_date(sir_data):
    """tests/testthat/test-particle-filter.R:1126-1150."""
    dat = example_sir(sir_data)
    for compare in (None, dat["compare"]):
        ll = []
        for offset in (0, 100):
            np.random.seed(1)
            p = particle_filter(_nested(sir_data, offset=offset), OracleGenerator("sir"), 42,
                                compare, index=dat["index"], seed=1)
            ll.append(p.run(PARS))
        assert ll[0].shape == (2,) and np.array_equal(ll[0], ll[1])


def test_nested_fork_smc2(sir_data):
    """tests/testthat/test-particle-filter.R:935-968: the fork equals a fresh nested filter
    seeded with the parent's pre-fork RNG state, and the parent takes the child's RNG."""
    dat = example_sir(sir_data)
    data = _nested(sir_data, n_days=30)
    for compare in (None, dat["compare"]):
        np.random.seed(1)
        p1 = particle_filter(data, OracleGenerator("sir"), 42, compare, index=dat["index"], seed=1)
        obj = p1.run_begin(PARS, save_history=True)
        for i in range(1, 11):
            obj.step(i)
        before = obj.model.rng_state()
        np.random.seed(1)
        res = obj.fork_smc2(PARS)
        assert res.model.rng_state() == obj.model.rng_state() != before
        p2 = particle_filter(data, OracleGenerator("sir"), 42, compare, index=dat["index"],
                             seed=before)
        np.random.seed(1)
        cmp = p2.run_begin(PARS, save_history=True)
        cmp.step(10)
        assert np.array_equal(res.log_likelihood, cmp.log_likelihood)
        assert np.array_equal(res.model.state(), cmp.model.state())


def test_constant_log_likelihood_leaves_history_alone(sir_data):
    """tests/testthat/test-particle-filter.R:1356-1402: per-population constants add to the
    likelihoods; history and state are those of the plain filter."""
    dat = example_sir(sir_data)
    data = _nested(sir_data, n_days=30)
    out = []
    for const in (None, lambda pars: -10.0 * pars["beta"]):
        np.random.seed(1)
        p = particle_filter(data, OracleGenerator("sir"), 42, dat["compare"], index=dat["index"],
                            seed=1, constant_log_likelihood=const)
        out.append((p.run(PARS, save_history=True), p.history(), p.state()))
    assert np.allclose(out[1][0], out[0][0] + np.array([-2.0, -3.0]), rtol=0, atol=1e-12)
    assert np.array_equal(out[0][1], out[1][1]) and np.array_equal(out[0][2], out[1][2])


def test_nested_initial_states(sir_data):
    """tests/testthat/test-particle-filter.R:1096-1123 (a matrix per population becomes the
    first slice of the history), :1170-1183 (a vector is recycled over the particles)."""
    n = 100
    dat = example_sir(sir_data)
    day = sir_data["day"][:20]
    data = particle_filter_data({"day": np.concatenate([day, day]),
                                 "incidence": np.full(40, np.nan),       # no shuffling occurs
                                 "population": np.repeat(["a", "b"], 20)},
                                "day", 4, 0, population="population")

    def initial(info, n_particles, pars):
        y = np.zeros((5, n_particles))
        i0 = np.random.default_rng(1).poisson(pars["I0"], n_particles)
        y[0] = 1100 - i0
        y[1] = i0
        return y

    pars = [{"I0": 10, "beta": 0.2, "gamma": 0.1}, {"I0": 20, "beta": 0.3, "gamma": 0.1}]
    for compare in (dat["compare"], None):
        p = particle_filter(data, OracleGenerator("sir"), n, compare, index=dat["index"],
                            initial=initial, seed=1)
        p.run(pars, save_history=True)
        h = p.history()
        assert np.array_equal(h[:, :, 0, 0], initial(None, n, pars[0])[:3])
        assert np.array_equal(h[:, :, 1, 0], initial(None, n, pars[1])[:3])
    p = particle_filter(_nested(sir_data), OracleGenerator("sir"), 42, None, index=dat["index"],
                        initial=lambda info, n_particles, pars: [1000.0, 0, 0, 0, 0], seed=1)
    ll = p.run(PARS)
    assert ll.shape == (2,)
    assert np.all(p.state()[0] == 1000.0) and np.all(p.state()[1:4] == 0)   # nobody infected
