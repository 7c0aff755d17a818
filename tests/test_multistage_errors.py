"""Error behaviour of the multistage filter that tests/test_multistage.py does not hold
(tests/testthat/test-particle-filter-multistage.R:265-279, 323-356, 567-598), on the
host mirror with the oracle-backed generator."""
import numpy as np
import pytest

from mcstate_b200.particle_filter import multistage_epoch, multistage_parameters, particle_filter
from test_multistage import PARS, named_index
from test_particle_filter import OracleGenerator, example_sir


def _filter(dat, index=named_index, compare="r"):
    return particle_filter(dat["data"], OracleGenerator("sir"), 42,
                           dat["compare"] if compare == "r" else None, index=index, seed=1)


def test_named_index_required_to_save_history(sir_data):
    """:265-279."""
    dat = example_sir(sir_data)
    f = _filter(dat, index=lambda info: {"run": [5], "state": [1, 2, 3]})
    pars = multistage_parameters(PARS, [multistage_epoch(10), multistage_epoch(20)])
    with pytest.raises(ValueError, match="Named index required"):
        f.run(pars, save_history=True)
    assert np.isfinite(f.run(pars))                       # fine without the history


@pytest.mark.parametrize("compare", ["r", None])
def test_no_restart_on_epoch_changes(sir_data, compare):
    """:323-356."""
    dat = example_sir(sir_data)
    pars = multistage_parameters(PARS, [multistage_epoch(10), multistage_epoch(25)])
    f = _filter(dat, compare=compare)
    for ask, bad in (([10], "10"), ([5, 10, 15], "10"), ([5, 10, 15, 20, 25, 30], "10, 25")):
        with pytest.raises(ValueError,
                           match="save_restart cannot include epoch change: error for %s" % bad):
            f.run(pars, save_restart=ask)
    f.run(pars, save_restart=[5, 15, 30])
    assert f.restart_state().shape == (5, 42, 3)


@pytest.mark.parametrize("compare", ["r", None])
def test_number_of_stages_is_fixed_after_the_first_run(sir_data, compare):
    """:567-598."""
    dat = example_sir(sir_data)
    pars0 = PARS
    pars1 = multistage_parameters(PARS, [])
    pars2 = multistage_parameters(PARS, [multistage_epoch(10)])
    f1 = _filter(dat, compare=compare)
    f1.run(pars0)
    f1.run(pars1)
    with pytest.raises(ValueError, match=r"Expected single-stage parameters \(but given one "
                                         r"with 2 stages\)"):
        f1.run(pars2)
    f2 = _filter(dat, compare=compare)
    f2.run(pars2)
    for p in (pars0, pars1):
        with pytest.raises(ValueError, match=r"Expected multistage_pars with 2 stages \(but "
                                             r"given one with 1\)"):
            f2.run(p)
    f2.run(pars2)                                          # and again (:601-627)
