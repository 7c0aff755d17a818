"""The product's ``particle_filter_data`` against tests/testthat/test-particle-filter-data.R
(validation messages, layouts, non-unit and irregular times, the missing-initial-time
warning).  tests/test_oracle.py holds the same layouts for the oracle's restatement."""
import warnings

import numpy as np
import pytest

from mcstate_b200.particle_filter_data import particle_filter_data, particle_filter_data_split


def _d(shift=0.0):
    return {"t": np.arange(1, 12) + shift, "y": np.arange(0, 11) + shift}


def test_validates_time():
    """tests/testthat/test-particle-filter-data.R:4-31."""
    with pytest.raises(TypeError, match="'data' must be a data.frame"):
        particle_filter_data(None, "t", 10, 0)
    with pytest.raises(ValueError, match="Did not find column 'time', representing time, in data"):
        particle_filter_data(_d(), "time", 10, 0)
    with pytest.raises(ValueError, match=r"'data\$t' must be an integer"):
        particle_filter_data(_d(0.5), "t", 10, 0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with pytest.raises(ValueError,
                           match=r"The first time must be at least 1 \(but was given 0\)"):
            particle_filter_data(_d(-1), "t", 10)
    with pytest.raises(ValueError, match="All times must be non-negative"):
        particle_filter_data(_d(-2), "t", 10, 0)
    with pytest.raises(ValueError, match="'initial_time' must be non-negative"):
        particle_filter_data(_d(), "t", 10, -1)
    with pytest.raises(ValueError, match="'initial_time' must be <= 1"):
        particle_filter_data(_d(), "t", 10, 2)


def test_reserved_names_rate_and_initial_time():
    """tests/testthat/test-particle-filter-data.R:34-66."""
    for name in ("time", "step", "model_time"):
        with pytest.raises(ValueError, match="The time column cannot be called '%s'" % name):
            particle_filter_data({name: np.arange(1, 11)}, name, 1, 0)
    with pytest.raises(ValueError, match="'rate' must be an integer"):
        particle_filter_data(_d(), "t", 2.3, 0)
    with pytest.raises(ValueError, match="'initial_time' must be non-negative"):
        particle_filter_data(_d(), "t", 2, -10)
    with pytest.raises(ValueError, match="'initial_time' must be <= 1"):
        particle_filter_data(_d(), "t", 2, 2)
    with pytest.raises(ValueError, match="'initial_time' must be an integer"):
        particle_filter_data(_d(), "t", 2, 0.5)


def test_creates_data_and_offsets():
    """tests/testthat/test-particle-filter-data.R:69-110."""
    res = particle_filter_data({"day": np.arange(1, 12), "data": np.arange(0, 1.01, 0.1)},
                               "day", 10, 0)
    assert res.names() == ["day_start", "day_end", "time_start", "time_end", "data"]
    assert list(res["day_start"]) == list(range(0, 11)) and list(res["day_end"]) == list(range(1, 12))
    assert list(res["time_start"]) == [10 * i for i in range(0, 11)]
    assert list(res["time_end"]) == [10 * i for i in range(1, 12)]
    assert (res.rate, res.time, res.population) == (10, "day", None)
    assert np.array_equal(res.model_times, np.stack([np.arange(0, 11), np.arange(1, 12)], 1))
    assert np.array_equal(res.times, res.model_times * 10)
    res = particle_filter_data({"hour": np.arange(11, 21), "a": np.arange(10)}, "hour", 4, 1)
    assert list(res["hour_start"]) == [1] + list(range(11, 20))
    assert list(res["time_start"]) == [4] + [4 * i for i in range(11, 20)]
    assert list(res["time_end"]) == [4 * i for i in range(11, 21)]


def test_one_observation():
    """tests/testthat/test-particle-filter-data.R:115-123."""
    d = {"hour": np.array([1, 2]), "a": np.array([2, 3]), "b": np.array([3, 4])}
    df1 = particle_filter_data({k: v[:1] for k, v in d.items()}, "hour", 10, 0)
    df2 = particle_filter_data(d, "hour", 10, 0)
    assert df1.names() == df2.names() and df1.nrow() == 1
    assert all(df1[k][0] == df2[k][0] for k in df1.names())


def test_populations():
    """tests/testthat/test-particle-filter-data.R:124-170."""
    rng = np.random.default_rng(3)
    data = rng.uniform(size=22)
    day, group = np.tile(np.arange(1, 12), 2), np.repeat(["a", "b"], 11)
    perm = rng.permutation(22)
    res = particle_filter_data({"day": day[perm], "data": data[perm], "group": group[perm]},
                               "day", 10, 0, population="group")
    assert res.is_nested and res.populations == ["a", "b"] and res.nrow() == 22
    assert list(res["day_start"]) == list(range(0, 11)) * 2
    assert list(res["time_end"]) == [10 * i for i in range(1, 12)] * 2
    assert list(res["group"]) == ["a"] * 11 + ["b"] * 11
    assert np.array_equal(res["data"], data)
    split = particle_filter_data_split(res, compiled_compare=False)
    assert len(split) == 11 and [r["group"] for r in split[4]] == ["a", "b"]
    assert split[4][1]["data"] == data[11 + 4]
    with pytest.raises(ValueError, match="Unequal time between populations"):
        particle_filter_data({"day": np.concatenate([np.arange(1, 10, 2), np.arange(1, 11)]),
                              "data1": rng.uniform(size=15),
                              "population": np.repeat(["a", "b"], [5, 10])},
                             "day", 10, 0, population="population")
    with pytest.raises(ValueError,
                       match="Did not find column 'grp', representing population, in data"):
        particle_filter_data({"day": day, "data": data, "group": group}, "day", 1, 0,
                             population="grp")


def test_non_unit_and_irregular_times(sir_data):
    """tests/testthat/test-particle-filter-data.R:231-275: dropping rows keeps the ends
    of the rows that stay and starts each interval where the previous kept one ended."""
    day, inc = sir_data["day"], sir_data["incidence"].copy()
    full = particle_filter_data({"day": day, "incidence": inc}, "day", 4, 0)
    for keep in (np.arange(100) % 2 == 1,
                 np.concatenate([np.random.default_rng(1).uniform(size=99) >= 0.5, [True]])):
        sub = particle_filter_data({"day": day[keep], "incidence": inc[keep]}, "day", 4, 0)
        i = np.flatnonzero(keep)
        assert np.array_equal(sub["day_end"], full["day_end"][i])
        assert np.array_equal(sub["time_end"], full["time_end"][i])
        assert np.array_equal(sub["day_start"], np.concatenate([[0], sub["day_end"][:-1]]))
        assert np.array_equal(sub["time_start"], np.concatenate([[0], sub["time_end"][:-1]]))
        assert np.array_equal(sub["incidence"], inc[i])
        assert (sub.rate, sub.time) == (full.rate, full.time)


def test_warns_without_initial_time():
    """tests/testthat/test-particle-filter-data.R:278-284."""
    d = {"day": np.arange(2, 13), "data": np.arange(0, 1.01, 0.1)}
    with pytest.warns(UserWarning, match="'initial_time' should be provided. I'm assuming '1'"):
        res = particle_filter_data(d, "day", 10)
    ref = particle_filter_data(d, "day", 10, 1)
    assert all(np.array_equal(res[k], ref[k]) for k in ref.names())
    assert np.array_equal(res.times, ref.times)
