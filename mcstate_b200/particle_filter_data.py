"""``particle_filter_data()``: observations -> the interval schedule the filter runs.

Mirror of R/particle_filter_data.R:76-181 (construction and validation, same error
messages) and :184-207 (``particle_filter_data_split``).  Discrete-time models only:
the engine's models are discrete (``time_type() == "discrete"``).
"""
from __future__ import annotations

import warnings

import numpy as np

from .dust import dust_data


class particle_filter_data:
    """The classed data.frame R returns, as columns + attributes.

    ``columns`` holds ``<time>_start``, ``<time>_end``, ``time_start``, ``time_end``
    and every data column (and the population column when nested), row order as in R
    (sorted by population, then time).  Attributes: ``rate``, ``time``, ``times``,
    ``model_times``, ``population``, ``populations``.
    """

    def __init__(self, data, time, rate, initial_time=None, population=None):
        cols = _as_columns(data)
        if not isinstance(time, str):
            raise TypeError("'time' must be a character")
        if time not in cols:
            raise ValueError("Did not find column '%s', representing time, in data" % time)
        if time in ("step", "time", "model_time"):
            raise ValueError("The time column cannot be called '%s'" % time)
        if rate is None:
            raise ValueError("continuous-time data (rate = NULL) is not supported by this engine")
        if int(rate) != rate:
            raise ValueError("'rate' must be an integer")
        if rate < 1:
            raise ValueError("'rate' must be at least 1")
        rate = int(rate)

        n = len(cols[time])
        order = np.arange(n)
        if population is None:
            populations = None
            model_time_end = np.asarray(cols[time])
        else:
            if population not in cols:
                raise ValueError(
                    "Did not find column '%s', representing population, in data" % population)
            pop = np.asarray(cols[population])
            populations = sorted(set(pop.tolist()))
            order = np.lexsort((np.asarray(cols[time]), pop))
            tsplit = [np.asarray(cols[time])[order][pop[order] == p] for p in populations]
            if any(len(s) != len(tsplit[0]) or np.any(s != tsplit[0]) for s in tsplit):
                raise ValueError("Unequal time between populations")
            model_time_end = tsplit[0]

        mte = np.asarray(model_time_end, dtype=np.float64)
        if np.any(mte != np.floor(mte)):
            raise ValueError("'data$%s' must be an integer" % time)
        mte = mte.astype(np.int64)
        if initial_time is None:
            initial_time = int(mte[0]) - 1
            warnings.warn(
                "'initial_time' should be provided. I'm assuming '%d' which is one time unit "
                "before the first time in your data (%d), but this might not be appropriate. "
                "This will become an error in a future version of mcstate"
                % (initial_time, mte[0]))
            if mte[0] < 1:
                raise ValueError("The first time must be at least 1 (but was given %d)" % mte[0])
        if np.any(mte < 0):
            raise ValueError("All times must be non-negative")
        if initial_time != np.floor(initial_time):
            raise ValueError("'initial_time' must be an integer")
        initial_time = int(initial_time)
        if initial_time < 0:
            raise ValueError("'initial_time' must be non-negative")
        if initial_time > mte[0]:
            raise ValueError("'initial_time' must be <= %d" % mte[0])

        mts = np.concatenate([[initial_time], mte[:-1]])
        self.model_times = np.stack([mts, mte], axis=1)
        self.times = self.model_times * rate
        self.rate = rate
        self.time = time
        self.population = population
        self.populations = populations
        reps = 1 if populations is None else len(populations)
        out = {
            time + "_start": np.tile(self.model_times[:, 0], reps),
            time + "_end": np.tile(self.model_times[:, 1], reps),
            "time_start": np.tile(self.times[:, 0], reps),
            "time_end": np.tile(self.times[:, 1], reps),
        }
        for k, v in cols.items():
            if k != time:
                out[k] = np.asarray(v)[order]
        self.columns = out

    @property
    def is_nested(self):
        return self.populations is not None

    def nrow(self):
        return len(self.columns["time_end"])

    def __getitem__(self, name):
        return self.columns[name]

    def names(self):
        return list(self.columns)

    def rows(self, keep):
        """Row subset, like ``data[i, ]`` (R/particle_filter_data.R:211-227): the time
        attributes follow the rows that are kept."""
        keep = np.arange(self.nrow())[keep]
        new = object.__new__(particle_filter_data)
        new.__dict__.update(self.__dict__)
        new.columns = {k: v[keep] for k, v in self.columns.items()}
        k = keep[keep < len(self.times)]
        new.model_times = self.model_times[k]
        new.times = self.times[k]
        return new


def subset_populations(data, populations):
    """The nested data restricted to some of its populations (all time rows kept): what a
    shard of a multi-population filter holds."""
    if not data.is_nested:
        raise ValueError("'data' is not nested")
    populations = list(populations)
    if any(p not in data.populations for p in populations):
        raise ValueError("unknown population")
    keep = np.isin(np.asarray(data.columns[data.population]), populations)
    new = object.__new__(particle_filter_data)
    new.__dict__.update(data.__dict__)
    new.columns = {k: v[keep] for k, v in data.columns.items()}
    new.populations = [p for p in data.populations if p in populations]
    return new


def _as_columns(data):
    if hasattr(data, "to_dict") and hasattr(data, "columns"):  # pandas.DataFrame
        return {c: np.asarray(data[c]) for c in data.columns}
    if isinstance(data, dict):
        return {k: np.asarray(v) for k, v in data.items()}
    raise TypeError("'data' must be a data.frame")


def particle_filter_data_split(data, compiled_compare):
    """R/particle_filter_data.R:184-207: one entry per interval."""
    cols = dict(data.columns)
    if compiled_compare:
        return dust_data(cols, "time_end", data.population)
    n = data.nrow()
    rows = [{k: (v[i].item() if hasattr(v[i], "item") else v[i]) for k, v in cols.items()}
            for i in range(n)]
    if data.population is None:
        return rows
    groups = [[r for r in rows if r[data.population] == p] for p in data.populations]
    return [[g[i] for g in groups] for i in range(n // len(data.populations))]
