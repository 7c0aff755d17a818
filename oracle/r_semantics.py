"""numpy restatement of the pure-R numerics on the hot path -- TEST INFRASTRUCTURE ONLY.

Each function follows one reference function and says which.  Indices are
1-based where the R code's are, so the reference's known answers can be typed
in unchanged.
"""
from __future__ import annotations

import numpy as np


def scale_log_weights(log_weights):
    """R/particle_filter.R:570-587."""
    lw = np.array(log_weights, dtype=np.float64)
    lw[np.isnan(lw)] = -np.inf
    m = lw.max()
    if not np.isfinite(m):
        return {"weights": np.full(lw.shape, np.nan), "average": -np.inf}
    w = np.exp(lw - m)
    return {"weights": w, "average": float(np.log(w.mean()) + m)}


def particle_resample(weights, u0):
    """R/particle_filter.R:515-523 with ``runif(1, 0, 1/n)`` supplied as ``u0``.

    ``findInterval(u, cw)`` = number of ``cw_j <= u``; matrix input is handled
    column by column (``:516-518``) with one ``u0`` per column.
    """
    w = np.asarray(weights, dtype=np.float64)
    if w.ndim == 2:
        u0 = np.broadcast_to(u0, (w.shape[1],))
        return np.stack([particle_resample(w[:, j], u0[j]) for j in range(w.shape[1])], axis=1)
    n = w.size
    u = u0 + np.arange(n) * (1.0 / n)
    cw = np.cumsum(w / w.sum())
    return np.searchsorted(cw, u, side="right") + 1


def particle_filter_early_exit(log_likelihood, min_log_likelihood):
    """R/particle_filter_state.R:463-474."""
    ll = np.atleast_1d(np.asarray(log_likelihood, dtype=np.float64))
    mn = np.atleast_1d(np.asarray(min_log_likelihood, dtype=np.float64))
    if np.any(ll == -np.inf):
        return True
    if ll.size == 1:
        return bool(ll[0] < mn[0])
    if mn.size == 1:
        return bool(ll.sum() < mn[0])
    return bool(np.all(ll < mn))


def history_single(history_value, history_order, index_particle=None):
    """R/particle_filter.R:616-646. value ``(ny, np, nt)``, order ``(np, nt)`` 1-based."""
    hv = np.asarray(history_value)
    if history_order is None:
        return hv if index_particle is None else hv[:, np.asarray(index_particle) - 1, :]
    ho = np.asarray(history_order)
    ny, _, nt = hv.shape
    ip = np.arange(1, hv.shape[1] + 1) if index_particle is None else np.asarray(index_particle)
    out = np.empty((ny, ip.size, nt), dtype=hv.dtype)
    for t in range(nt - 1, -1, -1):
        ip = ho[ip - 1, t]
        out[:, :, t] = hv[:, ip - 1, t]
    return out


def particle_filter_data(time, columns, rate, initial_time, population=None):
    """R/particle_filter_data.R:76-181, discrete-time branch.

    ``time``: integer vector; ``columns``: dict name -> vector; ``population``:
    optional vector of labels (the factor column).  Returns a dict with the
    columns the R data.frame has plus the ``times`` / ``model_times``
    attributes and the sorted population levels.
    """
    time = np.asarray(time)
    cols = {k: np.asarray(v) for k, v in columns.items()}
    if population is None:
        populations = None
        model_time_end = time
        order = np.arange(time.size)
    else:
        pop = np.asarray(population)
        populations = sorted(set(pop.tolist()))
        order = np.lexsort((time, pop))
        split = [time[order][pop[order] == p] for p in populations]
        if any(len(s) != len(split[0]) or np.any(s != split[0]) for s in split):
            raise ValueError("Unequal time between populations")
        model_time_end = split[0]
    if np.any(model_time_end != np.floor(model_time_end)):
        raise ValueError("time must be an integer")
    if initial_time is None:
        initial_time = model_time_end[0] - 1
        if model_time_end[0] < 1:
            raise ValueError("The first time must be at least 1")
    if np.any(model_time_end < 0):
        raise ValueError("All times must be non-negative")
    if initial_time != np.floor(initial_time):
        raise ValueError("'initial_time' must be an integer")
    if initial_time < 0:
        raise ValueError("'initial_time' must be non-negative")
    if initial_time > model_time_end[0]:
        raise ValueError("'initial_time' must be <= %d" % model_time_end[0])
    model_time_start = np.concatenate([[initial_time], model_time_end[:-1]])
    model_times = np.stack([model_time_start, model_time_end], axis=1).astype(np.int64)
    times = model_times * rate
    reps = 1 if populations is None else len(populations)
    out = {
        "time_start": np.tile(times[:, 0], reps),
        "time_end": np.tile(times[:, 1], reps),
        "model_time_start": np.tile(model_times[:, 0], reps),
        "model_time_end": np.tile(model_times[:, 1], reps),
        "columns": {k: v[order] for k, v in cols.items()},
        "population": None if population is None else np.asarray(population)[order],
        "populations": populations,
        "rate": rate,
        "times": times,
        "model_times": model_times,
    }
    return out
