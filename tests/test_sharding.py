"""Sharded nested filter (mcstate_b200.sharding): the populations of one multi-parameter
model split over ranks give bit-for-bit the unsharded filter's log-likelihoods
(stream partition pinned by tests/testthat/test-particle-filter.R:1316-1318).
World-size-2 gloo processes on the CPU (oracle-backed generator); `-m gpu` runs the same two
ranks against the B200 engine (both on device 0 when the box has one GPU)."""
import os
import pickle
import subprocess
import sys
import textwrap

import numpy as np
import pytest
from conftest import free_port

from mcstate_b200.particle_filter import particle_filter
from mcstate_b200.particle_filter_data import particle_filter_data, subset_populations
from mcstate_b200.sharding import block_range
from oracle_generator import OracleGenerator

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def nested_data(n_pop=3, with_na=True):
    """inst/nested_sir_incidence.csv (populations A, B) plus synthetic ones made from A."""
    rows = [l.strip().split(",") for l in open(os.path.join(ROOT, "tests", "golden",
                                                            "nested_sir_incidence.csv"))][1:]
    cases = np.array([float(r[0]) for r in rows])
    day = np.array([int(r[1]) for r in rows])
    pop = np.array([r[2].strip('"') for r in rows])
    cols = {"day": list(day[pop == "A"]) + list(day[pop == "B"]),
            "incidence": list(cases[pop == "A"]) + list(cases[pop == "B"]),
            "population": ["A"] * 100 + ["B"] * 100}
    for k in range(2, n_pop):
        extra = np.roll(cases[pop == "A"], 3 * k) + (k % 2)
        cols["day"] += list(day[pop == "A"])
        cols["incidence"] += list(extra)
        cols["population"] += [chr(ord("A") + k)] * 100
    inc = np.array(cols["incidence"], dtype=float)
    if with_na:  # missing observations at different rows of different populations
        inc[7] = np.nan
        inc[100 + 30] = np.nan
        inc[100 * (n_pop - 1) + 30] = np.nan
        inc[100 * (n_pop - 1) + 55] = np.nan
    cols["incidence"] = inc
    return particle_filter_data(cols, "day", 4, 0, population="population")


def pars_for(n_pop):
    return [{"beta": 0.2 + 0.02 * k, "gamma": 0.1 + 0.005 * k} for k in range(n_pop)]


def test_block_range_and_subset():
    assert [block_range(3, r, 2) for r in range(2)] == [(0, 2), (2, 3)]
    assert [block_range(8, r, 4) for r in range(4)] == [(0, 2), (2, 4), (4, 6), (6, 8)]
    assert [block_range(5, r, 3) for r in range(3)] == [(0, 2), (2, 4), (4, 5)]
    d = nested_data(3)
    s = subset_populations(d, ["B", "C"])
    assert s.populations == ["B", "C"] and s.nrow() == 200
    assert np.array_equal(s.times, d.times)
    with pytest.raises(ValueError):
        subset_populations(d, ["Z"])


_WORKER = textwrap.dedent("""
    import os, sys, pickle
    import numpy as np
    import torch.distributed as dist
    sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
    from test_sharding import nested_data, pars_for, make_generator, impossible
    from mcstate_b200.sharding import nested_filter_sharded
    kind, n_pop, n_particles = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    backend = "gloo"
    if kind == "gpu":   # one GPU per rank and NCCL when the box has the GPUs, else gloo
        import torch
        backend = os.environ.get("MCB_TEST_BACKEND") or (
            "nccl" if torch.cuda.device_count() >= int(os.environ["WORLD_SIZE"]) else "gloo")
    dev = None
    if kind == "gpu":
        import torch
        dev = int(os.environ["LOCAL_RANK"]) % torch.cuda.device_count()
        torch.cuda.set_device(dev)
    import datetime
    dist.init_process_group(backend, timeout=datetime.timedelta(seconds=120))
    f = nested_filter_sharded(nested_data(n_pop), make_generator(kind), n_particles, seed=7,
                              gpu_config=dev)
    p = pars_for(n_pop)
    out = [f.run(p), f.run(p), f.run(list(reversed(p)))]
    if os.environ.get("MCB_TEST_IMPOSSIBLE"):
        # the last population cannot have produced its data (no infections, no noise): the
        # whole vector is -Inf on every rank, as from the unsharded filter
        # (R/particle_filter_state.R:463-474), and the next evaluation is a normal one again
        out = [f.run(impossible(p)), f.run(p)]
    if dist.get_rank() == 0:
        with open(sys.argv[4], "wb") as fh:
            pickle.dump(out, fh)
    dist.barrier()
    dist.destroy_process_group()
""")


def make_generator(kind):
    if kind == "oracle":
        return OracleGenerator("sir")
    from mcstate_b200 import dust

    g = dust.dust_example("sir")
    if not g.gpu_info()["has_cuda"]:
        raise RuntimeError("no CUDA device visible: the engine has no CPU fallback")
    return g


def _sharded(tmp_path, kind, n_pop, n_particles, world=2):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    out = tmp_path / "res.pkl"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
           "--nproc-per-node=%d" % world, "--master-addr", "127.0.0.1", "--master-port",
           str(free_port()), str(script), kind, str(n_pop), str(n_particles),
           str(out)]
    p = subprocess.run(cmd, env=dict(os.environ, MASTER_ADDR="127.0.0.1"), capture_output=True,
                       text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-3000:]
    with open(out, "rb") as fh:
        return pickle.load(fh)


def impossible(pars):
    """the last population without infections or observation noise: modelled incidence is
    exactly 0, every positive observation has log density -Inf"""
    return list(pars[:-1]) + [dict(pars[-1], exp_noise=float("inf"), I0=0)]


def _unsharded(kind, n_pop, n_particles):
    f = particle_filter(nested_data(n_pop), make_generator(kind), n_particles, None, seed=7,
                        gpu_config=0 if kind == "gpu" else None)
    p = pars_for(n_pop)
    return [f.run(p), f.run(p), f.run(list(reversed(p)))]


@pytest.mark.parametrize("kind", [pytest.param("oracle", id="oracle"),
                                  pytest.param("gpu", id="b200", marks=pytest.mark.gpu)])
def test_sharded_nested_filter_equals_unsharded(tmp_path, kind):
    """3 populations over 2 ranks (blocks of 2 and 1), NA rows in different populations,
    three consecutive runs (the RNG streams carry over between them)."""
    n_particles = 96 if kind == "oracle" else 4096
    want = _unsharded(kind, 3, n_particles)
    got = _sharded(tmp_path, kind, 3, n_particles)
    for w, g in zip(want, got):
        assert np.array_equal(np.asarray(w), np.asarray(g)), (w, g)
    assert not np.array_equal(want[0], want[1])  # the second run continued the streams


def test_sharded_nested_filter_with_an_impossible_population(tmp_path, monkeypatch):
    """ADVICE (round 1): when one population is impossible only its rank stops early; the
    gathered RESULT must still be the unsharded filter's (all -Inf), on every rank, and the
    filter must be usable afterwards."""
    monkeypatch.setenv("MCB_TEST_IMPOSSIBLE", "1")
    got = _sharded(tmp_path, "oracle", 3, 40)
    f = particle_filter(nested_data(3), make_generator("oracle"), 40, None, seed=7)
    p = pars_for(3)
    want_bad = f.run(impossible(p))
    assert np.all(np.asarray(want_bad) == -np.inf)
    assert np.array_equal(np.asarray(got[0]), np.asarray(want_bad))
    assert np.all(np.isfinite(got[1])) and len(got[1]) == 3
