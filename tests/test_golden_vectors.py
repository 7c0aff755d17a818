"""The committed regression vectors (tests/golden/oracle_vectors.json, made by
tests/golden/make_oracle_vectors.py): the oracle must still compute them bit for bit,
and the engine must reproduce the whole-filter entries.  This catches the oracle and the
engine drifting together, which the on-the-fly GPU == oracle tests cannot."""
import importlib.util
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")


def _load():
    with open(os.path.join(GOLDEN, "oracle_vectors.json")) as fh:
        return json.load(fh)


def test_oracle_reproduces_the_committed_vectors(orc):
    spec = importlib.util.spec_from_file_location("make_oracle_vectors",
                                                  os.path.join(GOLDEN, "make_oracle_vectors.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    want, got = _load(), json.loads(json.dumps(mod.vectors()))
    assert sorted(got) == sorted(want)
    for key in want:
        assert got[key] == want[key], key


def test_fixture_files_are_the_reference_series():
    """inst/sir_incidence.csv: 100 days, column `cases`."""
    with open(os.path.join(GOLDEN, "sir_incidence.csv")) as fh:
        head = fh.readline().strip().replace('"', "")
        rows = fh.readlines()
    assert head == "cases,day" and len(rows) == 100
    with open(os.path.join(GOLDEN, "nested_sir_incidence.csv")) as fh:
        head = fh.readline().strip().replace('"', "")
    assert head.split(",") == ["cases", "day", "population"]


@pytest.mark.gpu
def test_engine_reproduces_the_committed_filter_vectors(orc):
    from mcstate_b200 import dust

    d = np.loadtxt(os.path.join(GOLDEN, "sir_incidence.csv"), delimiter=",", skiprows=1)
    t, x = d[:, 1].astype(np.uint64) * 4, d[:, 0]
    want = _load()["filters"]
    cases = [("sir_n100_seed1", "sir", 100, 1, t, x, "incidence"),
             ("sir_n4096_seed1", "sir", 4096, 1, t, x, "incidence"),
             ("sirs_n64_seed3", "sirs", 64, 3, t, x, "incidence"),
             ("volatility_n50_seed5", "volatility", 50, 5, np.arange(1, 41, dtype=np.uint64),
              np.sin(np.arange(1, 41) * 0.37) * 2.0, "observed")]
    for name, model, n, seed, tt, xx, col in cases:
        for single_launch in (True, False):
            m = dust.dust_example(model).new({}, 0, n, seed=seed)
            m.set_persistent(single_launch)
            m.set_data(dust.dust_data({"time_end": tt, col: xx}), False)
            ll = m.filter()["log_likelihood"][0]
            assert mod_bits(ll) == want[name]["log_likelihood"], (name, single_launch, ll)
            assert mod_bits(float(m.state().sum())) == want[name]["state_sum"]
            words = np.frombuffer(m.rng_state(), dtype=np.uint64)
            assert int(words[0]) == want[name]["rng_word_0"]


def mod_bits(x):
    import struct

    return "%016x" % struct.unpack("<Q", struct.pack("<d", float(x)))[0]
