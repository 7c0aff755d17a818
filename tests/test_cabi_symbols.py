"""The C-ABI library loads and exports every symbol include/mcstate_b200.h declares
(no compute: runs without a GPU)."""
import ctypes as C
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "mcstate_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mcb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from mcstate_b200 import _lib

    _lib.build()
    L = C.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 35
    for n in names:
        assert hasattr(L, n), n
    assert sorted(_lib.SIGNATURES) == names


def test_host_only_entry_points_work_without_gpu(orc):
    from mcstate_b200 import _lib, dust

    L = _lib.lib()
    assert L.mcb_version() >= 100
    # model registry
    g = dust.dust_example("sir")
    assert g.state_names == ["S", "I", "R", "cases_cumul", "cases_inc"]
    assert g.par_names[:4] == ["beta", "gamma", "I0", "exp_noise"]
    assert g.par_defaults[:4] == [0.2, 0.1, 10.0, 1e6]
    assert dust.dust_generator.name == "dust_generator" and g.has_compare()
    # seeding plumbing equals the oracle's
    s = np.zeros(4, dtype=np.uint64)
    assert L.mcb_rng_seed(42, s.ctypes.data_as(_lib.u64p)) == 0
    assert np.array_equal(s, orc.seed_state(42))
    got = dust.dust_rng_distributed_state(1, 5, 3)
    want = orc.rng_distributed_state(1, 5, 3)
    assert [np.frombuffer(b, dtype=np.uint64).reshape(5, 4).tolist() for b in got] == want.tolist()
    r = dust.dust_rng(7).long_jump()
    base = orc.seed_state(7)
    orc.lib().orc_long_jump(base.ctypes.data_as(C.POINTER(C.c_uint64)))
    u = r.random_real(3)
    want_u = [orc.lib().orc_random_real(base.ctypes.data_as(C.POINTER(C.c_uint64)), 0)
              for _ in range(3)]
    assert list(u) == want_u


def test_errors_are_reported_not_fatal():
    from mcstate_b200 import _lib

    L = _lib.lib()
    mid = C.c_int()
    rc = L.mcb_model_lookup(b"nope", C.byref(mid), None, None, None)
    assert rc == _lib.ERR_ARG and b"unknown model" in L.mcb_last_error()
    n = C.c_int(-1)
    rc = L.mcb_device_count(C.byref(n))
    assert rc in (_lib.OK, _lib.ERR_CUDA) and n.value >= 0
