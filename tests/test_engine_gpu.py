"""GPU parity: the CUDA engine, called through the C ABI (via mcstate_b200.dust),
against the CPU oracle on identical seeds and inputs.  Everything is compared
bit for bit: states, RNG words, log-weights, ancestor indices, log-likelihoods
(the reference's own CPU == GPU test demands identical results,
tests/testthat/test-particle-filter.R:1480-1504).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dust():
    from mcstate_b200 import dust as d

    n = d.dust_example("sir").gpu_info()
    if not n["has_cuda"]:
        pytest.fail("no CUDA device visible: the engine has no CPU fallback")
    return d


def _data(sir_data, n=None, rate=4):
    day = sir_data["day"][:n]
    inc = sir_data["incidence"][:n]
    return day * rate, inc


def _set_data(model, t, x):
    from mcstate_b200 import dust as d

    model.set_data(d.dust_data({"time_end": t, "incidence": x}), False)


def _rng_words(model):
    return np.frombuffer(model.rng_state(), dtype=np.uint64).reshape(-1, 4)


@pytest.mark.parametrize("n", [1, 31, 1000, (1 << 16) + 3])
def test_stream_derivation_matches_serial_jumps(dust, orc, n):
    m = dust.dust_example("sir").new({}, 0, n, seed=1)
    assert np.array_equal(_rng_words(m), orc.rng_streams(1, n + 1))
    # raw seed with several streams: verbatim, remainder jumped from the last one
    raw = orc.rng_streams(42, 3)
    m = dust.dust_example("sir").new({}, 0, 50, seed=raw.tobytes())
    assert np.array_equal(_rng_words(m), orc.rng_streams(raw.ravel(), 51))


@pytest.mark.parametrize("model,mid", [("sir", 0), ("sirs", 1)])
@pytest.mark.parametrize("alg,scr", [("xoshiro256plus", 0), ("xoshiro256starstar", 1)])
def test_run_state_and_rng_bit_exact(dust, orc, model, mid, alg, scr):
    n = 5000
    pars = {"beta": 0.3, "gamma": 0.12, "I0": 25}
    g = dust.dust_example(model, alg)
    m = g.new(pars, 0, n, seed=7)
    vec = g._pars_vector(pars, False)
    o = orc.OracleModel(pars=vec, n_particles=n, seed=7, model=mid, scrambler=scr, n_threads=8)
    assert np.array_equal(m.state(), o.state())
    for t_end in (1, 4, 5, 400):
        got = m.run(t_end)
        want = o.run(t_end)
        assert np.array_equal(got, want), t_end
        assert np.array_equal(_rng_words(m), o.rng_state()), t_end
        assert m.time() == t_end


def test_run_covers_both_binomial_algorithms(dust, orc):
    # large populations push n*q above 10 (BTRS); tiny ones stay on inversion
    n = 4096
    for pars in ({"S0": 1e6, "I0": 5000.0, "beta": 0.5}, {"S0": 50.0, "I0": 2.0}):
        g = dust.dust_example("sir")
        m = g.new(pars, 0, n, seed=3)
        o = orc.OracleModel(pars=g._pars_vector(pars, False), n_particles=n, seed=3, n_threads=8)
        assert np.array_equal(m.run(80), o.run(80))
        assert np.array_equal(_rng_words(m), o.rng_state())


def test_deterministic_mode(dust, orc):
    g = dust.dust_example("sir")
    m = g.new({}, 0, 3, seed=1, deterministic=True)
    o = orc.OracleModel(n_particles=3, seed=1, deterministic=True)
    assert np.array_equal(m.run(100), o.run(100))
    assert np.array_equal(_rng_words(m), orc.rng_streams(1, 4))  # no draws consumed


def test_index_update_state_reorder_simulate(dust, orc):
    n = 300
    g = dust.dust_example("sir")
    m = g.new({}, 0, n, seed=2)
    o = orc.OracleModel(n_particles=n, seed=2)
    m.set_index([5, 1])
    got = m.run(8)
    want = o.run(8)
    assert got.shape == (2, n) and np.array_equal(got, want[[4, 0]])
    assert np.array_equal(m.state([2, 3]), want[[1, 2]])
    m.set_index(None)
    # reorder: 1-based indices
    kappa = np.random.default_rng(0).integers(1, n + 1, n)
    m.reorder(kappa)
    o.reorder(kappa - 1)
    assert np.array_equal(m.state(), o.state())
    with pytest.raises(Exception, match="index"):
        m.reorder(np.zeros(n, dtype=int))
    # update_state: vector recycled, matrix, pars + time reset; RNG untouched
    before = _rng_words(m).copy()
    m.update_state(state=[10, 20, 30, 40, 50])
    assert np.array_equal(m.state(), np.tile(np.array([[10, 20, 30, 40, 50.0]]).T, (1, n)))
    mat = np.arange(5 * n, dtype=float).reshape(5, n)
    m.update_state(state=mat)
    assert np.array_equal(m.state(), mat)
    m.update_state(pars={"I0": 33}, time=0)
    assert m.time() == 0
    assert np.array_equal(m.state()[:, 0], [1000, 33, 0, 0, 0])
    assert np.array_equal(_rng_words(m), before)
    # simulate
    o.update_state(pars=g._pars_vector({"I0": 33}, False), time=0)
    sim = m.simulate([4, 8, 20])
    assert sim.shape == (5, n, 3)
    for k, t in enumerate((4, 8, 20)):
        assert np.array_equal(sim[:, :, k], o.run(t))


def test_state_of_chosen_particles(dust):
    m = dust.dust_example("sir").new({"I0": 30}, 0, 5000, seed=2)
    m.run(40)
    full = m.state()
    pick = [1, 5000, 77, 77]
    assert np.array_equal(m.state_particles(pick), full[:, [0, 4999, 76, 76]])
    with pytest.raises(IndexError):
        m.state_particles([5001])


def test_rng_pointer_matches_a_model_s_streams(dust, orc):
    """dust_rng_pointer$new(seed, n)$state() (tests/testthat/test-particle-filter.R:1316)."""
    raw = dust.dust_rng_pointer(9, 6).state()
    assert np.array_equal(np.frombuffer(raw, dtype=np.uint64).reshape(-1, 4), orc.rng_streams(9, 6))
    m = dust.dust_example("sir").new({}, 0, 5, seed=9)
    assert m.rng_state() == raw


def test_rng_state_roundtrip(dust, orc):
    g = dust.dust_example("sir")
    m = g.new({}, 0, 10, seed=1)
    s = m.rng_state()
    assert len(s) == 11 * 32 and m.rng_state(first_only=True) == s[:32]
    m.run(12)
    m2 = g.new({}, 0, 10, seed=99)
    m2.set_rng_state(s)
    assert np.array_equal(m2.run(12), m.state())
    with pytest.raises(Exception, match="rng_state"):
        m2.set_rng_state(s[:64])


def test_compare_data_bit_exact(dust, orc, sir_data):
    n = 2000
    t, x = _data(sir_data, 10)
    g = dust.dust_example("sir")
    m = g.new({}, 0, n, seed=5)
    o = orc.OracleModel(n_particles=n, seed=5)
    _set_data(m, t, x)
    o.set_data(t, x)
    assert m.compare_data() is None  # no datum at time 0
    m.run(4); o.run(4)
    got, want = m.compare_data(), o.compare_data()
    assert np.array_equal(got, want)
    assert np.array_equal(_rng_words(m), o.rng_state())  # one exponential draw each


WEIGHT_KINDS = ["uniform", "onehot", "sparse", "heavy", "tiny", "real", "onehot_last", "band"]


def _weights(kind, n, rng):
    if kind == "uniform":
        return np.ones(n)
    if kind == "onehot":
        w = np.zeros(n); w[(2 * n) // 3] = 1.0
        return w
    if kind == "onehot_last":       # every slot walks to the last tile
        w = np.zeros(n); w[n - 1] = 1.0
        return w
    if kind == "sparse":
        w = rng.uniform(size=n); w[rng.uniform(size=n) < 0.9] = 0.0; w[n - 1] = 1.0
        return w
    if kind == "heavy":
        w = np.exp(rng.normal(0, 10, n))
        return w / w.max()
    if kind == "tiny":
        w = np.full(n, 1e-300); w[0] = 1.0
        return w
    if kind == "band":              # near one-hot over a dense band of weights in [2^-46, 2^-40]
        w = 2.0 ** rng.uniform(-46, -40, n); w[n // 2] = 1.0
        return w
    lw = -0.5 * rng.normal(0, 3, n) ** 2
    return np.exp(lw - lw.max())


@pytest.mark.parametrize("kind", ["uniform", "onehot_last", "heavy", "band", "real"])
@pytest.mark.parametrize("n", [1 << 20, (1 << 21) + 5])
def test_resample_indices_bit_exact_headline_sizes(dust, orc, kind, n):
    """The ancestors kernel at config C2's size (512 weight tiles, prefix staged in shared
    memory) and beyond it (1025 tiles: prefix searched in global memory)."""
    rng = np.random.default_rng(n + 1)
    w = _weights(kind, n, rng)
    g = dust.dust_example("volatility")
    m = g.new({}, 0, n, seed=11)
    m.update_state(state=np.arange(n, dtype=float).reshape(1, n))
    u_words = _rng_words(m)[-1].copy()
    got = m.resample(w)
    u = _first_uniform(u_words)
    want = orc.resample_weight(w, u, orc.EXACT) + 1
    assert np.array_equal(got, want)
    assert np.array_equal(m.state()[0], (want - 1).astype(float))


@pytest.mark.parametrize("kind", ["cluster40", "two_in_a_thread", "heavy_then_light", "steps"])
def test_resample_heavy_particles_and_block_written_ranges(dust, orc, kind):
    """The ancestors kernel's special paths (csrc/kernels.cuh: k_ancestors): a particle that
    owns >= 2048 output slots is written by its whole block (a list of at most 32 per tile),
    the others by their thread, which must skip the block-written ranges:
      cluster40         40 equal heavy particles side by side in one tile (the list overflows:
                        eight of them fall back to their threads)
      two_in_a_thread   two heavy particles among the 8 of one thread, light ones between
      heavy_then_light  one particle with half the slots, then thousands with one or none
      steps             weights falling by 2^-3 per particle: ranges of every size"""
    n = 1 << 18
    rng = np.random.default_rng(17)
    w = np.zeros(n)
    if kind == "cluster40":
        w[5000:5040] = 1.0
    elif kind == "two_in_a_thread":
        w[:] = 1e-7 * rng.uniform(size=n)
        w[16001] = 1.0; w[16006] = 0.75; w[16003] = 2e-6
    elif kind == "heavy_then_light":
        w[:] = rng.uniform(size=n) / n
        w[777] = 1.0
    else:
        w[1000:1040] = 2.0 ** (-3.0 * np.arange(40))
    g = dust.dust_example("volatility")
    m = g.new({}, 0, n, seed=3)
    m.update_state(state=np.arange(n, dtype=float).reshape(1, n))
    for _ in range(2):
        u = _first_uniform(_rng_words(m)[-1].copy())
        got = m.resample(w)
        want = orc.resample_weight(w, u, orc.EXACT) + 1
        assert np.array_equal(got, want), kind
    # a multi-set model: sets that do not start on a multiple of 32, one of them one-hot
    n, n_pars = 3001, 5
    gs = dust.dust_example("sir")
    ms = gs.new([{"beta": 0.2 + 0.01 * p} for p in range(n_pars)], 0, n, seed=4, pars_multi=True)
    ws = rng.uniform(size=(n, n_pars)) ** 6
    ws[:, 2] = 0.0; ws[1234, 2] = 1.0
    ws /= ws.max(axis=0)
    u_words = _rng_words(ms)[-1].copy()
    got = ms.resample(ws)
    st = [int(x) for x in u_words]
    for p in range(n_pars):     # one uniform per set from the shared filter stream, in set order
        r = (st[0] + st[3]) & 0xFFFFFFFFFFFFFFFF
        u = float(r >> 11) * 2.0 ** -53
        want = orc.resample_weight(np.ascontiguousarray(ws[:, p]), u, orc.EXACT) + 1
        assert np.array_equal(got[:, p], want), p
        t = (st[1] << 17) & 0xFFFFFFFFFFFFFFFF
        st[2] ^= st[0]; st[3] ^= st[1]; st[1] ^= st[2]; st[0] ^= st[3]; st[2] ^= t
        st[3] = ((st[3] << 45) | (st[3] >> 19)) & 0xFFFFFFFFFFFFFFFF


def _first_uniform(words):
    """first uniform of a stream given as four words (xoshiro256+, the default scrambler)"""
    s = [int(x) for x in words]
    return float(((s[0] + s[3]) & 0xFFFFFFFFFFFFFFFF) >> 11) * 2.0 ** -53


@pytest.mark.parametrize("kind", WEIGHT_KINDS)
@pytest.mark.parametrize("n", [1, 7, 1024, 1025, 5000, (1 << 17) + 11])
def test_resample_indices_bit_exact(dust, orc, kind, n):
    rng = np.random.default_rng(n)
    if kind in ("onehot_last", "band"):
        w = _weights(kind, n, rng)
    elif kind == "uniform":
        w = np.ones(n)
    elif kind == "onehot":
        w = np.zeros(n); w[(2 * n) // 3] = 1.0
    elif kind == "sparse":
        w = rng.uniform(size=n); w[rng.uniform(size=n) < 0.9] = 0.0; w[n - 1] = 1.0
    elif kind == "heavy":
        w = np.exp(rng.normal(0, 10, n)); w /= w.max()
    elif kind == "tiny":
        w = np.full(n, 1e-300); w[0] = 1.0
    else:
        lw = -0.5 * rng.normal(0, 3, n) ** 2
        w = np.exp(lw - lw.max())
    g = dust.dust_example("sir")
    m = g.new({}, 0, n, seed=11)
    o = orc.OracleModel(n_particles=n, seed=11)
    m.run(4); o.run(4)
    for _ in range(2):  # consecutive draws from the filter stream
        got = m.resample(w)
        want = o.resample(w) + 1
        assert np.array_equal(got, want)
        assert np.array_equal(m.state(), o.state())
    assert np.array_equal(_rng_words(m)[-1], o.rng_state()[-1])


def _filter_pair(dust, orc, n, sir_data, n_rows=None, pars=None, seed=1, n_pars=0, **kw):
    t, x = _data(sir_data, n_rows)
    g = dust.dust_example("sir")
    if n_pars:
        plist = pars or [{"beta": 0.2 + 0.02 * p} for p in range(n_pars)]
        m = g.new(plist, 0, n, seed=seed, pars_multi=True)
        o = orc.OracleModel(pars=g._pars_vector(plist, True), n_particles=n, n_pars=n_pars,
                            seed=seed, n_threads=8)
    else:
        m = g.new(pars or {}, 0, n, seed=seed)
        o = orc.OracleModel(pars=g._pars_vector(pars or {}, False), n_particles=n, seed=seed,
                            n_threads=8)
    return m, o, t, x


@pytest.mark.parametrize("n", [1, 100, 1000, 4097, 1 << 16])
def test_filter_log_likelihood_state_rng_bit_exact(dust, orc, sir_data, n):
    m, o, t, x = _filter_pair(dust, orc, n, sir_data)
    _set_data(m, t, x)
    o.set_data(t, x)
    got = m.filter()
    want = o.filter()
    assert np.array_equal(got["log_likelihood"], want["log_likelihood"])
    assert np.isfinite(got["log_likelihood"]).all()
    assert np.array_equal(m.state(), o.state())
    assert np.array_equal(_rng_words(m), o.rng_state())
    assert m.time() == 400
    # the model object (and its RNG) persists: a second evaluation continues the streams
    m.update_state(pars={}, time=0)
    o.update_state(pars=o_default(orc), time=0)
    got2, want2 = m.filter(), o.filter()
    assert np.array_equal(got2["log_likelihood"], want2["log_likelihood"])
    assert got2["log_likelihood"][0] != got["log_likelihood"][0]


def o_default(orc):
    return orc.default_pars(orc.SIR)


def test_filter_trajectories_and_snapshots(dust, orc, sir_data):
    n = 333
    m, o, t, x = _filter_pair(dust, orc, n, sir_data, n_rows=30, seed=4)
    _set_data(m, t, x)
    o.set_data(t, x)
    m.set_index([1, 2, 3])
    snaps = [t[4], t[17], t[29]]
    got = m.filter(save_trajectories=True, time_snapshot=snaps)
    want = o.filter(save_trajectories=True, index=[0, 1, 2], snapshot_times=snaps)
    assert got["trajectories"].shape == (3, n, 31)
    assert np.array_equal(got["trajectories"], want["trajectories"])
    assert np.array_equal(got["snapshots"], want["snapshots"])
    assert np.array_equal(got["log_likelihood"], want["log_likelihood"])
    tr = got["trajectories"]
    assert np.all(np.diff(tr[0], axis=1) <= 0) and np.all(np.diff(tr[2], axis=1) >= 0)
    assert np.array_equal(got["snapshots"][:, :, 2], m.state())


def test_filter_history_of_chosen_particles_traced_on_device(dust, orc, sir_data):
    """the history stays in HBM and single lineages are traced there
    (history_single with an index, R/particle_filter.R:616-646; what pMCMC keeps,
    R/pmcmc_state.R:32-51).  Same values as the full traced array."""
    n = 1000
    t, x = _data(sir_data, 40)
    g = dust.dust_example("sir")
    full = g.new({}, 0, n, seed=8)
    lazy = g.new({}, 0, n, seed=8)
    for m in (full, lazy):
        _set_data(m, t, x)
        m.set_index([1, 2, 5])
    want = full.filter(save_trajectories=True)
    got = lazy.filter(save_trajectories="device")
    assert got["trajectories"] is None
    assert np.array_equal(got["log_likelihood"], want["log_likelihood"])
    pick = np.array([1, 500, 1000, 37, 37])
    h = lazy.filter_history(pick)
    assert h.shape == (3, 5, 41)
    assert np.array_equal(h, want["trajectories"][:, pick - 1, :])
    assert np.array_equal(lazy.filter_history(np.arange(1, n + 1)), want["trajectories"])
    with pytest.raises(IndexError):
        lazy.filter_history([n + 1])
    # a run without history invalidates it
    lazy.update_state(pars={}, time=0)
    lazy.filter()
    with pytest.raises(Exception, match="save_trajectories"):
        lazy.filter_history([1])

    # several parameter sets: indices are per set
    sets = [{"beta": 0.2}, {"beta": 0.3}, {"beta": 0.25}]
    full = g.new(sets, 0, 200, seed=3, pars_multi=True)
    lazy = g.new(sets, 0, 200, seed=3, pars_multi=True)
    for m in (full, lazy):
        m.set_data(dust.dust_data({"time_end": t, "incidence": x}), True)
        m.set_index([2, 5])
    want = full.filter(save_trajectories=True)["trajectories"]     # [2, 200, 3, 41]
    lazy.filter(save_trajectories="device")
    ip = np.array([[1, 200, 7], [50, 50, 50]])
    h = lazy.filter_history(ip)
    assert h.shape == (2, 2, 3, 41)
    for j in range(3):
        assert np.array_equal(h[:, :, j, :], want[:, ip[:, j] - 1, j, :])


def test_filter_history_after_early_stop(dust, sir_data):
    n = 256
    t, x = _data(sir_data, 30)
    x = x.copy()
    x[:6] = 0.0          # no infections, no noise: the first positive datum is impossible
    pars = {"exp_noise": float("inf"), "I0": 0}
    g = dust.dust_example("sir")
    a, b = g.new(pars, 0, n, seed=2), g.new(pars, 0, n, seed=2)
    for m in (a, b):
        _set_data(m, t, x)
    want = a.filter(save_trajectories=True)
    b.filter(save_trajectories="device")
    assert want["log_likelihood"][0] == -np.inf
    got = b.filter_history([3, 200])
    assert np.array_equal(got, want["trajectories"][:, [2, 199], :], equal_nan=True)
    assert np.isnan(got[:, :, -1]).all()


def test_filter_incremental_equals_one_shot(dust, orc, sir_data):
    # tests/testthat/test-particle-filter.R:604-658
    n = 500
    t, x = _data(sir_data, 40)
    g = dust.dust_example("sir")
    a = g.new({}, 0, n, seed=9)
    b = g.new({}, 0, n, seed=9)
    _set_data(a, t, x); _set_data(b, t, x)
    one = a.filter()["log_likelihood"]
    parts = b.filter(t[9])["log_likelihood"] + b.filter(t[10])["log_likelihood"] + \
        b.filter(t[39])["log_likelihood"]
    o = orc.OracleModel(n_particles=n, seed=9)
    o.set_data(t, x)
    p1 = o.filter(t[9])["log_likelihood"]; p2 = o.filter(t[10])["log_likelihood"]
    p3 = o.filter(t[39])["log_likelihood"]
    assert np.array_equal(parts, p1 + p2 + p3)
    assert abs(one[0] - parts[0]) < 1e-9 * abs(one[0])
    assert np.array_equal(a.state(), b.state())


def test_filter_missing_data_rows_equal_dropped_rows(dust, orc, sir_data):
    # tests/testthat/test-particle-filter.R:1698-1776
    n = 400
    t, x = _data(sir_data, 50)
    x_na = x.copy()
    x_na[::2] = np.nan
    keep = ~np.isnan(x_na)
    g = dust.dust_example("sir")
    a = g.new({}, 0, n, seed=3)
    b = g.new({}, 0, n, seed=3)
    _set_data(a, t, x_na)
    _set_data(b, t[keep], x_na[keep])
    la, lb = a.filter()["log_likelihood"], b.filter()["log_likelihood"]
    assert np.array_equal(la, lb)
    assert np.array_equal(a.state(), b.state())
    o = orc.OracleModel(n_particles=n, seed=3)
    o.set_data(t, x_na)
    assert np.array_equal(la, o.filter()["log_likelihood"])
    # trailing / leading NA rows
    x2 = x.copy(); x2[:3] = np.nan; x2[-2:] = np.nan
    c = g.new({}, 0, n, seed=3)
    _set_data(c, t, x2)
    o = orc.OracleModel(n_particles=n, seed=3)
    o.set_data(t, x2)
    assert np.array_equal(c.filter(save_trajectories=True)["trajectories"],
                          o.filter(save_trajectories=True)["trajectories"])
    assert np.array_equal(c.state(), o.state()) and c.time() == o.time()


def test_filter_impossible_data_stops(dust, orc, sir_data):
    # tests/testthat/test-particle-filter.R:71-93: -Inf, later history NA.
    # No infections and no noise => modelled incidence is exactly 0: a zero
    # observation has log density 0, a positive one -Inf for every particle.
    n = 64
    t, x = _data(sir_data, 20)
    x = x.copy()
    x[:6] = 0.0
    pars = {"exp_noise": float("inf"), "I0": 0}
    g = dust.dust_example("sir")
    m = g.new(pars, 0, n, seed=1)
    o = orc.OracleModel(pars=g._pars_vector(pars, False), n_particles=n, seed=1)
    _set_data(m, t, x); o.set_data(t, x)
    got = m.filter(save_trajectories=True)
    want = o.filter(save_trajectories=True)
    assert got["log_likelihood"][0] == -np.inf and want["log_likelihood"][0] == -np.inf
    stop = want["n_intervals"]
    assert stop == 7
    assert np.array_equal(np.isnan(got["trajectories"]), np.isnan(want["trajectories"]))
    assert np.array_equal(np.nan_to_num(got["trajectories"]), np.nan_to_num(want["trajectories"]))
    assert np.isnan(got["trajectories"][:, :, stop + 1:]).all()
    assert not np.isnan(got["trajectories"][:, :, :stop + 1]).any()
    assert m.time() == o.time() == t[6]
    assert np.array_equal(m.state(), o.state())
    assert np.array_equal(_rng_words(m), o.rng_state())


def test_filter_min_log_likelihood_early_exit(dust, orc, sir_data):
    # R/particle_filter_state.R:463-474
    n = 256
    m, o, t, x = _filter_pair(dust, orc, n, sir_data, seed=2)
    _set_data(m, t, x); o.set_data(t, x)
    got = m.filter(min_log_likelihood=-100.0)
    want = o.filter(min_log_likelihood=-100.0)
    assert got["log_likelihood"][0] == -np.inf == want["log_likelihood"][0]
    assert m.time() == o.time() and m.time() < 400
    assert np.array_equal(m.state(), o.state())
    assert np.array_equal(_rng_words(m), o.rng_state())
    # a bound that never binds changes nothing
    m2, o2, _, _ = _filter_pair(dust, orc, n, sir_data, seed=2)
    _set_data(m2, t, x); o2.set_data(t, x)
    assert np.array_equal(m2.filter(min_log_likelihood=-1e9)["log_likelihood"],
                          o2.filter()["log_likelihood"])


@pytest.mark.parametrize("n,n_pars", [(100, 2), (1500, 3), (37, 8)])
def test_multi_parameter_filter(dust, orc, sir_data, n, n_pars):
    m, o, t, x = _filter_pair(dust, orc, n, sir_data, n_rows=40, n_pars=n_pars, seed=6)
    from mcstate_b200 import dust as d
    # shared data: one stream of observations for every parameter set
    m.set_data(d.dust_data({"time_end": t, "incidence": x}), True)
    o.set_data(t, x, shared=True)
    got = m.filter(save_trajectories=True)
    want = o.filter(save_trajectories=True)
    assert got["log_likelihood"].shape == (n_pars,)
    assert np.array_equal(got["log_likelihood"], want["log_likelihood"])
    tr = want["trajectories"].reshape(5, n_pars, n, -1).transpose(0, 2, 1, 3)
    assert got["trajectories"].shape == (5, n, n_pars, 41)
    assert np.array_equal(got["trajectories"], tr)
    assert np.array_equal(_rng_words(m), o.rng_state())
    st = o.state().reshape(5, n_pars, n).transpose(0, 2, 1)
    assert np.array_equal(m.state(), st)


@pytest.mark.parametrize("n,n_pars", [(2049, 37), (4097, 5), (300, 64), (70001, 3)])
def test_scheduling_stress_many_sets_odd_tiles_repeated(dust, orc, sir_data, n, n_pars):
    """In place of compute-sanitizer's racecheck (closed on this pool):
    the cross-block protocols of the per-interval kernels -- K3's per-set counter and
    early-exit rule, the ancestor slots K2 empties and K3 fills from other blocks' tiles,
    the warp-level atomic maximum of K1 -- are run with many small parameter sets and tile
    counts that do not divide the set (a last tile of 1 particle, sets that straddle
    warps), several times over, while a second handle keeps the GPU busy on another stream
    so that blocks are scheduled differently each time.  Every repetition must give the
    oracle's bits."""
    from mcstate_b200 import dust as d
    import torch

    m0, o, t, x = _filter_pair(dust, orc, n, sir_data, n_rows=25, n_pars=n_pars, seed=9)
    del m0
    o.set_data(t, x, shared=True)
    rng0 = orc.rng_streams(9, n * n_pars + 1).tobytes()
    want = o.filter()["log_likelihood"]
    want_state = o.state().reshape(5, n_pars, n).transpose(0, 2, 1)
    want_rng = o.rng_state()
    g = dust.dust_example("sir")
    noise = g.new({}, 0, 1 << 18, seed=3)                     # the other stream's work
    noise.set_persistent(False)
    _set_data(noise, t, x)
    plist = [{"beta": 0.2 + 0.02 * p} for p in range(n_pars)]
    for rep in range(6):
        m = g.new(plist, 0, n, seed=9, pars_multi=True)
        m.set_data(d.dust_data({"time_end": t, "incidence": x}), True)
        if rep % 2:
            noise.update_state(pars={}, time=0)
            noise.filter_enqueue()
        got = m.filter()["log_likelihood"]
        if rep % 2:
            noise.filter_collect()
        assert np.array_equal(got, want), rep
        assert np.array_equal(m.state(), want_state), rep
        assert np.array_equal(_rng_words(m), want_rng), rep
        # the same handle again from the start (the second use of a row range replays a graph)
        for _ in range(2):
            m.update_state(pars=plist, time=0)
            m.set_rng_state(rng0)
            assert np.array_equal(m.filter()["log_likelihood"], want), rep
    torch.cuda.synchronize()


def test_nested_equals_independent_filters_on_split_streams(dust, orc, sir_data):
    # tests/testthat/test-particle-filter.R:1285-1353: streams 1..n -> population 1,
    # n+1..2n -> population 2; the filter stream is shared and drawn in set order
    n = 200
    t, x = _data(sir_data, 30)
    x2 = np.roll(x, 3)
    from mcstate_b200 import dust as d
    g = d.dust_example("sir")
    plist = [{"beta": 0.2}, {"beta": 0.25}]
    m = g.new(plist, 0, n, seed=1, pars_multi=True)
    rows = [(int(t[i]), [{"incidence": x[i]}, {"incidence": x2[i]}]) for i in range(len(t))]
    m.set_data(rows, False)
    ll = m.filter()["log_likelihood"]
    o = orc.OracleModel(pars=g._pars_vector(plist, True), n_particles=n, n_pars=2, seed=1)
    o.set_data(t, np.stack([x, x2], axis=1))
    assert np.array_equal(ll, o.filter()["log_likelihood"])
    # each population alone, seeded with its slice of the streams
    streams = orc.rng_streams(1, 2 * n + 1)
    for p, (pars, xs) in enumerate(zip(plist, (x, x2))):
        s = np.concatenate([streams[p * n:(p + 1) * n], streams[-1:]])
        single = g.new(pars, 0, n, seed=s.tobytes())
        single.set_data(d.dust_data({"time_end": t, "incidence": xs}), False)
        # the first interval's likelihood depends on the particle streams only (the
        # shared filter stream is first used after it), so it must agree exactly
        l1 = single.filter(t[0])["log_likelihood"][0]
        m1 = g.new(plist, 0, n, seed=1, pars_multi=True)
        m1.set_data(rows, False)
        assert l1 == m1.filter(t[0])["log_likelihood"][p]


def test_full_size_filter_2_20(dust, orc, sir_data):
    """BASELINE config C2: 2^20 particles, bit-exact against the oracle, plus
    size-independent properties."""
    n = 1 << 20
    m, o, t, x = _filter_pair(dust, orc, n, sir_data, seed=1)
    _set_data(m, t, x); o.set_data(t, x)
    got = m.filter()["log_likelihood"]
    want = o.filter()["log_likelihood"]
    assert np.array_equal(got, want)
    st = m.state()
    assert np.array_equal(st, o.state())
    assert np.array_equal(st[0] + st[1] + st[2], np.full(n, 1010.0))  # population conserved
    assert np.all(st >= 0)
    # identical seeds => identical results; continuing the object continues the RNG
    m2 = dust.dust_example("sir").new({}, 0, n, seed=1)
    _set_data(m2, t, x)
    assert np.array_equal(m2.filter()["log_likelihood"], got)
    m2.update_state(pars={}, time=0)
    again = m2.filter()["log_likelihood"]
    assert again[0] != got[0] and abs(again[0] - got[0]) < 0.5


def test_filter_graph_follows_freq_and_time(dust, orc, sir_data):
    """The captured filter graph bakes in the step the call starts at and the incidence-reset
    schedule of every launch: a later call with another `freq`, or from another step, must
    not replay it (two-launch path: more particles than one wave holds)."""
    n = 160_000
    m, o, t, x = _filter_pair(dust, orc, n, sir_data, n_rows=6, pars={"freq": 4.0})
    m.set_persistent(False)
    _set_data(m, t, x)
    o.set_data(t, x)
    g = dust.dust_example("sir")
    assert np.array_equal(m.filter()["log_likelihood"], o.filter()["log_likelihood"])
    # three passes: a row range is launched directly the first time, captured as a graph the
    # second and replayed from the third on
    for _ in range(3):
        for pars, time in (({"freq": 4.0}, 0), ({"freq": 2.0}, 0), ({"freq": 2.0}, 3),
                           ({"freq": 4.0}, 1), ({"freq": 3.0}, 0)):
            m.update_state(pars=pars, time=time)
            o.update_state(pars=g._pars_vector(pars, False), time=time)
            assert np.array_equal(m.filter()["log_likelihood"],
                                  o.filter()["log_likelihood"]), (pars, time)
            assert np.array_equal(m.state(), o.state()), (pars, time)


def test_independent_sets_a_dead_one_does_not_stop_the_others(dust, sir_data):
    """Per-set seeds = independent filters batched in one handle (what smc2 holds,
    R/smc2.R:72-78,89-108): a parameter set whose weights all vanish gets -Inf, the others
    carry on exactly as the separate filters would.  Set 1 cannot explain a positive count
    (beta = 0, no observation noise)."""
    from mcstate_b200.dust import dust_rng_distributed_state
    from oracle_generator import OracleGenerator

    n = 512
    t, x = _data(sir_data, 12)
    data = dust.dust_data({"time_end": t, "incidence": x})
    plist = [{"beta": 0.2}, {"beta": 0.0, "exp_noise": float("inf")}, {"beta": 0.3}]
    seeds = b"".join(dust_rng_distributed_state(7, 1, len(plist)))
    m = dust.dust_example("sir").new(plist, 0, n, seed=seeds, pars_multi=True, seed_per_set=True,
                                     gpu_config=0)
    o = OracleGenerator("sir").new(plist, 0, n, seed=seeds, pars_multi=True, seed_per_set=True)
    m.set_data(data, True)
    o.set_data(data, True)
    got, want = m.filter()["log_likelihood"], o.filter()["log_likelihood"]
    assert got[1] == -np.inf and want[1] == -np.inf
    assert np.isfinite(got[[0, 2]]).all()
    assert np.array_equal(got, want)
    assert np.array_equal(m.state()[:, :, [0, 2]], o.state()[:, :, [0, 2]])
    # one data row at a time (how smc2 drives the handle): the dead set's step is -Inf every
    # time, the others' steps add up to the one-shot value
    m.update_state(pars=plist, time=0)
    total = np.zeros(3)
    for row in range(len(t)):
        total += m.filter(int(t[row]))["log_likelihood"]
    assert total[1] == -np.inf and np.isfinite(total[[0, 2]]).all()


def test_thousand_replicate_filters_agree_in_distribution(dust, orc, sir_data):
    """BASELINE's criterion for runs whose random draws differ: mean and variance of the
    log-likelihood over 1000 replicate filters.  The engine runs its 1000 replicates as
    one batched handle with a seed per set; the oracle runs 1000 separate filters on
    other seeds, so nothing is shared but the algorithm."""
    from mcstate_b200.dust import dust_rng_distributed_state

    n, reps = 256, 1000
    t, x = _data(sir_data)
    data = dust.dust_data({"time_end": t, "incidence": x})
    seeds = dust_rng_distributed_state(101, 1, reps)
    m = dust.dust_example("sir").new([{}] * reps, 0, n, seed=b"".join(seeds), pars_multi=True,
                                     seed_per_set=True, gpu_config=0)
    m.set_data(data, True)
    got = m.filter()["log_likelihood"]
    assert got.shape == (reps,) and np.isfinite(got).all()
    want = np.empty(reps)
    for k in range(reps):
        o = orc.OracleModel(n_particles=n, seed=5000 + k, n_threads=8)
        o.set_data(t, x)
        want[k] = o.filter()["log_likelihood"][0]
    se = np.sqrt(got.var(ddof=1) / reps + want.var(ddof=1) / reps)
    assert abs(got.mean() - want.mean()) < 4 * se, (got.mean(), want.mean(), se)
    ratio = got.var(ddof=1) / want.var(ddof=1)
    assert 0.75 < ratio < 1.33, ratio           # sd of a variance ratio at 1000 reps ~ 0.07
    # and both sit below the large-N value by about half their variance (log of an
    # unbiased likelihood estimate)
    big = dust.dust_example("sir").new({}, 0, 1 << 18, seed=3)
    big.set_data(data, False)
    ll_big = big.filter()["log_likelihood"][0]
    assert abs(got.mean() + 0.5 * got.var(ddof=1) - ll_big) < 0.15


def test_large_filter_2_24_properties(dust, sir_data):
    """Sixteen times config C2 in one handle (2^24 particles, 8192 weight tiles, 1.9 GB of
    state and streams): too long for the CPU oracle, so size-independent properties --
    determinism, conservation, and agreement of the likelihood estimate with the 2^20 run
    (its standard error shrinks with 1/sqrt(N))."""
    n = 1 << 24
    t, x = _data(sir_data, 40)
    g = dust.dust_example("sir")
    m = g.new({}, 0, n, seed=1)
    _set_data(m, t, x)
    ll = m.filter()["log_likelihood"][0]
    st = m.state()
    assert st.shape == (5, n) and np.all(st >= 0)
    assert np.array_equal(st[0] + st[1] + st[2], np.full(n, 1010.0))
    small = g.new({}, 0, 1 << 20, seed=2)
    _set_data(small, t, x)
    assert abs(ll - small.filter()["log_likelihood"][0]) < 0.02
    # the first 2^20 streams are those of a 2^20 model: same first step for those slots
    a, b = g.new({}, 0, n, seed=1), g.new({}, 0, 1 << 20, seed=1)
    assert np.array_equal(a.run(4)[:, :1 << 20], b.run(4))
    del a, b
    m.update_state(pars={}, time=0)
    m2 = g.new({}, 0, n, seed=1)
    _set_data(m2, t, x)
    assert m2.filter()["log_likelihood"][0] == ll


# ---- the lean (integer-state, table-driven) path of K1 against the general one -------------
# Every case is bit-exact against the oracle, which only knows the general algorithm.
@pytest.mark.parametrize("pars,label", [
    ({"S0": 1e6, "I0": 5000.0, "beta": 0.5}, "population beyond the memo tables"),
    ({"S0": 70000.0, "I0": 30.0}, "n_tot just above 65536 entries"),
    ({"beta": 1.5, "gamma": 0.4}, "S q >= 10 mid-interval: lean steps then BTRS"),
    ({"beta": 0.0}, "probability exactly zero: no draw, no uniform"),
    ({"gamma": 0.0}, "recovery probability zero"),
    ({"freq": 1.0, "gamma": 0.6931471805599453}, "p_IR = 1/2: r = 1, stalled walks"),
    ({"freq": 1.0, "gamma": 50.0, "beta": 30.0}, "probabilities above one half (flipped)"),
    ({"freq": 3.0}, "freq 3: incidence resets off the data grid"),
    ({"S0": 3.0, "I0": 1.0}, "tiny population"),
    ({"S0": 0.0, "I0": 40.0}, "no susceptibles"),
])
def test_lean_path_special_cases(dust, orc, pars, label):
    n = 3000
    g = dust.dust_example("sir")
    m = g.new(pars, 0, n, seed=9)
    o = orc.OracleModel(pars=g._pars_vector(pars, False), n_particles=n, seed=9, n_threads=8)
    for t_end in (3, 4, 50, 51, 130):   # 46 and 79 steps in one launch: beyond one 32-bit mask
        assert np.array_equal(m.run(t_end), o.run(t_end)), (label, t_end)
        assert np.array_equal(_rng_words(m), o.rng_state()), (label, t_end)


def test_random_parameter_sweep(dust, orc, sir_data):
    """Eighty random parameter sets (tiny to large populations, probabilities from ~0 to
    beyond one half, 1-8 steps per day, both models, both scramblers, both execution
    paths) through a short filter: every one bit-exact against the oracle.  Exercises the
    borders of the lean path -- counts below the written-out walk steps, walk tables that
    end, BTRS hand-overs mid-interval, flipped probabilities, stalled walks."""
    rs = np.random.RandomState(20261020)
    t_day, x = sir_data["day"][:25], sir_data["incidence"][:25]
    for case in range(80):
        model = "sir" if case % 4 else "sirs"
        scr = int(rs.randint(2))
        freq = float(rs.choice([1, 2, 3, 4, 8]))
        s0 = float(rs.choice([3, 11, 12, 13, 40, 1000, 5000, 70000]))
        pars = {"beta": float(rs.choice([0.0, 0.05, 0.2, 0.9, 3.0, 40.0])),
                "gamma": float(rs.choice([0.0, 0.02, 0.1, 0.6931471805599453, 2.5])),
                "S0": s0, "I0": float(rs.randint(0, max(2, int(min(s0, 200))))),
                "freq": freq}
        if model == "sirs":
            pars["alpha"] = float(rs.choice([0.0, 0.01, 0.3, 5.0]))
        n = int(rs.choice([1, 33, 200, 1000]))
        seed = int(rs.randint(1, 1 << 30))
        t = (t_day * freq).astype(np.uint64)
        g = dust.dust_example(model, ["xoshiro256plus", "xoshiro256starstar"][scr])
        m = g.new(pars, 0, n, seed=seed)
        m.set_persistent(bool(case & 1))
        o = orc.OracleModel(pars=g._pars_vector(pars, False), n_particles=n, seed=seed,
                            model=0 if model == "sir" else 1, scrambler=scr, n_threads=4)
        _set_data(m, t, x)
        o.set_data(t, x)
        got, want = m.filter()["log_likelihood"], o.filter()["log_likelihood"]
        assert np.array_equal(got, want, equal_nan=True), (case, model, pars, n, got, want)
        assert np.array_equal(m.state(), o.state()), (case, model, pars)
        assert np.array_equal(_rng_words(m), o.rng_state()), (case, model, pars)


def test_lean_path_leaves_non_integer_states_to_the_general_code(dust, orc):
    n = 500
    g = dust.dust_example("sir")
    m = g.new({}, 0, n, seed=4)
    o = orc.OracleModel(n_particles=n, seed=4, n_threads=4)
    st = np.tile(np.array([[999.5, 10.0, 0.5, 0.0, 0.0]]).T, (1, n))
    st[:, ::3] = np.array([[1000.0, 10.0, 0.0, 0.0, 0.0]]).T     # a third stays integral
    st[:, 1::7] = np.array([[900.0, 10.0, 0.0, 0.0, 0.0]]).T     # N differs from the tables'
    st[0, 5] = -0.0                                              # the sign of a zero survives
    m.update_state(state=st)
    o.update_state(state=st, set_initial_state=False)
    for t_end in (1, 8, 60):
        assert np.array_equal(m.run(t_end), o.run(t_end))
        assert np.array_equal(_rng_words(m), o.rng_state())
    assert np.array_equal(np.signbit(m.state()), np.signbit(o.state()))


@pytest.mark.parametrize("alg,scr", [("xoshiro256plus", 0), ("xoshiro256starstar", 1)])
def test_sirs_filter_bit_exact(dust, orc, sir_data, alg, scr):
    """The second model family (three draws per step, waning immunity) through the filter."""
    n = 2048
    t, x = _data(sir_data, 60)
    pars = {"alpha": 0.05, "beta": 0.3, "gamma": 0.1, "I0": 15}
    g = dust.dust_example("sirs", alg)
    m = g.new(pars, 0, n, seed=21)
    _set_data(m, t, x)
    o = orc.OracleModel(pars=g._pars_vector(pars, False), n_particles=n, seed=21, model=1,
                        scrambler=scr, n_threads=8)
    o.set_data(t, x)
    assert np.array_equal(m.filter()["log_likelihood"], o.filter()["log_likelihood"])
    assert np.array_equal(m.state(), o.state())
    assert np.array_equal(_rng_words(m), o.rng_state())


# ---- the single-launch filter (csrc/persistent.cuh) == the two-launch path ---------
def _pair(dust, name, pars, n, seed, **kw):
    g = dust.dust_example(name)
    a, b = g.new(pars, 0, n, seed=seed, **kw), g.new(pars, 0, n, seed=seed, **kw)
    b.set_persistent(False)
    return a, b


def _same(a, b):
    assert np.array_equal(a.state(), b.state())
    assert np.array_equal(_rng_words(a), _rng_words(b))
    assert a.time() == b.time()


@pytest.mark.parametrize("n", [1, 100, 255, 256, 257, 4097, 1 << 16, 150000])
def test_persistent_filter_equals_two_launch(dust, orc, sir_data, n):
    t, x = _data(sir_data)
    a, b = _pair(dust, "sir", {}, n, 3)
    for m in (a, b):
        _set_data(m, t, x)
    la, lb = a.filter()["log_likelihood"], b.filter()["log_likelihood"]
    assert np.array_equal(la, lb)
    _same(a, b)
    # uses one launch for the rows (plus uniforms and the stream update)
    assert a.launch_count() < b.launch_count() - 150
    if n <= 4097:
        o = orc.OracleModel(n_particles=n, seed=3, n_threads=8)
        o.set_data(t, x)
        assert np.array_equal(la, o.filter()["log_likelihood"])
        assert np.array_equal(a.state(), o.state())
        assert np.array_equal(_rng_words(a), o.rng_state())
    # a second evaluation on the same handles (the generator carries over)
    for m in (a, b):
        m.update_state(pars={"beta": 0.25}, time=0)
    assert np.array_equal(a.filter()["log_likelihood"], b.filter()["log_likelihood"])
    _same(a, b)


def test_persistent_filter_trajectories(dust, sir_data):
    n = 5000
    t, x = _data(sir_data, 50)
    xm = x.copy()
    xm[[3, 4, 20]] = np.nan
    for data_x, kw in ((x, {}), (xm, {}), (x, {"min_log_likelihood": -70.0})):
        a, b = _pair(dust, "sir", {}, n, 9)
        for m in (a, b):
            _set_data(m, t, data_x)
            m.set_index([1, 3, 5])
        ra = a.filter(save_trajectories=True, **kw)
        rb = b.filter(save_trajectories=True, **kw)
        assert np.array_equal(ra["log_likelihood"], rb["log_likelihood"])
        assert np.array_equal(ra["trajectories"], rb["trajectories"], equal_nan=True)
        _same(a, b)
    # the lazily traced lineages too
    a, b = _pair(dust, "sir", {}, n, 10)
    for m in (a, b):
        _set_data(m, t, x)
        m.filter(save_trajectories="device")
    pick = [1, 77, n]
    assert np.array_equal(a.filter_history(pick), b.filter_history(pick))


def test_persistent_filter_special_rows(dust, sir_data):
    n = 3000
    t, x = _data(sir_data, 40)
    # missing data rows (no weights, no resampling there), incremental calls, early exit
    xm = x.copy()
    xm[[0, 5, 6, 17, 39]] = np.nan
    a, b = _pair(dust, "sir", {}, n, 5)
    for m in (a, b):
        _set_data(m, t, xm)
    for t_end in (t[9], t[10], t[39]):
        assert np.array_equal(a.filter(t_end)["log_likelihood"], b.filter(t_end)["log_likelihood"])
        _same(a, b)
    # min_log_likelihood: stops at the same row with the same state and streams
    a, b = _pair(dust, "sir", {}, n, 6)
    for m in (a, b):
        _set_data(m, t, x)
    ra, rb = a.filter(min_log_likelihood=-60.0), b.filter(min_log_likelihood=-60.0)
    assert ra["log_likelihood"][0] == -np.inf and rb["log_likelihood"][0] == -np.inf
    _same(a, b)
    assert a.time() < t[-1]
    # impossible data: every weight is zero
    x0 = x.copy()
    x0[:6] = 0.0
    a, b = _pair(dust, "sir", {"exp_noise": float("inf"), "I0": 0}, 500, 1)
    for m in (a, b):
        _set_data(m, t, x0)
    assert a.filter()["log_likelihood"][0] == -np.inf == b.filter()["log_likelihood"][0]
    _same(a, b)
    # deterministic mode, the second scrambler, the other models
    a, b = _pair(dust, "sir", {}, 777, 2, deterministic=True)
    for m in (a, b):
        _set_data(m, t, x)
    assert np.array_equal(a.filter()["log_likelihood"], b.filter()["log_likelihood"])
    _same(a, b)
    g = dust.dust_example("sirs", "xoshiro256starstar")
    a, b = g.new({"alpha": 0.05}, 0, 2000, seed=4), g.new({"alpha": 0.05}, 0, 2000, seed=4)
    b.set_persistent(False)
    for m in (a, b):
        _set_data(m, t, x)
    assert np.array_equal(a.filter()["log_likelihood"], b.filter()["log_likelihood"])
    _same(a, b)
    g = dust.dust_example("volatility")
    a, b = g.new({}, 0, 5000, seed=4), g.new({}, 0, 5000, seed=4)
    b.set_persistent(False)
    rs = np.random.RandomState(2)
    for m in (a, b):
        m.set_data(dust.dust_data({"time_end": np.arange(1, 51, dtype=np.uint64),
                                   "observed": rs.normal(size=50)}), False)
        rs = np.random.RandomState(2)
    assert np.array_equal(a.filter()["log_likelihood"], b.filter()["log_likelihood"])
    _same(a, b)
