"""Pins the CPU oracle before anything is compared against it (CPU only).

Sources of the expected values:
  * the reference's own known-answer tests (file:line cited per test);
  * published generator vectors (splitmix64: the widely reproduced seed 1234567
    sequence; xoshiro256+ / xoshiro256**: state {1,2,3,4} sequences from the
    generators' reference test-suites);
  * relations the reference's test-suite asserts.
"""
import ctypes as C
import math

import numpy as np
import pytest

from oracle import r_semantics as R


# --------------------------------------------------------------------- RNG
def test_splitmix64_published_vector(orc):
    st = C.c_uint64(1234567)
    got = [orc.lib().orc_splitmix64(C.byref(st)) for _ in range(5)]
    assert got == [6457827717110365317, 3203168211198807973, 9817491932198370423,
                   4593380528125082431, 16408922859458223821]


def test_xoshiro256_published_vectors(orc):
    s = np.array([1, 2, 3, 4], dtype=np.uint64)
    plus = [int(x) for x in orc.next_words(s.copy(), 10, orc.PLUS)]
    assert plus == [5, 211106232532999, 211106635186183, 9223759065350669058,
                    9250833439874351877, 13862484359527728515, 2346507365006083650,
                    1168864526675804870, 34095955243042024, 3466914240207415127]
    ss = [int(x) for x in orc.next_words(s.copy(), 10, orc.STARSTAR)]
    assert ss == [11520, 0, 1509978240, 1215971899390074240, 1216172134540287360,
                  607988272756665600, 16172922978634559625, 8476171486693032832,
                  10595114339597558777, 2904607092377533576]


def test_random_real_is_53_bit_fraction(orc):
    s = np.array([1, 2, 3, 4], dtype=np.uint64)
    u = orc.lib().orc_random_real(s.ctypes.data_as(C.POINTER(C.c_uint64)), orc.PLUS)
    assert u == (5 >> 11) * 2.0 ** -53
    s = orc.seed_state(42)
    w = orc.next_words(s.copy(), 3)
    s2 = s.copy()
    us = [orc.lib().orc_random_real(s2.ctypes.data_as(C.POINTER(C.c_uint64)), orc.PLUS)
          for _ in range(3)]
    assert us == [(int(x) >> 11) * 2.0 ** -53 for x in w]
    assert all(0.0 <= x < 1.0 for x in us)


def test_jump_is_linear_and_commutes(orc):
    # jump/long_jump are GF(2)-linear maps of the state and powers of the same
    # transition matrix, so they commute and distribute over xor
    a, b = orc.seed_state(1), orc.seed_state(2)
    L = orc.lib()
    p = lambda x: x.ctypes.data_as(C.POINTER(C.c_uint64))
    ja, jb, jab = a.copy(), b.copy(), a ^ b
    L.orc_jump(p(ja)); L.orc_jump(p(jb)); L.orc_jump(p(jab))
    assert np.array_equal(ja ^ jb, jab)
    x, y = a.copy(), a.copy()
    L.orc_jump(p(x)); L.orc_long_jump(p(x))
    L.orc_long_jump(p(y)); L.orc_jump(p(y))
    assert np.array_equal(x, y)
    # a jump also commutes with stepping the generator
    x, y = a.copy(), a.copy()
    L.orc_next(p(x), 0); L.orc_jump(p(x))
    L.orc_jump(p(y)); L.orc_next(p(y), 0)
    assert np.array_equal(x, y)


def test_streams_prefix_stable_and_raw_seed(orc):
    # tests/testthat/test-pmcmc-parallel.R:40-60: 32 bytes per stream, prefix
    # stable in the number of streams, node 1 equals the input seed
    s5 = orc.rng_streams(1, 5)
    s9 = orc.rng_streams(1, 9)
    assert s5.shape == (5, 4) and s5.nbytes == 5 * 32
    assert np.array_equal(s5, s9[:5])
    assert np.array_equal(s5[0], orc.seed_state(1))
    # raw seed holding 3 streams: verbatim, remainder jumped from the last
    again = orc.rng_streams(s9[:3].ravel(), 9)
    assert np.array_equal(again, s9)
    d = orc.rng_distributed_state(1, 4, 3)
    assert np.array_equal(d[0], s9[:4])
    base = orc.seed_state(1)
    orc.lib().orc_long_jump(base.ctypes.data_as(C.POINTER(C.c_uint64)))
    assert np.array_equal(d[1, 0], base)
    d2 = orc.rng_distributed_state(1, 4, 2)
    assert np.array_equal(d[:2], d2)


# ------------------------------------------------------------ det exp / log
def _ulps(a, b):
    return np.abs(np.asarray(a).view(np.int64) - np.asarray(b).view(np.int64))


def test_det_exp_log_within_one_ulp_of_libm(orc):
    L = orc.lib()
    rng = np.random.default_rng(7)
    xs = np.concatenate([rng.uniform(-745, 709, 20000), rng.normal(0, 1, 20000),
                         rng.uniform(-1e-3, 1e-3, 5000)])
    got = np.array([L.orc_exp(x) for x in xs])
    assert _ulps(got, np.exp(xs)).max() <= 1
    ys = np.concatenate([np.exp(rng.uniform(-700, 700, 20000)), rng.uniform(0.5, 2, 20000),
                         rng.uniform(1e-320, 1e-300, 500)])
    got = np.array([L.orc_log(y) for y in ys])
    assert _ulps(got, np.log(ys)).max() <= 1


def test_det_exp_log_special_values(orc):
    L = orc.lib()
    assert L.orc_exp(0.0) == 1.0
    assert L.orc_exp(-800.0) == 0.0
    assert L.orc_exp(710.0) == math.inf
    assert math.isnan(L.orc_exp(math.nan))
    assert L.orc_exp(-math.inf) == 0.0
    assert L.orc_exp(-745.0) == 5e-324
    assert L.orc_log(1.0) == 0.0
    assert L.orc_log(0.0) == -math.inf
    assert math.isnan(L.orc_log(-1.0))
    assert L.orc_log(math.inf) == math.inf
    assert abs(L.orc_log(5e-324) - math.log(5e-324)) < 1e-12


# ---------------------------------------------------------------- samplers
def _draws(orc, fn, n, *args, seed=11):
    s = orc.seed_state(seed)
    p = s.ctypes.data_as(C.POINTER(C.c_uint64))
    return np.array([fn(p, orc.PLUS, *args) for _ in range(n)])


@pytest.mark.parametrize("n,p", [(10, 0.1), (1000, 0.002), (50, 0.5), (1000, 0.3),
                                 (400, 0.9), (30, 0.97), (100000, 0.4)])
def test_binomial_moments(orc, n, p):
    x = _draws(orc, orc.lib().orc_binomial, 40000, float(n), p)
    assert np.all(x == np.floor(x)) and x.min() >= 0 and x.max() <= n
    se = math.sqrt(n * p * (1 - p) / x.size)
    assert abs(x.mean() - n * p) < 5 * se
    assert abs(x.var() - n * p * (1 - p)) < 0.06 * n * p * (1 - p) + 1e-9


def test_binomial_edge_cases(orc):
    L = orc.lib()
    s = orc.seed_state(3)
    p = s.ctypes.data_as(C.POINTER(C.c_uint64))
    before = s.copy()
    assert L.orc_binomial(p, 0, 0.0, 0.5) == 0.0
    assert L.orc_binomial(p, 0, 10.0, 0.0) == 0.0
    assert L.orc_binomial(p, 0, 10.0, 1.0) == 10.0
    assert np.array_equal(s, before)  # no uniforms consumed on the trivial branches


@pytest.mark.parametrize("lam", [0.5, 4.0, 9.9, 10.0, 55.5, 1000.0])
def test_poisson_moments(orc, lam):
    x = _draws(orc, orc.lib().orc_poisson, 40000, lam)
    assert abs(x.mean() - lam) < 5 * math.sqrt(lam / x.size)
    assert abs(x.var() - lam) < 0.06 * lam + 0.05


def test_lfact_is_log_factorial(orc):
    """log(k!) without libm's lgamma (table below 16, Stirling's series from there): the
    Poisson sampler's acceptance test must give the same bits on the CPU and on the GPU."""
    L = orc.lib()
    for k in list(range(0, 400)) + [1000, 12345, 10 ** 6, 10 ** 9]:
        want = math.lgamma(k + 1.0)
        assert abs(L.orc_lfact(float(k)) - want) <= 1e-15 * max(1.0, want), k
    s = orc.seed_state(3)
    p = s.ctypes.data_as(C.POINTER(C.c_uint64))
    before = s.copy()
    assert L.orc_poisson(p, 0, 0.0) == 0.0
    assert np.array_equal(s, before)   # lambda == 0: no uniform consumed


def test_exponential_and_density(orc):
    x = _draws(orc, orc.lib().orc_exponential, 40000, 4.0)
    assert abs(x.mean() - 0.25) < 0.01
    L = orc.lib()
    assert L.orc_density_poisson_log(0.0, 0.0, 0.0) == 0.0
    got = L.orc_density_poisson_log(3.0, math.lgamma(4.0), 2.5)
    assert abs(got - (3 * math.log(2.5) - 2.5 - math.lgamma(4.0))) < 1e-14


# --------------------------------------------- weights: reference known answers
def test_scale_log_weights_reference_known_answers(orc):
    """tests/testthat/test-particle-filter.R:174-187, typed in unchanged."""
    inf, nan, e = math.inf, math.nan, math.e
    cases = [
        ([-inf, -inf], [nan, nan], -inf),
        ([-inf, 1], [0, 1], math.log(e / 2)),
        ([-inf, 1, 1], [0, 1, 1], math.log(e * 2 / 3)),
        ([nan, nan], [nan, nan], -inf),
        ([nan, 1], [0, 1], math.log(e / 2)),
    ]
    for lw, w_exp, avg_exp in cases:
        for w, avg in (orc.scale_log_weights_r(lw),
                       (lambda r: (r["weights"], r["average"]))(R.scale_log_weights(lw)),
                       orc.scale_log_weights(lw, orc.DUST_FP64),
                       orc.scale_log_weights(lw, orc.EXACT)):
            np.testing.assert_array_equal(np.isnan(w), np.isnan(w_exp))
            np.testing.assert_allclose(np.nan_to_num(w), np.nan_to_num(np.array(w_exp, float)),
                                       rtol=1e-15)
            if math.isinf(avg_exp):
                assert avg == avg_exp
            else:
                assert abs(avg - avg_exp) < 1e-14


def test_scale_log_weights_flavours_agree(orc):
    rng = np.random.default_rng(0)
    for n in (1, 2, 17, 1000, 65536):
        lw = rng.normal(-300, 40, n)
        lw[rng.uniform(size=n) < 0.05] = -np.inf
        lw[0] = -250.0
        w_r, a_r = orc.scale_log_weights_r(lw)
        w_d, a_d = orc.scale_log_weights(lw, orc.DUST_FP64)
        w_e, a_e = orc.scale_log_weights(lw, orc.EXACT)
        assert _ulps(w_r, w_d).max() <= 1  # libm exp vs deterministic exp
        assert np.array_equal(w_d, w_e)
        assert abs(a_r - a_d) < 1e-12 * abs(a_r)
        assert abs(a_d - a_e) < 1e-10 * abs(a_d)  # north_star tolerance (fp64, 1e-10 relative)


@pytest.mark.parametrize("n", [1 << 20, 1 << 24])
def test_scale_log_weights_flavours_agree_at_headline_sizes(orc, n):
    """The engine's weight sums (fixed point, 52 fractional bits, 128-bit accumulators:
    csrc/kernels.cuh) against dust's left-to-right fp64 sum at the sizes the bench runs
    (2^20) and beyond (2^24), on the cases where a fixed-point sum could lose something: a
    near-one-hot set with a dense band of weights in [2^-46, 2^-40] (each keeps 6+ bits)
    and one in [2^-60, 2^-50] (most round to zero -- together they are 2^-30 of the total).
    The log-mean-exp agrees within the north_star's 1e-10 relative everywhere; the chosen
    ancestors differ in at most one slot in a million and never by more than a few
    particles (PARITY_NOTES.md #14 quotes the measured figures)."""
    rng = np.random.default_rng(0)
    cases = {}
    lw = rng.normal(-300, 40, n); lw[rng.uniform(size=n) < 0.05] = -np.inf; lw[0] = -250.0
    cases["normal"] = lw
    lw = np.log(2.0) * rng.uniform(-46, -40, n); lw[n // 2] = 0.0
    cases["band_46_40"] = lw
    lw = np.log(2.0) * rng.uniform(-60, -50, n); lw[n // 2] = 0.0
    cases["band_60_50"] = lw
    cases["uniform"] = np.zeros(n)
    cases["mild"] = rng.normal(0, 3, n)
    for name, lw in cases.items():
        w_d, a_d = orc.scale_log_weights(lw, orc.DUST_FP64)
        w_e, a_e = orc.scale_log_weights(lw, orc.EXACT)
        assert np.array_equal(w_d, w_e), name
        assert abs(a_d - a_e) <= 1e-10 * max(abs(a_d), 1.0), (name, a_d, a_e)
        k_e = orc.resample_weight(w_e, 0.37, orc.EXACT)
        k_d = orc.resample_weight(w_d, 0.37, orc.DUST_FP64)
        assert np.mean(k_e != k_d) < 2e-6, name
        assert np.abs(k_e - k_d).max() <= 16, name


# ---------------------------------------------------------------- resampling
def test_particle_resample_r_matches_numpy_restatement(orc):
    rng = np.random.default_rng(1)
    for n in (1, 2, 7, 100, 4096):
        w = rng.uniform(size=n) ** 4
        u0 = rng.uniform(0, 1 / n)
        k_c = orc.particle_resample_r(w, u0)
        k_np = R.particle_resample(w, u0)
        assert np.array_equal(k_c, k_np)
        assert k_c.min() >= 1 and k_c.max() <= n + 1 and np.all(np.diff(k_c) >= 0)
    # column-wise for matrices (R/particle_filter.R:516-518)
    w = rng.uniform(size=(50, 3))
    k = R.particle_resample(w, [0.001, 0.01, 0.019])
    for j in range(3):
        assert np.array_equal(k[:, j], R.particle_resample(w[:, j], [0.001, 0.01, 0.019][j]))


def test_resample_conventions_differ_only_by_tie_rule(orc):
    # equal weights: R path (right-open findInterval) vs dust path (first j with cw_j >= uu)
    n = 8
    w = np.ones(n)
    # u0 = 0: every threshold sits exactly on a cumulative weight
    k_r = orc.particle_resample_r(w, 0.0) - 1
    k_d = orc.resample_weight(w, 0.0, orc.DUST_FP64)
    assert list(k_r) == list(range(n))
    assert list(k_d) == [0] + list(range(n - 1))  # ties go to the current particle
    # generic u: identical
    assert np.array_equal(orc.particle_resample_r(w, 0.3 / n) - 1,
                          orc.resample_weight(w, 0.3, orc.DUST_FP64))


@pytest.mark.parametrize("kind", ["uniform", "onehot", "zeros", "heavy", "tiny"])
def test_resample_exact_vs_fp64(orc, kind):
    rng = np.random.default_rng(5)
    for n in (1, 3, 64, 1000, 1 << 16):
        if kind == "uniform":
            w = np.ones(n)
        elif kind == "onehot":
            w = np.zeros(n); w[n // 3] = 1.0
        elif kind == "zeros":
            w = rng.uniform(size=n); w[rng.uniform(size=n) < 0.7] = 0.0; w[n - 1] = 1.0
        elif kind == "heavy":
            w = np.exp(rng.normal(0, 10, n)); w /= w.max()
        else:
            w = np.full(n, 1e-300); w[0] = 1.0
        for u in (0.0, 0.37, 1 - 2.0 ** -53):
            k_e = orc.resample_weight(w, u, orc.EXACT)
            k_d = orc.resample_weight(w, u, orc.DUST_FP64)
            assert k_e.min() >= 0 and k_e.max() <= n - 1 and np.all(np.diff(k_e) >= 0)
            if kind in ("uniform",) or u == 0.0:
                # thresholds sit on (or within rounding of) cumulative weights:
                # the two accumulations may break ties differently, by one slot
                assert np.abs(k_e - k_d).max() <= 1
            else:
                assert np.mean(k_e != k_d) < 1e-3
            # every chosen particle has positive weight (u > 0)
            if u > 0:
                assert np.all(w[k_e] > 0)


def test_resample_exact_offspring_counts_unbiased(orc):
    rng = np.random.default_rng(9)
    n = 2000
    w = rng.uniform(size=n) ** 3
    w /= w.max()
    counts = np.zeros(n)
    reps = 400
    for r in range(reps):
        counts += np.bincount(orc.resample_weight(w, rng.uniform(), orc.EXACT), minlength=n)
    expect = reps * n * w / w.sum()
    # systematic resampling: each count is floor or ceil of its expectation per draw
    assert np.all(np.abs(counts - expect) <= reps)
    assert abs(np.corrcoef(counts, expect)[0, 1] - 1) < 1e-3


def test_history_single_matches_numpy_restatement(orc):
    rng = np.random.default_rng(2)
    ny, npart, nt = 3, 11, 6
    value = rng.normal(size=(ny, npart, nt))
    order = rng.integers(1, npart + 1, size=(npart, nt))
    order[:, 0] = np.arange(1, npart + 1)
    want = R.history_single(value, order)
    v_c = np.ascontiguousarray(value.transpose(2, 1, 0))
    o_c = np.ascontiguousarray(order.T.astype(np.int32))
    out = np.zeros_like(v_c)
    orc.lib().orc_history_single(v_c.ctypes.data_as(C.POINTER(C.c_double)),
                                 o_c.ctypes.data_as(C.POINTER(C.c_int)), ny, npart, nt,
                                 out.ctypes.data_as(C.POINTER(C.c_double)))
    assert np.array_equal(out.transpose(2, 1, 0), want)


def test_early_exit_rules():
    """R/particle_filter_state.R:463-474; cases of tests/testthat/test-particle-filter.R:1220-1282."""
    assert R.particle_filter_early_exit([-np.inf, 1.0], [-1e9, -1e9])
    assert R.particle_filter_early_exit(-10.0, -5.0)
    assert not R.particle_filter_early_exit(-10.0, -15.0)
    assert R.particle_filter_early_exit([-10.0, -10.0], -15.0)       # summed
    assert not R.particle_filter_early_exit([-10.0, -10.0], -25.0)
    assert not R.particle_filter_early_exit([-10.0, -10.0], [-15.0, -5.0])  # needs all
    assert R.particle_filter_early_exit([-10.0, -10.0], [-5.0, -5.0])


# ------------------------------------------------------- particle_filter_data
def test_particle_filter_data_reference_layouts():
    """tests/testthat/test-particle-filter-data.R:69-110."""
    res = R.particle_filter_data(np.arange(1, 12), {"data": np.arange(0, 1.01, 0.1)}, 10, 0)
    assert list(res["model_time_start"]) == list(range(0, 11))
    assert list(res["model_time_end"]) == list(range(1, 12))
    assert list(res["time_start"]) == [10 * i for i in range(0, 11)]
    assert list(res["time_end"]) == [10 * i for i in range(1, 12)]
    assert np.array_equal(res["model_times"], np.stack([np.arange(0, 11), np.arange(1, 12)], 1))
    assert np.array_equal(res["times"], res["model_times"] * 10)
    # offset initial time (:93-110)
    res = R.particle_filter_data(np.arange(11, 21), {"a": np.arange(10)}, 4, 1)
    assert list(res["model_time_start"]) == [1] + list(range(11, 20))
    assert list(res["time_start"]) == [4] + [4 * i for i in range(11, 20)]
    assert list(res["time_end"]) == [4 * i for i in range(11, 21)]
    with pytest.raises(ValueError, match="'initial_time' must be <= 1"):
        R.particle_filter_data(np.arange(1, 5), {"a": np.arange(4)}, 2, 2)
    with pytest.raises(ValueError, match="must be an integer"):
        R.particle_filter_data(np.arange(1, 5), {"a": np.arange(4)}, 2, 0.5)


def test_particle_filter_data_nested_layout():
    """tests/testthat/test-particle-filter-data.R:124-155."""
    rng = np.random.default_rng(3)
    data = rng.uniform(size=22)
    day = np.tile(np.arange(1, 12), 2)
    group = np.repeat(["a", "b"], 11)
    perm = rng.permutation(22)
    res = R.particle_filter_data(day[perm], {"data": data[perm]}, 10, 0, population=group[perm])
    assert list(res["model_time_start"]) == list(range(0, 11)) * 2
    assert list(res["time_end"]) == [10 * i for i in range(1, 12)] * 2
    assert list(res["population"]) == ["a"] * 11 + ["b"] * 11
    assert np.array_equal(res["columns"]["data"], data)
    assert res["populations"] == ["a", "b"]
    assert np.array_equal(res["times"], np.stack([np.arange(0, 11), np.arange(1, 12)], 1) * 10)
