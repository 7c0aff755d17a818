"""csrc/samplers.cuh: walk_fast -- the certified fast walk of the state-dependent binomial
draw.  It evaluates the CDF approximately (table-driven exponential for f_0, one fma per
factor) and accepts the k it finds only when the uniform is further than 2^-32 from both ends
of the slice; otherwise the exact code decides.  The engine's claim: an ACCEPTED k is the
exact walk's k.  On the GPU the bit-exact state tests cover it over ~10^9 draws; here the same
operation sequence, restated in C (gcc, fma, no contraction), runs beside the oracle's exact
inversion (orc_binomial) on the same uniforms over random (n, q)."""
import ctypes as C
import subprocess
import textwrap

import numpy as np

SRC = textwrap.dedent(r"""
    #include <math.h>
    #include <stdint.h>
    #include <string.h>
    static double exp2_tab[32], recip[64];
    void init(void) {
      for (int j = 0; j < 32; ++j) exp2_tab[j] = exp2((double)j / 32.0);
      for (int k = 1; k < 64; ++k) recip[k] = 1.0 / (double)k;
    }
    static const double c_fexp[7] = {46.16624130844683, 0x1.62e42fefa0000p-6,
                                     0x1.cf79abc9e3b3ap-45, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5};
    static double fast_exp_neg(double x) {
      const double magic = 6755399441055744.0;
      const double tm = fma(x, c_fexp[0], magic);
      int64_t bits; memcpy(&bits, &tm, 8);
      const int ki = (int)(int32_t)(bits & 0xffffffff);
      const double kd = tm - magic;
      double r = fma(-kd, c_fexp[1], x);
      r = fma(-kd, c_fexp[2], r);
      double p = fma(r, c_fexp[3], c_fexp[4]);
      p = fma(p, r, c_fexp[5]);
      p = fma(p, r, c_fexp[6]);
      p = fma(p, r, 1.0);
      p = fma(p, r, 1.0);
      const double v = exp2_tab[ki & 31] * p;
      int64_t vb; memcpy(&vb, &v, 8);
      vb += (int64_t)(ki >> 5) << 52;
      double out; memcpy(&out, &vb, 8);
      return out;
    }
    /* -1: the exact code must decide */
    int walk_fast(int n, double q, double r, double lq, double u) {
      const double nd = (double)n, nd1 = nd + 1.0;
      const double mode = q * nd1;
      const double mode_r = (mode + 6755399441055744.0) - 6755399441055744.0;
      const double f0 = fast_exp_neg(nd * lq);
      const double g = r * nd1;
      if (fabs(mode - mode_r) <= mode * 0x1p-30) return -1;
      double F = f0, f = f0;
      int k = 0;
      const int kmax = n < 63 ? n : 63;
      while (u >= F) {
        if (k >= kmax) return -1;
        ++k;
        f *= fma(g, recip[k], -r);
        F += f;
      }
      return ((F - u > 0x1p-32) & (u - (F - f) >= 0x1p-32)) ? k : -1;
    }
""")


def test_an_accepted_fast_walk_is_the_exact_walk(orc, tmp_path):
    src = tmp_path / "walk.c"
    src.write_text(SRC)
    so = tmp_path / "libwalk.so"
    subprocess.run(["gcc", "-O2", "-mfma", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(so),
                    str(src), "-lm"], check=True)
    W = C.CDLL(str(so))
    W.init()
    W.walk_fast.restype = C.c_int
    W.walk_fast.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double]
    L = orc.lib()
    rng = np.random.default_rng(11)
    state = orc.seed_state(5).copy()
    p_state = state.ctypes.data_as(orc._u64p)
    accepted = handed_back = 0
    for it in range(60000):
        n = int(rng.integers(1, 1 << 16)) if it % 3 else int(rng.integers(1, 40))
        mean = rng.uniform(0.01, 9.99)
        q = min(0.45 * rng.uniform(0.9, 1.0), mean / n)   # (q = 1/2 with odd n is a mode on an
        if n * q >= 10.0:                                  #  integer: the second test's case)
            continue
        r, lq = q / (1.0 - q), L.orc_log(1.0 - q)
        peek = state.copy()
        u = L.orc_random_real(peek.ctypes.data_as(orc._u64p), orc.PLUS)
        k_exact = L.orc_binomial(p_state, orc.PLUS, float(n), q)
        one_draw = np.array_equal(state, peek)          # the exact walk kept its first attempt
        k_fast = W.walk_fast(n, q, r, lq, u)
        if k_fast >= 0:
            accepted += 1
            assert one_draw, (n, q, u)
            assert k_fast == int(k_exact), (n, q, u, k_fast, k_exact)
        else:
            handed_back += 1
    assert accepted > 50000 and handed_back < 20       # the certificate hardly ever refuses


def test_near_integer_mode_is_handed_back(tmp_path):
    """q (n + 1) an integer: the exact chain's factor passes exactly one there, f may stall and
    the attempt be abandoned -- the fast walk must leave such a draw to the exact code."""
    src = tmp_path / "walk.c"
    src.write_text(SRC)
    so = tmp_path / "libwalk.so"
    subprocess.run(["gcc", "-O2", "-mfma", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(so),
                    str(src), "-lm"], check=True)
    W = C.CDLL(str(so))
    W.init()
    W.walk_fast.restype = C.c_int
    W.walk_fast.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double]
    import math
    for n, q in ((7, 0.25), (15, 0.125), (99, 0.03), (3, 0.5)):
        assert abs(q * (n + 1) - round(q * (n + 1))) < 1e-12
        assert W.walk_fast(n, q, q / (1 - q), math.log1p(-q), 0.4) == -1
