// Per-particle samplers and densities for the compartmental models.
//
// Algorithms are those of the reference's native dependency dust [SURVEY.md
// A.2]: binomial by CDF inversion (n q < 10) or Hoermann's BTRS transformed
// rejection (draw_binomial), Poisson by Knuth's product (lambda < 10) or PTRS
// (draw_poisson), normal by Box-Muller, exponential by inversion.  The number of
// uniforms a draw consumes is data dependent, which is why every particle owns
// a stream; the order of draws and of floating-point operations is part of the
// contract (bit-exact against the oracle).
#pragma once
#include "detmath.cuh"
#include "rng.cuh"

namespace mcb {

template <int SCR>
__device__ __forceinline__ double draw_exponential(Xoshiro &r, double rate) {
  return -det_log(xo_real<SCR>(r)) / rate;
}

// a / b for a divisor the parameter set fixes, correctly rounded WITHOUT the division
// sequence -- the bits of `a / b`: y = RN(1/b) (prepared once per parameter set by
// reciprocal_for), q0 = RN(a y), then two residual corrections q <- RN(q + RN(a - b q) y).
// The residuals are exact up to a rounding that cannot matter, the first correction makes
// q faithful, and Markstein's theorem gives RN(a/b) from the second -- for every b whose
// significand is not all ones, with a, b and the quotient away from the ends of the
// exponent range.  reciprocal_for returns 0 where that is not so, and the caller divides.
__device__ __forceinline__ double reciprocal_for(double b) {
  const uint64_t bits = double_to_bits(b);
  const int e = (int)((bits >> 52) & 0x7ff);
  const bool ok = !(bits >> 63) && e > 1023 - 200 && e < 1023 + 200 &&
                  (bits & 0xfffffffffffffULL) != 0xfffffffffffffULL;
  return ok ? 1.0 / b : 0.0;
}
__device__ __forceinline__ double div_by_fixed(double a, double b, double y) {
  if (y == 0.0 || !(fabs(a) > 1e-200 && fabs(a) < 1e200)) return a / b;
  double q = a * y;
  double r = fma(-q, b, a);
  q = fma(r, y, q);
  r = fma(-q, b, a);
  return fma(r, y, q);
}
// the exponential draw with the rate's reciprocal prepared: same bits as draw_exponential
template <int SCR>
__device__ __forceinline__ double draw_exponential_fixed(Xoshiro &r, double rate, double inv) {
  return div_by_fixed(-det_log(xo_real<SCR>(r)), rate, inv);
}

// Box-Muller, first variate only; a uniform at or below machine epsilon is drawn
// again (both of the pair) so log never sees it.  sqrt is IEEE (correctly rounded).
template <int SCR>
__device__ __forceinline__ double draw_normal(Xoshiro &r, double mean, double sd) {
  double u1, u2;
  do {
    u1 = xo_real<SCR>(r);
    u2 = xo_real<SCR>(r);
  } while (u1 <= 2.220446049250313e-16);
  const double z = __dsqrt_rn(-2.0 * det_log(u1)) * det_cos2pi(u2);
  return z * sd + mean;
}

// dnorm(x, mu, sd, log = TRUE); log_sd = det_log(sd), prepared per parameter set
__device__ __forceinline__ double density_normal_log(double x, double mu, double sd,
                                                     double log_sd) {
  const double dx = x - mu;
  return -(dx * dx) / (2.0 * sd * sd) - 0.918938533204672741780329736406 - log_sd;
}

// x^n by squaring, multiplying in the low bits first (not pow())
__device__ __forceinline__ double pow_int(double x, int n) {
  double acc = 1.0;
  while (n != 0) {
    if (n & 1) acc *= x;
    n >>= 1;
    if (n != 0) x *= x;
  }
  return acc;
}

// {(double)k, RN(1/k)} for k = 0..kInvSmall-1 (entry 0 unused), filled by the host
// with IEEE division: the walk reads its divisor and reciprocal with one uniform load
constexpr int kInvSmall = 64;
__constant__ double2 c_inv_small[kInvSmall];
__constant__ double c_stirling_tail[10] = {
    0.0810614667953272,  0.0413406959554092,  0.0276779256849983, 0.02079067210376509,
    0.0166446911898211,  0.0138761288230707,  0.0118967099458917, 0.0104112652619720,
    0.00925546218271273, 0.00833056343336287};

// a / k for a small integer k, correctly rounded WITHOUT the generic division
// sequence: y = RN(1/k) from the table, q0 = RN(a y), then two residual
// corrections q <- RN(q + RN(a - k q) y).  The residuals are exact (fma), the
// first correction makes q faithful, and Markstein's theorem then gives
// RN(a/k) from the second -- so the bits equal those of `a / k`.  Needs a away
// from overflow/underflow; the caller checks that once per draw.
__device__ __forceinline__ double div_small_int(double a, double kd, double y) {
  double q = a * y;
  double r = fma(-q, kd, a);
  q = fma(r, y, q);
  r = fma(-q, kd, a);
  return fma(r, y, q);
}

// CDF walk from k = 0 given r = q/(1-q) and f0 = (1-q)^n: stop at the first k
// with u < f_k, where f_k = f_{k-1} (g/k - r).  An attempt is abandoned (fresh
// uniform) when f stops changing or k runs past n.
// Every lane of a warp is at the same k in the same iteration, so the table
// entry is a uniform constant-bank load.  The first loop covers k up to
// min(n, kInvSmall - 1) with the table division and no inner branch (k <= n
// holds by construction there); the second is the general form.
template <int SCR>
__device__ __forceinline__ double inversion_walk(Xoshiro &rng, int n, double r, double f0) {
  const double g = r * (double)(n + 1);
  const bool safe = g > 1e-280 && g < 1e280;
  const int k_fast = safe ? min(n, kInvSmall - 1) : 0;
  for (;;) {
    double u = xo_real<SCR>(rng);
    double f = f0;
    int k = 0;
    bool ok = true;
    while (u >= f && k < k_fast) {
      u -= f;
      ++k;
      const double2 e = c_inv_small[k];
      const double fn = f * (div_small_int(g, e.x, e.y) - r);
      if (fn == f) { ok = false; break; }
      f = fn;
    }
    if (ok) {
      while (u >= f) {
        u -= f;
        ++k;
        const double fn = f * (g / (double)k - r);
        if (fn == f || k > n) { ok = false; break; }
        f = fn;
      }
    }
    if (ok) return (double)k;
  }
}

template <int SCR>
__device__ __forceinline__ double binomial_inversion(Xoshiro &rng, int n, double q) {
  const double qq = 1.0 - q;
  return inversion_walk<SCR>(rng, n, q / qq, pow_int(qq, n));
}

// ---------------------------------------------------------------------------
// The lean path of the run kernel (models.cuh): the SAME draws as the functions
// above -- same bits -- computed another way.
//
// Certified fast walk.  The inversion walk above is a chain of dependent,
// individually rounded operations (integer power, correctly rounded divisions,
// u -= f): exact to reproduce, slow to run.  But its OUTCOME is just "the k
// whose slice [F_{k-1}, F_k) of the CDF holds u".  So the lean path evaluates
// the CDF approximately with cheap operations -- f_0 = exp(n log(1-q)) from a
// table-driven exponential, f_k = f_{k-1} (g (1/k) - r) with one fma per factor
// -- and accepts the k it finds only when u is further than kWalkEps from both
// ends of the slice, a margin far above the worst-case difference between the
// approximate and the exact chain (bound below).  Anything closer, or anything
// the exact walk treats specially (a factor that may round to one -> a stalled
// f; a walk beyond the table), is handed back to the exact code with the
// generator rewound.  The result is therefore ALWAYS the exact walk's; only
// the cost differs (a draw is handed back about once in 10^9).
//
// Error bound (n q < 10, q <= 1/2, k <= 63, n < 2^16; ulp = 2^-52):
//   exact chain   u_k = u - F*_{k-1} + d,  |d| <= k 2^-54 (each u -= f rounds a
//                 value below one), so it walks on at k iff u >= F*_k - d;
//   f_0           n log(1-q) carries <= 14.5 * 2 ulp absolute, the exponential
//                 <= 2^-47 relative, pow_int's own roundings <= 32 * 2^-53:
//                 together <= 2^-45 relative;
//   factors       g (1/k) - r against RN(RN(g/k) - r): <= 2 ulp of g/k, i.e.
//                 <= 2 ulp (n+1)/(n+1-k) relative to the factor, <= 2^-45 for the
//                 k <= min(n, 63) walked; 63 of them compound to <= 2^-39;
//   sums          F_k <= 1 + 2^-30, adding <= 64 terms rounds <= 2^-47.
// Total |F_k (approximate) - F*_k (exact)| < 2^-38; kWalkEps = 2^-32 leaves a
// factor 64.
// ---------------------------------------------------------------------------
__device__ __forceinline__ long long as_ll(double x) { return __double_as_longlong(x); }

// One step back of the xoshiro256 state walk (the map is linear and invertible):
// s1 = A ^ (s1 << 17) with A = s1' ^ s2' unrolls to four shifted terms.
__device__ __forceinline__ void xo_rewind(Xoshiro &r) {
  const uint64_t s3m = (r.s3 >> 45) | (r.s3 << 19);   // rotr 45
  const uint64_t a = r.s1 ^ r.s2;
  const uint64_t s1 = a ^ (a << 17) ^ (a << 34) ^ (a << 51);
  const uint64_t s0 = r.s0 ^ s3m;
  const uint64_t s2m = r.s2 ^ (s1 << 17);
  r.s0 = s0;
  r.s1 = s1;
  r.s2 = s2m ^ s0;
  r.s3 = s3m ^ s1;
}

__device__ __forceinline__ double2 ldg_f64x2(const double *p) {
  double2 e;
  asm("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(e.x), "=d"(e.y) : "l"(p));
  return e;
}
// one 32-byte load (LDG.256): four consecutive table entries (an evict-last hint on these
// loads measured the same: profiles/ab_r02_v3.txt)
__device__ __forceinline__ double4 ldg_f64x4(const double *p) {
  double4 e;
  asm("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];"
      : "=d"(e.x), "=d"(e.y), "=d"(e.z), "=d"(e.w) : "l"(p));
  return e;
}

// 2^(j/32), j = 0..31, filled by the host.  Read straight from global memory: its two
// lines never leave L1, and a shared-memory copy costs an address kept or rebuilt at
// every step (measured 0.9 us per launch slower: profiles/ab_r02_v6.txt)
constexpr int kExp2Tab = 32;
__device__ double g_exp2_tab[kExp2Tab];
// 1/k as constant-bank operands (entry 0 unused); approximate values are all the fast
// walk needs
__constant__ double c_recip[kInvSmall];
__constant__ double c_fexp[7] = {
    46.16624130844683,           // 32 / ln 2
    0x1.62e42fefa0000p-6,        // ln 2 / 32, leading 37 bits (k * this is exact for |k| < 2^16)
    0x1.cf79abc9e3b3ap-45,       // ln 2 / 32, the rest
    1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5};

// exp(x) for x in [-700, 0], relative error below 2^-47: k = rint(32 x / ln 2),
// r = x - k ln2/32 (|r| <= 0.0109), exp(x) = 2^(k >> 5) 2^((k & 31)/32) (1 + r + ... + r^5/120).
// Not reproducible to the last bit across implementations and not meant to be: it
// only proposes what the certificate of the fast walk then accepts or rejects.
__device__ __forceinline__ double fast_exp_neg(double x) {
  const double magic = 6755399441055744.0;   // 1.5 * 2^52: adding it rounds to an integer
  const double tm = fma(x, c_fexp[0], magic);
  const int ki = __double2loint(tm);
  const double kd = tm - magic;
  double r = fma(-kd, c_fexp[1], x);
  r = fma(-kd, c_fexp[2], r);
  double p = fma(r, c_fexp[3], c_fexp[4]);
  p = fma(p, r, c_fexp[5]);
  p = fma(p, r, c_fexp[6]);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  const double v = __ldg(g_exp2_tab + (ki & (kExp2Tab - 1))) * p;
  return __hiloint2double(__double2hiint(v) + ((ki >> 5) << 20), __double2loint(v));
}

constexpr double kWalkEps = 0x1p-32;

// Binomial(n, q) by the certified fast walk: n >= 1, q in (0, 1/2] with n q < 10,
// r = q/(1-q), lq = log(1-q) (both from the set's table).  Draws ONE uniform.  Returns
// the exact walk's k, or -1 when the exact code must decide (the caller rewinds the
// generator by that one uniform).
template <int SCR>
__device__ __forceinline__ int walk_fast(Xoshiro &rng, int n, double q, double r, double lq) {
  const double nd = (double)n, nd1 = nd + 1.0;
  // the factor g/k - r passes one at k = q (n + 1): if that is (nearly) an integer the
  // exact chain may meet f_k == f_{k-1} there and abandon the attempt -- its call
  const double mode = q * nd1;
  const double mode_r = (mode + 6755399441055744.0) - 6755399441055744.0;
  const double f0 = fast_exp_neg(nd * lq);
  const double g = r * nd1;
  const double u = xo_real<SCR>(rng);
  if (fabs(mode - mode_r) <= mode * 0x1p-30) return -1;
  double F = f0, f = f0;   // F = F_k, the CDF up to and including k
  int k;
  if (n >= 8) {
    // the usual case: the first eight steps written out, each one compare, one fma for
    // the factor (1/k straight from the constant bank), one product, one sum
#define MCB_FAST_STEP(K)                               \
    if (!(u >= F)) { k = (K) - 1; goto done; }         \
    f *= fma(g, c_recip[K], -r);                       \
    F += f;
    MCB_FAST_STEP(1) MCB_FAST_STEP(2) MCB_FAST_STEP(3) MCB_FAST_STEP(4)
    MCB_FAST_STEP(5) MCB_FAST_STEP(6) MCB_FAST_STEP(7) MCB_FAST_STEP(8)
#undef MCB_FAST_STEP
    k = 8;
    const int kmax = min(n, kInvSmall - 1);
    while (u >= F) {
      if (k >= kmax) return -1;
      ++k;
      f *= fma(g, c_recip[k], -r);
      F += f;
    }
  } else {
    k = 0;
    while (u >= F) {
      if (k >= n) return -1;
      ++k;
      f *= fma(g, c_recip[k], -r);
      F += f;
    }
  }
done:
  // F_{k-1} is F - f to within a rounding, far inside the margin (k = 0: the slice starts
  // at zero, and a u below the margin is simply handed back)
  return ((F - u > kWalkEps) & (u - (F - f) >= kWalkEps)) ? k : -1;
}

// ---------------------------------------------------------------------------
// Walk tables: a draw whose probability is fixed by the parameter set
// (recovery, waning) has a CDF walk that depends on the count n alone, so the
// whole sequence f_0 .. f_{kWalkK-1} of every count is prepared once per
// parameter set (k_prepare_sets) by the very operations of inversion_walk --
// same bits -- and a draw is then one uniform, 32-byte loads and a
// subtract-and-compare chain: no division, no product.
//   row n: f_k >= 0 while the walk may stand at k; -1 from the first k the walk
//   would abandon the attempt at (f stalled, k > n) to the end of the row; the
//   whole row -1 when count n is not a plain inversion draw (n q >= 10 -> BTRS,
//   p outside (0, 1/2], g near the ends of the exponent range).
// ---------------------------------------------------------------------------
constexpr int kWalkK = 32;

__device__ __forceinline__ void fill_walk_row(double p, double q, double qq, double r, int n,
                                              double *__restrict__ row) {
  const double g = r * (double)(n + 1);
  const bool plain = p > 0.0 && p <= 0.5 && !((double)n * q >= 10.0) && g > 1e-280 && g < 1e280;
  double f = plain ? pow_int(qq, n) : -1.0;
  bool dead = !plain;
  row[0] = f;
  for (int k = 1; k < kWalkK; ++k) {
    if (!dead) {
      const double fn = f * (g / (double)k - r);
      if (fn == f || k > n) dead = true; else f = fn;
    }
    row[k] = dead ? -1.0 : f;
  }
}

// Binomial(n, p) from row n of a walk table.  Returns the draw, or -1 with the
// generator exactly as it was on entry (not a plain draw, an abandoned attempt, or
// a walk past the row): the caller hands the rest of the interval to the general
// code.  A -1 entry keeps the chain going (u >= -1) to the end of the row.
template <int SCR>
__device__ __forceinline__ int draw_binomial_table(Xoshiro &rng, const double *__restrict__ row,
                                                   double4 F) {   // F = the row's first four entries
  if (F.x < 0.0) return -1;
  double u = xo_real<SCR>(rng);
  int k = 0, c3;
  // one way out of the loop (the lanes of a warp leave it at different rounds: with a
  // single exit they meet again once, and what follows runs once)
  bool more;
  do {
    // four steps of the chain: k counts the leading steps that go on (u >= f_k),
    // each comparison and-ed with the one before inside the DSETP itself
    asm("{\n\t.reg .pred p0, p1, p2, p3;\n\t"
        "setp.ge.f64 p0, %1, %3;\n\t"
        "sub.rn.f64 %1, %1, %3;\n\t"
        "setp.ge.and.f64 p1, %1, %4, p0;\n\t"
        "sub.rn.f64 %1, %1, %4;\n\t"
        "setp.ge.and.f64 p2, %1, %5, p1;\n\t"
        "sub.rn.f64 %1, %1, %5;\n\t"
        "setp.ge.and.f64 p3, %1, %6, p2;\n\t"
        "sub.rn.f64 %1, %1, %6;\n\t"
        "@p0 add.s32 %0, %0, 1;\n\t"
        "@p1 add.s32 %0, %0, 1;\n\t"
        "@p2 add.s32 %0, %0, 1;\n\t"
        "@p3 add.s32 %0, %0, 1;\n\t"
        "selp.s32 %2, 1, 0, p3;\n\t}"
        : "+r"(k), "+d"(u), "=r"(c3)
        : "d"(F.x), "d"(F.y), "d"(F.z), "d"(F.w));
    more = c3 != 0 && k < kWalkK;
    if (more) F = ldg_f64x4(row + k);
  } while (more);
  if (c3) {   // walked past the row
    xo_rewind(rng);
    return -1;
  }
  return k;
}

__device__ __forceinline__ double stirling_tail(double k) {
  if (k <= 9.0) return c_stirling_tail[(int)k];
  const double kp1sq = (k + 1.0) * (k + 1.0);
  return (1.0 / 12.0 - (1.0 / 360.0 - 1.0 / 1260.0 / kp1sq) / kp1sq) / (k + 1.0);
}

#ifndef MCB_BTRS_INLINE
#define MCB_BTRS_INLINE __forceinline__
#endif
template <int SCR>
__device__ MCB_BTRS_INLINE double binomial_btrs(Xoshiro &rng, double n, double q) {
  const double sd = sqrt(n * q * (1.0 - q));
  const double b = 1.15 + 2.53 * sd;
  const double a = -0.0873 + 0.0248 * b + 0.01 * q;
  const double c = n * q + 0.5;
  const double v_r = 0.92 - 4.2 / b;
  const double r = q / (1.0 - q);
  const double alpha = (2.83 + 5.1 / b) * sd;
  const double m = floor((n + 1.0) * q);
  for (;;) {
    double u = xo_real<SCR>(rng);
    double v = xo_real<SCR>(rng);
    u -= 0.5;
    const double us = 0.5 - fabs(u);
    const double k = floor((2.0 * a / us + b) * u + c);
    if (us >= 0.07 && v <= v_r) return k;
    if (k < 0.0 || k > n) continue;
    v = det_log(v * alpha / (a / (us * us) + b));
    const double bound =
        (m + 0.5) * det_log((m + 1.0) / (r * (n - m + 1.0))) +
        (n + 1.0) * det_log((n - m + 1.0) / (n - k + 1.0)) +
        (k + 0.5) * det_log(r * (n - k + 1.0) / (k + 1.0)) +
        stirling_tail(m) + stirling_tail(n - m) - stirling_tail(k) -
        stirling_tail(n - k);
    if (v <= bound) return k;
  }
}

// deterministic = true replaces the draw by its mean (dust deterministic mode,
// R/deterministic_state.R:190-203)
template <int SCR>
__device__ __forceinline__ double draw_binomial(Xoshiro &rng, bool deterministic, double n,
                                                double p) {
  if (deterministic) return n * p;
  if (n == 0.0 || p == 0.0) return 0.0;
  if (p == 1.0) return n;
  const double q = p > 0.5 ? 1.0 - p : p;
  const double draw = (n * q >= 10.0) ? binomial_btrs<SCR>(rng, n, q)
                                      : binomial_inversion<SCR>(rng, (int)n, q);
  return p > 0.5 ? n - draw : draw;
}

// A binomial whose probability is fixed by the parameter set (recovery, waning):
// q, 1-q, q/(1-q) and the table n -> (1-q)^n are prepared once per parameter
// set (k_prepare_sets), so the per-draw setup -- a division and the integer
// power -- becomes one cached load.  Same operations, same bits.
struct BinomialConst {
  double p, q, qq, r;   // q = min(p, 1-p), qq = 1-q, r = q/qq
  bool flip;            // p > 0.5
};

template <int SCR>
__device__ __forceinline__ double draw_binomial_const(Xoshiro &rng, bool deterministic, double n,
                                                      const BinomialConst &c) {
  if (deterministic) return n * c.p;
  if (n == 0.0 || c.p == 0.0) return 0.0;
  if (c.p == 1.0) return n;
  double draw;
  if (n * c.q >= 10.0) {
    draw = binomial_btrs<SCR>(rng, n, c.q);
  } else {
    const int ni = (int)n;
    draw = inversion_walk<SCR>(rng, ni, c.r, pow_int(c.qq, ni));
  }
  return c.flip ? n - draw : draw;
}

// Poisson(lambda) [SURVEY.md A.2]: lambda == 0 -> 0 without a draw; below 10 Knuth's
// product of uniforms against exp(-lambda); from 10 up Hoermann's PTRS transformed
// rejection (two uniforms per attempt), log(k!) from det_lfact.  deterministic = true
// gives the mean.  Oracle twin: orc_poisson.
template <int SCR>
__device__ __noinline__ double poisson_ptrs(Xoshiro &rng, double lambda) {
  const double b = 0.931 + 2.53 * sqrt(lambda);
  const double a = -0.059 + 0.02483 * b;
  const double inv_alpha = 1.1239 + 1.1328 / (b - 3.4);
  const double v_r = 0.9277 - 3.6224 / (b - 2.0);
  const double log_lambda = det_log(lambda);
  for (;;) {
    const double u = xo_real<SCR>(rng) - 0.5;
    const double v = xo_real<SCR>(rng);
    const double us = 0.5 - fabs(u);
    const double k = floor((2.0 * a / us + b) * u + lambda + 0.43);
    if (us >= 0.07 && v <= v_r) return k;
    if (k < 0.0 || (us < 0.013 && v > us)) continue;
    if (det_log(v) + det_log(inv_alpha) - det_log(a / (us * us) + b) <=
        -lambda + k * log_lambda - det_lfact(k))
      return k;
  }
}

template <int SCR>
__device__ __forceinline__ double draw_poisson(Xoshiro &rng, bool deterministic, double lambda) {
  if (deterministic) return lambda;
  if (lambda == 0.0) return 0.0;
  if (lambda < 10.0) {
    const double lim = det_exp(-lambda);
    double prod = 1.0, k = 0.0;
    for (;;) {
      prod *= xo_real<SCR>(rng);
      if (prod <= lim) return k;
      k += 1.0;
    }
  }
  return poisson_ptrs<SCR>(rng, lambda);
}

// x log(lambda) - lambda - lgamma(x + 1); lgamma(x + 1) depends on the datum
// only and is computed once per datum on the host (mcb_set_data).
__device__ __forceinline__ double density_poisson_log(double x, double lgamma_x1,
                                                      double lambda) {
  if (x == 0.0 && lambda == 0.0) return 0.0;
  return x * det_log(lambda) - lambda - lgamma_x1;
}

}  // namespace mcb
