// Per-particle samplers and densities for the compartmental models.
//
// Algorithms are those of the reference's native dependency dust [SURVEY.md
// A.2]: binomial by CDF inversion (n q < 10) or Hoermann's BTRS transformed
// rejection, Poisson by Knuth's product (lambda < 10) or PTRS.  The number of
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
// the same table in global memory: a block stages it in shared memory with one
// coalesced load per thread (per-thread indexing of the constant bank serialises);
// one spare entry (the lean walk looks two entries ahead)
__device__ double2 g_inv_small[kInvSmall + 1];
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
// Lean twins used by the integer-state path of the run kernel (models.cuh).
// Same operations in the same order as the functions above -- same bits -- with
// the bookkeeping stripped: comparisons of non-negative doubles are done on
// their bit patterns (integer pipe instead of the half-rate fp64 pipe), the
// {k, 1/k} table is read from shared memory with one 16-byte load, and every
// rare event (walk past the table, stalled f, k > n) leaves to a cold function.
// ---------------------------------------------------------------------------
__device__ __forceinline__ long long as_ll(double x) { return __double_as_longlong(x); }

// pow_int with the first multiplication peeled (1.0 * x == x) and the squaring
// moved to the top of the loop: the same products in the same order (a product
// by a square the count has no bit for is skipped, as pow_int never forms it).
// The conditional product is one predicated DMUL.
__device__ __forceinline__ void mul_if(double &acc, double x, int bit) {
  asm("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t@p mul.rn.f64 %0, %0, %1;\n\t}"
      : "+d"(acc) : "d"(x), "r"(bit));
}
// `groups`: how many groups of three bits beyond bit 0 the largest count can have
// (warp-uniform, from the table size): the loop runs on it, not on the lane's n,
// so there is no per-lane loop control.  Squarings past a lane's top bit are
// computed and never multiplied in.
__device__ __forceinline__ double pow_int_lean(double x, int n, int groups) {
  double acc = (n & 1) ? x : 1.0;
#pragma unroll 1
  for (int g = 0; g < groups; ++g) {
    x *= x;
    mul_if(acc, x, n & 2);
    x *= x;
    mul_if(acc, x, n & 4);
    x *= x;
    mul_if(acc, x, n & 8);
    n >>= 3;
  }
  return acc;
}

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

// The CDF walk for the lean path.  Returns k >= 0, or -1 when anything unusual
// happened (the walk ran past min(n, table), or f stopped changing): the caller
// then rewinds the generator by the one uniform drawn here and repeats the draw
// with the general code, which handles those cases by the book.
// s_inv: the {k, RN(1/k)} table in shared memory; g = r (n + 1) must be away from
// overflow/underflow (the caller's table entry guarantees it).
__device__ __forceinline__ double2 lds_f64x2(uint32_t addr) {
  double2 e;
  asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(e.x), "=d"(e.y) : "r"(addr));
  return e;
}
__device__ __forceinline__ double2 ldg_f64x2(const double *p) {
  double2 e;
  asm("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(e.x), "=d"(e.y) : "l"(p));
  return e;
}
// one 32-byte load (LDG.256): four consecutive table entries
__device__ __forceinline__ double4 ldg_f64x4(const double *p) {
  double4 e;
  asm("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];"
      : "=d"(e.x), "=d"(e.y), "=d"(e.z), "=d"(e.w) : "l"(p));
  return e;
}

// s_inv_addr: shared-window address of the {k, RN(1/k)} table.
// One predicate drives each step: keep walking while u >= f, f changed, and the
// table has another entry; the two rare reasons are told apart after the walk
// (the order of the reference's tests is kept: a stalled f abandons the attempt
// whatever u is).
// Every lane of a warp is at the same k in the same step, so the first MCB_PEEL
// steps are written out with their divisor known at compile time: g/1, g/2, g/4,
// g/8 are exact scalings (g is away from underflow), the others take k and its
// reciprocal as immediates (no table load, no address arithmetic).  A warp walks
// as far as its slowest lane (about 8 steps at the benchmark's parameters), so
// these steps are what nearly every warp executes.  The loop after them is
// unrolled by two so f and its successor swap roles without moves.
#ifndef MCB_PEEL
#define MCB_PEEL 12   // measured at 2^20: 4 -> 90.1 us, 8 -> 88.0, 12 -> 87.2, 16 -> 87.4
#endif
// RN(1/k) for the written-out steps, as constant-bank operands (a 64-bit immediate
// would cost two uniform moves per use)
__constant__ double c_recip[17] = {0.0,        1.0,        1.0 / 2.0,  1.0 / 3.0,  1.0 / 4.0,
                                   1.0 / 5.0,  1.0 / 6.0,  1.0 / 7.0,  1.0 / 8.0,  1.0 / 9.0,
                                   1.0 / 10.0, 1.0 / 11.0, 1.0 / 12.0, 1.0 / 13.0, 1.0 / 14.0,
                                   1.0 / 15.0, 1.0 / 16.0};
// One written-out step.  The two ways out need nothing but k (a stalled f abandons
// the attempt whatever u is -- the order of the reference's tests), so no value has to
// be kept in a fixed register for them and the steps chain without moves.  (Measured:
// one compound exit to a common label 87.2 us, these two returns 86.3, per-step
// landing pads 88.6.)
#define MCB_WALK_STEP(K, FACTOR)                 \
  u -= f;                                        \
  fn = f * ((FACTOR) - r);                       \
  if (fn == f) return -1;                        \
  if (!(u >= fn)) return (K);                    \
  f = fn;
#define MCB_WALK_DIV(K) div_small_int(g, (double)(K), c_recip[K])
template <int SCR>
__device__ __forceinline__ int walk_lean(Xoshiro &rng, int n, double r, double f0, double g,
                                         uint32_t s_inv_addr) {
  const int kmax = min(n, kInvSmall - 1);
  double u = xo_real<SCR>(rng);
  double f = f0, fn;
  if (!(u >= f)) return 0;   // n >= 1: there is a k = 1 to move to
  uint32_t a = s_inv_addr;
  if (kmax > MCB_PEEL) {     // the usual case: every written-out step has a successor
    MCB_WALK_STEP(1, g)
    MCB_WALK_STEP(2, g * 0.5)
    // g/6, g/12 and g/10 are g/3 and g/5 scaled by a power of two: the same bits as
    // dividing (RN commutes with an exact scaling, g is away from underflow)
    const double g3 = MCB_WALK_DIV(3);
    MCB_WALK_STEP(3, g3)
    MCB_WALK_STEP(4, g * 0.25)
#if MCB_PEEL >= 6
    const double g5 = MCB_WALK_DIV(5);
    MCB_WALK_STEP(5, g5)
    MCB_WALK_STEP(6, g3 * 0.5)
#endif
#if MCB_PEEL >= 8
    MCB_WALK_STEP(7, MCB_WALK_DIV(7))
    MCB_WALK_STEP(8, g * 0.125)
#endif
#if MCB_PEEL >= 10
    MCB_WALK_STEP(9, MCB_WALK_DIV(9))
    MCB_WALK_STEP(10, g5 * 0.5)
#endif
#if MCB_PEEL >= 12
    MCB_WALK_STEP(11, MCB_WALK_DIV(11))
    MCB_WALK_STEP(12, g3 * 0.25)
#endif
#if MCB_PEEL >= 16
    MCB_WALK_STEP(13, MCB_WALK_DIV(13))
    MCB_WALK_STEP(14, MCB_WALK_DIV(14))
    MCB_WALK_STEP(15, MCB_WALK_DIV(15))
    MCB_WALK_STEP(16, g * 0.0625)
#endif
    a += 16u * (uint32_t)(MCB_PEEL < 6 ? 4 : MCB_PEEL);
  }
  // whatever is left (and counts too small for the written-out steps): the loop
  const uint32_t a_end = s_inv_addr + 16u * (uint32_t)kmax;
  for (;;) {
    // The factors c_k = g/k - r do not depend on f or u: the next two are computed
    // side by side (two independent dependency chains), then applied in order.
    // The second may go unused; the table has one spare entry for its load.
    const double2 e1 = lds_f64x2(a + 16);
    const double2 e2 = lds_f64x2(a + 32);
    double q1 = g * e1.y, q2 = g * e2.y;
    double r1 = fma(-q1, e1.x, g), r2 = fma(-q2, e2.x, g);
    q1 = fma(r1, e1.y, q1); q2 = fma(r2, e2.y, q2);
    r1 = fma(-q1, e1.x, g); r2 = fma(-q2, e2.x, g);
    const double c1 = fma(r1, e1.y, q1) - r, c2 = fma(r2, e2.y, q2) - r;
    u -= f;
    fn = f * c1;
    a += 16;
    if (!((u >= fn) & (fn != f) & (a < a_end))) break;
    u -= fn;
    f = fn;
    fn = f * c2;
    a += 16;
    if (!((u >= fn) & (fn != f) & (a < a_end))) break;
    f = fn;
  }
  // stalled, or out of table with u >= f still
  return ((fn == f) | (u >= fn)) ? -1 : (int)((a - s_inv_addr) >> 4);
}
#undef MCB_WALK_STEP
#undef MCB_WALK_DIV

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
  int k = 0;
  for (;;) {
    // four steps of the chain: k counts the leading steps that go on (u >= f_k),
    // each comparison and-ed with the one before inside the DSETP itself
    int c3;
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
    if (!c3) break;
    if (k == kWalkK) {
      xo_rewind(rng);
      return -1;
    }
    F = ldg_f64x4(row + k);
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

// x log(lambda) - lambda - lgamma(x + 1); lgamma(x + 1) depends on the datum
// only and is computed once per datum on the host (mcb_set_data).
__device__ __forceinline__ double density_poisson_log(double x, double lgamma_x1,
                                                      double lambda) {
  if (x == 0.0 && lambda == 0.0) return 0.0;
  return x * det_log(lambda) - lambda - lgamma_x1;
}

}  // namespace mcb
