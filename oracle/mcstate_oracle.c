/*
 * mcstate_oracle.c -- CPU ORACLE, TEST INFRASTRUCTURE ONLY (see mcstate_oracle.h).
 *
 * Restates, in plain C with an OpenMP particle loop, what
 * particle_filter$run(pars) computes on the compiled-compare path:
 *   R/particle_filter_state.R:138-162 (step_compiled -> model$filter) with the
 *   per-interval order of R/particle_filter_state.R:63-117
 *   (run -> save history -> compare -> scale_log_weights -> accumulate ->
 *    early exit -> resample -> reorder -> record order -> snapshot).
 * Everything below the model-object boundary is the absent dependency dust
 * (>= 0.13.12); it follows SURVEY.md Appendix A and is PARITY UNPINNED --
 * every such choice is a single named function here and a line in
 * PARITY_NOTES.md.
 *
 * Build: see oracle/Makefile (gcc -O2 -fopenmp -ffp-contract=off -mfma).
 * -ffp-contract=off matters: fused multiply-adds happen only where fma() is
 * written, so the GPU kernels (nvcc -fmad=false) round identically.
 */
#include "mcstate_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ */
/* bit casts                                                           */
/* ------------------------------------------------------------------ */
static inline uint64_t d2u(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
static inline double u2d(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }

/* ------------------------------------------------------------------ */
/* RNG: xoshiro256 with + or ** scrambler              */
/* algorithm name "xoshiro256plus": tests/testthat/test-pmcmc-parallel.R:41 */
/* 32 bytes per stream: R/pmcmc_parallel.R:130                          */
/* ------------------------------------------------------------------ */
uint64_t orc_splitmix64(uint64_t *state) {
  uint64_t z = (*state += 0x9e3779b97f4a7c15ULL);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
  return z ^ (z >> 31);
}

/* [dust, unverified] the output of each splitmix round is fed back as the
 * next round's input (not the canonical "keep the counter" seeding). */
void orc_seed(uint64_t seed, uint64_t s[4]) {
  for (int i = 0; i < 4; ++i) {
    uint64_t st = seed;
    seed = orc_splitmix64(&st);
    s[i] = seed;
  }
}

static inline uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }

uint64_t orc_next(uint64_t s[4], int scrambler) {
  const uint64_t out = scrambler == ORC_SCRAMBLER_PLUS
                           ? s[0] + s[3]
                           : rotl64(s[1] * 5, 7) * 9;
  const uint64_t t = s[1] << 17;
  s[2] ^= s[0];
  s[3] ^= s[1];
  s[1] ^= s[2];
  s[0] ^= s[3];
  s[2] ^= t;
  s[3] = rotl64(s[3], 45);
  return out;
}

static void jump_with(uint64_t s[4], const uint64_t poly[4]) {
  uint64_t acc[4] = {0, 0, 0, 0};
  for (int i = 0; i < 4; ++i) {
    for (int b = 0; b < 64; ++b) {
      if (poly[i] & (1ULL << b)) {
        for (int k = 0; k < 4; ++k) acc[k] ^= s[k];
      }
      orc_next(s, ORC_SCRAMBLER_PLUS); /* state advance is scrambler independent */
    }
  }
  memcpy(s, acc, sizeof acc);
}

void orc_jump(uint64_t s[4]) {
  static const uint64_t poly[4] = {0x180ec6d33cfd0abaULL, 0xd5a61266f0c9392cULL,
                                   0xa9582618e03fc9aaULL, 0x39abdc4529b1661cULL};
  jump_with(s, poly);
}

void orc_long_jump(uint64_t s[4]) {
  static const uint64_t poly[4] = {0x76e15d3efefdcbbfULL, 0xc5004e441c522fb3ULL,
                                   0x77710069854ee241ULL, 0x39109bb02acbe635ULL};
  jump_with(s, poly);
}

double orc_random_real(uint64_t s[4], int scrambler) {
  return (double)(orc_next(s, scrambler) >> 11) * 0x1.0p-53;
}

/* First n_seed_streams streams verbatim, the rest by jumping on from the last
 * supplied one (A.1 "raw seed of 32k bytes"). */
void orc_rng_streams(const uint64_t *seed_state, size_t n_seed_streams,
                     size_t n_streams, uint64_t *out) {
  size_t have = n_seed_streams < n_streams ? n_seed_streams : n_streams;
  memcpy(out, seed_state, have * 4 * sizeof(uint64_t));
  for (size_t i = have; i < n_streams; ++i) {
    memcpy(out + 4 * i, out + 4 * (i - 1), 4 * sizeof(uint64_t));
    orc_jump(out + 4 * i);
  }
}

/* R/pmcmc_parallel.R:125-151, R/smc2.R:68-71: node k starts from the base
 * state long-jumped k times; chain 1 equals the input
 * (tests/testthat/test-pmcmc-parallel.R:47-48,58). out: [n_nodes][n_streams][4] */
void orc_rng_distributed_state(const uint64_t seed_state[4], size_t n_streams,
                               size_t n_nodes, uint64_t *out) {
  uint64_t base[4];
  memcpy(base, seed_state, sizeof base);
  for (size_t k = 0; k < n_nodes; ++k) {
    orc_rng_streams(base, 1, n_streams, out + k * n_streams * 4);
    orc_long_jump(base);
  }
}

/* ------------------------------------------------------------------ */
/* Deterministic exp / log.  Only IEEE add/mul/div/fma and integer ops, */
/* so the CUDA kernels can reproduce every bit (libm and CUDA's own     */
/* exp/log differ in the last place now and then).  Both are accurate   */
/* to < 1 ulp; tests/test_oracle_math.py checks them against libm.      */
/* ------------------------------------------------------------------ */
double orc_exp(double x) {
  if (x != x) return x;
  if (x > 709.782712893384) return INFINITY;
  if (x < -745.1332191019412) return 0.0;
  const double kd = rint(x * 1.4426950408889634074);
  /* Cody-Waite reduction with ln2 split in two doubles; fma keeps k*ln2_hi exact */
  double r = fma(-kd, 6.93147180559945286e-01, x);
  r = fma(-kd, 2.31904681384629956e-17, r);
  /* Taylor series to r^13 on |r| <= ln2/2: truncation < 5e-18 */
  double p = 1.6059043836821613e-10;     /* 1/13! */
  p = fma(p, r, 2.08767569878681e-09);   /* 1/12! */
  p = fma(p, r, 2.505210838544172e-08);  /* 1/11! */
  p = fma(p, r, 2.755731922398589e-07);  /* 1/10! */
  p = fma(p, r, 2.7557319223985893e-06); /* 1/9!  */
  p = fma(p, r, 2.48015873015873e-05);   /* 1/8!  */
  p = fma(p, r, 1.984126984126984e-04);  /* 1/7!  */
  p = fma(p, r, 1.388888888888889e-03);  /* 1/6!  */
  p = fma(p, r, 8.333333333333333e-03);  /* 1/5!  */
  p = fma(p, r, 4.1666666666666664e-02); /* 1/4!  */
  p = fma(p, r, 1.6666666666666666e-01); /* 1/3!  */
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  /* scale by 2^k in two exact-or-once-rounded steps (covers subnormal results) */
  const int k = (int)kd;
  const int k1 = k >> 1, k2 = k - k1;
  const double s1 = u2d((uint64_t)(1023 + k1) << 52);
  const double s2 = u2d((uint64_t)(1023 + k2) << 52);
  return (p * s1) * s2;
}

/* Argument reduction to [sqrt(1/2), sqrt(2)) and the s = f/(2+f) series with
 * the classic 7-term minimax polynomial (Sun fdlibm e_log.c algorithm). */
double orc_log(double x) {
  uint64_t ix = d2u(x);
  uint32_t hx = (uint32_t)(ix >> 32);
  int k = 0;
  if (hx < 0x00100000u || (hx >> 31)) {
    if ((ix << 1) == 0) return -INFINITY;
    if (hx >> 31) return NAN;
    k -= 54;
    x *= 0x1p54;
    ix = d2u(x);
    hx = (uint32_t)(ix >> 32);
  } else if (hx >= 0x7ff00000u) {
    return x;
  } else if (hx == 0x3ff00000u && (ix << 32) == 0) {
    return 0.0;
  }
  hx += 0x3ff00000u - 0x3fe6a09eu;
  k += (int)(hx >> 20) - 0x3ff;
  hx = (hx & 0x000fffffu) + 0x3fe6a09eu;
  x = u2d(((uint64_t)hx << 32) | (ix & 0xffffffffULL));

  const double f = x - 1.0;
  const double hfsq = 0.5 * f * f;
  const double s = f / (2.0 + f);
  const double z = s * s;
  const double w = z * z;
  const double t1 = w * fma(w, fma(w, 1.531383769920937332e-01, 2.222219843214978396e-01),
                            3.999999999940941908e-01);
  const double t2 = z * fma(w, fma(w, fma(w, 1.479819860511658591e-01,
                                          1.818357216161805012e-01),
                                   2.857142874366239149e-01),
                            6.666666666666735130e-01);
  const double R = t2 + t1;
  const double dk = (double)k;
  return s * (hfsq + R) + dk * 1.90821492927058770002e-10 - hfsq + f +
         dk * 6.93147180369123816490e-01;
}

/* cos(2 pi u) for u in [0, 1).  The reduction is exact: x = 4u = q + r with q the
 * quadrant and r in [0, 1), r <- 1 - r above one half, so only a = (pi/2) r in
 * [0, pi/4] is rounded; then the fdlibm kernel polynomials for sin and cos, written
 * with explicit fma.  cos(pi/2 (q + r)) = {cos, -sin, -cos, sin}(pi/2 r) for q = 0..3. */
double orc_cos2pi(double u) {
  const double x = 4.0 * u;
  const double qd = floor(x);
  double r = x - qd;
  const int q = (int)qd;
  const int swap = r > 0.5;
  if (swap) r = 1.0 - r;
  const double a = r * 1.57079632679489661923;
  const double z = a * a;
  const int use_sin = (q & 1) ^ swap;
  double val;
  if (use_sin) {
    double t = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
    t = fma(z, t, 2.75573137070700676789e-06);
    t = fma(z, t, -1.98412698298579493134e-04);
    t = fma(z, t, 8.33333333332248946124e-03);
    t = fma(z, t, -1.66666666666666324348e-01);
    val = fma(z * a, t, a);
  } else {
    double t = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
    t = fma(z, t, -2.75573143513906633035e-07);
    t = fma(z, t, 2.48015872894767294178e-05);
    t = fma(z, t, -1.38888888888741095749e-03);
    t = fma(z, t, 4.16666666666666019037e-02);
    val = 1.0 - (0.5 * z - z * (z * t));
  }
  return (q == 1 || q == 2) ? -val : val;
}

/* ------------------------------------------------------------------ */
/* Samplers and densities [dust, unverified]           */
/* ------------------------------------------------------------------ */
/* Box-Muller, first variate only; a uniform at or below machine epsilon is drawn
 * again (both of the pair) so log() never sees it [dust random/normal, unverified] */
double orc_normal(uint64_t s[4], int scrambler, double mean, double sd) {
  double u1, u2;
  do {
    u1 = orc_random_real(s, scrambler);
    u2 = orc_random_real(s, scrambler);
  } while (u1 <= 2.220446049250313e-16);
  const double z = sqrt(-2.0 * orc_log(u1)) * orc_cos2pi(u2);
  return z * sd + mean;
}

/* dnorm(x, mu, sd, log = TRUE) = -(x - mu)^2 / (2 sd^2) - log(sqrt(2 pi)) - log(sd) */
double orc_density_normal_log(double x, double mu, double sd) {
  const double dx = x - mu;
  return -(dx * dx) / (2.0 * sd * sd) - 0.918938533204672741780329736406 - orc_log(sd);
}

double orc_exponential(uint64_t s[4], int scrambler, double rate) {
  return -orc_log(orc_random_real(s, scrambler)) / rate;
}

/* integer power by repeated squaring -- deliberately not pow() */
static double pow_int(double x, int n) {
  double acc = 1.0;
  while (n != 0) {
    if (n & 1) acc *= x;
    n >>= 1;
    if (n != 0) x *= x;
  }
  return acc;
}

/* CDF walk from k = 0; one uniform per attempt; an attempt is abandoned
 * (and a fresh uniform drawn) when f stops changing or k runs past n. */
static double binomial_inversion(uint64_t s[4], int scrambler, int n, double q) {
  const double qq = 1.0 - q;
  const double r = q / qq;
  const double g = r * (double)(n + 1);
  const double f0 = pow_int(qq, n);
  for (;;) {
    double u = orc_random_real(s, scrambler);
    double f = f0, f_prev = f0;
    int k = 0, ok = 1;
    while (u >= f) {
      u -= f;
      ++k;
      f *= (g / (double)k - r);
      if (f == f_prev || k > n) { ok = 0; break; }
      f_prev = f;
    }
    if (ok) return (double)k;
  }
}

static double stirling_tail(double k) {
  static const double tab[10] = {
      0.0810614667953272,  0.0413406959554092,  0.0276779256849983,
      0.02079067210376509, 0.0166446911898211,  0.0138761288230707,
      0.0118967099458917,  0.0104112652619720,  0.00925546218271273,
      0.00833056343336287};
  if (k <= 9.0) return tab[(int)k];
  const double kp1sq = (k + 1.0) * (k + 1.0);
  return (1.0 / 12.0 - (1.0 / 360.0 - 1.0 / 1260.0 / kp1sq) / kp1sq) / (k + 1.0);
}

/* Hoermann 1993 transformed rejection (BTRS); two uniforms per attempt */
static double binomial_btrs(uint64_t s[4], int scrambler, double n, double q) {
  const double sd = sqrt(n * q * (1.0 - q));
  const double b = 1.15 + 2.53 * sd;
  const double a = -0.0873 + 0.0248 * b + 0.01 * q;
  const double c = n * q + 0.5;
  const double v_r = 0.92 - 4.2 / b;
  const double r = q / (1.0 - q);
  const double alpha = (2.83 + 5.1 / b) * sd;
  const double m = floor((n + 1.0) * q);
  for (;;) {
    double u = orc_random_real(s, scrambler);
    double v = orc_random_real(s, scrambler);
    u -= 0.5;
    const double us = 0.5 - fabs(u);
    const double k = floor((2.0 * a / us + b) * u + c);
    if (us >= 0.07 && v <= v_r) return k;
    if (k < 0.0 || k > n) continue;
    v = orc_log(v * alpha / (a / (us * us) + b));
    const double bound =
        (m + 0.5) * orc_log((m + 1.0) / (r * (n - m + 1.0))) +
        (n + 1.0) * orc_log((n - m + 1.0) / (n - k + 1.0)) +
        (k + 0.5) * orc_log(r * (n - k + 1.0) / (k + 1.0)) +
        stirling_tail(m) + stirling_tail(n - m) - stirling_tail(k) -
        stirling_tail(n - k);
    if (v <= bound) return k;
  }
}

double orc_binomial(uint64_t s[4], int scrambler, double n, double p) {
  if (n == 0.0 || p == 0.0) return 0.0;
  if (p == 1.0) return n;
  const double q = p > 0.5 ? 1.0 - p : p;
  const double draw = (n * q >= 10.0) ? binomial_btrs(s, scrambler, n, q)
                                      : binomial_inversion(s, scrambler, (int)n, q);
  return p > 0.5 ? n - draw : draw;
}

/* log(k!) for an integer-valued k >= 0, reproducible to the bit (engine twin:
 * det_lfact, csrc/detmath.cuh): a table below 16, from there Stirling's series for
 * lgamma(x), x = k + 1, with four correction terms (truncation below 1e-14 at
 * x = 17) -- IEEE operations and orc_log only, no libm lgamma. */
static const double k_log_fact[16] = {
    0.0, 0.0, 0.693147180559945, 1.7917594692280554, 3.178053830347945, 4.787491742782047,
    6.579251212010102, 8.525161361065415, 10.604602902745249, 12.801827480081467,
    15.104412573075514, 17.502307845873887, 19.987214495661885, 22.55216385312342,
    25.191221182738683, 27.89927138384089};
double orc_lfact(double k) {
  if (k < 16.0) return k_log_fact[(int)k];
  const double x = k + 1.0;
  const double xi = 1.0 / x, x2 = xi * xi;
  const double ser =
      (1.0 / 12.0 - x2 * (1.0 / 360.0 - x2 * (1.0 / 1260.0 - x2 * (1.0 / 1680.0)))) * xi;
  return (x - 0.5) * orc_log(x) - x + 0.91893853320467274178 + ser;
}

/* lambda == 0 -> 0 (no draw); Knuth's product below 10; Hoermann's
 * PTRS from 10 up, with log(k!) from orc_lfact */
double orc_poisson(uint64_t s[4], int scrambler, double lambda) {
  if (lambda == 0.0) return 0.0;
  if (lambda < 10.0) { /* Knuth: multiply uniforms until below exp(-lambda) */
    const double lim = orc_exp(-lambda);
    double prod = 1.0, k = 0.0;
    for (;;) {
      prod *= orc_random_real(s, scrambler);
      if (prod <= lim) return k;
      k += 1.0;
    }
  }
  /* Hoermann PTRS */
  const double b = 0.931 + 2.53 * sqrt(lambda);
  const double a = -0.059 + 0.02483 * b;
  const double inv_alpha = 1.1239 + 1.1328 / (b - 3.4);
  const double v_r = 0.9277 - 3.6224 / (b - 2.0);
  const double log_lambda = orc_log(lambda);
  for (;;) {
    double u = orc_random_real(s, scrambler) - 0.5;
    const double v = orc_random_real(s, scrambler);
    const double us = 0.5 - fabs(u);
    const double k = floor((2.0 * a / us + b) * u + lambda + 0.43);
    if (us >= 0.07 && v <= v_r) return k;
    if (k < 0.0 || (us < 0.013 && v > us)) continue;
    if (orc_log(v) + orc_log(inv_alpha) - orc_log(a / (us * us) + b) <=
        -lambda + k * log_lambda - orc_lfact(k))
      return k;
  }
}

/* x log(lambda) - lambda - lgamma(x+1); lgamma(x+1) depends on the datum only
 * and is supplied by the caller (host libm, once per datum). */
double orc_density_poisson_log(double x, double lgamma_x1, double lambda) {
  if (x == 0.0 && lambda == 0.0) return 0.0;
  return x * orc_log(lambda) - lambda - lgamma_x1;
}

/* ------------------------------------------------------------------ */
/* Weights                                                             */
/* ------------------------------------------------------------------ */
/* R/particle_filter.R:570-587, libm exp/log exactly as R would use */
double orc_scale_log_weights_r(const double *logw, size_t n, double *weights_out) {
  double mx = -INFINITY;
  for (size_t i = 0; i < n; ++i) {
    const double v = isnan(logw[i]) ? -INFINITY : logw[i];
    if (v > mx) mx = v;
  }
  if (!isfinite(mx)) {
    if (weights_out) for (size_t i = 0; i < n; ++i) weights_out[i] = NAN;
    return -INFINITY;
  }
  double tot = 0.0;
  for (size_t i = 0; i < n; ++i) {
    const double v = isnan(logw[i]) ? -INFINITY : logw[i];
    const double w = exp(v - mx);
    if (weights_out) weights_out[i] = w;
    tot += w;
  }
  return log(tot / (double)n) + mx;
}

/* Fixed-point weights.  A weight in [0, 1] becomes the integer rint(w 2^52): every
 * weight keeps at least the precision a double sum of weights in [0, 1] has, whatever
 * the number of particles (a rounding error of at most 2^-53 per weight, against up to
 * n 2^-53 per addition of a left-to-right fp64 sum).  Sums are 128-bit integers (n
 * below 2^31 of them stay below 2^83), so they are exact and independent of the order
 * of summation -- a tiled parallel scan and this serial loop give the same bits. */
typedef unsigned __int128 u128;
#define ORC_WEIGHT_SHIFT 52
int orc_weight_shift(size_t n) {
  (void)n;
  return ORC_WEIGHT_SHIFT;
}

static inline uint64_t weight_to_fixed(double w) {
  return (uint64_t)llrint(w * 0x1p52);
}

/* the double nearest to x (ties to even): the top 64 bits with a sticky last bit
 * are converted (one correctly rounded conversion), then scaled exactly */
static inline double u128_to_double(u128 x) {
  const uint64_t hi = (uint64_t)(x >> 64), lo = (uint64_t)x;
  if (hi == 0) return (double)lo;
  const int s = 64 - __builtin_clzll(hi);       /* bits of hi in use: 1..63 here */
  uint64_t v = (hi << (64 - s)) | (lo >> s);
  if ((lo << (64 - s)) != 0) v |= 1ULL;
  return (double)v * u2d((uint64_t)(1023 + s) << 52);
}

/* ceil(d) as an integer, d >= 0 and finite (doubles from 2^52 up are integers) */
static inline u128 ceil_to_u128(double d) {
  if (d < 0x1p63) return (u128)(uint64_t)ceil(d);
  const uint64_t b = d2u(d);
  const int e = (int)(b >> 52) - 1075;
  const uint64_t m = (b & 0xfffffffffffffULL) | 0x10000000000000ULL;
  return (u128)m << e;
}

/* Compiled-filter flavour: NaN -> -Inf, weights exponentiated in place with the
 * deterministic exp.  mode selects how the mean is accumulated. */
double orc_scale_log_weights(double *w, size_t n, int mode) {
  double mx = -INFINITY;
  for (size_t i = 0; i < n; ++i) {
    if (isnan(w[i])) w[i] = -INFINITY;
    if (w[i] > mx) mx = w[i];
  }
  if (!isfinite(mx)) {
    for (size_t i = 0; i < n; ++i) w[i] = NAN;
    return -INFINITY;
  }
  if (mode == ORC_RESAMPLE_EXACT) {
    u128 tot = 0;
    for (size_t i = 0; i < n; ++i) {
      w[i] = orc_exp(w[i] - mx);
      tot += weight_to_fixed(w[i]);
    }
    const double mean = (u128_to_double(tot) * 0x1p-52) / (double)n;
    return orc_log(mean) + mx;
  }
  double tot = 0.0;
  for (size_t i = 0; i < n; ++i) {
    w[i] = orc_exp(w[i] - mx);
    tot += w[i];
  }
  return orc_log(tot / (double)n) + mx;
}

/* ------------------------------------------------------------------ */
/* Resampling                                                          */
/* ------------------------------------------------------------------ */
/* R/particle_filter.R:515-523.  u0 is the value runif(1, 0, 1/n) returned.
 * findInterval(u, cw) counts cw_j <= u (right-open), +1 for 1-based. */
void orc_particle_resample_r(const double *weights, size_t n, double u0, int *kappa) {
  double tot = 0.0;
  for (size_t i = 0; i < n; ++i) tot += weights[i];
  double *cw = (double *)malloc(n * sizeof(double));
  double acc = 0.0;
  for (size_t i = 0; i < n; ++i) { acc += weights[i] / tot; cw[i] = acc; }
  size_t j = 0;
  for (size_t i = 0; i < n; ++i) {
    /* seq(0, by = 1/n, length.out = n)[i] is 0 + i * by in R */
    const double u = u0 + (double)i * (1.0 / (double)n);
    while (j < n && cw[j] <= u) ++j;
    kappa[i] = (int)j + 1;
  }
  free(cw);
}

/* dust resample_weight: thresholds tot*u/n + i*tot/n; particle
 * i takes the first j whose inclusive cumulative weight reaches its threshold.
 * EXACT mode sums the weights as fixed-point integers, so the result does not
 * depend on summation order and a parallel scan reproduces it bit for bit. */
void orc_resample_weight(const double *w, size_t n, double u, int mode, int *kappa) {
  if (mode == ORC_RESAMPLE_EXACT) {
    u128 tot = 0;
    for (size_t i = 0; i < n; ++i) tot += weight_to_fixed(w[i]);
    const double tot_d = u128_to_double(tot);
    const double uu0 = tot_d * u / (double)n, du = tot_d / (double)n;
    u128 ww = 0;
    size_t j = 0;
    for (size_t i = 0; i < n; ++i) {
      const u128 thr = ceil_to_u128(uu0 + (double)i * du);
      while (ww < thr && j < n) { ww += weight_to_fixed(w[j]); ++j; }
      kappa[i] = j == 0 ? 0 : (int)(j - 1);
    }
    return;
  }
  double tot = 0.0;
  for (size_t i = 0; i < n; ++i) tot += w[i];
  const double uu0 = tot * u / (double)n, du = tot / (double)n;
  double ww = 0.0;
  size_t j = 0;
  for (size_t i = 0; i < n; ++i) {
    const double uu = uu0 + (double)i * du;
    while (ww < uu && j < n) { ww += w[j]; ++j; }
    kappa[i] = j == 0 ? 0 : (int)(j - 1);
  }
}

/* R/particle_filter.R:616-646 with index_particle = NULL.
 * value: [nt][np][ny] (R dims ny,np,nt), order: [nt][np] 1-based, out like value */
void orc_history_single(const double *value, const int *order, size_t ny,
                        size_t np, size_t nt, double *out) {
  int *cur = (int *)malloc(np * sizeof(int));
  for (size_t i = 0; i < np; ++i) cur[i] = (int)i + 1;
  for (size_t t = nt; t-- > 0;) {
    for (size_t i = 0; i < np; ++i) {
      cur[i] = order[t * np + (size_t)(cur[i] - 1)];
      memcpy(out + (t * np + i) * ny, value + (t * np + (size_t)(cur[i] - 1)) * ny,
             ny * sizeof(double));
    }
  }
  free(cur);
}

/* ------------------------------------------------------------------ */
/* Models [dust inst/examples, unverified]             */
/* parameter vectors:                                                  */
/*   sir : beta gamma I0 exp_noise S0 R0 freq                          */
/*   sirs: alpha beta gamma I0 exp_noise S0 R0 freq                    */
/*   volatility: alpha sigma gamma tau x0                              */
/*     (tests/testthat/helper-mcstate.R:98-146: x' = alpha x + sigma N(0,1),  */
/*      observed ~ N(gamma x, tau); one real state, the Kalman filter is exact) */
/*   seir: beta sigma gamma I0 exp_noise S0 E0 freq import             */
/*     (not a dust example: the twin of mcstate_b200/examples/seir.cuh, the   */
/*      user-supplied model the engine compiles the way dust::dust() would)   */
/* ------------------------------------------------------------------ */
#define ORC_MAX_PAR 12
#define ORC_MAX_STATE 8

size_t orc_model_n_state(int model) {
  return model == ORC_MODEL_SIR ? 5 : model == ORC_MODEL_SIRS ? 4 : model == ORC_MODEL_SEIR ? 6
       : 1;   /* volatility, walk: 1 */
}
size_t orc_model_n_par(int model) {
  return model == ORC_MODEL_SIR ? 7 : model == ORC_MODEL_SIRS ? 8 : model == ORC_MODEL_SEIR ? 9
       : model == ORC_MODEL_VOLATILITY ? 5 : 1;
}
size_t orc_model_n_data(int model) { (void)model; return 1; }

void orc_model_default_pars(int model, double *p) {
  if (model == ORC_MODEL_SIR) {
    /* scripts/generate_vignette_data.R:11-12, tests/testthat/helper-mcstate.R:11-15 */
    p[0] = 0.2; p[1] = 0.1; p[2] = 10.0; p[3] = 1e6; p[4] = 1000.0; p[5] = 0.0; p[6] = 4.0;
  } else if (model == ORC_MODEL_SIRS) {
    p[0] = 0.1; p[1] = 0.2; p[2] = 0.1; p[3] = 10.0; p[4] = 1e6; p[5] = 1000.0;
    p[6] = 0.0; p[7] = 4.0;
  } else if (model == ORC_MODEL_VOLATILITY) {
    /* tests/testthat/helper-mcstate.R:99 */
    p[0] = 0.91; p[1] = 1.0; p[2] = 1.0; p[3] = 1.0; p[4] = 0.0;
  } else if (model == ORC_MODEL_SEIR) {
    p[0] = 0.3; p[1] = 0.25; p[2] = 0.1; p[3] = 10.0; p[4] = 1e6; p[5] = 1000.0; p[6] = 0.0;
    p[7] = 4.0; p[8] = 0.0;
  } else {
    p[0] = 1.0;   /* walk: sd */
  }
}

static void model_initial(int model, const double *p, double *y) {
  if (model == ORC_MODEL_SIR) {
    y[0] = p[4]; y[1] = p[2]; y[2] = p[5]; y[3] = 0.0; y[4] = 0.0;
  } else if (model == ORC_MODEL_SIRS) {
    y[0] = p[5]; y[1] = p[3]; y[2] = p[6]; y[3] = 0.0;
  } else if (model == ORC_MODEL_VOLATILITY) {
    y[0] = p[4];
  } else if (model == ORC_MODEL_SEIR) {
    y[0] = p[5]; y[1] = p[6]; y[2] = p[3]; y[3] = 0.0; y[4] = 0.0; y[5] = 0.0;
  } else {
    y[0] = 0.0;
  }
}

static inline double draw_binomial(uint64_t s[4], int scr, int det, double n, double p) {
  return det ? n * p : orc_binomial(s, scr, n, p);
}

/* vignettes/sir_models.Rmd:44-54; n_IR is drawn before n_SI [unverified] */
static void sir_update(uint64_t time, const double *y, uint64_t s[4], int scr, int det,
                       const double *p, double *yn) {
  const double S = y[0], I = y[1], R = y[2], cum = y[3];
  const double beta = p[0], gamma = p[1], freq = p[6], dt = 1.0 / freq;
  const double N = S + I + R;
  const double p_SI = 1.0 - orc_exp(-beta * I / N);
  const double p_IR = 1.0 - orc_exp(-gamma);
  const double n_IR = draw_binomial(s, scr, det, I, p_IR * dt);
  const double n_SI = draw_binomial(s, scr, det, S, p_SI * dt);
  yn[0] = S - n_SI;
  yn[1] = I + n_SI - n_IR;
  yn[2] = R + n_IR;
  yn[3] = cum + n_SI;
  yn[4] = (time % (uint64_t)freq == 0) ? n_SI : y[4] + n_SI;
}

static void sirs_update(uint64_t time, const double *y, uint64_t s[4], int scr, int det,
                        const double *p, double *yn) {
  const double S = y[0], I = y[1], R = y[2];
  const double alpha = p[0], beta = p[1], gamma = p[2], freq = p[7], dt = 1.0 / freq;
  const double N = S + I + R;
  const double p_SI = 1.0 - orc_exp(-beta * I / N);
  const double p_IR = 1.0 - orc_exp(-gamma);
  const double p_RS = 1.0 - orc_exp(-alpha);
  const double n_SI = draw_binomial(s, scr, det, S, p_SI * dt);
  const double n_IR = draw_binomial(s, scr, det, I, p_IR * dt);
  const double n_RS = draw_binomial(s, scr, det, R, p_RS * dt);
  yn[0] = S - n_SI + n_RS;
  yn[1] = I + n_SI - n_IR;
  yn[2] = R + n_IR - n_RS;
  yn[3] = (time % (uint64_t)freq == 0) ? n_SI : y[3] + n_SI;
}

/* SIR with a latent compartment (vignettes/sir_models.Rmd:44-54 extended): draws in the
 * order S->E, E->I, I->R, then Poisson(import dt) exposed arrivals from outside (no draw
 * when import == 0); incidence counts E->I */
static void seir_update(uint64_t time, const double *y, uint64_t s[4], int scr, int det,
                        const double *p, double *yn) {
  const double S = y[0], E = y[1], I = y[2], R = y[3];
  const double beta = p[0], sigma = p[1], gamma = p[2], freq = p[7], dt = 1.0 / freq;
  const double N = S + E + I + R;
  const double p_SE = 1.0 - orc_exp(-beta * I / N);
  const double p_EI = 1.0 - orc_exp(-sigma);
  const double p_IR = 1.0 - orc_exp(-gamma);
  const double n_SE = draw_binomial(s, scr, det, S, p_SE * dt);
  const double n_EI = draw_binomial(s, scr, det, E, p_EI * dt);
  const double n_IR = draw_binomial(s, scr, det, I, p_IR * dt);
  const double lambda = p[8] * dt;
  const double n_imp = det ? lambda : orc_poisson(s, scr, lambda);
  yn[0] = S - n_SE;
  yn[1] = E + n_SE - n_EI + n_imp;
  yn[2] = I + n_EI - n_IR;
  yn[3] = R + n_IR;
  yn[4] = y[4] + n_EI;
  yn[5] = (time % (uint64_t)freq == 0) ? n_EI : y[5] + n_EI;
}

/* x' = alpha x + sigma N(0, 1); deterministic mode puts the mean in place of the draw */
static void volatility_update(const double *y, uint64_t s[4], int scr, int det, const double *p,
                              double *yn) {
  const double z = det ? 0.0 : orc_normal(s, scr, 0.0, 1.0);
  yn[0] = p[0] * y[0] + p[1] * z;
}

void orc_model_update(int model, uint64_t time, const double *y, uint64_t s[4],
                      int scrambler, int deterministic, const double *pars,
                      double *y_next) {
  if (model == ORC_MODEL_SIR) sir_update(time, y, s, scrambler, deterministic, pars, y_next);
  else if (model == ORC_MODEL_SIRS) sirs_update(time, y, s, scrambler, deterministic, pars, y_next);
  else if (model == ORC_MODEL_VOLATILITY) volatility_update(y, s, scrambler, deterministic, pars, y_next);
  else if (model == ORC_MODEL_SEIR) seir_update(time, y, s, scrambler, deterministic, pars, y_next);
  else y_next[0] = y[0] + (deterministic ? 0.0 : orc_normal(s, scrambler, 0.0, pars[0]));   /* walk */
}

/* tests/testthat/helper-mcstate.R:7-22 (R twin): Poisson density of the
 * observed incidence around modelled incidence + Exp(exp_noise) noise drawn
 * from the particle's own stream. datum = {incidence, lgamma(incidence+1)} */
static double model_compare(int model, const double *y, const double *datum,
                            uint64_t s[4], int scr, int det, const double *p) {
  if (model == ORC_MODEL_VOLATILITY) return orc_density_normal_log(datum[0], p[2] * y[0], p[3]);
  const double modelled = model == ORC_MODEL_SIR ? y[4] : model == ORC_MODEL_SEIR ? y[5] : y[3];
  const double rate = model == ORC_MODEL_SIR ? p[3] : p[4];   /* sirs, seir: slot 4 */
  const double noise = det ? 1.0 / rate : orc_exponential(s, scr, rate);
  return orc_density_poisson_log(datum[0], datum[1], modelled + noise);
}

/* ------------------------------------------------------------------ */
/* The model object                                                    */
/* ------------------------------------------------------------------ */
struct orc_model {
  int model, scrambler, deterministic, n_threads, resample_mode;
  size_t n_state, n_par, n_particles, n_pars, n_sets, n_total;
  uint64_t time;
  double *y, *y_next;   /* [n_total][n_state] */
  uint64_t *rng;        /* [n_total + 1][4]; the last stream is the filter's */
  double *pars;         /* [n_sets][n_par] */
  size_t n_data, data_cursor;
  uint64_t *data_time;
  double *data;         /* [n_data][n_data_sets][2] = {value, lgamma(value+1)} */
  int data_shared;
  /* one multi-parameter model sharded over several handles: this one holds sets
   * [set_offset, set_offset + n_sets) of n_sets_global and replays the whole draw order
   * of the shared filter stream (include/mcstate_b200.h: mcb_set_stream_sharing) */
  size_t n_sets_global, set_offset;
  unsigned char *na_global; /* [n_data][n_sets_global] or NULL */
};

orc_model *orc_alloc(int model, const double *pars, size_t n_particles, size_t n_pars,
                     uint64_t time, const uint64_t *seed_state, size_t n_seed_streams,
                     int scrambler, int deterministic, int n_threads) {
  orc_model *m = (orc_model *)calloc(1, sizeof *m);
  m->model = model;
  m->scrambler = scrambler;
  m->deterministic = deterministic;
  m->n_threads = n_threads > 0 ? n_threads : 1;
  m->resample_mode = ORC_RESAMPLE_EXACT;
  m->n_state = orc_model_n_state(model);
  m->n_par = orc_model_n_par(model);
  m->n_particles = n_particles;
  m->n_pars = n_pars;
  m->n_sets = n_pars == 0 ? 1 : n_pars;
  m->n_total = n_particles * m->n_sets;
  m->n_sets_global = m->n_sets;
  m->set_offset = 0;
  m->na_global = NULL;
  m->time = time;
  m->y = (double *)malloc(m->n_total * m->n_state * sizeof(double));
  m->y_next = (double *)malloc(m->n_total * m->n_state * sizeof(double));
  m->rng = (uint64_t *)malloc((m->n_total + 1) * 4 * sizeof(uint64_t));
  m->pars = (double *)malloc(m->n_sets * m->n_par * sizeof(double));
  orc_rng_streams(seed_state, n_seed_streams, m->n_total + 1, m->rng);
  orc_update_state(m, pars, NULL, 0, 1, time, 1);
  return m;
}

void orc_free(orc_model *m) {
  if (!m) return;
  free(m->y); free(m->y_next); free(m->rng); free(m->pars);
  free(m->data_time); free(m->data); free(m->na_global);
  free(m);
}

void orc_set_resample_mode(orc_model *m, int mode) { m->resample_mode = mode; }
void orc_set_n_threads(orc_model *m, int n) { m->n_threads = n > 0 ? n : 1; }
size_t orc_n_total(const orc_model *m) { return m->n_total; }
uint64_t orc_time(const orc_model *m) { return m->time; }

/* R/particle_filter_state.R:248,253: pars + time => new parameters and the
 * model's initial conditions; state => explicit state; RNG never touched. */
void orc_update_state(orc_model *m, const double *pars, const double *state,
                      size_t state_len, int has_time, uint64_t time,
                      int set_initial_state) {
  if (has_time) {
    m->time = time;
    m->data_cursor = 0;
    while (m->data_cursor < m->n_data && m->data_time[m->data_cursor] <= time)
      ++m->data_cursor;
  }
  if (pars) {
    memcpy(m->pars, pars, m->n_sets * m->n_par * sizeof(double));
    if (set_initial_state && !state) {
      for (size_t i = 0; i < m->n_total; ++i)
        model_initial(m->model, m->pars + (i / m->n_particles) * m->n_par,
                      m->y + i * m->n_state);
    }
  }
  if (state) {
    if (state_len == m->n_state) {
      for (size_t i = 0; i < m->n_total; ++i)
        memcpy(m->y + i * m->n_state, state, m->n_state * sizeof(double));
    } else {
      memcpy(m->y, state, m->n_total * m->n_state * sizeof(double));
    }
  }
}

void orc_run(orc_model *m, uint64_t time_end) {
  const size_t ns = m->n_state;
#pragma omp parallel for schedule(static) num_threads(m->n_threads)
  for (size_t i = 0; i < m->n_total; ++i) {
    double a[ORC_MAX_STATE], b[ORC_MAX_STATE];
    double *cur = a, *nxt = b;
    memcpy(cur, m->y + i * ns, ns * sizeof(double));
    const double *p = m->pars + (i / m->n_particles) * m->n_par;
    uint64_t *s = m->rng + 4 * i;
    for (uint64_t t = m->time; t < time_end; ++t) {
      orc_model_update(m->model, t, cur, s, m->scrambler, m->deterministic, p, nxt);
      double *tmp = cur; cur = nxt; nxt = tmp;
    }
    memcpy(m->y + i * ns, cur, ns * sizeof(double));
  }
  if (time_end > m->time) m->time = time_end;
}

void orc_state(const orc_model *m, const size_t *index, size_t n_index, double *out) {
  if (!index) {
    memcpy(out, m->y, m->n_total * m->n_state * sizeof(double));
    return;
  }
  for (size_t i = 0; i < m->n_total; ++i)
    for (size_t k = 0; k < n_index; ++k)
      out[i * n_index + k] = m->y[i * m->n_state + index[k]];
}

/* state rows follow kappa; RNG streams stay with their slot */
void orc_reorder(orc_model *m, const int *kappa) {
  const size_t ns = m->n_state;
#pragma omp parallel for schedule(static) num_threads(m->n_threads)
  for (size_t i = 0; i < m->n_total; ++i)
    memcpy(m->y_next + i * ns, m->y + (size_t)kappa[i] * ns, ns * sizeof(double));
  double *tmp = m->y; m->y = m->y_next; m->y_next = tmp;
}

void orc_rng_state(const orc_model *m, uint64_t *out) {
  memcpy(out, m->rng, (m->n_total + 1) * 4 * sizeof(uint64_t));
}

void orc_set_rng_state(orc_model *m, const uint64_t *in) {
  memcpy(m->rng, in, (m->n_total + 1) * 4 * sizeof(uint64_t));
}

void orc_set_data(orc_model *m, size_t n_data, const uint64_t *time_end,
                  const double *values, int shared) {
  free(m->data_time); free(m->data); free(m->na_global);
  const size_t sets = shared ? 1 : m->n_sets;
  m->n_data = n_data;
  m->data_shared = shared;
  m->data_time = (uint64_t *)malloc(n_data * sizeof(uint64_t));
  m->data = (double *)malloc(n_data * sets * 2 * sizeof(double));
  memcpy(m->data_time, time_end, n_data * sizeof(uint64_t));
  for (size_t i = 0; i < n_data * sets; ++i) {
    m->data[2 * i] = values[i];
    m->data[2 * i + 1] = isnan(values[i]) ? NAN : lgamma(values[i] + 1.0);
  }
  m->data_cursor = 0;
  while (m->data_cursor < n_data && m->data_time[m->data_cursor] <= m->time)
    ++m->data_cursor;
}

static const double *datum_for(const orc_model *m, size_t row, size_t set) {
  const size_t sets = m->data_shared ? 1 : m->n_sets;
  return m->data + (row * sets + (m->data_shared ? 0 : set)) * 2;
}

static void compare_row(orc_model *m, size_t row, double *logw) {
#pragma omp parallel for schedule(static) num_threads(m->n_threads)
  for (size_t i = 0; i < m->n_total; ++i) {
    const size_t set = i / m->n_particles;
    const double *d = datum_for(m, row, set);
    if (isnan(d[0])) { logw[i] = 0.0; continue; } /* NA datum: no draw, no weight */
    logw[i] = model_compare(m->model, m->y + i * m->n_state, d, m->rng + 4 * i,
                            m->scrambler, m->deterministic,
                            m->pars + set * m->n_par);
  }
}

int orc_compare_data(orc_model *m, double *logw) {
  for (size_t row = 0; row < m->n_data; ++row) {
    if (m->data_time[row] == m->time) { compare_row(m, row, logw); return 1; }
  }
  return 0;
}

void orc_set_filter_stream(orc_model *m, const uint64_t state[4]) {
  memcpy(m->rng + 4 * m->n_total, state, 4 * sizeof(uint64_t));
}

void orc_set_stream_sharing(orc_model *m, size_t n_sets_global, size_t set_offset,
                            const unsigned char *na_global) {
  free(m->na_global);
  m->na_global = (unsigned char *)malloc(m->n_data * n_sets_global);
  memcpy(m->na_global, na_global, m->n_data * n_sets_global);
  m->n_sets_global = n_sets_global;
  m->set_offset = set_offset;
}

/* one uniform per parameter set, in set order, from the LAST stream (A.4 step 6) */
void orc_resample(orc_model *m, const double *weights, int *kappa) {
  uint64_t *fs = m->rng + 4 * m->n_total;
  for (size_t pg = 0; pg < m->n_sets_global; ++pg) {
    const double u = orc_random_real(fs, m->scrambler);
    if (pg < m->set_offset || pg >= m->set_offset + m->n_sets) continue;
    const size_t p = pg - m->set_offset;
    int *kp = kappa + p * m->n_particles;
    orc_resample_weight(weights + p * m->n_particles, m->n_particles, u,
                        m->resample_mode, kp);
    for (size_t i = 0; i < m->n_particles; ++i) kp[i] += (int)(p * m->n_particles);
  }
  orc_reorder(m, kappa);
}

/* R/particle_filter_state.R:463-474 */
static int early_exit(const double *ll, size_t n, const double *min_ll, size_t n_min) {
  for (size_t i = 0; i < n; ++i)
    if (ll[i] == -INFINITY) return 1;
  if (n_min == 0) return 0;
  if (n == 1) return ll[0] < min_ll[0];
  if (n_min == 1) {
    double tot = 0.0;
    for (size_t i = 0; i < n; ++i) tot += ll[i];
    return tot < min_ll[0];
  }
  for (size_t i = 0; i < n; ++i)
    if (!(ll[i] < min_ll[i])) return 0;
  return 1;
}

size_t orc_filter(orc_model *m, uint64_t time_end, double *log_likelihood,
                  const size_t *index, size_t n_index, double *trajectories,
                  const uint64_t *snapshot_times, size_t n_snap, double *snapshots,
                  const double *min_log_likelihood, size_t n_min) {
  const size_t nt = m->n_total, np = m->n_particles;
  size_t first = m->data_cursor, last = first;
  while (last < m->n_data && m->data_time[last] <= time_end) ++last;
  const size_t n_rows = last - first;
  const size_t ni = index ? n_index : m->n_state;

  double *w = (double *)malloc(nt * sizeof(double));
  int *kappa = (int *)malloc(nt * sizeof(int));
  double *hist_value = NULL;
  int *hist_order = NULL;
  if (trajectories) {
    hist_value = (double *)malloc((n_rows + 1) * nt * ni * sizeof(double));
    hist_order = (int *)malloc((n_rows + 1) * nt * sizeof(int));
    orc_state(m, index, ni, hist_value);
    for (size_t i = 0; i < nt; ++i) hist_order[i] = (int)i;
  }
  for (size_t p = 0; p < m->n_sets; ++p) log_likelihood[p] = 0.0;

  size_t done = 0, snap = 0;
  int stopped = 0;
  for (size_t row = first; row < last; ++row, ++done) {
    orc_run(m, m->data_time[row]);
    if (trajectories)
      orc_state(m, index, ni, hist_value + (done + 1) * nt * ni);
    compare_row(m, row, w);

    int any_data = 0;
    for (size_t p = 0; p < m->n_sets; ++p) {
      if (isnan(datum_for(m, row, p)[0])) continue;
      any_data = 1;
      log_likelihood[p] += orc_scale_log_weights(w + p * np, np, m->resample_mode);
    }
    if (any_data && early_exit(log_likelihood, m->n_sets, min_log_likelihood, n_min)) {
      for (size_t p = 0; p < m->n_sets; ++p) log_likelihood[p] = -INFINITY;
      stopped = 1;
      ++done;
      break;
    }
    /* resample set by set; a set whose datum is NA keeps its particles */
    uint64_t *fs = m->rng + 4 * nt;
    for (size_t pg = 0; pg < m->n_sets_global; ++pg) {
      const int mine = pg >= m->set_offset && pg < m->set_offset + m->n_sets;
      const size_t p = pg - m->set_offset;
      const int na = m->na_global ? m->na_global[row * m->n_sets_global + pg]
                                  : (mine && isnan(datum_for(m, row, p)[0]));
      if (na) {
        if (mine)
          for (size_t i = 0; i < np; ++i) kappa[p * np + i] = (int)(p * np + i);
        continue;
      }
      const double u = orc_random_real(fs, m->scrambler);
      if (!mine) continue;
      int *kp = kappa + p * np;
      orc_resample_weight(w + p * np, np, u, m->resample_mode, kp);
      for (size_t i = 0; i < np; ++i) kp[i] += (int)(p * np);
    }
    if (any_data) orc_reorder(m, kappa);
    if (trajectories)
      memcpy(hist_order + (done + 1) * nt, kappa, nt * sizeof(int));
    while (snap < n_snap && snapshot_times[snap] < m->data_time[row]) ++snap;
    if (snapshots && snap < n_snap && snapshot_times[snap] == m->data_time[row]) {
      memcpy(snapshots + snap * nt * m->n_state, m->y, nt * m->n_state * sizeof(double));
      ++snap;
    }
  }
  m->data_cursor = first + done;

  if (trajectories) {
    /* ancestor tracing as R/particle_filter.R:635-642; rows after an early
     * stop stay NA (tests/testthat/test-particle-filter.R:71-93) */
    const size_t T1 = n_rows + 1;
    for (size_t i = 0; i < T1 * nt * ni; ++i) trajectories[i] = NAN;
    const size_t filled = done + 1; /* time slices holding values */
    int *cur = (int *)malloc(nt * sizeof(int));
    for (size_t i = 0; i < nt; ++i) cur[i] = (int)i;
    for (size_t t = filled; t-- > 0;) {
      for (size_t i = 0; i < nt; ++i) {
        /* the slice of an early stop was never resampled: identity order */
        if (!(stopped && t == filled - 1)) cur[i] = hist_order[t * nt + (size_t)cur[i]];
        memcpy(trajectories + (t * nt + i) * ni,
               hist_value + (t * nt + (size_t)cur[i]) * ni, ni * sizeof(double));
      }
    }
    free(cur);
  }
  free(w); free(kappa); free(hist_value); free(hist_order);
  return done;
}
