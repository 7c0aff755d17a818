// Reproducible fp64 exp / log for the model and weight kernels.
//
// The reference asserts that its CPU and GPU back ends return bit-identical
// log-likelihoods (tests/testthat/test-particle-filter.R:1480-1504).  CUDA's
// exp()/log() and a host libm differ in the last place now and then, so these
// two functions use nothing but IEEE add/mul/div, explicit fma and integer
// ops: any IEEE machine computing the same sequence gets the same bits.  The
// file is compiled with -fmad=false, so a fused multiply-add happens only
// where fma() is written.  Both are within 1 ulp of the correctly rounded
// result.
#pragma once
#include <cstdint>

namespace mcb {

__device__ __forceinline__ double bits_to_double(uint64_t u) { return __longlong_as_double((long long)u); }
__device__ __forceinline__ uint64_t double_to_bits(double x) { return (uint64_t)__double_as_longlong(x); }

// The coefficients live in constant memory: a 64-bit immediate costs two uniform
// moves per use, a constant-bank operand none.
__constant__ double c_exp_poly[12] = {
    1.6059043836821613e-10, 2.08767569878681e-09,   2.505210838544172e-08,
    2.755731922398589e-07,  2.7557319223985893e-06, 2.48015873015873e-05,
    1.984126984126984e-04,  1.388888888888889e-03,  8.333333333333333e-03,
    4.1666666666666664e-02, 1.6666666666666666e-01, 0.5};
__constant__ double c_exp_red[3] = {1.4426950408889634074, 6.93147180559945286e-01,
                                    2.31904681384629956e-17};
__constant__ double c_log_poly[7] = {
    1.531383769920937332e-01, 2.222219843214978396e-01, 3.999999999940941908e-01,
    1.479819860511658591e-01, 1.818357216161805012e-01, 2.857142874366239149e-01,
    6.666666666666735130e-01};
__constant__ double c_log_ln2[2] = {1.90821492927058770002e-10, 6.93147180369123816490e-01};

// exp(x): k = rint(x / ln 2); Cody-Waite reduction r = x - k ln2 with ln2 split
// in two doubles (the fma keeps k*ln2_hi exact); degree-13 Taylor polynomial on
// |r| <= ln2/2 (truncation < 5e-18); scaling by 2^k in two steps so a
// subnormal result is rounded once.
__device__ __forceinline__ double det_exp(double x) {
  if (x != x) return x;
  if (x > 709.782712893384) return __longlong_as_double(0x7ff0000000000000LL);
  if (x < -745.1332191019412) return 0.0;
  const double kd = rint(x * c_exp_red[0]);
  double r = fma(-kd, c_exp_red[1], x);
  r = fma(-kd, c_exp_red[2], r);
  double p = c_exp_poly[0];
#pragma unroll
  for (int j = 1; j < 12; ++j) p = fma(p, r, c_exp_poly[j]);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  const int k = (int)kd;
  const int k1 = k >> 1, k2 = k - k1;
  const double s1 = bits_to_double((uint64_t)(1023 + k1) << 52);
  const double s2 = bits_to_double((uint64_t)(1023 + k2) << 52);
  return (p * s1) * s2;
}

// det_exp for x <= 0 (or -inf; a NaN gives 0): the same operations on the same values --
// the same bits -- without the branches: the argument is clamped where the result is
// zero anyway, and the underflow rule is a select.  For the weights, exp(logw - max).
__device__ __forceinline__ double det_exp_nonpos(double x) {
  const double xc = fmax(x, -746.0);
  const double kd = rint(xc * c_exp_red[0]);
  double r = fma(-kd, c_exp_red[1], xc);
  r = fma(-kd, c_exp_red[2], r);
  double p = c_exp_poly[0];
#pragma unroll
  for (int j = 1; j < 12; ++j) p = fma(p, r, c_exp_poly[j]);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  const int k = (int)kd;
  const int k1 = k >> 1, k2 = k - k1;
  const double s1 = bits_to_double((uint64_t)(1023 + k1) << 52);
  const double s2 = bits_to_double((uint64_t)(1023 + k2) << 52);
  const double val = (p * s1) * s2;
  return xc < -745.1332191019412 ? 0.0 : val;
}

// log(x): x = 2^k m with m in [sqrt(1/2), sqrt(2)); f = m - 1, s = f/(2+f);
// log(1+f) = f - f^2/2 + s (f^2/2 + R(s^2)) with the classic 7-term minimax R;
// ln2 is split so k*ln2_hi is exact.
__device__ __forceinline__ double det_log(double x) {
  uint64_t ix = double_to_bits(x);
  uint32_t hx = (uint32_t)(ix >> 32);
  int k = 0;
  if (hx < 0x00100000u || (hx >> 31)) {
    if ((ix << 1) == 0) return __longlong_as_double(0xfff0000000000000LL);
    if (hx >> 31) return __longlong_as_double(0x7ff8000000000000LL);
    k -= 54;
    x *= 0x1p54;
    ix = double_to_bits(x);
    hx = (uint32_t)(ix >> 32);
  } else if (hx >= 0x7ff00000u) {
    return x;
  } else if (hx == 0x3ff00000u && (ix << 32) == 0) {
    return 0.0;
  }
  hx += 0x3ff00000u - 0x3fe6a09eu;
  k += (int)(hx >> 20) - 0x3ff;
  hx = (hx & 0x000fffffu) + 0x3fe6a09eu;
  x = bits_to_double(((uint64_t)hx << 32) | (ix & 0xffffffffULL));

  const double f = x - 1.0;
  const double hfsq = 0.5 * f * f;
  const double s = f / (2.0 + f);
  const double z = s * s;
  const double w = z * z;
  const double t1 = w * fma(w, fma(w, c_log_poly[0], c_log_poly[1]), c_log_poly[2]);
  const double t2 = z * fma(w, fma(w, fma(w, c_log_poly[3], c_log_poly[4]), c_log_poly[5]),
                            c_log_poly[6]);
  const double R = t2 + t1;
  const double dk = (double)k;
  return s * (hfsq + R) + dk * c_log_ln2[0] - hfsq + f + dk * c_log_ln2[1];
}

// log(k!) for an integer-valued k >= 0 (the Poisson sampler's acceptance test): a
// table below 16, from there Stirling's series for lgamma(x), x = k + 1, with four
// correction terms (truncation below 1e-14 at x = 17).  IEEE operations and det_log
// only: the same bits as the oracle's orc_lfact, which no libm lgamma would give.
__constant__ double c_log_fact[16] = {
    0.0, 0.0, 0.693147180559945, 1.7917594692280554, 3.178053830347945, 4.787491742782047,
    6.579251212010102, 8.525161361065415, 10.604602902745249, 12.801827480081467,
    15.104412573075514, 17.502307845873887, 19.987214495661885, 22.55216385312342,
    25.191221182738683, 27.89927138384089};
__device__ __forceinline__ double det_lfact(double k) {
  if (k < 16.0) return c_log_fact[(int)k];
  const double x = k + 1.0;
  const double xi = 1.0 / x, x2 = xi * xi;
  const double ser =
      (1.0 / 12.0 - x2 * (1.0 / 360.0 - x2 * (1.0 / 1260.0 - x2 * (1.0 / 1680.0)))) * xi;
  return (x - 0.5) * det_log(x) - x + 0.91893853320467274178 + ser;
}

// cos(2 pi u) for u in [0, 1).  Exact reduction: x = 4u = q + r (quadrant q, r in
// [0, 1)), r <- 1 - r above one half, so only a = (pi/2) r in [0, pi/4] is
// rounded; then the fdlibm kernel polynomials for sin and cos with explicit fma.
// cos(pi/2 (q + r)) = {cos, -sin, -cos, sin}(pi/2 r) for q = 0..3.
__constant__ double c_sin_poly[6] = {1.58969099521155010221e-10, -2.50507602534068634195e-08,
                                     2.75573137070700676789e-06, -1.98412698298579493134e-04,
                                     8.33333333332248946124e-03, -1.66666666666666324348e-01};
__constant__ double c_cos_poly[6] = {-1.13596475577881948265e-11, 2.08757232129817482790e-09,
                                     -2.75573143513906633035e-07, 2.48015872894767294178e-05,
                                     -1.38888888888741095749e-03, 4.16666666666666019037e-02};
__device__ __forceinline__ double det_cos2pi(double u) {
  const double x = 4.0 * u;
  const double qd = floor(x);
  double r = x - qd;
  const int q = (int)qd;
  const bool swap = r > 0.5;
  if (swap) r = 1.0 - r;
  const double a = r * 1.57079632679489661923;
  const double z = a * a;
  const bool use_sin = ((q & 1) != 0) != swap;
  double val;
  if (use_sin) {
    double t = fma(z, c_sin_poly[0], c_sin_poly[1]);
#pragma unroll
    for (int j = 2; j < 6; ++j) t = fma(z, t, c_sin_poly[j]);
    val = fma(z * a, t, a);
  } else {
    double t = fma(z, c_cos_poly[0], c_cos_poly[1]);
#pragma unroll
    for (int j = 2; j < 6; ++j) t = fma(z, t, c_cos_poly[j]);
    val = 1.0 - (0.5 * z - z * (z * t));
  }
  return (q == 1 || q == 2) ? -val : val;
}

}  // namespace mcb
