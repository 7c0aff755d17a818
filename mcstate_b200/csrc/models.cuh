// The SIR-type example models (dust::dust_example("sir") / ("sirs")).
//
// Maths: vignettes/sir_models.Rmd:44-54; compare: the R twin at
// tests/testthat/helper-mcstate.R:7-22; state order S, I, R, cumulative,
// interval incidence (tests/testthat/test-particle-filter.R:14,223); defaults
// scripts/generate_vignette_data.R:11-12.  Draw order follows dust's example
// sources [SURVEY.md A.3].
//
// A model is a struct with kState / kPar / kData, and
//   derive()   per parameter set, once per update_state(pars): constants that
//              depend on the parameters only (kMaxDerived doubles in HBM);
//   table()    per parameter set: memo tables indexed by a compartment count
//              (pure functions of a small integer and the parameters);
//   context()  per thread, per launch: pointers to the above;
//   initial(), update(), compare()  the general form, state as doubles;
//   lean_load(), lean_update(), lean_store()  the same arithmetic for the usual
//              case -- every compartment a non-negative integer, the closed
//              population equal to the one the tables were built for -- with the
//              state held as int32 and the per-draw setup read from the tables.
// State lives in registers for all `rate` steps of an interval.  A memo table is
// only consulted where its value is bit-identical to evaluating the expression;
// anything else takes the general path.
#pragma once
#include "samplers.cuh"

namespace mcb {

constexpr int kMaxState = 8;
constexpr int kMaxPar = 8;
constexpr int kMaxDerived = 16;
constexpr int kMaxTables = 3;

// where a thread finds its parameter set's prepared data
struct SetData {
  const double *derived;   // [kMaxDerived]
  const double2 *tables;   // [kMaxTables][n_tab]
  int n_tab;
};

// ---- memo-table entries ----------------------------------------------------
// A draw with a probability fixed by the parameter set (recovery, waning):
// entry n = {(1-q)^n, r (n+1)}, or {-1, .} when count n is not a plain
// inversion draw with this probability.
__device__ __forceinline__ double2 const_binomial_entry(double p, double q, double qq, double r,
                                                        int n) {
  const double g = r * (double)(n + 1);
  const bool plain = p > 0.0 && p <= 0.5 && !((double)n * q >= 10.0) && g > 1e-280 && g < 1e280;
  return make_double2(plain ? pow_int(qq, n) : -1.0, g);
}
// Infection: entry I = {q, q/(1-q)} with q = p_SI(I) dt, {0, 0} when the
// probability is zero (no draw), {-1, .} when it is not in (0, 1/2] or r is
// near the ends of the exponent range.
__device__ __forceinline__ double2 infection_entry(double beta, double n_tot, double dt, int I) {
  const double p = (1.0 - det_exp(-beta * (double)I / n_tot)) * dt;
  if (p == 0.0) return make_double2(0.0, 0.0);
  const double r = p / (1.0 - p);
  const bool plain = p > 0.0 && p <= 0.5 && r > 1e-270 && r < 1e270;
  return make_double2(plain ? p : -1.0, r);
}

// y (doubles) -> int32 compartments; false unless every value is an integer in
// [0, 2^30) with a clear sign bit (-0.0 stays on the general path: its sign
// would be lost)
template <int N>
__device__ __forceinline__ bool to_counts(const double *y, int *yi) {
  bool ok = true;
#pragma unroll
  for (int s = 0; s < N; ++s) {
    const unsigned u = __double2uint_rz(y[s]);
    ok = ok & (__double2hiint(y[s]) >= 0) & ((double)u == y[s]) & (u < (1u << 30));
    yi[s] = (int)u;
  }
  return ok;
}

// Binomial(S, q) for a state-dependent probability: e = entry I of the set's
// infection table.  Returns the draw, or -1 (generator untouched) when the
// general code must take over: probability outside (0, 1/2], S q >= 10 (BTRS), or
// a walk that met one of its rare exits.
template <int SCR>
__device__ __forceinline__ int draw_infection_lean(Xoshiro &rng, int S, double2 e,
                                                   uint32_t s_inv) {
  if (S == 0 || e.x == 0.0) return 0;
  const double Sd1 = (double)(S + 1);
  if (e.x < 0.0 || (Sd1 - 1.0) * e.x >= 10.0) return -1;
  const int k = walk_lean<SCR>(rng, S, e.y, pow_int_lean(1.0 - e.x, S), e.y * Sd1, s_inv);
  if (k < 0) xo_rewind(rng);
  return k;
}

struct SirModel {
  static constexpr int kState = 5;
  static constexpr int kPar = 7;  // beta gamma I0 exp_noise S0 R0 freq
  static constexpr int kData = 1; // incidence
  static constexpr int kTables = 2;
  enum { D_BETA, D_DT, D_EXP_NOISE, D_FREQ, D_NTOT, D_P_IR, D_Q_IR, D_QQ_IR, D_R_IR, D_FLIP_IR };
  enum { T_IR, T_SI };

  __device__ __forceinline__ static void derive(const double *__restrict__ p, double *d) {
    const double dt = 1.0 / p[6];
    const double p_ir = (1.0 - det_exp(-p[1])) * dt;
    const double q = p_ir > 0.5 ? 1.0 - p_ir : p_ir;
    const double qq = 1.0 - q;
    d[D_BETA] = p[0];
    d[D_DT] = dt;
    d[D_EXP_NOISE] = p[3];
    d[D_FREQ] = p[6];
    d[D_NTOT] = p[4] + p[2] + p[5];  // S0 + I0 + R0: the closed population
    d[D_P_IR] = p_ir;
    d[D_Q_IR] = q;
    d[D_QQ_IR] = qq;
    d[D_R_IR] = q / qq;
    d[D_FLIP_IR] = p_ir > 0.5 ? 1.0 : 0.0;
  }
  // entry n of table t
  __device__ __forceinline__ static double2 table(int t, int n, const double *d) {
    if (t == T_IR) return const_binomial_entry(d[D_P_IR], d[D_Q_IR], d[D_QQ_IR], d[D_R_IR], n);
    return infection_entry(d[D_BETA], d[D_NTOT], d[D_DT], n);
  }

  struct Ctx {
    const double *derived;
    BinomialConst ir;
    double beta, dt, n_tot;
    const double2 *tab_ir, *tab_si;
    int n_tab;
    uint32_t freq;
  };
  __device__ __forceinline__ static Ctx context(const SetData &sd) {
    const double *d = sd.derived;
    Ctx c;
    c.derived = d;
    c.n_tab = sd.n_tab;
    c.beta = __ldg(d + D_BETA);
    c.dt = __ldg(d + D_DT);
    c.n_tot = __ldg(d + D_NTOT);
    c.freq = (uint32_t)__ldg(d + D_FREQ);
    c.ir.p = __ldg(d + D_P_IR);
    c.ir.q = __ldg(d + D_Q_IR);
    c.ir.qq = __ldg(d + D_QQ_IR);
    c.ir.r = __ldg(d + D_R_IR);
    c.ir.flip = __ldg(d + D_FLIP_IR) != 0.0;
    c.ir.n_tab = sd.n_tab;
    c.tab_ir = sd.tables + (size_t)T_IR * sd.n_tab;
    c.tab_si = sd.tables + (size_t)T_SI * sd.n_tab;
    return c;
  }
  __device__ __forceinline__ static void initial(const double *__restrict__ p, double *y) {
    y[0] = __ldg(p + 4); y[1] = __ldg(p + 2); y[2] = __ldg(p + 5); y[3] = 0.0; y[4] = 0.0;
  }
  template <int SCR>
  __device__ __forceinline__ static void update(uint64_t time, double *y, Xoshiro &rng,
                                                const Ctx &c, bool det) {
    const double S = y[0], I = y[1], R = y[2];
    const double N = S + I + R;
    const double p_SI_dt = (1.0 - det_exp(-c.beta * I / N)) * c.dt;
    const double n_IR = draw_binomial_const<SCR>(rng, det, I, c.ir);
    const double n_SI = draw_binomial<SCR>(rng, det, S, p_SI_dt);
    y[0] = S - n_SI;
    y[1] = I + n_SI - n_IR;
    y[2] = R + n_IR;
    y[3] = y[3] + n_SI;
    y[4] = (time % c.freq == 0) ? n_SI : y[4] + n_SI;
  }
  template <int SCR>
  __device__ __forceinline__ static double compare(const double *y, double observed,
                                                   double lgamma_obs1, Xoshiro &rng,
                                                   const Ctx &c, bool det) {
    const double rate = __ldg(c.derived + D_EXP_NOISE);
    const double noise = det ? 1.0 / rate : draw_exponential<SCR>(rng, rate);
    return density_poisson_log(observed, lgamma_obs1, y[4] + noise);
  }

  // ---- the lean path ------------------------------------------------------
  // Register diet: the hot loop keeps r = q/(1-q) of the recovery draw, one
  // pointer to the set's tables and four counts (R is implied by the closed
  // population); it makes no calls.  A step that meets anything unusual is
  // abandoned with state and generator as they were at its start.
  struct Lean {
    double r_ir;
    uint32_t off;  // entry offset of the set's tables: set * kMaxTables * n_tab
  };
  static constexpr int kLeanState = 4;  // S, I, cumulative, interval incidence
  __device__ __forceinline__ static bool lean_load(const SetData &sd, const double *y, int *yi,
                                                   Lean &c) {
    c.r_ir = __ldg(sd.derived + D_R_IR);
    int all[kState];
    if (!to_counts<kState>(y, all)) return false;
    yi[0] = all[0]; yi[1] = all[1]; yi[2] = all[3]; yi[3] = all[4];
    // the tables are built for N = n_tot and cover counts < n_tab
    const double n_tot = __ldg(sd.derived + D_NTOT);
    return y[0] + y[1] + y[2] == n_tot && n_tot < (double)sd.n_tab;
  }
  __device__ __forceinline__ static void lean_store(const SetData &sd, const int *yi, double *y) {
    const double n_tot = __ldg(sd.derived + D_NTOT);
    y[0] = (double)yi[0];
    y[1] = (double)yi[1];
    y[2] = n_tot - (double)(yi[0] + yi[1]);   // exact: integers below 2^31
    y[3] = (double)yi[2];
    y[4] = (double)yi[3];
  }
  // One model step; `reset`: it starts a new incidence interval (time % freq == 0).
  // false: nothing changed, the general code must run this step.
  template <int SCR>
  __device__ __forceinline__ static bool lean_update(bool reset, int *y, Xoshiro &rng,
                                                     const Lean &c, const double2 *__restrict__ tables,
                                                     uint32_t n_tab, uint32_t s_inv) {
    const int S = y[0], I = y[1];
    int n_IR = 0;
    if (I != 0) {
      n_IR = draw_binomial_const_lean<SCR>(rng, I, c.r_ir, __ldg(tables + (c.off + T_IR * n_tab + (uint32_t)I)), s_inv);
      if (n_IR < 0) return false;
    }
    const int n_SI = draw_infection_lean<SCR>(rng, S, __ldg(tables + (c.off + T_SI * n_tab + (uint32_t)I)), s_inv);
    if (n_SI < 0) {
      if (I != 0) xo_rewind(rng);  // the recovery draw's uniform
      return false;
    }
    y[0] = S - n_SI;
    y[1] = I + n_SI - n_IR;
    y[2] = y[2] + n_SI;
    y[3] = reset ? n_SI : y[3] + n_SI;
    return true;
  }
};

struct SirsModel {
  static constexpr int kState = 4;
  static constexpr int kPar = 8;  // alpha beta gamma I0 exp_noise S0 R0 freq
  static constexpr int kData = 1;
  static constexpr int kTables = 3;
  enum { D_BETA, D_DT, D_EXP_NOISE, D_FREQ, D_NTOT, D_P_IR, D_Q_IR, D_QQ_IR, D_R_IR, D_FLIP_IR,
         D_P_RS, D_Q_RS, D_QQ_RS, D_R_RS, D_FLIP_RS };
  enum { T_IR, T_SI, T_RS };

  __device__ __forceinline__ static void derive(const double *__restrict__ p, double *d) {
    const double dt = 1.0 / p[7];
    d[D_BETA] = p[1];
    d[D_DT] = dt;
    d[D_EXP_NOISE] = p[4];
    d[D_FREQ] = p[7];
    d[D_NTOT] = p[5] + p[3] + p[6];
    const double p_ir = (1.0 - det_exp(-p[2])) * dt;
    const double q1 = p_ir > 0.5 ? 1.0 - p_ir : p_ir;
    d[D_P_IR] = p_ir; d[D_Q_IR] = q1; d[D_QQ_IR] = 1.0 - q1; d[D_R_IR] = q1 / (1.0 - q1);
    d[D_FLIP_IR] = p_ir > 0.5 ? 1.0 : 0.0;
    const double p_rs = (1.0 - det_exp(-p[0])) * dt;
    const double q2 = p_rs > 0.5 ? 1.0 - p_rs : p_rs;
    d[D_P_RS] = p_rs; d[D_Q_RS] = q2; d[D_QQ_RS] = 1.0 - q2; d[D_R_RS] = q2 / (1.0 - q2);
    d[D_FLIP_RS] = p_rs > 0.5 ? 1.0 : 0.0;
  }
  __device__ __forceinline__ static double2 table(int t, int n, const double *d) {
    if (t == T_IR) return const_binomial_entry(d[D_P_IR], d[D_Q_IR], d[D_QQ_IR], d[D_R_IR], n);
    if (t == T_RS) return const_binomial_entry(d[D_P_RS], d[D_Q_RS], d[D_QQ_RS], d[D_R_RS], n);
    return infection_entry(d[D_BETA], d[D_NTOT], d[D_DT], n);
  }

  struct Ctx {
    const double *derived;
    BinomialConst ir, rs;
    double beta, dt, n_tot;
    const double2 *tab_ir, *tab_si, *tab_rs;
    int n_tab;
    uint32_t freq;
  };
  __device__ __forceinline__ static Ctx context(const SetData &sd) {
    const double *d = sd.derived;
    Ctx c;
    c.derived = d;
    c.n_tab = sd.n_tab;
    c.beta = __ldg(d + D_BETA);
    c.dt = __ldg(d + D_DT);
    c.n_tot = __ldg(d + D_NTOT);
    c.freq = (uint32_t)__ldg(d + D_FREQ);
    c.ir.p = __ldg(d + D_P_IR); c.ir.q = __ldg(d + D_Q_IR); c.ir.qq = __ldg(d + D_QQ_IR);
    c.ir.r = __ldg(d + D_R_IR); c.ir.flip = __ldg(d + D_FLIP_IR) != 0.0;
    c.ir.n_tab = sd.n_tab;
    c.rs.p = __ldg(d + D_P_RS); c.rs.q = __ldg(d + D_Q_RS); c.rs.qq = __ldg(d + D_QQ_RS);
    c.rs.r = __ldg(d + D_R_RS); c.rs.flip = __ldg(d + D_FLIP_RS) != 0.0;
    c.rs.n_tab = sd.n_tab;
    c.tab_ir = sd.tables + (size_t)T_IR * sd.n_tab;
    c.tab_si = sd.tables + (size_t)T_SI * sd.n_tab;
    c.tab_rs = sd.tables + (size_t)T_RS * sd.n_tab;
    return c;
  }
  __device__ __forceinline__ static void initial(const double *__restrict__ p, double *y) {
    y[0] = __ldg(p + 5); y[1] = __ldg(p + 3); y[2] = __ldg(p + 6); y[3] = 0.0;
  }
  template <int SCR>
  __device__ __forceinline__ static void update(uint64_t time, double *y, Xoshiro &rng,
                                                const Ctx &c, bool det) {
    const double S = y[0], I = y[1], R = y[2];
    const double N = S + I + R;
    const double p_SI_dt = (1.0 - det_exp(-c.beta * I / N)) * c.dt;
    const double n_SI = draw_binomial<SCR>(rng, det, S, p_SI_dt);
    const double n_IR = draw_binomial_const<SCR>(rng, det, I, c.ir);
    const double n_RS = draw_binomial_const<SCR>(rng, det, R, c.rs);
    y[0] = S - n_SI + n_RS;
    y[1] = I + n_SI - n_IR;
    y[2] = R + n_IR - n_RS;
    y[3] = (time % c.freq == 0) ? n_SI : y[3] + n_SI;
  }
  template <int SCR>
  __device__ __forceinline__ static double compare(const double *y, double observed,
                                                   double lgamma_obs1, Xoshiro &rng,
                                                   const Ctx &c, bool det) {
    const double rate = __ldg(c.derived + D_EXP_NOISE);
    const double noise = det ? 1.0 / rate : draw_exponential<SCR>(rng, rate);
    return density_poisson_log(observed, lgamma_obs1, y[3] + noise);
  }

  struct Lean {
    double r_ir, r_rs;
    uint32_t off;
  };
  static constexpr int kLeanState = 4;
  __device__ __forceinline__ static bool lean_load(const SetData &sd, const double *y, int *yi,
                                                   Lean &c) {
    c.r_ir = __ldg(sd.derived + D_R_IR);
    c.r_rs = __ldg(sd.derived + D_R_RS);
    if (!to_counts<kState>(y, yi)) return false;
    const double n_tot = __ldg(sd.derived + D_NTOT);
    return y[0] + y[1] + y[2] == n_tot && n_tot < (double)sd.n_tab;
  }
  __device__ __forceinline__ static void lean_store(const SetData &, const int *yi, double *y) {
#pragma unroll
    for (int s = 0; s < kState; ++s) y[s] = (double)yi[s];
  }
  template <int SCR>
  __device__ __forceinline__ static bool lean_update(bool reset, int *y, Xoshiro &rng,
                                                     const Lean &c, const double2 *__restrict__ tables,
                                                     uint32_t n_tab, uint32_t s_inv) {
    const int S = y[0], I = y[1], R = y[2];
    const int n_SI = draw_infection_lean<SCR>(rng, S, __ldg(tables + (c.off + T_SI * n_tab + (uint32_t)I)), s_inv);
    if (n_SI < 0) return false;
    // did the infection draw take a uniform?  (it returns 0 without one for S = 0 or p = 0)
    const bool si_drew = S != 0 && I != 0 && __ldg(tables + (c.off + T_SI * n_tab + (uint32_t)I)).x != 0.0;
    int n_IR = 0;
    if (I != 0) {
      n_IR = draw_binomial_const_lean<SCR>(rng, I, c.r_ir, __ldg(tables + (c.off + T_IR * n_tab + (uint32_t)I)), s_inv);
      if (n_IR < 0) {
        if (si_drew) xo_rewind(rng);
        return false;
      }
    }
    int n_RS = 0;
    if (R != 0) {
      n_RS = draw_binomial_const_lean<SCR>(rng, R, c.r_rs, __ldg(tables + (c.off + T_RS * n_tab + (uint32_t)R)), s_inv);
      if (n_RS < 0) {
        if (I != 0) xo_rewind(rng);
        if (si_drew) xo_rewind(rng);
        return false;
      }
    }
    y[0] = S - n_SI + n_RS;
    y[1] = I + n_SI - n_IR;
    y[2] = R + n_IR - n_RS;
    y[3] = reset ? n_SI : y[3] + n_SI;
    return true;
  }
};

}  // namespace mcb
