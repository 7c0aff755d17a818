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
//   walk_row(), inf_entry()  per parameter set: what the memo tables are
//              built from -- a walk table per draw whose probability the
//              parameters fix (samplers.cuh), and one entry per infectious count
//              {q, q/(1-q)} for the state-dependent draw;
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
constexpr int kMaxPar = 12;
constexpr int kMaxDerived = 24;
constexpr int kMaxConst = 2;     // draws with a parameter-only probability, per model

// where a thread finds its parameter set's prepared data
struct SetData {
  const double *derived;   // [kMaxDerived]
  int n_tab;
};

// The memo tables as the lean path sees them: warp-uniform values that stay in the
// kernel's parameter bank (no registers spent on them between uses).
//   inf  [n_sets][n_tab]                     entries of the infection table
//   walk [n_sets][n_const][n_wrows][kWalkK]  walk tables (samplers.cuh)
struct LeanTabs {
  const double4 *inf;
  const double *walk;
  uint32_t n_tab, n_wrows;
};

// ---- memo-table entries ----------------------------------------------------
// Infection: entry I = {q, q/(1-q), log(1-q), 0} with q = p_SI(I) dt; q = 0 when the
// probability is zero (no draw), q = -1 when it is not in (0, 1/2] or r is near the
// ends of the exponent range.  q and r are the values the exact walk uses; the
// logarithm feeds the fast walk's first term (samplers.cuh).
__device__ __forceinline__ double4 infection_entry(double beta, double n_tot, double dt, int I) {
  const double p = (1.0 - det_exp(-beta * (double)I / n_tot)) * dt;
  if (p == 0.0) return make_double4(0.0, 0.0, 0.0, 0.0);
  const double r = p / (1.0 - p);
  const bool plain = p > 0.0 && p <= 0.5 && r > 1e-270 && r < 1e270;
  return make_double4(plain ? p : -1.0, r, plain ? det_log(1.0 - p) : 0.0, 0.0);
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
// a walk the certificate of walk_fast does not cover.
template <int SCR>
__device__ __forceinline__ int draw_infection_lean(Xoshiro &rng, int S, double4 e) {
  if (S == 0 || e.x == 0.0) return 0;
  if (e.x < 0.0 || (double)S * e.x >= 10.0) return -1;
  const int k = walk_fast<SCR>(rng, S, e.x, e.y, e.z);
  if (k < 0) xo_rewind(rng);
  return k;
}

// Binomial(n, p) for the set's constant draw `c`: row n of its walk table, or -1
// (generator untouched) when the table has no such row.
// row n of constant draw `set_const`'s walk table (n clamped into the table: the caller
// checks n < n_wrows before it uses what was loaded)
__device__ __forceinline__ const double *walk_row_ptr(int n, uint32_t set_const,
                                                      const LeanTabs &t) {
  return t.walk + (size_t)(set_const * t.n_wrows + min((uint32_t)n, t.n_wrows - 1)) * kWalkK;
}
template <int SCR>
__device__ __forceinline__ int draw_const_lean(Xoshiro &rng, int n, const double *row, double4 F,
                                               const LeanTabs &t) {
  if ((uint32_t)n >= t.n_wrows) return -1;
  return draw_binomial_table<SCR>(rng, row, F);
}

struct SirModel {
  static constexpr bool kHasLean = true;
  static constexpr int kState = 5;
  static constexpr int kPar = 7;  // beta gamma I0 exp_noise S0 R0 freq
  static constexpr int kData = 1; // incidence
  static constexpr int kConst = 1;  // recovery
  enum { D_BETA, D_DT, D_EXP_NOISE, D_FREQ, D_NTOT, D_P_IR, D_Q_IR, D_QQ_IR, D_R_IR, D_FLIP_IR,
         D_INV_EXP_NOISE };
  enum { C_IR };

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
    d[D_INV_EXP_NOISE] = reciprocal_for(p[3]);
  }
  // row n of the walk table of constant draw c; row I of the infection table
  __device__ __forceinline__ static void walk_row(int, int n, const double *d, double *row) {
    fill_walk_row(d[D_P_IR], d[D_Q_IR], d[D_QQ_IR], d[D_R_IR], n, row);
  }
  __device__ __forceinline__ static double4 inf_entry(int I, const double *d) {
    return infection_entry(d[D_BETA], d[D_NTOT], d[D_DT], I);
  }

  struct Ctx {
    const double *derived;
    BinomialConst ir;
    double beta, dt, n_tot;
    uint32_t freq;
  };
  __device__ __forceinline__ static Ctx context(const SetData &sd) {
    const double *d = sd.derived;
    Ctx c;
    c.derived = d;
    c.beta = __ldg(d + D_BETA);
    c.dt = __ldg(d + D_DT);
    c.n_tot = __ldg(d + D_NTOT);
    c.freq = (uint32_t)__ldg(d + D_FREQ);
    c.ir.p = __ldg(d + D_P_IR);
    c.ir.q = __ldg(d + D_Q_IR);
    c.ir.qq = __ldg(d + D_QQ_IR);
    c.ir.r = __ldg(d + D_R_IR);
    c.ir.flip = __ldg(d + D_FLIP_IR) != 0.0;
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
    const double noise = det ? 1.0 / rate
                             : draw_exponential_fixed<SCR>(rng, rate, __ldg(c.derived + D_INV_EXP_NOISE));
    return density_poisson_log(observed, lgamma_obs1, y[4] + noise);
  }

  // ---- the lean path ------------------------------------------------------
  // Register diet: the hot loop works on S, I and M; it makes no calls.  R is implied by
  // the closed population, and both incidence counters by S, which only ever falls by
  // the infections: cumulative + S never changes (C, set once), and the interval
  // incidence is M - S with M = S + incidence, reset to the S of the moment when a new
  // interval starts.  A step that meets anything unusual is abandoned with state and
  // generator as they were at its start.
  struct Lean {
    uint32_t set;
  };
  static constexpr int kLeanState = 4;  // S, I, M = S + interval incidence, C = S + cumulative
  __device__ __forceinline__ static bool lean_load(const SetData &sd, const double *y, int *yi,
                                                   Lean &c) {
    int all[kState];
    if (!to_counts<kState>(y, all)) return false;
    yi[0] = all[0]; yi[1] = all[1]; yi[2] = all[0] + all[4]; yi[3] = all[0] + all[3];
    // the tables are built for N = n_tot and cover counts < n_tab
    const double n_tot = __ldg(sd.derived + D_NTOT);
    return y[0] + y[1] + y[2] == n_tot && n_tot < (double)sd.n_tab;
  }
  __device__ __forceinline__ static void lean_store(const SetData &sd, const int *yi, double *y) {
    const double n_tot = __ldg(sd.derived + D_NTOT);
    y[0] = (double)yi[0];
    y[1] = (double)yi[1];
    y[2] = n_tot - (double)(yi[0] + yi[1]);   // exact: integers below 2^31
    y[3] = (double)(yi[3] - yi[0]);
    y[4] = (double)(yi[2] - yi[0]);
  }
  // One model step; `reset`: it starts a new incidence interval (time % freq == 0).
  // false: nothing changed, the general code must run this step.
  template <int SCR>
  __device__ __forceinline__ static bool lean_update(bool reset, int *y, Xoshiro &rng,
                                                     const Lean &c, const LeanTabs &t) {
    const int S = y[0], I = y[1];
    // both draws' table entries depend on the state at the start of the step alone:
    // ask for them together
    const double *row_ir = walk_row_ptr(I, c.set * kConst + C_IR, t);
    const double4 e_si =
        ldg_f64x4(reinterpret_cast<const double *>(t.inf + (c.set * t.n_tab + (uint32_t)I)));
    const double4 F_ir = ldg_f64x4(row_ir);
    int n_IR = 0;
    if (I != 0) {
      n_IR = draw_const_lean<SCR>(rng, I, row_ir, F_ir, t);
      if (n_IR < 0) return false;
    }
    const int n_SI = draw_infection_lean<SCR>(rng, S, e_si);
    if (n_SI < 0) {
      if (I != 0) xo_rewind(rng);  // the recovery draw's uniform
      return false;
    }
    y[0] = S - n_SI;
    y[1] = I + n_SI - n_IR;
    if (reset) y[2] = S;
    return true;
  }
};

struct SirsModel {
  static constexpr bool kHasLean = true;
  static constexpr int kState = 4;
  static constexpr int kPar = 8;  // alpha beta gamma I0 exp_noise S0 R0 freq
  static constexpr int kData = 1;
  static constexpr int kConst = 2;  // recovery, waning
  enum { D_BETA, D_DT, D_EXP_NOISE, D_FREQ, D_NTOT, D_P_IR, D_Q_IR, D_QQ_IR, D_R_IR, D_FLIP_IR,
         D_P_RS, D_Q_RS, D_QQ_RS, D_R_RS, D_FLIP_RS };
  enum { C_IR, C_RS };

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
  __device__ __forceinline__ static void walk_row(int c, int n, const double *d, double *row) {
    if (c == C_IR) fill_walk_row(d[D_P_IR], d[D_Q_IR], d[D_QQ_IR], d[D_R_IR], n, row);
    else fill_walk_row(d[D_P_RS], d[D_Q_RS], d[D_QQ_RS], d[D_R_RS], n, row);
  }
  __device__ __forceinline__ static double4 inf_entry(int I, const double *d) {
    return infection_entry(d[D_BETA], d[D_NTOT], d[D_DT], I);
  }

  struct Ctx {
    const double *derived;
    BinomialConst ir, rs;
    double beta, dt, n_tot;
    uint32_t freq;
  };
  __device__ __forceinline__ static Ctx context(const SetData &sd) {
    const double *d = sd.derived;
    Ctx c;
    c.derived = d;
    c.beta = __ldg(d + D_BETA);
    c.dt = __ldg(d + D_DT);
    c.n_tot = __ldg(d + D_NTOT);
    c.freq = (uint32_t)__ldg(d + D_FREQ);
    c.ir.p = __ldg(d + D_P_IR); c.ir.q = __ldg(d + D_Q_IR); c.ir.qq = __ldg(d + D_QQ_IR);
    c.ir.r = __ldg(d + D_R_IR); c.ir.flip = __ldg(d + D_FLIP_IR) != 0.0;
    c.rs.p = __ldg(d + D_P_RS); c.rs.q = __ldg(d + D_Q_RS); c.rs.qq = __ldg(d + D_QQ_RS);
    c.rs.r = __ldg(d + D_R_RS); c.rs.flip = __ldg(d + D_FLIP_RS) != 0.0;
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
    uint32_t set;
  };
  static constexpr int kLeanState = 4;
  __device__ __forceinline__ static bool lean_load(const SetData &sd, const double *y, int *yi,
                                                   Lean &c) {
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
                                                     const Lean &c, const LeanTabs &t) {
    const int S = y[0], I = y[1], R = y[2];
    const double4 e_si =
        ldg_f64x4(reinterpret_cast<const double *>(t.inf + (c.set * t.n_tab + (uint32_t)I)));
    const int n_SI = draw_infection_lean<SCR>(rng, S, e_si);
    if (n_SI < 0) return false;
    // did the infection draw take a uniform?  (it returns 0 without one for S = 0 or p = 0)
    const bool si_drew = S != 0 && e_si.x != 0.0;
    int n_IR = 0;
    if (I != 0) {
      const double *row = walk_row_ptr(I, c.set * kConst + C_IR, t);
      n_IR = draw_const_lean<SCR>(rng, I, row, ldg_f64x4(row), t);
      if (n_IR < 0) {
        if (si_drew) xo_rewind(rng);
        return false;
      }
    }
    int n_RS = 0;
    if (R != 0) {
      const double *row = walk_row_ptr(R, c.set * kConst + C_RS, t);
      n_RS = draw_const_lean<SCR>(rng, R, row, ldg_f64x4(row), t);
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

// dust::dust_example("volatility") (tests/testthat/helper-mcstate.R:98-146): one
// real-valued state, x' = alpha x + sigma N(0, 1), observed ~ N(gamma x, tau).
// Linear and Gaussian, so the Kalman filter gives the exact likelihood the
// particle filter estimates (helper-mcstate.R:116-142, test-if2.R:35-64).  No
// integer compartments: there is no lean path and no memo table.
struct VolatilityModel {
  static constexpr bool kHasLean = false;
  static constexpr int kState = 1;
  static constexpr int kPar = 5;  // alpha sigma gamma tau x0
  static constexpr int kData = 1; // observed
  static constexpr int kConst = 0;
  enum { D_ALPHA, D_SIGMA, D_GAMMA, D_TAU, D_LOG_TAU };

  __device__ __forceinline__ static void derive(const double *__restrict__ p, double *d) {
    d[D_ALPHA] = p[0];
    d[D_SIGMA] = p[1];
    d[D_GAMMA] = p[2];
    d[D_TAU] = p[3];
    d[D_LOG_TAU] = det_log(p[3]);
  }
  struct Ctx {
    double alpha, sigma, gamma, tau, log_tau;
  };
  __device__ __forceinline__ static Ctx context(const SetData &sd) {
    const double *d = sd.derived;
    return Ctx{__ldg(d + D_ALPHA), __ldg(d + D_SIGMA), __ldg(d + D_GAMMA), __ldg(d + D_TAU),
               __ldg(d + D_LOG_TAU)};
  }
  __device__ __forceinline__ static void initial(const double *__restrict__ p, double *y) {
    y[0] = __ldg(p + 4);
  }
  template <int SCR>
  __device__ __forceinline__ static void update(uint64_t, double *y, Xoshiro &rng, const Ctx &c,
                                                bool det) {
    const double z = det ? 0.0 : draw_normal<SCR>(rng, 0.0, 1.0);
    y[0] = c.alpha * y[0] + c.sigma * z;
  }
  template <int SCR>
  __device__ __forceinline__ static double compare(const double *y, double observed, double,
                                                   Xoshiro &, const Ctx &c, bool) {
    return density_normal_log(observed, c.gamma * y[0], c.tau, c.log_tau);
  }
};

// dust::dust_example("walk"): a random walk, x' = x + N(0, sd), with NO compiled
// compare -- the reference uses it to check that a filter without a compare function
// refuses such a model (tests/testthat/test-particle-filter.R:579-586); with an R-level
// compare it runs through the filter's R loop (model$run, $state, $reorder).
struct WalkModel {
  static constexpr bool kHasLean = false;
  static constexpr int kState = 1;
  static constexpr int kPar = 1;  // sd
  static constexpr int kData = 0;
  static constexpr int kConst = 0;
  enum { D_SD };
  __device__ __forceinline__ static void derive(const double *__restrict__ p, double *d) {
    d[D_SD] = p[0];
  }
  struct Ctx {
    double sd;
  };
  __device__ __forceinline__ static Ctx context(const SetData &sd) {
    return Ctx{__ldg(sd.derived + D_SD)};
  }
  __device__ __forceinline__ static void initial(const double *__restrict__, double *y) {
    y[0] = 0.0;
  }
  template <int SCR>
  __device__ __forceinline__ static void update(uint64_t, double *y, Xoshiro &rng, const Ctx &c,
                                                bool det) {
    y[0] = y[0] + (det ? 0.0 : draw_normal<SCR>(rng, 0.0, c.sd));
  }
  template <int SCR>
  __device__ __forceinline__ static double compare(const double *, double, double, Xoshiro &,
                                                   const Ctx &, bool) {
    return 0.0;   // never weighed: mcb_set_data refuses a model without a compare
  }
};

}  // namespace mcb
