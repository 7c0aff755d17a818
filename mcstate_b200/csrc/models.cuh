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
//   initial(), update(), compare().
// State lives in registers for all `rate` steps of an interval.  A memo table is
// only consulted where its value is bit-identical to evaluating the expression
// (the guards are in update()); counts beyond the table take the direct path.
#pragma once
#include "samplers.cuh"

namespace mcb {

constexpr int kMaxState = 8;
constexpr int kMaxPar = 8;
constexpr int kMaxDerived = 16;
constexpr int kMaxTables = 3;

// where a thread finds its parameter set's prepared data
struct SetData {
  const double *derived;  // [kMaxDerived]
  const double *tables;   // [kMaxTables][n_tab]
  int n_tab;
};

struct SirModel {
  static constexpr int kState = 5;
  static constexpr int kPar = 7;  // beta gamma I0 exp_noise S0 R0 freq
  static constexpr int kData = 1; // incidence
  static constexpr int kTables = 2;
  enum { D_BETA, D_DT, D_EXP_NOISE, D_FREQ, D_NTOT, D_P_IR, D_Q_IR, D_QQ_IR, D_R_IR, D_FLIP_IR };
  enum { T_F0_IR, T_PSI_DT };

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
  __device__ __forceinline__ static double table(int t, int n, const double *d) {
    if (t == T_F0_IR) return pow_int(d[D_QQ_IR], n);
    return (1.0 - det_exp(-d[D_BETA] * (double)n / d[D_NTOT])) * d[D_DT];
  }

  struct Ctx {
    const double *derived;
    BinomialConst ir;
    double beta, dt, n_tot;
    const double *psi_dt;
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
    c.ir.f0 = sd.tables + (size_t)T_F0_IR * sd.n_tab;
    c.ir.n_tab = sd.n_tab;
    c.psi_dt = sd.tables + (size_t)T_PSI_DT * sd.n_tab;
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
    // p_SI dt = (1 - exp(-beta I / N)) dt: from the memo table when I is an
    // in-range integer and N is the population the table was built for
    double p_SI_dt;
    const int Ii = (int)I;
    if (N == c.n_tot && Ii >= 0 && Ii < c.n_tab && (double)Ii == I) {
      p_SI_dt = __ldg(c.psi_dt + Ii);
    } else {
      p_SI_dt = (1.0 - det_exp(-c.beta * I / N)) * c.dt;
    }
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
};

struct SirsModel {
  static constexpr int kState = 4;
  static constexpr int kPar = 8;  // alpha beta gamma I0 exp_noise S0 R0 freq
  static constexpr int kData = 1;
  static constexpr int kTables = 3;
  enum { D_BETA, D_DT, D_EXP_NOISE, D_FREQ, D_NTOT, D_P_IR, D_Q_IR, D_QQ_IR, D_R_IR, D_FLIP_IR,
         D_P_RS, D_Q_RS, D_QQ_RS, D_R_RS, D_FLIP_RS };
  enum { T_F0_IR, T_PSI_DT, T_F0_RS };

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
  __device__ __forceinline__ static double table(int t, int n, const double *d) {
    if (t == T_F0_IR) return pow_int(d[D_QQ_IR], n);
    if (t == T_F0_RS) return pow_int(d[D_QQ_RS], n);
    return (1.0 - det_exp(-d[D_BETA] * (double)n / d[D_NTOT])) * d[D_DT];
  }

  struct Ctx {
    const double *derived;
    BinomialConst ir, rs;
    double beta, dt, n_tot;
    const double *psi_dt;
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
    c.ir.f0 = sd.tables + (size_t)T_F0_IR * sd.n_tab; c.ir.n_tab = sd.n_tab;
    c.rs.p = __ldg(d + D_P_RS); c.rs.q = __ldg(d + D_Q_RS); c.rs.qq = __ldg(d + D_QQ_RS);
    c.rs.r = __ldg(d + D_R_RS); c.rs.flip = __ldg(d + D_FLIP_RS) != 0.0;
    c.rs.f0 = sd.tables + (size_t)T_F0_RS * sd.n_tab; c.rs.n_tab = sd.n_tab;
    c.psi_dt = sd.tables + (size_t)T_PSI_DT * sd.n_tab;
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
    double p_SI_dt;
    const int Ii = (int)I;
    if (N == c.n_tot && Ii >= 0 && Ii < c.n_tab && (double)Ii == I) {
      p_SI_dt = __ldg(c.psi_dt + Ii);
    } else {
      p_SI_dt = (1.0 - det_exp(-c.beta * I / N)) * c.dt;
    }
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
};

}  // namespace mcb
