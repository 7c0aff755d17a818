// A user-supplied model for mcstate_b200.dust.dust(): the stochastic SEIR model.
//
// This file is what a user of the reference would hand to dust::dust() /
// odin.dust::odin_dust() (tests/testthat/test-pmcmc-parallel.R:197,
// tests/testthat/helper-mcstate.R:478-492): the model's update and its compiled compare.
// The maths extends the reference's SIR (vignettes/sir_models.Rmd:44-54) by a latent
// compartment; the compare is the SIR one (tests/testthat/helper-mcstate.R:7-22): Poisson
// density of the observed incidence around interval incidence + Exp(exp_noise) noise
// from the particle's own stream.
//
//   p_SE = 1 - exp(-beta I / N)   p_EI = 1 - exp(-sigma)   p_IR = 1 - exp(-gamma)
//   n_SE ~ Binomial(S, p_SE dt)   n_EI ~ Binomial(E, p_EI dt)   n_IR ~ Binomial(I, p_IR dt)
//   n_imp ~ Poisson(import dt)    exposed arrivals from outside the population
//
// drawn in that order (import = 0, the default: no arrivals and no draw); incidence
// counts the E -> I transitions.  The CPU twin the parity
// tests check it against is oracle/mcstate_oracle.c: seir_update.
//
// Contract (models.cuh): kState / kPar / kData, derive(), context(), initial(),
// update(), compare(); MCB_USER_MODEL_DESC names the model, its parameters (with
// defaults), its state variables and its data field.  That is all a model needs.
//
// OPTIONAL, and what this file adds to show it: the lean path (kHasLean = true).  The
// same arithmetic for the usual case -- every compartment a non-negative integer, the
// closed population the memo tables were built for, no importation -- with the state held
// as int32 and the per-draw setup read from tables the engine fills once per parameter
// set from walk_row() / inf_entry(): the two draws whose probability the parameters fix
// (E -> I, I -> R) become a table walk, the infection draw a certified fast walk
// (samplers.cuh).  Same draws, same bits as update(): whatever the lean step cannot take
// it hands back untouched and update() runs.  MCB_USER_MODEL_TABLES tells the engine which
// parameters add up to the closed population (to size the tables), how many
// fixed-probability draws there are, and which parameter is the reset period.
#pragma once

namespace mcb {

struct UserModel {
  static constexpr bool kHasLean = true;
  static constexpr int kState = 6;  // S E I R cases_cumul cases_inc
  static constexpr int kPar = 9;    // beta sigma gamma I0 exp_noise S0 E0 freq import
  static constexpr int kData = 1;   // incidence
  static constexpr int kConst = 2;  // draws with a parameter-only probability: E -> I, I -> R
  enum { D_BETA, D_P_EI_DT, D_P_IR_DT, D_DT, D_EXP_NOISE, D_FREQ, D_IMPORT_DT, D_NTOT,
         D_Q_EI, D_QQ_EI, D_R_EI, D_Q_IR, D_QQ_IR, D_R_IR, D_INV_EXP_NOISE };
  enum { C_EI, C_IR };

  // once per parameter set and update_state(pars): constants the parameters fix
  __device__ __forceinline__ static void derive(const double *__restrict__ p, double *d) {
    const double dt = 1.0 / p[7];
    d[D_BETA] = p[0];
    d[D_P_EI_DT] = (1.0 - det_exp(-p[1])) * dt;
    d[D_P_IR_DT] = (1.0 - det_exp(-p[2])) * dt;
    d[D_DT] = dt;
    d[D_EXP_NOISE] = p[4];
    d[D_FREQ] = p[7];
    d[D_IMPORT_DT] = p[8] * dt;
    // for the lean path: the closed population, and q = min(p, 1-p), 1-q, q/(1-q) of the
    // two fixed-probability draws exactly as draw_binomial() forms them
    d[D_NTOT] = p[5] + p[6] + p[3];
    const double q_ei = d[D_P_EI_DT] > 0.5 ? 1.0 - d[D_P_EI_DT] : d[D_P_EI_DT];
    d[D_Q_EI] = q_ei; d[D_QQ_EI] = 1.0 - q_ei; d[D_R_EI] = q_ei / (1.0 - q_ei);
    const double q_ir = d[D_P_IR_DT] > 0.5 ? 1.0 - d[D_P_IR_DT] : d[D_P_IR_DT];
    d[D_Q_IR] = q_ir; d[D_QQ_IR] = 1.0 - q_ir; d[D_R_IR] = q_ir / (1.0 - q_ir);
    d[D_INV_EXP_NOISE] = reciprocal_for(p[4]);
  }
  // row n of the walk table of fixed-probability draw c; entry I of the infection table
  __device__ __forceinline__ static void walk_row(int c, int n, const double *d, double *row) {
    if (c == C_EI) fill_walk_row(d[D_P_EI_DT], d[D_Q_EI], d[D_QQ_EI], d[D_R_EI], n, row);
    else fill_walk_row(d[D_P_IR_DT], d[D_Q_IR], d[D_QQ_IR], d[D_R_IR], n, row);
  }
  __device__ __forceinline__ static double4 inf_entry(int I, const double *d) {
    return infection_entry(d[D_BETA], d[D_NTOT], d[D_DT], I);
  }
  struct Ctx {
    double beta, p_ei_dt, p_ir_dt, dt, exp_noise, import_dt, inv_exp_noise;
    uint32_t freq;
  };
  __device__ __forceinline__ static Ctx context(const SetData &sd) {
    const double *d = sd.derived;
    return Ctx{__ldg(d + D_BETA), __ldg(d + D_P_EI_DT), __ldg(d + D_P_IR_DT), __ldg(d + D_DT),
               __ldg(d + D_EXP_NOISE), __ldg(d + D_IMPORT_DT), __ldg(d + D_INV_EXP_NOISE),
               (uint32_t)__ldg(d + D_FREQ)};
  }
  __device__ __forceinline__ static void initial(const double *__restrict__ p, double *y) {
    y[0] = __ldg(p + 5); y[1] = __ldg(p + 6); y[2] = __ldg(p + 3);
    y[3] = 0.0; y[4] = 0.0; y[5] = 0.0;
  }
  template <int SCR>
  __device__ __forceinline__ static void update(uint64_t time, double *y, Xoshiro &rng,
                                                const Ctx &c, bool det) {
    const double S = y[0], E = y[1], I = y[2], R = y[3];
    const double N = S + E + I + R;
    const double p_SE_dt = (1.0 - det_exp(-c.beta * I / N)) * c.dt;
    const double n_SE = draw_binomial<SCR>(rng, det, S, p_SE_dt);
    const double n_EI = draw_binomial<SCR>(rng, det, E, c.p_ei_dt);
    const double n_IR = draw_binomial<SCR>(rng, det, I, c.p_ir_dt);
    const double n_imp = draw_poisson<SCR>(rng, det, c.import_dt);
    y[0] = S - n_SE;
    y[1] = E + n_SE - n_EI + n_imp;
    y[2] = I + n_EI - n_IR;
    y[3] = R + n_IR;
    y[4] = y[4] + n_EI;
    y[5] = (time % c.freq == 0) ? n_EI : y[5] + n_EI;
  }
  template <int SCR>
  __device__ __forceinline__ static double compare(const double *y, double observed,
                                                   double lgamma_obs1, Xoshiro &rng,
                                                   const Ctx &c, bool det) {
    const double noise = det ? 1.0 / c.exp_noise
                             : draw_exponential_fixed<SCR>(rng, c.exp_noise, c.inv_exp_noise);
    return density_poisson_log(observed, lgamma_obs1, y[5] + noise);
  }

  // ---- the lean path (optional) -------------------------------------------------
  // The loop works on S, E, I and two sums that make the incidence counters free: S + E
  // only ever falls, by the E -> I transitions, so cumulative + S + E never changes (C),
  // and the interval incidence is M - (S + E), M reset to S + E when an interval starts.
  struct Lean {
    uint32_t set;
  };
  static constexpr int kLeanState = 5;  // S, E, I, M, C
  __device__ __forceinline__ static bool lean_load(const SetData &sd, const double *y, int *yi,
                                                   Lean &) {
    int all[kState];
    if (!to_counts<kState>(y, all)) return false;
    yi[0] = all[0]; yi[1] = all[1]; yi[2] = all[2];
    yi[3] = all[0] + all[1] + all[5];
    yi[4] = all[0] + all[1] + all[4];
    const double n_tot = __ldg(sd.derived + D_NTOT);
    // the tables are built for N = n_tot and cover counts < n_tab; arrivals change N
    return y[0] + y[1] + y[2] + y[3] == n_tot && n_tot < (double)sd.n_tab &&
           __ldg(sd.derived + D_IMPORT_DT) == 0.0;
  }
  __device__ __forceinline__ static void lean_store(const SetData &sd, const int *yi, double *y) {
    const double n_tot = __ldg(sd.derived + D_NTOT);
    const int q = yi[0] + yi[1];
    y[0] = (double)yi[0];
    y[1] = (double)yi[1];
    y[2] = (double)yi[2];
    y[3] = n_tot - (double)(q + yi[2]);   // exact: integers below 2^31
    y[4] = (double)(yi[4] - q);
    y[5] = (double)(yi[3] - q);
  }
  // One model step; `reset`: it starts a new incidence interval.  false: nothing changed
  // (state and generator as they were), update() must run this step.
  template <int SCR>
  __device__ __forceinline__ static bool lean_update(bool reset, int *y, Xoshiro &rng,
                                                     const Lean &c, const LeanTabs &t) {
    const int S = y[0], E = y[1], I = y[2];
    const double4 e_se =
        ldg_f64x4(reinterpret_cast<const double *>(t.inf + (c.set * t.n_tab + (uint32_t)I)));
    const double *row_ei = walk_row_ptr(E, c.set * kConst + C_EI, t);
    const double *row_ir = walk_row_ptr(I, c.set * kConst + C_IR, t);
    const double4 F_ei = ldg_f64x4(row_ei), F_ir = ldg_f64x4(row_ir);
    const int n_SE = draw_infection_lean<SCR>(rng, S, e_se);
    if (n_SE < 0) return false;
    const bool se_drew = S != 0 && e_se.x != 0.0;   // (no uniform for S = 0 or p = 0)
    int n_EI = 0;
    if (E != 0) {
      n_EI = draw_const_lean<SCR>(rng, E, row_ei, F_ei, t);
      if (n_EI < 0) {
        if (se_drew) xo_rewind(rng);
        return false;
      }
    }
    int n_IR = 0;
    if (I != 0) {
      n_IR = draw_const_lean<SCR>(rng, I, row_ir, F_ir, t);
      if (n_IR < 0) {
        if (E != 0) xo_rewind(rng);
        if (se_drew) xo_rewind(rng);
        return false;
      }
    }
    if (reset) y[3] = S + E;
    y[0] = S - n_SE;
    y[1] = E + n_SE - n_EI;
    y[2] = I + n_EI - n_IR;
    return true;
  }
};

}  // namespace mcb

#define MCB_USER_MODEL_DESC                                                          \
  {"seir", 6, 9, 1,                                                                   \
   {"beta", "sigma", "gamma", "I0", "exp_noise", "S0", "E0", "freq", "import"},       \
   {0.3, 0.25, 0.1, 10.0, 1e6, 1000.0, 0.0, 4.0, 0.0},                                \
   {"S", "E", "I", "R", "cases_cumul", "cases_inc"},                                  \
   {"incidence"}}

// lean path only: {parameter slots adding up to the closed population (S0, E0, I0)},
// number of fixed-probability draws, parameter slot of the reset period (freq)
#define MCB_USER_MODEL_TABLES {{5, 6, 3}, mcb::UserModel::kConst, 7}
