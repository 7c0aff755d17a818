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
// defaults), its state variables and its data field.
#pragma once

namespace mcb {

struct UserModel {
  static constexpr bool kHasLean = false;
  static constexpr int kState = 6;  // S E I R cases_cumul cases_inc
  static constexpr int kPar = 9;    // beta sigma gamma I0 exp_noise S0 E0 freq import
  static constexpr int kData = 1;   // incidence
  static constexpr int kConst = 0;
  enum { D_BETA, D_P_EI_DT, D_P_IR_DT, D_DT, D_EXP_NOISE, D_FREQ, D_IMPORT_DT };

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
  }
  struct Ctx {
    double beta, p_ei_dt, p_ir_dt, dt, exp_noise, import_dt;
    uint32_t freq;
  };
  __device__ __forceinline__ static Ctx context(const SetData &sd) {
    const double *d = sd.derived;
    return Ctx{__ldg(d + D_BETA), __ldg(d + D_P_EI_DT), __ldg(d + D_P_IR_DT), __ldg(d + D_DT),
               __ldg(d + D_EXP_NOISE), __ldg(d + D_IMPORT_DT), (uint32_t)__ldg(d + D_FREQ)};
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
    const double noise = det ? 1.0 / c.exp_noise : draw_exponential<SCR>(rng, c.exp_noise);
    return density_poisson_log(observed, lgamma_obs1, y[5] + noise);
  }
};

}  // namespace mcb

#define MCB_USER_MODEL_DESC                                                          \
  {"seir", 6, 9, 1,                                                                   \
   {"beta", "sigma", "gamma", "I0", "exp_noise", "S0", "E0", "freq", "import"},       \
   {0.3, 0.25, 0.1, 10.0, 1e6, 1000.0, 0.0, 4.0, 0.0},                                \
   {"S", "E", "I", "R", "cases_cumul", "cases_inc"},                                  \
   {"incidence"}}
