// The SIR-type example models (dust::dust_example("sir") / ("sirs")).
//
// Maths: vignettes/sir_models.Rmd:44-54; compare: the R twin at
// tests/testthat/helper-mcstate.R:7-22; state order S, I, R, cumulative,
// interval incidence (tests/testthat/test-particle-filter.R:14,223); defaults
// scripts/generate_vignette_data.R:11-12.  Draw order follows dust's example
// sources [SURVEY.md A.3].
//
// A model is a struct with: kState, kPar, kData, names/defaults for the host
// binding, initial(), update() and compare().  State lives in registers for
// all `rate` steps of an interval; parameters are a per-set vector in global
// memory read through the read-only path.
#pragma once
#include "samplers.cuh"

namespace mcb {

constexpr int kMaxState = 8;
constexpr int kMaxPar = 8;

struct SirModel {
  static constexpr int kState = 5;
  static constexpr int kPar = 7;  // beta gamma I0 exp_noise S0 R0 freq
  static constexpr int kData = 1; // incidence
  static constexpr int kCompareRow = 4;

  // per-thread constants derived once per launch
  struct Ctx {
    double beta, p_IR_dt, dt, exp_noise;
    uint32_t freq;
  };
  __device__ __forceinline__ static Ctx context(const double *__restrict__ p) {
    Ctx c;
    c.beta = __ldg(p + 0);
    c.freq = (uint32_t)__ldg(p + 6);
    c.dt = 1.0 / __ldg(p + 6);
    c.p_IR_dt = (1.0 - det_exp(-__ldg(p + 1))) * c.dt;
    c.exp_noise = __ldg(p + 3);
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
    const double p_SI = 1.0 - det_exp(-c.beta * I / N);
    const double n_IR = draw_binomial<SCR>(rng, det, I, c.p_IR_dt);
    const double n_SI = draw_binomial<SCR>(rng, det, S, p_SI * c.dt);
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
    const double noise = det ? 1.0 / c.exp_noise : draw_exponential<SCR>(rng, c.exp_noise);
    return density_poisson_log(observed, lgamma_obs1, y[4] + noise);
  }
};

struct SirsModel {
  static constexpr int kState = 4;
  static constexpr int kPar = 8;  // alpha beta gamma I0 exp_noise S0 R0 freq
  static constexpr int kData = 1;
  static constexpr int kCompareRow = 3;

  struct Ctx {
    double beta, p_IR_dt, p_RS_dt, dt, exp_noise;
    uint32_t freq;
  };
  __device__ __forceinline__ static Ctx context(const double *__restrict__ p) {
    Ctx c;
    c.beta = __ldg(p + 1);
    c.freq = (uint32_t)__ldg(p + 7);
    c.dt = 1.0 / __ldg(p + 7);
    c.p_IR_dt = (1.0 - det_exp(-__ldg(p + 2))) * c.dt;
    c.p_RS_dt = (1.0 - det_exp(-__ldg(p + 0))) * c.dt;
    c.exp_noise = __ldg(p + 4);
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
    const double p_SI = 1.0 - det_exp(-c.beta * I / N);
    const double n_SI = draw_binomial<SCR>(rng, det, S, p_SI * c.dt);
    const double n_IR = draw_binomial<SCR>(rng, det, I, c.p_IR_dt);
    const double n_RS = draw_binomial<SCR>(rng, det, R, c.p_RS_dt);
    y[0] = S - n_SI + n_RS;
    y[1] = I + n_SI - n_IR;
    y[2] = R + n_IR - n_RS;
    y[3] = (time % c.freq == 0) ? n_SI : y[3] + n_SI;
  }
  template <int SCR>
  __device__ __forceinline__ static double compare(const double *y, double observed,
                                                   double lgamma_obs1, Xoshiro &rng,
                                                   const Ctx &c, bool det) {
    const double noise = det ? 1.0 / c.exp_noise : draw_exponential<SCR>(rng, c.exp_noise);
    return density_poisson_log(observed, lgamma_obs1, y[3] + noise);
  }
};

}  // namespace mcb
