/*
 * mcstate_oracle.h -- CPU ORACLE, TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the particle-filter hot path that mcstate drives
 * through the `dust` model object.  Nothing in the
 * product (mcstate_b200/, include/) may include, link or call this; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
 * arm use it, and only as the checker or the timed CPU baseline.
 *
 * PARITY STATUS
 *   - pinned against the reference tree: scale_log_weights known answers
 *     (tests/testthat/test-particle-filter.R:174-187), particle_filter_data
 *     layouts (tests/testthat/test-particle-filter-data.R:69-155), the loop
 *     order of R/particle_filter_state.R:63-117, the R resampler
 *     R/particle_filter.R:515-523, trajectory tracing R/particle_filter.R:616-646.
 *   - pinned against published vectors: splitmix64 and xoshiro256 (plus, starstar) outputs.
 *   - PARITY UNPINNED for everything that lives in the absent third-party
 *     dependency dust (>= 0.13.12, DESCRIPTION:25): seeding idiom, stream
 *     layout, binomial/Poisson algorithms, SIR draw order, dust's
 *     resample_weight tie rule.  Those follow SURVEY.md Appendix A and are
 *     listed one by one in PARITY_NOTES.md.
 */
#ifndef MCSTATE_ORACLE_H
#define MCSTATE_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_SCRAMBLER_PLUS = 0, ORC_SCRAMBLER_STARSTAR = 1 };
enum { ORC_MODEL_SIR = 0, ORC_MODEL_SIRS = 1, ORC_MODEL_VOLATILITY = 2, ORC_MODEL_WALK = 3,
       ORC_MODEL_SEIR = 4 };   /* SEIR: the twin of the user-supplied example model */
/* how sum(w) and the cumulative sum are accumulated when resampling */
enum {
  ORC_RESAMPLE_DUST_FP64 = 0, /* left-to-right fp64, as dust's resample_weight   */
  ORC_RESAMPLE_EXACT = 1      /* fixed-point integer sums: order-independent      */
};

/* ---- random numbers ---- */
uint64_t orc_splitmix64(uint64_t *state);          /* canonical generator step */
void orc_seed(uint64_t seed, uint64_t s[4]);       /* dust seeding idiom       */
uint64_t orc_next(uint64_t s[4], int scrambler);
void orc_jump(uint64_t s[4]);
void orc_long_jump(uint64_t s[4]);
double orc_random_real(uint64_t s[4], int scrambler);
/* n streams, 4 words each, stream i = stream 0 jumped i times */
void orc_rng_streams(const uint64_t *seed_state, size_t n_seed_streams,
                     size_t n_streams, uint64_t *out);
/* dust_rng_distributed_state: node k = base long-jumped k times */
void orc_rng_distributed_state(const uint64_t seed_state[4], size_t n_streams,
                               size_t n_nodes, uint64_t *out);

/* ---- deterministic elementary functions shared (by algorithm) with the GPU ---- */
double orc_exp(double x);
double orc_log(double x);

/* ---- samplers and densities ---- */
double orc_exponential(uint64_t s[4], int scrambler, double rate);
double orc_cos2pi(double u);
double orc_normal(uint64_t s[4], int scrambler, double mean, double sd);
double orc_density_normal_log(double x, double mu, double sd);
double orc_binomial(uint64_t s[4], int scrambler, double n, double p);
double orc_poisson(uint64_t s[4], int scrambler, double lambda);
double orc_lfact(double k);   /* log(k!), reproducible (engine twin: det_lfact) */
double orc_density_poisson_log(double x, double lgamma_x1, double lambda);

/* ---- weights and resampling ---- */
/* R/particle_filter.R:570-587; weights_out may be NULL */
double orc_scale_log_weights_r(const double *logw, size_t n, double *weights_out);
/* compiled-filter flavour (A.4 step 3): in place, returns the average */
double orc_scale_log_weights(double *w, size_t n, int mode);
int orc_weight_shift(size_t n);
/* R/particle_filter.R:515-523 with the uniform supplied; kappa is 1-based */
void orc_particle_resample_r(const double *weights, size_t n, double u0, int *kappa);
/* dust resample_weight (R2b); kappa 0-based, relative to the block */
void orc_resample_weight(const double *w, size_t n, double u, int mode, int *kappa);
/* R/particle_filter.R:616-646; order/value as the R arrays, 1-based order */
void orc_history_single(const double *value, const int *order, size_t ny,
                        size_t np, size_t nt, double *out);

/* ---- the model object (dust_generator protocol) ---- */
typedef struct orc_model orc_model;

size_t orc_model_n_state(int model);
size_t orc_model_n_par(int model);
size_t orc_model_n_data(int model);
void orc_model_default_pars(int model, double *pars);

/* n_pars == 0 means pars_multi = FALSE (one parameter set) */
orc_model *orc_alloc(int model, const double *pars, size_t n_particles,
                     size_t n_pars, uint64_t time, const uint64_t *seed_state,
                     size_t n_seed_streams, int scrambler, int deterministic,
                     int n_threads);
void orc_free(orc_model *m);
void orc_set_resample_mode(orc_model *m, int mode);
void orc_set_n_threads(orc_model *m, int n);
size_t orc_n_total(const orc_model *m);
uint64_t orc_time(const orc_model *m);
void orc_run(orc_model *m, uint64_t time_end);
/* out is [n_total][n_index] (R layout: n_index fastest); index NULL = all rows */
void orc_state(const orc_model *m, const size_t *index, size_t n_index, double *out);
/* pars may be NULL; state may be NULL; state_len = n_state (recycled) or n_state*n_total */
void orc_update_state(orc_model *m, const double *pars, const double *state,
                      size_t state_len, int has_time, uint64_t time,
                      int set_initial_state);
void orc_reorder(orc_model *m, const int *kappa /* 0-based global */);
void orc_rng_state(const orc_model *m, uint64_t *out /* (n_total+1)*4 */);
void orc_set_rng_state(orc_model *m, const uint64_t *in);
/* data: n_data rows; values [n_data][n_sets][n_fields], n_sets = 1 if shared */
void orc_set_data(orc_model *m, size_t n_data, const uint64_t *time_end,
                  const double *values, int shared);
/* log weights of the datum whose time_end == current time; returns 0 if none */
int orc_compare_data(orc_model *m, double *logw);
/* weights in (length n_total), kappa out (0-based global); consumes the filter stream */
/* sharding twins of mcb_set_filter_stream / mcb_set_stream_sharing */
void orc_set_filter_stream(orc_model *m, const uint64_t state[4]);
void orc_set_stream_sharing(orc_model *m, size_t n_sets_global, size_t set_offset,
                            const unsigned char *na_global);
void orc_resample(orc_model *m, const double *weights, int *kappa);
/*
 * The compiled filter (A.4).  trajectories: [T+1][n_total][n_index] or NULL,
 * snapshots: [n_snap][n_total][n_state] or NULL.  Returns the number of data
 * intervals consumed.
 */
size_t orc_filter(orc_model *m, uint64_t time_end, double *log_likelihood,
                  const size_t *index, size_t n_index, double *trajectories,
                  const uint64_t *snapshot_times, size_t n_snap, double *snapshots,
                  const double *min_log_likelihood, size_t n_min);
/* one update() per call on raw arrays, for unit tests */
void orc_model_update(int model, uint64_t time, const double *y, uint64_t s[4],
                      int scrambler, int deterministic, const double *pars,
                      double *y_next);

#ifdef __cplusplus
}
#endif
#endif
