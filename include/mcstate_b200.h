/*
 * mcstate_b200.h -- C ABI of the B200-native particle-filter engine.
 *
 * This is the drop-in boundary of SURVEY.md section 8(b): the native layer
 * underneath the `dust_generator` model object that mcstate drives.  Every
 * entry point replaces one method mcstate calls on that object; the call
 * sites are cited as reference file:line (paths relative to the mcstate
 * tree).  The R binding a maintainer would add (`.Call` shims + an R6
 * generator) is in INTEGRATION.md.
 *
 * Conventions (mirroring dust's cpp11 entry points):
 *   - every function returns MCB_OK (0) or a negative error code and never
 *     aborts; mcb_last_error() gives the message (thread-local);
 *   - a handle is created by mcb_alloc, owned by the caller, freed by
 *     mcb_free; calls on one handle are synchronous and single-threaded,
 *     different handles are independent (SMC^2 holds up to 1024 of them,
 *     R/smc2.R:72-78);
 *   - arrays are HOST pointers borrowed for the call, in R's column-major
 *     layout: a state matrix [n_state, n_particles(, n_pars)] has n_state
 *     fastest; nothing here is a torch type;
 *   - all compute happens on the GPU: there is no CPU fallback, and
 *     mcb_alloc fails with MCB_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef MCSTATE_B200_H
#define MCSTATE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCB_OK 0
#define MCB_ERR_ARG (-1)    /* bad argument (R: stop() with a message)          */
#define MCB_ERR_CUDA (-2)   /* CUDA runtime failure, or no usable device         */
#define MCB_ERR_STATE (-3)  /* call not valid in the handle's current state      */

#define MCB_SCRAMBLER_PLUS 0      /* xoshiro256+  (dust's generator for double;  */
                                  /* tests/testthat/test-pmcmc-parallel.R:41)    */
#define MCB_SCRAMBLER_STARSTAR 1  /* xoshiro256** (BASELINE.json north_star)     */

typedef struct mcb_model mcb_model;

/* ---- library ---------------------------------------------------------- */
int mcb_version(void);
/* The message of the last failed call ON THIS THREAD (thread-local, so it takes no handle):
 * R calls the shim from its single interpreter thread, and a threaded host reads the
 * message on the thread whose call failed. */
const char *mcb_last_error(void);
/* replaces model$public_methods$has_gpu_support / gpu_info (R/particle_filter.R:212,241) */
int mcb_device_count(int *n_devices);
int mcb_device_info(int device_id, char *name, size_t name_len, int *cc_major,
                    int *cc_minor, int *n_sm, size_t *mem_bytes);

/* ---- model registry (dust::dust_example; tests/testthat/helper-mcstate.R:3) */
/* names: "sir", "sirs", "volatility", "walk" in the stock library.  Parameter vectors are
 * positional; names/defaults via the accessors so a host binding can map an R list and
 * ignore extra keys (R/particle_filter.R:283-284).
 * A user-supplied model (dust::dust(file) / odin.dust: tests/testthat/
 * test-pmcmc-parallel.R:197, helper-mcstate.R:478-492 + compare_variable.cpp) is a
 * library of its own with THIS SAME ABI, built from the engine sources with
 * -DMCB_USER_MODEL="<model header>"; it holds that one model as id 0, found by its name
 * (mcstate_b200/examples/seir.cuh). */
int mcb_model_lookup(const char *name, int *model_id, size_t *n_state, size_t *n_par,
                     size_t *n_data_fields);
const char *mcb_model_par_name(int model_id, size_t i);
double mcb_model_par_default(int model_id, size_t i);
const char *mcb_model_state_name(int model_id, size_t i);
/* name of data field i (the column the compiled compare reads: "incidence" for
 * sir/sirs -- tests/testthat/helper-mcstate.R:37 -- "observed" for volatility,
 * helper-mcstate.R:107) */
const char *mcb_model_data_name(int model_id, size_t i);

/* ---- RNG state arithmetic on the host (pure seeding plumbing) ----------- */
/* dust::dust_rng$new(seed): integer seed -> first stream */
int mcb_rng_seed(uint64_t seed, uint64_t state[4]);
/* $jump() / $long_jump()  (R/pmcmc_parallel.R:147) */
int mcb_rng_jump(uint64_t state[4]);
int mcb_rng_long_jump(uint64_t state[4]);
/* $random_real(n) on one stream (R/pmcmc_parallel.R:148) */
int mcb_rng_random_real(uint64_t state[4], int scrambler, size_t n, double *out);
/* n jumps at once (stream i of a seed = the seed jumped i times): how a shard finds
 * the first stream of its block of a larger model without walking there */
int mcb_rng_jump_n(uint64_t state[4], uint64_t n);
/* dust::dust_rng_distributed_state (R/pmcmc_parallel.R:133, R/smc2.R:68):
 * out[n_nodes][n_streams][4]; node k = seed long-jumped k times */
int mcb_rng_distributed_state(const uint64_t seed_state[4], size_t n_streams,
                              size_t n_nodes, uint64_t *out);

/* ---- the model object --------------------------------------------------- */
/*
 * generator$new(pars, time, n_particles, n_threads, seed, gpu_config, pars_multi)
 *   R/particle_filter_state.R:237-241
 * pars: [max(1,n_pars)][n_par]; n_pars == 0 <=> pars_multi = FALSE.
 * seed_state: 4*n_seed_streams words (an integer seed is expanded by the host
 * binding with mcb_rng_seed); the model owns n_particles*max(1,n_pars)+1
 * streams, the last one is the filter's resampling stream.
 */
int mcb_alloc(int model_id, const double *pars, size_t n_particles, size_t n_pars,
              uint64_t time, const uint64_t *seed_state, size_t n_seed_streams,
              int scrambler, int deterministic, int device_id, mcb_model **out);
/*
 * The same with the seeding made explicit.  MCB_SEED_SHARED is mcb_alloc (the dust
 * pars_multi layout: n_total particle streams and ONE filter stream shared by all
 * sets).  MCB_SEED_PER_SET takes one seed stream per parameter set and gives set p the
 * streams a separate model object seeded with it would hold (n_particles particle
 * streams and its own filter stream): independent filters batched in one handle, the
 * layout smc2 uses for its parameter particles (R/smc2.R:68-78).
 */
#define MCB_SEED_SHARED 0
#define MCB_SEED_PER_SET 1
int mcb_alloc_ex(int model_id, const double *pars, size_t n_particles, size_t n_pars,
                 uint64_t time, const uint64_t *seed_state, size_t n_seed_streams,
                 int seed_mode, int scrambler, int deterministic, int device_id,
                 mcb_model **out);
int mcb_free(mcb_model *m);
/* launch everything on this cudaStream_t (default: a stream owned by the handle) */
int mcb_set_stream(mcb_model *m, void *cuda_stream);
int mcb_synchronize(mcb_model *m);

/* $n_state() $n_particles() $n_pars() $time()  (R/particle_filter_state.R:268,288,373) */
int mcb_shape(const mcb_model *m, size_t *n_state, size_t *n_particles, size_t *n_pars,
              size_t *n_index);
int mcb_time(const mcb_model *m, uint64_t *time);
/* $pars(): the parameter vectors currently on the device */
int mcb_pars(const mcb_model *m, double *out);

/* $set_index(idx): 0-based row indices here (R/particle_filter_state.R:261,263) */
int mcb_set_index(mcb_model *m, const size_t *index, size_t n_index);
/* $run(time_end) -> [n_index, n_total]; out may be NULL (R/particle_filter_state.R:65) */
int mcb_run(mcb_model *m, uint64_t time_end, double *out);
/* $simulate(times) -> [n_index, n_total, n_times] (R/predict.R:79) */
int mcb_simulate(mcb_model *m, const uint64_t *times, size_t n_times, double *out);
/* $state(index) (R/particle_filter_state.R:79,81,114,272); index NULL = all rows */
int mcb_state(mcb_model *m, const size_t *index, size_t n_index, double *out);
/* The full state of n chosen particles only: out [n][n_state] (R: model$state()[, i],
 * what pmcmc keeps of an accepted step, R/pmcmc_state.R:40-43 -- one column of a
 * [n_state, n_total] matrix that need not cross PCIe).  particles: 0-based, set-major. */
int mcb_state_particles(mcb_model *m, const size_t *particles, size_t n, double *out);
/*
 * $update_state(pars=, state=, time=, set_initial_state=)
 *   R/particle_filter_state.R:248,253,512; R/smc2.R:141
 * pars NULL = keep; state NULL = keep; state_len is n_state (recycled over
 * particles) or n_state*n_total.  Never touches the RNG.
 */
int mcb_update_state(mcb_model *m, const double *pars, const double *state,
                     size_t state_len, int has_time, uint64_t time,
                     int set_initial_state);
/* $reorder(kappa): 1-based, within-block for [N,P] input (R/particle_filter_state.R:100) */
int mcb_reorder(mcb_model *m, const int *kappa);
/* $rng_state(first_only) / $set_rng_state(raw): 4 words (32 bytes) per stream
 * (R/particle_filter.R:813,877; R/particle_filter_state.R:363,369,406,420) */
int mcb_rng_state(mcb_model *m, int first_only, uint64_t *out);
int mcb_set_rng_state(mcb_model *m, const uint64_t *state, size_t n_words);
/*
 * $set_data(dust_data(...), shared)  (R/particle_filter_state.R:243-246)
 * values: [n_data][shared ? 1 : max(1,n_pars)][n_data_fields]; NaN = NA.
 */
int mcb_set_data(mcb_model *m, size_t n_data, const uint64_t *time_end,
                 const double *values, int shared);
/*
 * Sharding one multi-parameter model over several handles / GPUs (nested populations,
 * SURVEY.md 8e): the model's single filter stream draws one uniform per set per data
 * row, in set order.  A handle that holds sets [set_offset, set_offset + n_pars) of
 * n_sets_global replays the whole order on its own copy of the stream and keeps only
 * its own draws, so every shard's stream advances exactly as the unsharded model's
 * would.  na_global [n_data][n_sets_global] (1 = that set's datum is NA, no draw) is
 * needed because a shard does not hold the other sets' data.  Call after mcb_set_data;
 * mcb_set_filter_stream installs the shared stream (stream n_total_global of the seed).
 */
int mcb_set_filter_stream(mcb_model *m, const uint64_t state[4]);
int mcb_set_stream_sharing(mcb_model *m, size_t n_sets_global, size_t set_offset,
                           const unsigned char *na_global);
/*
 * smc2 replenishment on the device (R/smc2.R:174-189, `filter[accept] <- proposed$filter
 * [accept]`; R/particle_filter_state.R:420, the parent's RNG takes the child's): for every
 * parameter set p with take[p] != 0 copy the particle state and/or the RNG streams of
 * `src` into `dst` (same shape, same device).  RNG copies need MCB_SEED_PER_SET.
 */
int mcb_copy_sets(mcb_model *dst, mcb_model *src, const unsigned char *take, int copy_state,
                  int copy_rng);
/* $compare_data(): log-weights [n_total] for the datum at the current time;
 * *has_data = 0 (and out untouched) when there is none */
int mcb_compare_data(mcb_model *m, double *logw, int *has_data);
/* $resample(weights) -> kappa (1-based, within-block); draws from the filter stream */
int mcb_resample(mcb_model *m, const double *weights, int *kappa);

/*
 * $filter(time_end, save_trajectories, time_snapshot)   <- THE HOT PATH
 *   R/particle_filter_state.R:151 (step_compiled)
 * Runs every not-yet-consumed data row with time_end <= `time_end`; one CUDA
 * graph launch, no host synchronisation between intervals.
 *   log_likelihood [max(1,n_pars)]
 *   trajectories   [n_index, n_total, n_rows+1] ancestor-traced, or NULL
 *   snapshots      [n_state, n_total, n_snapshots] (state after resampling), or NULL
 *   min_log_likelihood: n_min = 0 (reference behaviour: never forwarded,
 *     R/particle_filter_state.R:151), 1 or max(1,n_pars) (rule of
 *     R/particle_filter_state.R:463-474)
 * mcb_filter_rows tells the caller how many rows a call would consume (to size
 * `trajectories`).
 */
int mcb_filter_rows(const mcb_model *m, uint64_t time_end, size_t *n_rows);
int mcb_filter(mcb_model *m, uint64_t time_end, int save_trajectories,
               const uint64_t *snapshot_times, size_t n_snapshots,
               const double *min_log_likelihood, size_t n_min,
               double *log_likelihood, double *trajectories, double *snapshots);

/* save_trajectories = MCB_TRAJECTORIES_ON_DEVICE: record the history (values and
 * ancestor order) but leave it in device memory; `trajectories` is ignored.  Then
 * mcb_filter_history traces the lineages of chosen final particles on the device
 * and returns only those: out [n_rows+1][n][n_index] (same layout as
 * `trajectories`, with n particles in place of n_total).  This is what pMCMC
 * consumes of a run -- one sampled particle: R/pmcmc_state.R:32-51 calls
 * filter$history(i), i.e. history_single(.., index_particle),
 * R/particle_filter.R:616-646 -- without moving the [n_index, n_total, T+1]
 * array (2.5 GB at 2^20 particles) across PCIe.  `particles`: 0-based indices
 * into n_total (set-major).  Valid until the next filter call on the handle. */
#define MCB_TRAJECTORIES_ON_DEVICE 2
int mcb_filter_history(mcb_model *m, const size_t *particles, size_t n, double *out);

/* A filter of one parameter set whose particles are all resident at once (up to
 * about 150 000 on a B200), run without snapshots, executes as ONE
 * cooperative kernel per mcb_filter call instead of three launches per data interval
 * (csrc/persistent.cuh); results are bit-identical either way.  enable = 0 keeps
 * this handle on the per-interval path (A/B measurements, tests); the environment
 * variable MCB_PERSISTENT=0 does the same for every handle. */
int mcb_set_persistent(mcb_model *m, int enable);
/* the same, split so many handles / GPUs can be in flight at once: enqueue
 * returns as soon as the work is queued, collect waits and copies out */
int mcb_filter_enqueue(mcb_model *m, uint64_t time_end);
int mcb_filter_collect(mcb_model *m, double *log_likelihood);
/* Device address of the log-likelihood vector [n_pars] the enqueued filter writes
 * (valid in stream order after mcb_filter_enqueue, until the next filter call): a
 * sharded caller hands it straight to its all-gather on the handle's stream instead
 * of a device -> host -> device round trip per data row (R/smc2.R:89-108 gathers one
 * step likelihood per parameter particle per row). */
int mcb_log_likelihood_device(mcb_model *m, double **device_ptr);
/* number of kernels launched by this handle so far (bench.py: gpu_launches) */
int mcb_launch_count(const mcb_model *m, uint64_t *n);
/*
 * Per-kernel device timing for the roofline report.  When enabled, the filter's
 * CUDA graph carries event records around the four per-interval kernels
 * (0: run+compare, 1: weights scan, 2: ancestors, 3: the final gather of a call);
 * mcb_kernel_timing returns the accumulated milliseconds and launch counts
 * since the last reset and clears them.
 */
int mcb_set_kernel_timing(mcb_model *m, int enable);
int mcb_kernel_timing(mcb_model *m, double ms[4], uint64_t launches[4]);

#ifdef __cplusplus
}
#endif
#endif
