/*
 * mcstate_b200_r.c -- the .Call layer between R and libmcstate_b200.so.
 *
 * One mcbr_* entry point per method mcstate calls on its model object (the
 * `dust_generator` interface: is_dust_generator, R/particle_filter.R:590-593; the calls
 * themselves, R/particle_filter_state.R:65-151, 229-263, 363-424; R/predict.R:79;
 * R/pmcmc_parallel.R:133).  Each forwards to the mcb_* function of
 * include/mcstate_b200.h that cites the same reference line.  Nothing is computed here:
 * arguments are unpacked, output vectors allocated in R's memory (the engine writes R's
 * column-major layout directly), error codes turned into R conditions.
 *
 * Threading contract: R calls into this file from its single interpreter thread, and
 * mcb_last_error() is thread-local, so the message fetched in chk() is always the one of
 * the call that just failed.
 *
 * Not compiled in the build image (no R there).  tests/test_r_binding.py type-checks it
 * with gcc against the stub headers in r/stub/ and the real include/mcstate_b200.h, and
 * checks the registration table against the exported mcb_* symbols.
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <stdint.h>
#include <string.h>

#include "mcstate_b200.h"

static void chk(int rc) {
  if (rc != MCB_OK) Rf_error("%s", mcb_last_error());
}

static mcb_model *handle(SEXP ptr) {
  mcb_model *m = (mcb_model *)R_ExternalPtrAddr(ptr);
  if (!m) Rf_error("Pointer has been invalidated (perhaps serialised?)");
  return m;
}

/* ownership: the handle (and its device memory) lives as long as the R object */
static void finalize(SEXP ptr) {
  mcb_model *m = (mcb_model *)R_ExternalPtrAddr(ptr);
  if (m) {
    mcb_free(m);
    R_ClearExternalPtr(ptr);
  }
}

/* R numerics hold model steps (integers below 2^53) */
static uint64_t as_u64(SEXP x) { return (uint64_t)Rf_asReal(x); }

struct shape {
  size_t n_state, n_particles, n_pars, n_index, n_sets, n_total;
};
static struct shape shape_of(mcb_model *m) {
  struct shape s;
  chk(mcb_shape(m, &s.n_state, &s.n_particles, &s.n_pars, &s.n_index));
  s.n_sets = s.n_pars ? s.n_pars : 1;
  s.n_total = s.n_particles * s.n_sets;
  return s;
}

/* 1-based R indices -> 0-based size_t; NULL -> NULL */
static size_t *index0(SEXP r_index, size_t *n) {
  if (Rf_isNull(r_index)) {
    *n = 0;
    return NULL;
  }
  *n = (size_t)XLENGTH(r_index);
  size_t *out = (size_t *)R_alloc(*n ? *n : 1, sizeof(size_t));
  for (size_t i = 0; i < *n; ++i)
    out[i] = (size_t)(Rf_isInteger(r_index) ? INTEGER(r_index)[i] : (int)REAL(r_index)[i]) - 1;
  return out;
}

/* ---- library / registry ------------------------------------------------------ */

/* model$public_methods$gpu_info() / has_gpu_support()  R/particle_filter.R:212,241 */
SEXP mcbr_gpu_info(void) {
  int n = 0;
  chk(mcb_device_count(&n));
  SEXP names = PROTECT(Rf_allocVector(STRSXP, n));
  SEXP mem = PROTECT(Rf_allocVector(REALSXP, n));
  SEXP sms = PROTECT(Rf_allocVector(INTSXP, n));
  SEXP cc = PROTECT(Rf_allocVector(INTSXP, n));
  for (int i = 0; i < n; ++i) {
    char name[256];
    int major = 0, minor = 0, n_sm = 0;
    size_t bytes = 0;
    chk(mcb_device_info(i, name, sizeof name, &major, &minor, &n_sm, &bytes));
    SET_STRING_ELT(names, i, Rf_mkChar(name));
    REAL(mem)[i] = (double)bytes;
    INTEGER(sms)[i] = n_sm;
    INTEGER(cc)[i] = major * 10 + minor;
  }
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 4));
  SET_VECTOR_ELT(out, 0, names);
  SET_VECTOR_ELT(out, 1, mem);
  SET_VECTOR_ELT(out, 2, sms);
  SET_VECTOR_ELT(out, 3, cc);
  UNPROTECT(5);
  return out;
}

/* names and defaults of a model's parameters, its state and data-field names: what the R
 * side needs to turn `pars` lists into positional vectors (extra keys ignored,
 * R/particle_filter.R:283-284) */
SEXP mcbr_model_info(SEXP r_model) {
  int id = 0;
  size_t n_state = 0, n_par = 0, n_data = 0;
  chk(mcb_model_lookup(CHAR(STRING_ELT(r_model, 0)), &id, &n_state, &n_par, &n_data));
  SEXP par_names = PROTECT(Rf_allocVector(STRSXP, n_par));
  SEXP par_defaults = PROTECT(Rf_allocVector(REALSXP, n_par));
  SEXP state_names = PROTECT(Rf_allocVector(STRSXP, n_state));
  SEXP data_names = PROTECT(Rf_allocVector(STRSXP, n_data));
  for (size_t i = 0; i < n_par; ++i) {
    SET_STRING_ELT(par_names, i, Rf_mkChar(mcb_model_par_name(id, i)));
    REAL(par_defaults)[i] = mcb_model_par_default(id, i);
  }
  for (size_t i = 0; i < n_state; ++i)
    SET_STRING_ELT(state_names, i, Rf_mkChar(mcb_model_state_name(id, i)));
  for (size_t i = 0; i < n_data; ++i)
    SET_STRING_ELT(data_names, i, Rf_mkChar(mcb_model_data_name(id, i)));
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 5));
  SET_VECTOR_ELT(out, 0, Rf_ScalarInteger(id));
  SET_VECTOR_ELT(out, 1, par_names);
  SET_VECTOR_ELT(out, 2, par_defaults);
  SET_VECTOR_ELT(out, 3, state_names);
  SET_VECTOR_ELT(out, 4, data_names);
  UNPROTECT(5);
  return out;
}

/* ---- RNG plumbing on the host ------------------------------------------------ */

/* dust::dust_rng$new(seed)$state(): integer seed -> raw(32) */
SEXP mcbr_rng_seed(SEXP r_seed) {
  SEXP out = PROTECT(Rf_allocVector(RAWSXP, 32));
  chk(mcb_rng_seed(as_u64(r_seed), (uint64_t *)RAW(out)));
  UNPROTECT(1);
  return out;
}

/* dust::dust_rng_distributed_state(seed, n_streams, n_nodes)  R/pmcmc_parallel.R:133,
 * R/smc2.R:68 -> raw(32 * n_streams * n_nodes), node-major */
SEXP mcbr_rng_distributed_state(SEXP r_seed_state, SEXP r_n_streams, SEXP r_n_nodes) {
  const size_t n_streams = (size_t)Rf_asInteger(r_n_streams);
  const size_t n_nodes = (size_t)Rf_asInteger(r_n_nodes);
  if (XLENGTH(r_seed_state) < 32) Rf_error("'seed' must be a raw vector of at least 32 bytes");
  SEXP out = PROTECT(Rf_allocVector(RAWSXP, (R_xlen_t)(32 * n_streams * n_nodes)));
  chk(mcb_rng_distributed_state((const uint64_t *)RAW(r_seed_state), n_streams, n_nodes,
                                (uint64_t *)RAW(out)));
  UNPROTECT(1);
  return out;
}

/* ---- the model object --------------------------------------------------------- */

/* generator$new(pars, time, n_particles, n_threads, seed, gpu_config, pars_multi)
 * R/particle_filter_state.R:237-241.  r_pars: numeric matrix [n_par, max(1, n_pars)] (the R
 * side builds it from the list(s)); r_n_pars: 0 <=> pars_multi = FALSE; r_seed: raw(32 k) */
SEXP mcbr_alloc(SEXP r_model, SEXP r_pars, SEXP r_n_particles, SEXP r_n_pars, SEXP r_time,
                SEXP r_seed, SEXP r_scrambler, SEXP r_deterministic, SEXP r_device) {
  int id = 0;
  size_t n_state = 0, n_par = 0, n_data = 0;
  chk(mcb_model_lookup(CHAR(STRING_ELT(r_model, 0)), &id, &n_state, &n_par, &n_data));
  const size_t n_pars = (size_t)Rf_asInteger(r_n_pars);
  if ((size_t)XLENGTH(r_pars) != n_par * (n_pars ? n_pars : 1))
    Rf_error("'pars' must hold %d values per parameter set", (int)n_par);
  if (XLENGTH(r_seed) < 32 || XLENGTH(r_seed) % 32 != 0)
    Rf_error("'seed' must be a raw vector of 32 bytes per stream");
  mcb_model *m = NULL;
  chk(mcb_alloc(id, REAL(r_pars), (size_t)Rf_asReal(r_n_particles), n_pars, as_u64(r_time),
                (const uint64_t *)RAW(r_seed), (size_t)XLENGTH(r_seed) / 32,
                Rf_asInteger(r_scrambler), Rf_asLogical(r_deterministic),
                Rf_asInteger(r_device), &m));
  SEXP ptr = PROTECT(R_MakeExternalPtr(m, R_NilValue, R_NilValue));
  R_RegisterCFinalizerEx(ptr, finalize, TRUE);
  UNPROTECT(1);
  return ptr;
}

/* $n_state() $n_particles() $n_pars()  R/particle_filter_state.R:268,288,373 */
SEXP mcbr_shape(SEXP ptr) {
  const struct shape s = shape_of(handle(ptr));
  SEXP out = PROTECT(Rf_allocVector(REALSXP, 4));
  REAL(out)[0] = (double)s.n_state;
  REAL(out)[1] = (double)s.n_particles;
  REAL(out)[2] = (double)s.n_pars;
  REAL(out)[3] = (double)s.n_index;
  UNPROTECT(1);
  return out;
}

/* $time() */
SEXP mcbr_time(SEXP ptr) {
  uint64_t t = 0;
  chk(mcb_time(handle(ptr), &t));
  return Rf_ScalarReal((double)t);
}

/* $set_index(index)  R/particle_filter_state.R:261,263 (NULL: every state) */
SEXP mcbr_set_index(SEXP ptr, SEXP r_index) {
  size_t n = 0;
  const size_t *idx = index0(r_index, &n);
  chk(mcb_set_index(handle(ptr), idx, n));
  return R_NilValue;
}

/* $run(time_end) -> [n_index, n_total]  R/particle_filter_state.R:65 */
SEXP mcbr_run(SEXP ptr, SEXP r_time_end) {
  mcb_model *m = handle(ptr);
  const struct shape s = shape_of(m);
  SEXP out = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)(s.n_index * s.n_total)));
  chk(mcb_run(m, as_u64(r_time_end), REAL(out)));
  UNPROTECT(1);
  return out;
}

/* $simulate(time_end) -> [n_index, n_total, n_times]  R/predict.R:79 */
SEXP mcbr_simulate(SEXP ptr, SEXP r_times) {
  mcb_model *m = handle(ptr);
  const struct shape s = shape_of(m);
  const size_t n_times = (size_t)XLENGTH(r_times);
  uint64_t *times = (uint64_t *)R_alloc(n_times ? n_times : 1, sizeof(uint64_t));
  for (size_t i = 0; i < n_times; ++i) times[i] = (uint64_t)REAL(r_times)[i];
  SEXP out = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)(s.n_index * s.n_total * n_times)));
  chk(mcb_simulate(m, times, n_times, REAL(out)));
  UNPROTECT(1);
  return out;
}

/* $state(index) -> [n_rows, n_total]  R/particle_filter_state.R:79,81,114,272 */
SEXP mcbr_state(SEXP ptr, SEXP r_index) {
  mcb_model *m = handle(ptr);
  const struct shape s = shape_of(m);
  size_t n = 0;
  const size_t *idx = index0(r_index, &n);
  const size_t n_rows = idx ? n : s.n_state;
  SEXP out = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)(n_rows * s.n_total)));
  chk(mcb_state(m, idx, n, REAL(out)));
  UNPROTECT(1);
  return out;
}

/* $state()[, i] for a few particles: what pmcmc keeps of a step  R/pmcmc_state.R:40-43 */
SEXP mcbr_state_particles(SEXP ptr, SEXP r_particles) {
  mcb_model *m = handle(ptr);
  const struct shape s = shape_of(m);
  size_t n = 0;
  const size_t *idx = index0(r_particles, &n);
  SEXP out = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)(s.n_state * n)));
  chk(mcb_state_particles(m, idx, n, REAL(out)));
  UNPROTECT(1);
  return out;
}

/* $update_state(pars, state, time, set_initial_state)  R/particle_filter_state.R:248,253,512;
 * R/smc2.R:141.  r_pars: NULL or matrix [n_par, max(1, n_pars)]; r_state: NULL, a vector
 * [n_state] or a matrix [n_state, n_total]; r_time: NULL or a step */
SEXP mcbr_update_state(SEXP ptr, SEXP r_pars, SEXP r_state, SEXP r_time, SEXP r_set_initial) {
  chk(mcb_update_state(handle(ptr), Rf_isNull(r_pars) ? NULL : REAL(r_pars),
                       Rf_isNull(r_state) ? NULL : REAL(r_state),
                       Rf_isNull(r_state) ? 0 : (size_t)XLENGTH(r_state), !Rf_isNull(r_time),
                       Rf_isNull(r_time) ? 0 : as_u64(r_time), Rf_asLogical(r_set_initial)));
  return R_NilValue;
}

/* $pars(): the parameter vectors on the device, [n_par, max(1, n_pars)] */
SEXP mcbr_pars(SEXP ptr, SEXP r_n_par) {
  mcb_model *m = handle(ptr);
  const struct shape s = shape_of(m);
  SEXP out = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)((size_t)Rf_asInteger(r_n_par) * s.n_sets)));
  chk(mcb_pars(m, REAL(out)));
  UNPROTECT(1);
  return out;
}

/* $reorder(kappa): 1-based, within-block for [N, P] input  R/particle_filter_state.R:100 */
SEXP mcbr_reorder(SEXP ptr, SEXP r_kappa) {
  mcb_model *m = handle(ptr);
  const struct shape s = shape_of(m);
  if ((size_t)XLENGTH(r_kappa) != s.n_total) Rf_error("'index' must have one entry per particle");
  chk(mcb_reorder(m, INTEGER(r_kappa)));
  return R_NilValue;
}

/* $rng_state(first_only)  R/particle_filter.R:813,877; R/particle_filter_state.R:363,406 */
SEXP mcbr_rng_state(SEXP ptr, SEXP r_first_only) {
  mcb_model *m = handle(ptr);
  const struct shape s = shape_of(m);
  const int first = Rf_asLogical(r_first_only);
  SEXP out = PROTECT(Rf_allocVector(RAWSXP, (R_xlen_t)(32 * (first ? 1 : s.n_total + 1))));
  chk(mcb_rng_state(m, first, (uint64_t *)RAW(out)));
  UNPROTECT(1);
  return out;
}

/* $set_rng_state(raw)  R/particle_filter_state.R:369,420 */
SEXP mcbr_set_rng_state(SEXP ptr, SEXP r_state) {
  chk(mcb_set_rng_state(handle(ptr), (const uint64_t *)RAW(r_state),
                        (size_t)XLENGTH(r_state) / 8));
  return R_NilValue;
}

/* $set_data(dust_data(...), shared)  R/particle_filter_state.R:243-246.  r_time_end: the
 * steps; r_values: [n_data_fields, shared ? 1 : max(1, n_pars), n_data], NA for missing */
SEXP mcbr_set_data(SEXP ptr, SEXP r_time_end, SEXP r_values, SEXP r_shared) {
  const size_t n = (size_t)XLENGTH(r_time_end);
  uint64_t *t = (uint64_t *)R_alloc(n ? n : 1, sizeof(uint64_t));
  for (size_t i = 0; i < n; ++i) t[i] = (uint64_t)REAL(r_time_end)[i];
  chk(mcb_set_data(handle(ptr), n, t, REAL(r_values), Rf_asLogical(r_shared)));
  return R_NilValue;
}

/* $compare_data(): log-weights [n_total], or NULL when the current time has no datum
 * R/particle_filter_state.R:75-77 */
SEXP mcbr_compare_data(SEXP ptr) {
  mcb_model *m = handle(ptr);
  const struct shape s = shape_of(m);
  int has = 0;
  SEXP out = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)s.n_total));
  chk(mcb_compare_data(m, REAL(out), &has));
  UNPROTECT(1);
  return has ? out : R_NilValue;
}

/* $resample(weights) -> kappa (1-based, within-block) */
SEXP mcbr_resample(SEXP ptr, SEXP r_weights) {
  mcb_model *m = handle(ptr);
  const struct shape s = shape_of(m);
  if ((size_t)XLENGTH(r_weights) != s.n_total) Rf_error("'weights' must have one entry per particle");
  SEXP out = PROTECT(Rf_allocVector(INTSXP, (R_xlen_t)s.n_total));
  chk(mcb_resample(m, REAL(r_weights), INTEGER(out)));
  UNPROTECT(1);
  return out;
}

/* model$filter(time_end, save_trajectories, time_snapshot[, min_log_likelihood])
 * R/particle_filter_state.R:151 -- THE HOT PATH: one call, one CUDA graph launch (or one
 * cooperative kernel), 64 bytes in, 16 bytes out unless trajectories / snapshots are asked
 * for.  r_save: 0 / 1 / 2 (2 = keep the history on the device for mcbr_filter_history) */
SEXP mcbr_filter(SEXP ptr, SEXP r_time_end, SEXP r_save, SEXP r_snap, SEXP r_min_ll) {
  mcb_model *m = handle(ptr);
  const struct shape s = shape_of(m);
  const uint64_t time_end = as_u64(r_time_end);
  size_t n_rows = 0;
  chk(mcb_filter_rows(m, time_end, &n_rows));
  const size_t n_snap = Rf_isNull(r_snap) ? 0 : (size_t)XLENGTH(r_snap);
  const int save = Rf_asInteger(r_save);
  SEXP ll = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)s.n_sets));
  SEXP traj = PROTECT(save == 1 ? Rf_allocVector(REALSXP, (R_xlen_t)(s.n_index * s.n_total * (n_rows + 1)))
                                : R_NilValue);
  SEXP snaps = PROTECT(n_snap ? Rf_allocVector(REALSXP, (R_xlen_t)(s.n_state * s.n_total * n_snap))
                              : R_NilValue);
  uint64_t *snap_t = (uint64_t *)R_alloc(n_snap ? n_snap : 1, sizeof(uint64_t));
  for (size_t i = 0; i < n_snap; ++i) snap_t[i] = (uint64_t)REAL(r_snap)[i];
  chk(mcb_filter(m, time_end, save, snap_t, n_snap, Rf_isNull(r_min_ll) ? NULL : REAL(r_min_ll),
                 Rf_isNull(r_min_ll) ? 0 : (size_t)XLENGTH(r_min_ll), REAL(ll),
                 save == 1 ? REAL(traj) : NULL, n_snap ? REAL(snaps) : NULL));
  SEXP out = PROTECT(Rf_allocVector(VECSXP, 4));
  SET_VECTOR_ELT(out, 0, ll);
  SET_VECTOR_ELT(out, 1, traj);
  SET_VECTOR_ELT(out, 2, snaps);
  SET_VECTOR_ELT(out, 3, Rf_ScalarReal((double)n_rows));
  UNPROTECT(4);
  return out;
}

/* filter$history(index_particle) for a run that kept its history on the device:
 * [n_index, n, n_rows + 1]  R/pmcmc_state.R:32-51, R/particle_filter.R:616-646 */
SEXP mcbr_filter_history(SEXP ptr, SEXP r_particles, SEXP r_n_rows) {
  mcb_model *m = handle(ptr);
  const struct shape s = shape_of(m);
  size_t n = 0;
  const size_t *idx = index0(r_particles, &n);
  const size_t n_rows = (size_t)Rf_asReal(r_n_rows);
  SEXP out = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)(s.n_index * n * (n_rows + 1))));
  chk(mcb_filter_history(m, idx, n, REAL(out)));
  UNPROTECT(1);
  return out;
}

/* ---- sharding (nested populations over several GPUs) ---------------- */

SEXP mcbr_set_filter_stream(SEXP ptr, SEXP r_state) {
  if (XLENGTH(r_state) != 32) Rf_error("a filter stream is 32 bytes");
  chk(mcb_set_filter_stream(handle(ptr), (const uint64_t *)RAW(r_state)));
  return R_NilValue;
}

SEXP mcbr_set_stream_sharing(SEXP ptr, SEXP r_n_sets_global, SEXP r_set_offset, SEXP r_na) {
  chk(mcb_set_stream_sharing(handle(ptr), (size_t)Rf_asInteger(r_n_sets_global),
                             (size_t)Rf_asInteger(r_set_offset),
                             Rf_isNull(r_na) ? NULL : (const unsigned char *)RAW(r_na)));
  return R_NilValue;
}

/* smc2: filter[accept] <- proposed$filter[accept]  R/smc2.R:174-189 */
SEXP mcbr_copy_sets(SEXP dst, SEXP src, SEXP r_take, SEXP r_state, SEXP r_rng) {
  chk(mcb_copy_sets(handle(dst), handle(src), (const unsigned char *)RAW(r_take),
                    Rf_asLogical(r_state), Rf_asLogical(r_rng)));
  return R_NilValue;
}

/* ---- registration --------------------------------------------------------------- */
static const R_CallMethodDef calls[] = {
    {"mcbr_gpu_info", (DL_FUNC)&mcbr_gpu_info, 0},
    {"mcbr_model_info", (DL_FUNC)&mcbr_model_info, 1},
    {"mcbr_rng_seed", (DL_FUNC)&mcbr_rng_seed, 1},
    {"mcbr_rng_distributed_state", (DL_FUNC)&mcbr_rng_distributed_state, 3},
    {"mcbr_alloc", (DL_FUNC)&mcbr_alloc, 9},
    {"mcbr_shape", (DL_FUNC)&mcbr_shape, 1},
    {"mcbr_time", (DL_FUNC)&mcbr_time, 1},
    {"mcbr_set_index", (DL_FUNC)&mcbr_set_index, 2},
    {"mcbr_run", (DL_FUNC)&mcbr_run, 2},
    {"mcbr_simulate", (DL_FUNC)&mcbr_simulate, 2},
    {"mcbr_state", (DL_FUNC)&mcbr_state, 2},
    {"mcbr_state_particles", (DL_FUNC)&mcbr_state_particles, 2},
    {"mcbr_update_state", (DL_FUNC)&mcbr_update_state, 5},
    {"mcbr_pars", (DL_FUNC)&mcbr_pars, 2},
    {"mcbr_reorder", (DL_FUNC)&mcbr_reorder, 2},
    {"mcbr_rng_state", (DL_FUNC)&mcbr_rng_state, 2},
    {"mcbr_set_rng_state", (DL_FUNC)&mcbr_set_rng_state, 2},
    {"mcbr_set_data", (DL_FUNC)&mcbr_set_data, 4},
    {"mcbr_compare_data", (DL_FUNC)&mcbr_compare_data, 1},
    {"mcbr_resample", (DL_FUNC)&mcbr_resample, 2},
    {"mcbr_filter", (DL_FUNC)&mcbr_filter, 5},
    {"mcbr_filter_history", (DL_FUNC)&mcbr_filter_history, 3},
    {"mcbr_set_filter_stream", (DL_FUNC)&mcbr_set_filter_stream, 2},
    {"mcbr_set_stream_sharing", (DL_FUNC)&mcbr_set_stream_sharing, 4},
    {"mcbr_copy_sets", (DL_FUNC)&mcbr_copy_sets, 5},
    {NULL, NULL, 0}};

void R_init_mcstateb200(DllInfo *dll) {
  R_registerRoutines(dll, NULL, calls, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);
}
