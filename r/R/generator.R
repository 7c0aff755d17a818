## The `dust_generator` mcstate drives, over libmcstate_b200.so.
##
## mcstate touches its native engine only through this interface:
##   * `is_dust_generator(model)`                       R/particle_filter.R:590-593
##   * static capability queries read from `model$public_methods` WITHOUT an instance
##     (has_compare, has_gpu_support, time_type, ...)   R/particle_filter.R:212,241,991,1029
##   * `model$new(pars, time, n_particles, n_threads, seed, gpu_config, pars_multi)` and the
##     methods of the object it returns               R/particle_filter_state.R:65-151,229-263
## so `particle_filter$new(data, b200_example("sir"), n_particles, compare = NULL)$run(pars)`
## runs unchanged.  Every method is one `.Call` into src/mcstate_b200_r.c; nothing is
## computed in R.  The Python twin that IS tested in this repository is
## mcstate_b200/dust.py (same methods, same argument meaning, same error messages).

`%||%` <- function(a, b) if (is.null(a)) b else a

## integer or NULL seed, raw(32 k) or a dust_rng_pointer-like object -> raw(32 k)
seed_raw <- function(seed) {
  if (is.raw(seed)) {
    if (length(seed) == 0L || length(seed) %% 32L != 0L) {
      stop("Expected raw vector 'seed' to have a length that is a multiple of 32")
    }
    return(seed)
  }
  if (is.null(seed)) {
    seed <- sample.int(.Machine$integer.max, 1L)   # as dust: from R's RNG, so set.seed() works
  }
  if (!is.numeric(seed) || length(seed) != 1L) {
    stop("Invalid type for 'seed'")
  }
  .Call("mcbr_rng_seed", as.numeric(seed), PACKAGE = "mcstateb200")
}

## `pars` (a named list, or an unnamed list of them when pars_multi) -> matrix
## [n_par, n_sets] in the model's positional order; names the model does not know are
## ignored (R/particle_filter.R:283-284 passes the transformed list straight through)
pars_matrix <- function(info, pars, pars_multi) {
  one <- function(p) {
    if (!is.list(p) && !is.null(p)) stop("'pars' must be a list")
    v <- info$par_defaults
    hit <- intersect(names(p), info$par_names)
    v[match(hit, info$par_names)] <- vapply(p[hit], as.numeric, numeric(1))
    v
  }
  if (pars_multi) {
    if (!is.null(names(pars))) stop("Expected 'pars' to be an unnamed list")
    if (length(pars) == 0L) stop("Expected 'pars' to have at least one element")
    m <- vapply(pars, one, numeric(length(info$par_names)))
    matrix(m, nrow = length(info$par_names))
  } else {
    matrix(one(pars), ncol = 1L)
  }
}

#' dust::dust_data(): rows of a data.frame -> list(list(time_end, values), ...) as mcstate's
#' particle_filter_data_split produces (R/particle_filter_data.R:178-203)
#' @export
b200_data <- function(data, name_time = "time_end", multi = NULL) {
  rows <- lapply(seq_len(nrow(data)), function(i) as.list(data[i, , drop = FALSE]))
  if (is.null(multi)) {
    lapply(rows, function(r) list(r[[name_time]], r))
  } else {
    groups <- split(rows, vapply(rows, function(r) r[[name_time]], numeric(1)))
    lapply(unname(groups), function(g) c(list(g[[1L]][[name_time]]), g))
  }
}

#' dust::dust_rng_distributed_state (R/pmcmc_parallel.R:133, R/smc2.R:68)
#' @export
b200_rng_distributed_state <- function(seed = NULL, n_streams = 1L, n_nodes = 1L,
                                       algorithm = "xoshiro256plus") {
  raw <- .Call("mcbr_rng_distributed_state", seed_raw(seed)[1:32], as.integer(n_streams),
               as.integer(n_nodes), PACKAGE = "mcstateb200")
  len <- 32L * n_streams
  lapply(seq_len(n_nodes), function(k) raw[((k - 1L) * len + 1L):(k * len)])
}

#' The generator for one of the engine's stock models ("sir", "sirs", "volatility", "walk"):
#' the stand-in for dust::dust_example(name)
#' @export
b200_example <- function(name = "sir", algorithm = "xoshiro256plus") {
  b200_model(name, algorithm)
}

#' The generator of any model the loaded library holds (a user-supplied model is a library
#' of its own built with -DMCB_USER_MODEL, INTEGRATION.md section 6)
#' @export
b200_model <- function(name, algorithm = "xoshiro256plus") {
  info <- .Call("mcbr_model_info", name, PACKAGE = "mcstateb200")
  names(info) <- c("id", "par_names", "par_defaults", "state_names", "data_names")
  scrambler <- match(algorithm, c("xoshiro256plus", "xoshiro256starstar")) - 1L
  if (is.na(scrambler)) stop(sprintf("Unknown algorithm '%s'", algorithm))
  has_compare_ <- length(info$data_names) > 0L

  gen <- R6::R6Class(
    "dust",
    cloneable = FALSE,
    private = list(
      ptr = NULL, pars_ = NULL, pars_multi_ = FALSE, n_threads_ = 1L, index_names_ = NULL,
      n_rows_ = 0,
      shp = function() {
        s <- .Call("mcbr_shape", private$ptr, PACKAGE = "mcstateb200")
        list(n_state = s[[1L]], n_particles = s[[2L]], n_pars = s[[3L]], n_index = s[[4L]])
      },
      ## [n_rows, n_particles] or [n_rows, n_particles, n_pars] (+ trailing dims)
      dims = function(x, n_rows, extra = NULL) {
        s <- private$shp()
        d <- c(n_rows, s$n_particles, if (private$pars_multi_) s$n_pars, extra)
        dim(x) <- d
        if (!is.null(private$index_names_) && n_rows == length(private$index_names_)) {
          rownames(x) <- private$index_names_
        }
        x
      }),
    public = list(
      ## R/particle_filter_state.R:237-241
      initialize = function(pars, time, n_particles, n_threads = 1L, seed = NULL,
                            pars_multi = FALSE, deterministic = FALSE, gpu_config = NULL) {
        device <- if (is.list(gpu_config)) gpu_config$device_id else (gpu_config %||% 0L)
        m <- pars_matrix(info, pars, pars_multi)
        private$ptr <- .Call("mcbr_alloc", name, m, as.numeric(n_particles),
                             if (pars_multi) ncol(m) else 0L, as.numeric(time),
                             seed_raw(seed), scrambler, isTRUE(deterministic),
                             as.integer(device), PACKAGE = "mcstateb200")
        private$pars_ <- pars
        private$pars_multi_ <- pars_multi
        private$n_threads_ <- n_threads
      },
      name = function() name,
      param = function() NULL,
      ## R/particle_filter_state.R:65
      run = function(time_end) {
        y <- .Call("mcbr_run", private$ptr, as.numeric(time_end), PACKAGE = "mcstateb200")
        private$dims(y, private$shp()$n_index)
      },
      ## R/predict.R:79
      simulate = function(time_end) {
        y <- .Call("mcbr_simulate", private$ptr, as.numeric(time_end), PACKAGE = "mcstateb200")
        private$dims(y, private$shp()$n_index, length(time_end))
      },
      ## R/particle_filter_state.R:261,263
      set_index = function(index) {
        private$index_names_ <- names(index)
        .Call("mcbr_set_index", private$ptr, if (is.null(index)) NULL else as.integer(index),
              PACKAGE = "mcstateb200")
        invisible()
      },
      index = function() private$index_names_,
      ## R/particle_filter_state.R:79,81,114,272
      state = function(index = NULL) {
        y <- .Call("mcbr_state", private$ptr, if (is.null(index)) NULL else as.integer(index),
                   PACKAGE = "mcstateb200")
        n_rows <- if (is.null(index)) private$shp()$n_state else length(index)
        private$index_names_ -> keep
        private$index_names_ <- names(index)
        on.exit(private$index_names_ <- keep)
        private$dims(y, n_rows)
      },
      ## one column of state(): what pmcmc keeps of an accepted step (R/pmcmc_state.R:40-43)
      state_particles = function(index_particle) {
        y <- .Call("mcbr_state_particles", private$ptr, as.integer(index_particle),
                   PACKAGE = "mcstateb200")
        matrix(y, nrow = private$shp()$n_state)
      },
      ## R/particle_filter_state.R:248,253,512; R/smc2.R:141
      update_state = function(pars = NULL, state = NULL, time = NULL,
                              set_initial_state = NULL, index = NULL, reset_step_size = NULL) {
        if (!is.null(index)) stop("'index' is not supported by update_state() of this engine")
        if (is.null(set_initial_state)) {
          set_initial_state <- !is.null(pars) && is.null(state)
        }
        if (set_initial_state && !is.null(state)) {
          stop("Can't use both 'set_initial_state' and provide 'state'")
        }
        m <- NULL
        if (!is.null(pars)) {
          m <- pars_matrix(info, pars, private$pars_multi_)
          private$pars_ <- pars
        }
        .Call("mcbr_update_state", private$ptr, m,
              if (is.null(state)) NULL else as.numeric(state),
              if (is.null(time)) NULL else as.numeric(time), isTRUE(set_initial_state),
              PACKAGE = "mcstateb200")
        invisible()
      },
      ## R/particle_filter_state.R:100
      reorder = function(index) {
        .Call("mcbr_reorder", private$ptr, as.integer(index), PACKAGE = "mcstateb200")
        invisible()
      },
      ## R/particle_filter.R:813,877; R/particle_filter_state.R:363,406
      rng_state = function(first_only = FALSE, last_only = FALSE) {
        if (first_only && last_only) stop("Only one of 'first_only' or 'last_only' may be TRUE")
        raw <- .Call("mcbr_rng_state", private$ptr, isTRUE(first_only), PACKAGE = "mcstateb200")
        if (last_only) raw[(length(raw) - 31L):length(raw)] else raw
      },
      ## R/particle_filter_state.R:369,420
      set_rng_state = function(rng_state) {
        .Call("mcbr_set_rng_state", private$ptr, rng_state, PACKAGE = "mcstateb200")
        invisible()
      },
      ## R/particle_filter_state.R:243-246.  data: list(list(time_end, values...), ...)
      set_data = function(data, shared = FALSE) {
        field <- info$data_names
        time_end <- vapply(data, function(d) as.numeric(d[[1L]]), numeric(1))
        values <- vapply(data, function(d) {
          vapply(d[-1L], function(el) as.numeric(unlist(el[field])), numeric(length(field)))
        }, numeric(length(field) * (length(data[[1L]]) - 1L)))
        .Call("mcbr_set_data", private$ptr, time_end, as.numeric(values), isTRUE(shared),
              PACKAGE = "mcstateb200")
        invisible()
      },
      ## R/particle_filter_state.R:75-77
      compare_data = function() {
        .Call("mcbr_compare_data", private$ptr, PACKAGE = "mcstateb200")
      },
      resample = function(weights) {
        .Call("mcbr_resample", private$ptr, as.numeric(weights), PACKAGE = "mcstateb200")
      },
      ## R/particle_filter_state.R:151 -- the hot path
      filter = function(time_end = NULL, save_trajectories = FALSE, time_snapshot = NULL,
                        min_log_likelihood = NULL) {
        res <- .Call("mcbr_filter", private$ptr, as.numeric(time_end %||% 2^53),
                     as.integer(save_trajectories),
                     if (is.null(time_snapshot)) NULL else as.numeric(time_snapshot),
                     if (is.null(min_log_likelihood)) NULL else as.numeric(min_log_likelihood),
                     PACKAGE = "mcstateb200")
        names(res) <- c("log_likelihood", "trajectories", "snapshots", "n_rows")
        s <- private$shp()
        private$n_rows_ <- res$n_rows
        if (!is.null(res$trajectories)) {
          res$trajectories <- private$dims(res$trajectories, s$n_index, res$n_rows + 1)
        }
        if (!is.null(res$snapshots)) {
          keep <- private$index_names_
          private$index_names_ <- NULL
          res$snapshots <- private$dims(res$snapshots, s$n_state, length(time_snapshot))
          private$index_names_ <- keep
        }
        res["n_rows"] <- NULL
        res
      },
      ## the lineages of chosen particles from a history kept on the device
      ## (filter(save_trajectories = 2L); R/pmcmc_state.R:32-51)
      filter_history = function(index_particle) {
        y <- .Call("mcbr_filter_history", private$ptr, as.integer(index_particle),
                   private$n_rows_, PACKAGE = "mcstateb200")
        array(y, c(private$shp()$n_index, length(index_particle), private$n_rows_ + 1))
      },
      ## nested populations sharded over several GPUs (include/mcstate_b200.h)
      set_filter_stream = function(state) {
        .Call("mcbr_set_filter_stream", private$ptr, state, PACKAGE = "mcstateb200")
        invisible()
      },
      set_stream_sharing = function(n_sets_global, set_offset, na_global = NULL) {
        .Call("mcbr_set_stream_sharing", private$ptr, as.integer(n_sets_global),
              as.integer(set_offset),
              if (is.null(na_global)) NULL else as.raw(t(na_global)), PACKAGE = "mcstateb200")
        invisible()
      },
      ## smc2: filter[accept] <- proposed$filter[accept] (R/smc2.R:174-189)
      copy_sets_from = function(other, take, state = TRUE, rng = TRUE) {
        .Call("mcbr_copy_sets", private$ptr, other$.__enclos_env__$private$ptr, as.raw(take),
              isTRUE(state), isTRUE(rng), PACKAGE = "mcstateb200")
        invisible()
      },
      info = function() {
        one <- function(p) {
          v <- stats::setNames(as.list(info$par_defaults), info$par_names)
          hit <- intersect(names(p), info$par_names)
          v[hit] <- p[hit]
          list(vars = info$state_names, pars = v)
        }
        if (private$pars_multi_) lapply(private$pars_, one) else one(private$pars_)
      },
      pars = function() private$pars_,
      ## the positional parameter vectors as they sit on the device, [n_par, n_sets]
      pars_device = function() {
        m <- .Call("mcbr_pars", private$ptr, length(info$par_names), PACKAGE = "mcstateb200")
        matrix(m, nrow = length(info$par_names), dimnames = list(info$par_names, NULL))
      },
      shape = function() {
        s <- private$shp()
        c(s$n_particles, if (private$pars_multi_) s$n_pars)
      },
      n_state = function() private$shp()$n_state,
      n_particles = function() {
        s <- private$shp()
        s$n_particles * max(1, s$n_pars)
      },
      n_particles_each = function() private$shp()$n_particles,
      n_pars = function() if (private$pars_multi_) private$shp()$n_pars else 0L,
      time = function() .Call("mcbr_time", private$ptr, PACKAGE = "mcstateb200"),
      n_threads = function() private$n_threads_,
      set_n_threads = function(n_threads) {
        prev <- private$n_threads_
        private$n_threads_ <- n_threads
        invisible(prev)
      },
      ## static capability queries (read from $public_methods, no instance)
      has_compare = function() has_compare_,
      has_gpu_support = function(fake_gpu = FALSE) TRUE,
      has_openmp = function() FALSE,
      time_type = function() "discrete",
      real_size = function() 64L,
      rng_algorithm = function() algorithm,
      uses_gpu = function(fake_gpu = FALSE) TRUE,
      gpu_info = function() {
        g <- .Call("mcbr_gpu_info", PACKAGE = "mcstateb200")
        list(has_cuda = length(g[[1L]]) > 0L, cuda_version = NULL,
             devices = data.frame(id = seq_along(g[[1L]]) - 1L, name = g[[1L]],
                                  memory = g[[2L]] / 1024^2, version = g[[4L]],
                                  stringsAsFactors = FALSE))
      }))
  attr(gen, "name") <- "dust_generator"   # is_dust_generator, R/particle_filter.R:590-593
  gen
}
