// The per-interval kernels of the compiled particle filter
// (model$filter, R/particle_filter_state.R:151; per-interval order as
// R/particle_filter_state.R:63-117).
//
// Three launches per data interval, chained by programmatic dependent launch:
//   K1 k_run_compare   [gather of the previous interval's resampling fused into the
//                      load: slot i starts from particle kappa[i]] -> `rate` model
//                      steps -> compare -> state', rng', logw, max
//   K2 k_weights_scan  w = exp(logw - max) as fixed point, per-tile inclusive scan;
//                      the block that ends a parameter set turns the tile sums into a
//                      prefix, log-mean-exp -> ll, resampling thresholds; the last set
//                      to finish applies the early-exit rule
//   K3 k_ancestors     systematic resampling as a merge: one block per 2048 output
//                      slots stages the tile prefix and the two weight tiles its slots
//                      fall into in shared memory and writes kappa (4 B per slot), so
//                      K1 and every other state pass read their ancestor with one
//                      coalesced load and do no search
// plus k_finish_state once per filter call (the last interval's gather).
// Filters that fit the GPU in one wave run as ONE cooperative kernel per call instead
// (persistent.cuh); the model steps (run_model_steps) are shared.
//
// Data layout (HBM): state is SoA [n_state][ld] fp64, RNG is SoA [4][ld_rng]
// u64, both with ld a multiple of 32 so every row starts 256-byte aligned and a
// warp's access to one row is one contiguous 256-byte segment.
// The state ping-pongs between `y` and `y_tmp`, arranged so the last interval
// always writes `y_tmp` and k_finish_state lands the live state in `y`: the
// launch sequence is fixed per row range and replays as a CUDA graph.
//
// Weights are summed as fixed-point integers, rint(w 2^52), in 128 bits
// (scale_log_weights: R/particle_filter.R:570-587; resample_weight: SURVEY.md
// R2b).  Integer addition is associative, so the tiled parallel scan gives the
// same cumulative sums as a serial loop and the chosen ancestors do not depend
// on tile size, SM count or scheduling; 52 fractional bits whatever the number
// of particles keep every weight at least as precisely as a double sum would.
#pragma once
#include <cstdint>

#include "models.cuh"

namespace mcb {

#ifndef MCB_RUN_THREADS
#define MCB_RUN_THREADS 128
#endif
constexpr int kRunThreads = MCB_RUN_THREADS;
constexpr int kTileThreads = 256;
#ifndef MCB_TILE_ITEMS
#define MCB_TILE_ITEMS 8
#endif
constexpr int kTileItems = MCB_TILE_ITEMS;  // consecutive particles per thread in K2
constexpr int kTile = kTileThreads * kTileItems;  // particles per weight tile
constexpr int kGatherThreads = 256;

typedef unsigned __int128 u128;
constexpr int kShift = 52;   // fractional bits of a fixed-point weight

struct SetInfo {   // per parameter set, rewritten every interval by K2's finishing block
  double uu0, du;  // threshold_i = ceil(uu0 + i * du)   (fixed-point units)
  int dead;        // every weight of the interval was zero: the set is not resampled
  int pad;
};

struct View {
  double *y;          // [n_state][ld]   live state
  double *y_tmp;      // [n_state][ld]   post-update, pre-resample
  uint64_t *rng;      // [4][ld_rng]     stream n_total is the filter's
  const double *pars; // [n_sets][kMaxPar]
  double *derived;    // [n_sets][kMaxDerived]  parameter-only constants (Model::derive)
  double4 *tab_inf;   // [n_sets][n_tab]                     infection entries (Model::inf_entry)
  double *tab_walk;   // [n_sets][n_const][n_wrows][kWalkK]  walk tables (Model::walk_row)
  const double *data; // [n_data][data_sets][2] = {value, lgamma(value+1)}
  double *logw;       // [n_total]
  unsigned long long *cw;        // [n_total] inclusive fixed-point cumsum inside a tile
  unsigned long long *tile_sum;  // [n_sets][tiles_per_set] fixed-point weight of each tile
  u128 *tile_pre;                // [n_sets][tiles_per_set] inclusive prefix of the tile sums
  unsigned long long *wmax;      // [n_data][n_sets] order-preserving key of max logw
  const double *u_draws;         // [n_data][n_sets] resampling uniforms (filter stream)
  SetInfo *setinfo;   // [n_sets]
  double *ll;         // [n_sets] accumulated log-likelihood of this filter call
  const double *min_ll; // [n_min]
  int *flags;         // [0] stopped, [1] stop row
  unsigned int *ticket; // [n_sets + 2] K2 counters: tiles done per set, sets done, any set dead
  const int *index;   // [n_index] rows returned by run()/saved in trajectories
  size_t n_particles, n_sets, n_total, ld, ld_rng;
  int n_state, n_index, n_min, data_sets, tiles_per_set, deterministic, n_tab, n_wrows;
  // independent != 0: the sets are separate filters batched in one handle (per-set
  // seeds): a set whose weights all vanish gets -Inf and is not resampled, the others
  // go on.  0: the sets of one multi-parameter filter -- any dead set ends the call
  // (R/particle_filter_state.R:463-474).
  int independent;
  // filter streams (resampling uniforms).  n_fs == 1: one stream shared by every set
  // (the dust pars_multi layout), drawn in set order; when the sets of one model are
  // sharded over several handles, each handle replays the whole order
  // (n_sets_global draws per row, its own sets at [set_offset, set_offset + n_sets))
  // so the shared stream advances identically everywhere.  n_fs == n_sets: one stream
  // per set (independent filters batched in one handle).  Streams n_total .. n_total+n_fs-1.
  int n_fs, n_sets_global, set_offset;
  // bit k: step t0 + k of this launch starts a new incidence interval; valid when the
  // host could tell (same freq in every set, at most 32 steps)
  uint32_t reset_mask;
  int reset_mask_valid;
  const unsigned char *na_global;  // [n_data][n_sets_global] datum is NA, or nullptr
};

// ---- programmatic dependent launch (no-ops for a kernel launched the plain way) --
__device__ __forceinline__ void grid_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;");
}
__device__ __forceinline__ void grid_dependency_wait() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

// ---- order-preserving map double -> u64 so max is one integer atomic --------
__device__ __forceinline__ unsigned long long key_of(double x) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(x);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ULL);
}
__device__ __forceinline__ double key_to_double(unsigned long long k) {
  // key 0 (never written) decodes as -inf: a set nobody weighted is dead
  if (k == 0) return __longlong_as_double(0xfff0000000000000LL);
  const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffULL) : ~k;
  return __longlong_as_double((long long)b);
}

// which parameter set particle i belongs to: nothing to divide for one set, a 32-bit
// division when the indices fit (a 64-bit one is a ~60-instruction routine)
__device__ __forceinline__ size_t set_of(const View &v, size_t i) {
  if (v.n_sets == 1) return 0;
  if (v.n_total <= 0xffffffffULL) return (size_t)((uint32_t)i / (uint32_t)v.n_particles);
  return i / v.n_particles;
}

// is the datum of (row, set) missing?  The look-up itself:
__device__ __forceinline__ bool datum_missing(const View &v, int row, size_t set) {
  const size_t ds = v.data_sets == 1 ? 0 : set;
  const double x = __ldg(v.data + ((size_t)row * v.data_sets + ds) * 2);
  return x != x;
}
// ...and as the per-interval kernels ask it: they are only launched for rows the host
// weighs (row_all_na), so with one data stream for every set (data_sets == 1: a row is
// missing for all sets or for none) there is nothing to look up.
__device__ __forceinline__ bool datum_is_na(const View &v, int row, size_t set) {
  if (v.data_sets == 1) return false;
  const size_t ds = set;
  const double x = __ldg(v.data + ((size_t)row * v.data_sets + ds) * 2);
  return x != x;
}

__device__ __forceinline__ unsigned long long shfl_u64(unsigned long long x, int src) {
  return __shfl_sync(0xffffffffu, x, src);
}
__device__ __forceinline__ unsigned long long shfl_up_u64(unsigned long long x, int d) {
  return __shfl_up_sync(0xffffffffu, x, d);
}
__device__ __forceinline__ unsigned long long shfl_xor_u64(unsigned long long x, int m) {
  return __shfl_xor_sync(0xffffffffu, x, m);
}

__device__ __forceinline__ u128 shfl_up_u128(u128 x, int d) {
  const unsigned long long lo = shfl_up_u64((unsigned long long)x, d);
  const unsigned long long hi = shfl_up_u64((unsigned long long)(x >> 64), d);
  return ((u128)hi << 64) | lo;
}

// the double nearest to x (ties to even): the top 64 bits with a sticky last bit are
// converted (one correctly rounded conversion), then scaled exactly (oracle twin:
// u128_to_double)
__device__ __forceinline__ double u128_to_double(u128 x) {
  const unsigned long long hi = (unsigned long long)(x >> 64), lo = (unsigned long long)x;
  if (hi == 0) return __ull2double_rn(lo);
  const int s = 64 - __clzll((long long)hi);   // bits of hi in use: sums stay below 2^83
  unsigned long long v = (hi << (64 - s)) | (lo >> s);
  if ((lo << (64 - s)) != 0) v |= 1ULL;
  return __ull2double_rn(v) * bits_to_double((uint64_t)(1023 + s) << 52);
}

// ceil(d) as an integer, d >= 0 and finite (doubles from 2^52 up are integers)
__device__ __forceinline__ u128 ceil_to_u128(double d) {
  if (d < 9223372036854775808.0) return (u128)__double2ull_ru(d);
  const uint64_t b = double_to_bits(d);
  const int e = (int)(b >> 52) - 1075;
  const uint64_t m = (b & 0xfffffffffffffULL) | 0x10000000000000ULL;
  return (u128)m << e;
}

// ---------------------------------------------------------------------------
// Systematic resampling (dust resample_weight; the R twin is
// particle_resample, R/particle_filter.R:515-523): output slot `il` of a set
// takes the first particle whose inclusive cumulative weight reaches
// ceil(uu0 + il du), the last one if none does.  Compared exactly, as integers.
// ---------------------------------------------------------------------------
__device__ __forceinline__ u128 slot_threshold(const SetInfo &si, uint32_t il) {
  return ceil_to_u128(si.uu0 + (double)il * si.du);
}

// first index in [lo, n) with pre[idx] >= thr, else n - 1
__device__ __forceinline__ int tile_lower_bound(const u128 *pre, int lo, int n, u128 thr) {
  int hi = n - 1;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (pre[mid] >= thr) hi = mid; else lo = mid + 1;
  }
  return lo;
}

// ---------------------------------------------------------------------------
// K1: one thread per particle: [gather of the previous interval's resampling,
// fused into the load] -> `n_steps` model updates -> compare.  State (kState
// doubles) and the generator (4 words) stay in registers throughout.
//   kappa != nullptr: slot i starts from particle kappa[i], the ancestor K3 chose
//                  for it (reads y_in[kappa[i]]; y_in != y_out), so the resampled
//                  state is never written to HBM and read back;
//   kappa == nullptr: slot i continues its own particle (y_in may equal y_out);
//   row >= 0: weigh against data row `row`; row < 0: no compare (model$run).
// ---------------------------------------------------------------------------
// The general form of the model steps (state as doubles, every special case of
// the draws), out of line so the run kernel's registers are sized for the lean
// path: deterministic mode, non-integer or out-of-table states.
template <class M, int SCR>
__device__ __noinline__ void run_general(double *y, uint64_t *rs, uint64_t t0, uint32_t n_steps,
                                         const double *derived, int n_tab, bool det) {
  const SetData sd{derived, n_tab};
  const typename M::Ctx ctx = M::context(sd);
  Xoshiro rng{rs[0], rs[1], rs[2], rs[3]};
  double z[M::kState];
#pragma unroll
  for (int s = 0; s < M::kState; ++s) z[s] = y[s];
  for (uint32_t k = 0; k < n_steps; ++k) M::template update<SCR>(t0 + k, z, rng, ctx, det);
#pragma unroll
  for (int s = 0; s < M::kState; ++s) y[s] = z[s];
  rs[0] = rng.s0; rs[1] = rng.s1; rs[2] = rng.s2; rs[3] = rng.s3;
}

// `n_steps` model updates of one particle, state and generator in registers: the
// lean path where the model has one and the state allows it, the general code for
// whatever is left (run_general, out of line), or -- for a model without integer
// compartments -- its one form, inline.  s_exp2: the block's copy of g_exp2_tab.
template <class M, int SCR>
__device__ __forceinline__ void run_model_steps(const View &v, const SetData &sd,
                                                const typename M::Ctx &ctx, size_t set, double *y,
                                                Xoshiro &rng, uint64_t t0, uint32_t n_steps,
                                                bool det, const double *s_exp2) {
  if (n_steps == 0) return;
  uint32_t done = 0;
  if constexpr (M::kHasLean) {
    int yi[M::kLeanState];
    typename M::Lean lc;
    lc.set = (uint32_t)set;
    const LeanTabs tabs{v.tab_inf, v.tab_walk, (uint32_t)v.n_tab, (uint32_t)v.n_wrows};
    if (!det && M::lean_load(sd, y, yi, lc)) {
      // the usual case: integer compartments, setup from the memo tables, no calls
      const uint32_t s_exp2_addr = (uint32_t)__cvta_generic_to_shared(s_exp2);
      bool ok = true;
      uint32_t freq = 1, phase = 0;
      if (!v.reset_mask_valid) {
        freq = (uint32_t)__ldg(sd.derived + M::D_FREQ);
        phase = t0 <= 0xffffffffULL ? (uint32_t)t0 % freq : (uint32_t)(t0 % freq);
      }
      while (ok && done < n_steps) {
        // bit k: step done + k starts a new incidence interval
        const uint32_t nk = min(n_steps - done, 32u);
        uint32_t resets = v.reset_mask;
        if (!v.reset_mask_valid) {
          resets = 0;
          for (uint32_t k = 0; k < nk; ++k) {
            resets |= (phase == 0 ? 1u : 0u) << k;
            phase = phase + 1 == freq ? 0 : phase + 1;
          }
        }
        for (uint32_t k = 0; k < nk; ++k) {
          ok = M::template lean_update<SCR>((resets & 1u) != 0, yi, rng, lc, tabs, s_exp2_addr);
          if (!ok) break;
          resets >>= 1;
          ++done;
        }
      }
      M::lean_store(sd, yi, y);
    }
  }
  if constexpr (!M::kHasLean) {
    // a model without integer compartments has one form only: inline
    for (uint32_t k = 0; k < n_steps; ++k) M::template update<SCR>(t0 + k, y, rng, ctx, det);
  } else if (done < n_steps) {  // whatever the lean path did not take
    uint64_t rs[4] = {rng.s0, rng.s1, rng.s2, rng.s3};
    run_general<M, SCR>(y, rs, t0 + done, n_steps - done, sd.derived, sd.n_tab, det);
    rng.s0 = rs[0]; rng.s1 = rs[1]; rng.s2 = rs[2]; rng.s3 = rs[3];
  }
}

#ifndef MCB_K1_MINBLOCKS
#define MCB_K1_MINBLOCKS 12
#endif
// A model without a lean path runs its update as written (doubles, the general samplers)
// and spills at the 40 registers the lean SIR loop is tuned for (SEIR user model: 700 B of
// spill stores per thread).  More registers and fewer resident blocks were measured
// SLOWER all the same (2^20 particles, us per interval; profiles/ab_general_r01_v42.txt):
//   min blocks      12 (40 regs)   8 (64)   6 (80)   4 (125, no spills)
//   seir            210.5          210.9    225.5    272.0
//   volatility      46.6           48.6     55.3
// -- the spills hit L1 and occupancy hides the samplers' fp64 latency -- so the bound is
// shared; the switch stays for a model that behaves differently.
#ifndef MCB_K1_MINBLOCKS_GENERAL
#define MCB_K1_MINBLOCKS_GENERAL MCB_K1_MINBLOCKS
#endif
template <class M>
constexpr int k1_min_blocks() { return M::kHasLean ? MCB_K1_MINBLOCKS : MCB_K1_MINBLOCKS_GENERAL; }

template <class M, int SCR, bool HIST>
__global__ void __launch_bounds__(kRunThreads, k1_min_blocks<M>())
k_run_compare(View v, const double *__restrict__ y_in, double *__restrict__ y_out,
              uint64_t t0, uint32_t n_steps, const int *__restrict__ kappa, int row,
              double *__restrict__ hist) {
  __shared__ unsigned long long s_key[kRunThreads / 32];
  __shared__ double s_exp2[kExp2Tab];
  // programmatic dependent launch: let the next kernel be scheduled as this one
  // drains, and do not read anything an earlier kernel wrote before the wait
  grid_launch_dependents();
  if (threadIdx.x < kExp2Tab) s_exp2[threadIdx.x] = __ldg(g_exp2_tab + threadIdx.x);
  grid_dependency_wait();
  if (v.flags[0]) return;
  __syncthreads();
  const size_t i = (size_t)blockIdx.x * kRunThreads + threadIdx.x;
  const bool live = i < v.n_total;
  const size_t ii = live ? i : v.n_total - 1;
  const size_t set = set_of(v, ii);
  const SetData sd{v.derived + set * kMaxDerived, v.n_tab};
  const typename M::Ctx ctx = M::context(sd);
  const bool det = v.deterministic != 0;

  double y[M::kState];
  Xoshiro rng;
  unsigned long long key = 0;
  bool weighted = false;
  // which particle this slot continues: the ancestor systematic resampling chose
  const size_t src = (live && kappa) ? (size_t)__ldcs(kappa + i) : i;
  if (live) {
#pragma unroll
    // state and generator stream through once per launch: keep them out of L1's
    // way (evict-first), the memo tables and the search arrays live there
    for (int s = 0; s < M::kState; ++s) y[s] = __ldcs(y_in + (size_t)s * v.ld + src);
    rng.s0 = __ldcs(v.rng + 0 * v.ld_rng + i);
    rng.s1 = __ldcs(v.rng + 1 * v.ld_rng + i);
    rng.s2 = __ldcs(v.rng + 2 * v.ld_rng + i);
    rng.s3 = __ldcs(v.rng + 3 * v.ld_rng + i);

    run_model_steps<M, SCR>(v, sd, ctx, set, y, rng, t0, n_steps, det, s_exp2);

    if (n_steps > 0 || y_in != y_out) {
#pragma unroll
      for (int s = 0; s < M::kState; ++s) __stcs(y_out + (size_t)s * v.ld + i, y[s]);
    }
    if (HIST) {
      for (int k = 0; k < v.n_index; ++k) {
        const int s = v.index[k];
        double val = y[0];
#pragma unroll
        for (int q = 1; q < M::kState; ++q) val = (s == q) ? y[q] : val;
        hist[(size_t)k * v.ld + i] = val;
      }
    }
    if (row >= 0 && !datum_is_na(v, row, set)) {
      const size_t ds = v.data_sets == 1 ? 0 : set;
      const double *d = v.data + ((size_t)row * v.data_sets + ds) * 2;
      double lw = M::template compare<SCR>(y, __ldg(d), __ldg(d + 1), rng, ctx, det);
      if (lw != lw) lw = __longlong_as_double(0xfff0000000000000LL);  // NaN -> -Inf
      v.logw[i] = lw;
      key = key_of(lw);
      weighted = true;
    }
    if (n_steps > 0 || weighted) {
      __stcs(v.rng + 0 * v.ld_rng + i, rng.s0);
      __stcs(v.rng + 1 * v.ld_rng + i, rng.s1);
      __stcs(v.rng + 2 * v.ld_rng + i, rng.s2);
      __stcs(v.rng + 3 * v.ld_rng + i, rng.s3);
    }
  }
  if (row < 0) return;

  // max of logw per parameter set: block reduce when the block sits inside
  // one set (the common case), else one atomic per weighted thread
  const size_t first = (size_t)blockIdx.x * kRunThreads;
  size_t last = first + kRunThreads - 1;
  if (last >= v.n_total) last = v.n_total - 1;
  const size_t set_first = set_of(v, first);
  const bool one_set = v.n_sets == 1 || set_first == set_of(v, last);
  unsigned long long *slot = v.wmax + (size_t)row * v.n_sets;
  if (one_set) {
    // warp max of a 64-bit key with two 32-bit REDUX: high words, then the low
    // words of the lanes that hold the maximal high word
    const unsigned hi = (unsigned)(key >> 32);
    const unsigned m_hi = __reduce_max_sync(0xffffffffu, hi);
    const unsigned m_lo = __reduce_max_sync(0xffffffffu, hi == m_hi ? (unsigned)key : 0u);
    key = ((unsigned long long)m_hi << 32) | m_lo;
    if ((threadIdx.x & 31) == 0) s_key[threadIdx.x >> 5] = key;
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
      for (int w = 1; w < kRunThreads / 32; ++w) key = s_key[w] > key ? s_key[w] : key;
      if (key) atomicMax(slot + set_first, key);
    }
  } else if (weighted) {
    atomicMax(slot + set, key);
  }
}

// ---------------------------------------------------------------------------
// Per parameter set, once its total weight is known: log-mean-exp into ll
// (scale_log_weights, R/particle_filter.R:570-587) and the resampling
// thresholds.  One thread.  Returns true if the set is dead (all -Inf).
// ---------------------------------------------------------------------------
// u, mx, ll_old: the set's resampling uniform, max log-weight and likelihood so far,
// loaded by the caller at the start of the kernel (they do not change during it) so
// that no dependent global load sits on the tail of the weights stage.
__device__ __forceinline__ bool finish_set(const View &v, int row, size_t set, u128 tot, double u,
                                           double mx, double ll_old) {
  const double neg_inf = __longlong_as_double(0xfff0000000000000LL);
  const double n = (double)v.n_particles;
  const double tot_d = u128_to_double(tot);
  const bool dead = row >= 0 && (!(mx > neg_inf) || tot == 0);
  SetInfo si;
  si.uu0 = tot_d * u / n;
  si.du = tot_d / n;
  si.dead = dead ? 1 : 0;
  si.pad = 0;
  v.setinfo[set] = si;
  if (row < 0) return false;
  if (dead) {
    v.ll[set] = neg_inf;
    return true;
  }
  const double mean = (tot_d * 0x1p-52) / n;
  v.ll[set] = ll_old + (det_log(mean) + mx);
  return false;
}

// The weights stage finishes per parameter set: the block that completes a set's
// last tile (ticket counter of the set) turns the set's tile sums into an
// inclusive prefix and runs finish_set, so many small sets (nested populations,
// SMC^2) finish in parallel, each on the SM that happened to end it.  The last
// set to finish (a second counter) applies the early-exit rule of
// R/particle_filter_state.R:463-474 over all the sets.
//   ticket [n_sets] per-set counters, [n_sets] the counter of finished sets,
//          [n_sets + 1] "some set is dead" of this interval
__device__ __forceinline__ void finalize_set(const View &v, int row, size_t set, double pre_u,
                                             double pre_mx, double pre_ll) {
  __shared__ u128 s_tot[kTileThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int T = v.tiles_per_set;
  const unsigned long long *sums = v.tile_sum + set * T;
  u128 *pre = v.tile_pre + set * T;
  bool dead = false;
  const bool weighed = row < 0 || !datum_is_na(v, row, set);
  if (weighed && T <= 32) {
    // a small set: one warp
    if (warp == 0) {
      u128 incl = lane < T ? (u128)__ldcg(sums + lane) : (u128)0;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const u128 o = shfl_up_u128(incl, d);
        if (lane >= d) incl += o;
      }
      if (lane < T) pre[lane] = incl;
      const unsigned long long tot_lo = shfl_u64((unsigned long long)incl, T - 1);
      const unsigned long long tot_hi = shfl_u64((unsigned long long)(incl >> 64), T - 1);
      // lane 0 is thread 0: the one that holds the set's uniform and likelihood so far
      if (lane == 0)
        dead = finish_set(v, row, set, ((u128)tot_hi << 64) | tot_lo, pre_u, pre_mx, pre_ll);
    }
  } else if (weighed) {
    // a big set: the whole block scans its tiles, two consecutive tiles per thread: 512
    // tiles (1 Mi particles) per pass, one barrier each
    constexpr int kPer = 2;
    u128 carry = 0;   // the same value in every thread
    for (int base = 0; base < T; base += kTileThreads * kPer) {
      if (base) __syncthreads();   // s_tot is written again
      const int i0 = base + tid * kPer;
      const unsigned long long x0 = i0 < T ? __ldcg(sums + i0) : 0ULL;
      const unsigned long long x1 = i0 + 1 < T ? __ldcg(sums + i0 + 1) : 0ULL;
      const u128 local = (u128)x0 + x1;
      u128 incl = local;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const u128 o = shfl_up_u128(incl, d);
        if (lane >= d) incl += o;
      }
      if (lane == 31) s_tot[warp] = incl;
      __syncthreads();
      u128 offset = carry + incl - local, total = 0;
#pragma unroll
      for (int w = 0; w < kTileThreads / 32; ++w) {
        if (w < warp) offset += s_tot[w];
        total += s_tot[w];
      }
      if (i0 < T) pre[i0] = offset + x0;
      if (i0 + 1 < T) pre[i0 + 1] = offset + local;
      carry += total;
    }
    if (tid == 0) dead = finish_set(v, row, set, carry, pre_u, pre_mx, pre_ll);
  }
  if (tid != 0) return;
  bool any_dead = dead;
  if (v.n_sets > 1) {
    unsigned int *done = v.ticket + v.n_sets;
    if (dead) atomicExch(done + 1, 1u);
    __threadfence();
    if (atomicAdd(done, 1u) != (unsigned)v.n_sets - 1) return;
    // every set of this interval is finished: reset the counters, apply the stop rule
    __threadfence();
    any_dead = atomicExch(done + 1, 0u) != 0;
    *done = 0;
  }
  if (row < 0) return;
  // independent filters batched in one handle: a dead set has its -Inf and the others go on
  bool stop = any_dead && !v.independent;
  if (!stop && v.n_min > 0) {
    const volatile double *ll = v.ll;
    if (v.n_sets == 1) {
      stop = ll[0] < v.min_ll[0];
    } else if (v.n_min == 1) {
      double tot = 0.0;
      for (size_t p = 0; p < v.n_sets; ++p) tot += ll[p];
      stop = tot < v.min_ll[0];
    } else {
      stop = true;
      for (size_t p = 0; p < v.n_sets; ++p) stop = stop && (ll[p] < v.min_ll[p]);
    }
  }
  if (stop) {
    for (size_t p = 0; p < v.n_sets; ++p) v.ll[p] = __longlong_as_double(0xfff0000000000000LL);
    v.flags[1] = row;
    __threadfence();
    v.flags[0] = 1;
  }
}

// rint(w 2^52) for w in [0, 1] without the (quarter-rate) conversion: adding 2^52 lands
// in [2^52, 2^53], where doubles are the integers, so the addition itself rounds to
// nearest-even; the integer is then the distance of the bit patterns
__device__ __forceinline__ unsigned long long weight_to_fixed(double w) {
  const double t = fma(w, 0x1p52, 0x1p52);
  return (unsigned long long)(__double_as_longlong(t) - 0x4330000000000000LL);
}

// ---------------------------------------------------------------------------
// K2: one block per tile of kTile particles (tiles never straddle a set):
// w = exp(logw - max) -> fixed point, inclusive scan inside the tile, tile sum.
// The block that takes its set's last ticket then runs finalize_set, so the
// whole weights stage is one launch.  The tile sum is published (and the ticket
// taken) before the block stores its cumulative weights: only the sum is read
// by the finishing block, the cumulative weights are for the next launch.
// GIVEN = true: `logw` already holds weights in [0,1] (model$resample(weights)).
// ---------------------------------------------------------------------------
template <bool GIVEN>
__global__ void __launch_bounds__(kTileThreads, 4) k_weights_scan(View v, int row) {
  __shared__ unsigned long long s_warp[kTileThreads / 32];
  __shared__ bool s_last;
  grid_launch_dependents();
  grid_dependency_wait();
  if (v.flags[0]) return;
  const size_t set = v.n_sets == 1 ? 0 : blockIdx.x / (unsigned)v.tiles_per_set;
  const int tile = (int)(blockIdx.x - (unsigned)set * (unsigned)v.tiles_per_set);
  // what the set's finishing thread will need, asked for now (finish_set)
  double pre_u = 0.0, pre_ll = 0.0, mx = 0.0;
  if (threadIdx.x == 0) {
    pre_u = v.u_draws[(size_t)(row < 0 ? 0 : row) * v.n_sets + set];
    pre_ll = v.ll[set];
  }
  const bool have = GIVEN || !datum_is_na(v, row, set);   // block-uniform
  const size_t in_set = (size_t)tile * kTile + (size_t)threadIdx.x * kTileItems;
  const size_t base = set * v.n_particles + in_set;
  const bool full = in_set + kTileItems <= v.n_particles && (base & 1) == 0;
  unsigned long long q[kTileItems];
  unsigned long long offset = 0, total = 0;
  if (have) {
    if (!GIVEN) mx = key_to_double(v.wmax[(size_t)row * v.n_sets + set]);
    const bool dead = !GIVEN && !(mx > __longlong_as_double(0xfff0000000000000LL));
    double lw[kTileItems];
    if (full) {  // 16-byte vector loads: the 8 values of a thread are contiguous
      const double2 *src = reinterpret_cast<const double2 *>(v.logw + base);
#pragma unroll
      for (int j = 0; j < kTileItems / 2; ++j) {
        const double2 t = __ldcs(src + j);
        lw[2 * j] = t.x;
        lw[2 * j + 1] = t.y;
      }
    } else {
#pragma unroll
      for (int j = 0; j < kTileItems; ++j)
        lw[j] = in_set + j < v.n_particles ? v.logw[base + j] : __longlong_as_double(0xfff0000000000000LL);
    }
    unsigned long long run = 0;
#pragma unroll
    for (int j = 0; j < kTileItems; ++j) {
      unsigned long long x = 0;
      if (in_set + j < v.n_particles && !dead)
        x = weight_to_fixed(GIVEN ? lw[j] : det_exp_nonpos(lw[j] - mx));
      run += x;
      q[j] = run;
    }
    unsigned long long incl = run;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned long long o = shfl_up_u64(incl, d);
      if ((threadIdx.x & 31) >= d) incl += o;
    }
    if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
    __syncthreads();
    offset = incl - run;
#pragma unroll
    for (int w = 0; w < kTileThreads / 32; ++w) {
      if (w < (int)(threadIdx.x >> 5)) offset += s_warp[w];
      total += s_warp[w];
    }
  }
  // ---- publish the tile sum and take the set's ticket; the block that takes the last
  // one finishes the set
  if (threadIdx.x == 0) {
    if (have) v.tile_sum[blockIdx.x] = total;
    __threadfence();
    s_last = atomicAdd(v.ticket + set, 1u) == (unsigned)v.tiles_per_set - 1;
  }
  if (have) {
    if (full) {
      ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(v.cw + base);
#pragma unroll
      for (int j = 0; j < kTileItems / 2; ++j)
        dst[j] = make_ulonglong2(q[2 * j] + offset, q[2 * j + 1] + offset);
    } else {
#pragma unroll
      for (int j = 0; j < kTileItems; ++j)
        if (in_set + j < v.n_particles) v.cw[base + j] = q[j] + offset;
    }
  }
  __syncthreads();
  if (!s_last) return;
  if (threadIdx.x == 0) v.ticket[set] = 0;
  __threadfence();
  finalize_set(v, row, set, pre_u, mx, pre_ll);
}

// ---------------------------------------------------------------------------
// K3: the ancestors.  Systematic resampling is a merge of two sorted sequences --
// the thresholds ceil(uu0 + i du) of the output slots and the cumulative weights
// of the particles -- so one block takes 2048 consecutive output slots of a set,
// stages the set's tile prefix and a window of two weight tiles in shared memory
// (one coalesced round trip to L2 each), and every thread walks its 8 consecutive
// slots through the window: a halving search when it enters a tile, a short
// bounded search from the previous ancestor after that.  Work follows the OUTPUT,
// so the cost does not depend on how the weights are spread (uniform, degenerate,
// one particle holding all).  When slots of the block fall beyond the window, the
// window moves to the first tile still needed and the threads carry on where they
// stopped: as many rounds as the block's slots have tile pairs to visit (one or
// two unless the weights are very uneven).  A set without a datum in this row, or
// whose weights all vanished, keeps its particles (identity).
//   kappa [n_total]: global index of the particle each slot continues
// ---------------------------------------------------------------------------
constexpr int kAncWinTiles = 2;
constexpr int kAncMaxSmemTiles = 512;
constexpr int kAncNear = 64;   // how far past the previous ancestor the short search looks
static_assert(kTile == 2048, "the searches of k_ancestors run 11 halving steps");

// first index in [0, n) with cw[idx] >= rem, else n - 1; n <= 2048.  Eleven halving
// steps whatever the data (no divergence)
__device__ __forceinline__ int tile_search(const unsigned long long *cw, int n,
                                           unsigned long long rem) {
  int pos = 0;
#pragma unroll
  for (int half = kTile / 2; half >= 1; half >>= 1) {
    const int probe = pos + half - 1;
    if (probe < n && cw[probe] < rem) pos = probe + 1;
  }
  return min(pos, n - 1);
}
// the same for a key at or beyond the one that found `pos`: nothing to do when the
// previous ancestor still reaches it, six halving steps when the answer lies within
// kAncNear particles, the full search otherwise
__device__ __forceinline__ int tile_search_from(const unsigned long long *cw, int pos, int n,
                                                unsigned long long rem) {
  if (cw[pos] >= rem || pos >= n - 1) return pos;
  const int hi = min(pos + kAncNear, n - 1);
  if (cw[hi] < rem) return hi == n - 1 ? hi : tile_search(cw, n, rem);
  int lo = pos + 1;
#pragma unroll
  for (int half = kAncNear / 2; half >= 1; half >>= 1) {
    const int probe = lo + half - 1;
    if (probe < hi && cw[probe] < rem) lo = probe + 1;
  }
  return lo;
}

__global__ void __launch_bounds__(kTileThreads, 4)
k_ancestors(View v, int row, int *__restrict__ kappa) {
  __shared__ u128 s_pre[kAncMaxSmemTiles];
  __shared__ unsigned long long s_cw[kAncWinTiles * kTile];
  __shared__ int s_tile;   // the first tile of the window / the first tile still needed
  grid_launch_dependents();
  grid_dependency_wait();
  if (v.flags[0]) return;
  const int T = v.tiles_per_set;
  const size_t set = v.n_sets == 1 ? 0 : blockIdx.x / (unsigned)T;
  const int tile = (int)(blockIdx.x - (unsigned)set * (unsigned)T);
  const size_t gbase = set * v.n_particles;
  const uint32_t il0 = (uint32_t)tile * kTile + threadIdx.x * kTileItems;  // first slot of the thread
  const uint32_t n_p = (uint32_t)v.n_particles;
  const int n_mine = il0 >= n_p ? 0 : (int)min((uint32_t)kTileItems, n_p - il0);
  int *out = kappa + gbase + il0;
  const SetInfo si = v.setinfo[set];
  if ((row >= 0 && datum_is_na(v, row, set)) || si.dead) {   // block-uniform: identity
    for (int k = 0; k < n_mine; ++k) out[k] = (int)(gbase + il0 + k);
    return;
  }
  const u128 *pre_g = v.tile_pre + set * T;
  const bool pre_staged = T <= kAncMaxSmemTiles;
  const u128 *pre = pre_staged ? s_pre : pre_g;
  // ---- the tile the block's first slot falls into: with the prefix staged, every
  // thread looks at its own entries and the one at the lower bound reports it
  const u128 thr_first = slot_threshold(si, (uint32_t)tile * kTile);
  if (pre_staged) {
    if (threadIdx.x == 0) s_tile = T - 1;   // no prefix reaches the threshold: the last tile
    for (int t = threadIdx.x; t < T; t += kTileThreads) s_pre[t] = pre_g[t];
    __syncthreads();
    for (int t = threadIdx.x; t < T; t += kTileThreads)
      if (s_pre[t] >= thr_first && (t == 0 || s_pre[t - 1] < thr_first)) s_tile = t;
  } else if (threadIdx.x == 0) {
    s_tile = tile_lower_bound(pre_g, 0, T, thr_first);
  }
  __syncthreads();
  int t_w = s_tile;   // first tile of the window (block-uniform)
  const unsigned long long *cw_set = v.cw + gbase;
  // the thread's slots, in order: tile and position only move forward
  int k = 0, t = -1, pos = 0, tlen = 0;
  u128 before = 0, tile_end = 0;
  for (;;) {
    __syncthreads();   // everyone has read s_tile and is done with the previous window
    if (threadIdx.x == 0) s_tile = 0x7fffffff;
    {  // ---- stage tiles t_w and t_w + 1 of the set's in-tile cumulative weights
      const size_t w_base = (size_t)t_w * kTile;
      const int w_len = (int)min((size_t)(kAncWinTiles * kTile), v.n_particles - w_base);
      if (((gbase + w_base) & 1) == 0) {
        const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(cw_set + w_base);
        ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(s_cw);
        for (int j = threadIdx.x; 2 * j + 1 < w_len; j += kTileThreads) dst[j] = __ldcg(src + j);
        if ((w_len & 1) && threadIdx.x == 0) s_cw[w_len - 1] = __ldcg(cw_set + w_base + w_len - 1);
      } else {
        for (int j = threadIdx.x; j < w_len; j += kTileThreads) s_cw[j] = __ldcg(cw_set + w_base + j);
      }
    }
    __syncthreads();
    while (k < n_mine) {
      const u128 thr = slot_threshold(si, il0 + k);
      bool fresh = false;   // a tile this thread has not searched yet
      if (t < 0 || (thr > tile_end && t < T - 1)) {
        const int t_from = t < 0 ? t_w : t + 1;   // (t < 0 only in the first round: t_w = t_a)
        t = (t_from < T - 1 && thr > pre[t_from]) ? tile_lower_bound(pre, t_from + 1, T, thr)
                                                  : t_from;
        before = t ? pre[t - 1] : (u128)0;
        tile_end = pre[t];
        tlen = (int)min((size_t)kTile, v.n_particles - (size_t)t * kTile);
        fresh = true;
        pos = 0;
      }
      if (t >= t_w + kAncWinTiles) {   // beyond the window: ask for it, wait for the next round
        atomicMin(&s_tile, t);   // (pos == 0: the tile is searched afresh next round)
        break;
      }
      const u128 d = thr > before ? thr - before : (u128)0;
      const unsigned long long rem = (unsigned long long)(d >> 64) ? ~0ULL : (unsigned long long)d;
      const unsigned long long *a = s_cw + (t - t_w) * kTile;
      pos = (fresh || pos == 0) ? tile_search(a, tlen, rem) : tile_search_from(a, pos, tlen, rem);
      out[k] = (int)(gbase + (size_t)t * kTile + pos);
      ++k;
    }
    __syncthreads();
    const int t_next = s_tile;
    if (t_next == 0x7fffffff) break;   // every slot of the block has its ancestor
    t_w = t_next;
  }
}

// ---------------------------------------------------------------------------
// Stand-alone gather (src -> dst) by the ancestors K3 wrote, used where the
// resampled state must exist in memory: snapshots and model$resample.
// ---------------------------------------------------------------------------
template <int NS>
__global__ void __launch_bounds__(kGatherThreads)
k_resample_gather(View v, const int *__restrict__ kappa, const double *__restrict__ src_buf,
                  double *__restrict__ dst_buf) {
  if (v.flags[0]) return;
  const size_t i = (size_t)blockIdx.x * kGatherThreads + threadIdx.x;
  if (i >= v.n_total) return;
  const size_t src = (size_t)__ldcs(kappa + i);
#pragma unroll
  for (int s = 0; s < NS; ++s)
    __stcs(dst_buf + (size_t)s * v.ld + i, __ldcs(src_buf + (size_t)s * v.ld + src));
}

// End of a filter call over rows [first, last): bring the live state back into
// v.y.  The last row always wrote v.y_tmp.  Running: the resampling of row last-1
// (kappa != nullptr: it was weighed) on the way.  Stopped at row rs: the post-update
// state of rs, as it is (R breaks before resampling), from whichever buffer rs wrote.
template <int NS>
__global__ void __launch_bounds__(kGatherThreads)
k_finish_state(View v, int last, const int *__restrict__ kappa) {
  const size_t i = (size_t)blockIdx.x * kGatherThreads + threadIdx.x;
  if (i >= v.n_total) return;
  if (v.flags[0]) {
    const int rs = v.flags[1];
    if (((last - 1 - rs) & 1) == 0) {
#pragma unroll
      for (int s = 0; s < NS; ++s) v.y[(size_t)s * v.ld + i] = v.y_tmp[(size_t)s * v.ld + i];
    }
    return;
  }
  const size_t src = kappa ? (size_t)__ldcs(kappa + i) : i;
#pragma unroll
  for (int s = 0; s < NS; ++s)
    __stcs(v.y + (size_t)s * v.ld + i, __ldcs(v.y_tmp + (size_t)s * v.ld + src));
}


// ---- small kernels off the hot path ---------------------------------------

// resampling uniforms for every pending (row, set) pair, drawn in order from the
// filter stream(s); the state after each row is kept so an early stop can rewind to
// the draws that were really made.  Shared stream: one thread walks rows x global
// sets.  Per-set streams: one thread per set.
template <int SCR>
__global__ void k_draw_uniforms(View v, int row_first, int row_last, double *u_draws,
                                uint64_t *fs_after) {
  const size_t f = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= (size_t)v.n_fs) return;
  Xoshiro r;
  const size_t st = v.n_total + f;
  r.s0 = v.rng[0 * v.ld_rng + st]; r.s1 = v.rng[1 * v.ld_rng + st];
  r.s2 = v.rng[2 * v.ld_rng + st]; r.s3 = v.rng[3 * v.ld_rng + st];
  const bool per_set = v.n_fs > 1;
  const int n_glob = per_set ? 1 : v.n_sets_global;
  for (int row = row_first; row < row_last; ++row) {
    const size_t slot = (size_t)(row < 0 ? 0 : row) * v.n_sets;
    for (int pg = 0; pg < n_glob; ++pg) {
      const int p = per_set ? (int)f : pg - v.set_offset;   // local set index
      const bool mine = p >= 0 && p < (int)v.n_sets;
      bool na = false;
      if (row_first >= 0)
        na = (!per_set && v.na_global) ? v.na_global[(size_t)row * n_glob + pg] != 0
                                       : (mine && datum_missing(v, row, (size_t)p));
      if (na) continue;
      const double u = xo_real<SCR>(r);
      if (mine) u_draws[slot + p] = u;
    }
    if (row >= 0) {
      uint64_t *o = fs_after + ((size_t)row * v.n_fs + f) * 4;
      o[0] = r.s0; o[1] = r.s1; o[2] = r.s2; o[3] = r.s3;
    }
  }
  if (row_first < 0) {  // model$resample(): consume the draws immediately
    v.rng[0 * v.ld_rng + st] = r.s0; v.rng[1 * v.ld_rng + st] = r.s1;
    v.rng[2 * v.ld_rng + st] = r.s2; v.rng[3 * v.ld_rng + st] = r.s3;
  }
}

// The shared-stream case of the above as one warp: the lanes look up, 32 (row, set)
// pairs at a time, whether each pair's datum is missing (the only memory reads), and
// lane 0 then walks the pairs in order without waiting on memory -- the serial version
// pays one dependent load per row, ~36 us for 100 rows, which is 2 % of a filter
// evaluation at 2^16 particles.  Same draws in the same order.
template <int SCR>
__global__ void k_draw_uniforms_shared(View v, int row_first, int row_last, double *u_draws,
                                       uint64_t *fs_after) {
  const int lane = threadIdx.x;
  const size_t st = v.n_total;
  Xoshiro r{0, 0, 0, 0};
  if (lane == 0) {
    r.s0 = v.rng[0 * v.ld_rng + st]; r.s1 = v.rng[1 * v.ld_rng + st];
    r.s2 = v.rng[2 * v.ld_rng + st]; r.s3 = v.rng[3 * v.ld_rng + st];
  }
  const int n_glob = v.n_sets_global;
  const long long n_pairs = (long long)(row_last - row_first) * n_glob;
  int row0 = row_first, pg0 = 0;     // lane 0's position: counted, not divided
  for (long long base = 0; base < n_pairs; base += 32) {
    const long long q = base + lane;
    bool na = false;
    if (q < n_pairs && row_first >= 0) {
      const int row = row_first + (int)(q / n_glob), pg = (int)(q % n_glob);
      const int p = pg - v.set_offset;
      const bool mine = p >= 0 && p < (int)v.n_sets;
      na = v.na_global ? v.na_global[(size_t)row * n_glob + pg] != 0
                       : (mine && datum_missing(v, row, (size_t)p));
    }
    const unsigned mask = __ballot_sync(0xffffffffu, na);
    if (lane == 0) {
      const int n_here = (int)min(32LL, n_pairs - base);
      for (int j = 0; j < n_here; ++j) {
        if (!((mask >> j) & 1u)) {
          const double u = xo_real<SCR>(r);
          const int p = pg0 - v.set_offset;
          if (p >= 0 && p < (int)v.n_sets)
            u_draws[(size_t)(row0 < 0 ? 0 : row0) * v.n_sets + p] = u;
        }
        if (++pg0 == n_glob) {
          if (row0 >= 0) {
            uint64_t *o = fs_after + (size_t)row0 * 4;
            o[0] = r.s0; o[1] = r.s1; o[2] = r.s2; o[3] = r.s3;
          }
          pg0 = 0;
          ++row0;
        }
      }
    }
    __syncwarp();
  }
  if (lane == 0 && row_first < 0) {  // model$resample(): consume the draws immediately
    v.rng[0 * v.ld_rng + st] = r.s0; v.rng[1 * v.ld_rng + st] = r.s1;
    v.rng[2 * v.ld_rng + st] = r.s2; v.rng[3 * v.ld_rng + st] = r.s3;
  }
}

// end of a filter call: the filter streams advance past the draws of the rows
// that were resampled (a stopped row was not: R breaks before particle_resample)
__global__ void k_filter_finish(View v, int row_first, int row_last, const uint64_t *fs_after) {
  const size_t f = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= (size_t)v.n_fs) return;
  const int done_last = v.flags[0] ? v.flags[1] - 1 : row_last - 1;
  if (done_last < row_first) return;
  const size_t st = v.n_total + f;
  const uint64_t *o = fs_after + ((size_t)done_last * v.n_fs + f) * 4;
  v.rng[0 * v.ld_rng + st] = o[0];
  v.rng[1 * v.ld_rng + st] = o[1];
  v.rng[2 * v.ld_rng + st] = o[2];
  v.rng[3 * v.ld_rng + st] = o[3];
}

// Once per update_state(pars): the parameter-only constants and memo tables of
// every set (blockIdx.y = set).
template <class M>
__global__ void k_prepare_sets(View v) {
  const size_t set = blockIdx.y;
  double d[kMaxDerived];
#pragma unroll
  for (int k = 0; k < kMaxDerived; ++k) d[k] = 0.0;
  M::derive(v.pars + set * kMaxPar, d);
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e == 0) {
#pragma unroll
    for (int k = 0; k < kMaxDerived; ++k) v.derived[set * kMaxDerived + k] = d[k];
  }
  // one thread per entry or row: n_tab infection entries, then kConst x n_wrows walk rows
  if constexpr (M::kHasLean) {
    if (e < v.n_tab) {
      v.tab_inf[set * (size_t)v.n_tab + e] = M::inf_entry(e, d);
    } else if (e - v.n_tab < M::kConst * v.n_wrows) {
      const int w = e - v.n_tab, c = w / v.n_wrows, n = w - c * v.n_wrows;
      M::walk_row(c, n, d, v.tab_walk + ((set * M::kConst + c) * (size_t)v.n_wrows + n) * kWalkK);
    }
  }
}

template <class M>
__global__ void k_set_initial(View v) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= v.n_total) return;
  const size_t set = set_of(v, i);
  double y[M::kState];
  M::initial(v.pars + set * kMaxPar, y);
#pragma unroll
  for (int s = 0; s < M::kState; ++s) v.y[(size_t)s * v.ld + i] = y[s];
}

// SoA rows -> R layout [n_rows fastest][particle]; index == nullptr: all rows
__global__ void k_pack_rows(const double *__restrict__ y, size_t ld, size_t n_total,
                            const int *__restrict__ index, int n_rows, double *__restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_total) return;
  for (int k = 0; k < n_rows; ++k) {
    const int s = index ? index[k] : k;
    out[i * n_rows + k] = y[(size_t)s * ld + i];
  }
}

// the state of a few chosen particles (pMCMC keeps one: R/pmcmc_state.R:40-43):
// out [n][n_state]
__global__ void k_state_particles(const double *__restrict__ y, size_t ld, int n_state,
                                  const unsigned long long *__restrict__ particles, size_t n,
                                  double *__restrict__ out) {
  const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const size_t src = (size_t)particles[j];
  for (int s = 0; s < n_state; ++s) out[j * n_state + s] = y[(size_t)s * ld + src];
}

// model$simulate: n_slices saved slices [slice][n_rows][ld] -> R layout
// [n_rows fastest][particle][slice] in one pass
__global__ void k_pack_slices(const double *__restrict__ value, size_t ld, size_t n_total,
                              int n_rows, size_t n_slices, double *__restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t s = blockIdx.y;
  if (i >= n_total || s >= n_slices) return;
  for (int k = 0; k < n_rows; ++k)
    out[(s * n_total + i) * n_rows + k] = value[(s * n_rows + k) * ld + i];
}

// R layout (n_state fastest) -> SoA; recycle = 1: one state vector for everyone
__global__ void k_unpack_state(const double *__restrict__ in, int recycle, int n_state,
                               size_t ld, size_t n_total, double *__restrict__ y) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_total) return;
  for (int s = 0; s < n_state; ++s)
    y[(size_t)s * ld + i] = in[(recycle ? 0 : i * n_state) + s];
}

// model$reorder(kappa): kappa 1-based within the set (as R passes it)
__global__ void k_gather_by_kappa(View v, const int *__restrict__ kappa) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= v.n_total) return;
  const size_t set = set_of(v, i);
  const size_t src = set * v.n_particles + (size_t)(kappa[i] - 1);
  for (int s = 0; s < v.n_state; ++s) v.y_tmp[(size_t)s * v.ld + i] = v.y[(size_t)s * v.ld + src];
}

__global__ void k_copy_state_if_running(View v, const double *__restrict__ src,
                                        double *__restrict__ dst) {
  if (v.flags[0]) return;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= v.n_total) return;
  for (int s = 0; s < v.n_state; ++s) dst[(size_t)s * v.ld + i] = src[(size_t)s * v.ld + i];
}

__global__ void k_iota(int *out, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (int)i;
}

// ancestor tracing of the saved history, as history_single
// (R/particle_filter.R:635-642): walk the order arrays back from the final
// particles.  value [T1][n_index][ld], order [T1][n_total] (global, 0-based),
// out in R layout [n_index fastest][particle][time].
__global__ void k_trace_history(View v, const double *__restrict__ value,
                                const int *__restrict__ order, int rows, int row_first,
                                double *__restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= v.n_total) return;
  const int T1 = rows + 1;
  const bool stopped = v.flags[0] != 0;
  const int filled = stopped ? (v.flags[1] - row_first) + 2 : T1;
  const double na = __longlong_as_double(0x7ff8000000000000LL);
  size_t cur = i;
  for (int t = T1 - 1; t >= 0; --t) {
    double *o = out + ((size_t)t * v.n_total + i) * v.n_index;
    if (t >= filled) {
      for (int k = 0; k < v.n_index; ++k) o[k] = na;
      continue;
    }
    if (!(stopped && t == filled - 1)) cur = (size_t)order[(size_t)t * v.n_total + cur];
    for (int k = 0; k < v.n_index; ++k)
      o[k] = value[((size_t)t * v.n_index + k) * v.ld + cur];
  }
}

// The same walk for a few chosen final particles (what pMCMC keeps of a run: one
// sampled lineage, R/pmcmc_state.R:32-51): one thread per requested particle, so
// the history stays in HBM and only n x n_index x T1 doubles leave the device.
// out [T1][n][n_index].
__global__ void k_trace_particles(View v, const double *__restrict__ value,
                                  const int *__restrict__ order, int rows, int row_first,
                                  const unsigned long long *__restrict__ particles, size_t n,
                                  double *__restrict__ out) {
  const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int T1 = rows + 1;
  const bool stopped = v.flags[0] != 0;
  const int filled = stopped ? (v.flags[1] - row_first) + 2 : T1;
  const double na = __longlong_as_double(0x7ff8000000000000LL);
  size_t cur = (size_t)particles[j];
  for (int t = T1 - 1; t >= 0; --t) {
    double *o = out + ((size_t)t * n + j) * v.n_index;
    if (t >= filled) {
      for (int k = 0; k < v.n_index; ++k) o[k] = na;
      continue;
    }
    if (!(stopped && t == filled - 1)) cur = (size_t)order[(size_t)t * v.n_total + cur];
    for (int k = 0; k < v.n_index; ++k)
      o[k] = value[((size_t)t * v.n_index + k) * v.ld + cur];
  }
}

// stream i (i >= have) = J^(i - have + 1) applied to stream have-1, by the
// precomputed powers J^(2^l) (rng.cuh).  The column loads are warp-uniform.
__global__ void k_derive_streams(uint64_t *rng, size_t ld_rng, size_t have, size_t n_streams,
                                 const uint64_t *__restrict__ jump_cols) {
  const size_t i = have + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_streams) return;
  uint64_t s[4];
  for (int w = 0; w < 4; ++w) s[w] = rng[w * ld_rng + (have - 1)];
  const uint64_t e = (uint64_t)(i - have + 1);
  for (int l = 0; l < kJumpLevels; ++l) {
    if (!((e >> l) & 1ULL)) continue;
    const uint64_t *cols = jump_cols + (size_t)l * 256 * 4;
    uint64_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    for (int b = 0; b < 256; ++b) {
      const uint64_t m = 0ULL - ((s[b >> 6] >> (b & 63)) & 1ULL);
      a0 ^= m & __ldg(cols + b * 4 + 0);
      a1 ^= m & __ldg(cols + b * 4 + 1);
      a2 ^= m & __ldg(cols + b * 4 + 2);
      a3 ^= m & __ldg(cols + b * 4 + 3);
    }
    s[0] = a0; s[1] = a1; s[2] = a2; s[3] = a3;
  }
  for (int w = 0; w < 4; ++w) rng[w * ld_rng + i] = s[w];
}

// smc2 replenishment (R/smc2.R:174-189: filter[accept] <- proposed$filter[accept]): take
// the particle state and/or the RNG streams of the flagged sets from another handle of
// the same shape.
__global__ void k_copy_sets(View dst, View src, const unsigned char *__restrict__ take,
                            int copy_state, int copy_rng) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < dst.n_total) {
    const size_t set = i / dst.n_particles;
    if (take[set]) {
      if (copy_state)
        for (int s = 0; s < dst.n_state; ++s) dst.y[(size_t)s * dst.ld + i] = src.y[(size_t)s * src.ld + i];
      if (copy_rng)
        for (int w = 0; w < 4; ++w) dst.rng[w * dst.ld_rng + i] = src.rng[w * src.ld_rng + i];
    }
  }
  if (copy_rng && dst.n_fs > 1 && i < (size_t)dst.n_fs && take[i]) {  // the set's filter stream
    for (int w = 0; w < 4; ++w)
      dst.rng[w * dst.ld_rng + dst.n_total + i] = src.rng[w * src.ld_rng + src.n_total + i];
  }
}

// Per-set seeding (independent filters batched in one handle): set p's particle
// streams are J^j base_p, j < n_particles, and its filter stream J^n_particles base_p
// -- what n_sets separate model objects seeded with base_p would hold.
__global__ void k_derive_streams_per_set(uint64_t *rng, size_t ld_rng,
                                         const uint64_t *__restrict__ bases, size_t n_particles,
                                         size_t n_sets, const uint64_t *__restrict__ jump_cols) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t per = n_particles + 1;
  if (t >= per * n_sets) return;
  const size_t p = t / per, j = t - p * per;
  uint64_t s[4];
  for (int w = 0; w < 4; ++w) s[w] = bases[4 * p + w];
  for (int l = 0; l < kJumpLevels; ++l) {
    if (!((j >> l) & 1ULL)) continue;
    const uint64_t *cols = jump_cols + (size_t)l * 256 * 4;
    uint64_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    for (int b = 0; b < 256; ++b) {
      const uint64_t m = 0ULL - ((s[b >> 6] >> (b & 63)) & 1ULL);
      a0 ^= m & __ldg(cols + b * 4 + 0);
      a1 ^= m & __ldg(cols + b * 4 + 1);
      a2 ^= m & __ldg(cols + b * 4 + 2);
      a3 ^= m & __ldg(cols + b * 4 + 3);
    }
    s[0] = a0; s[1] = a1; s[2] = a2; s[3] = a3;
  }
  const size_t dst = j < n_particles ? p * n_particles + j : n_particles * n_sets + p;
  for (int w = 0; w < 4; ++w) rng[w * ld_rng + dst] = s[w];
}

}  // namespace mcb
