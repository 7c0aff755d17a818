// The per-interval kernels of the compiled particle filter
// (model$filter, R/particle_filter_state.R:151; per-interval order as
// R/particle_filter_state.R:63-117).
//
// Two launches per data interval:
//   K1 k_run_compare   [resample + gather of the previous interval, fused into the
//                      load] -> `rate` model steps -> compare -> state', rng', logw, max
//   K2 k_weights_scan  w = exp(logw - max) as fixed point, per-tile inclusive scan;
//                      the block that ends a parameter set adds its tile prefix,
//                      log-mean-exp -> ll, resampling thresholds; the last set to
//                      finish applies the early-exit rule
// plus k_finish_state once per filter call (the last interval's resample).
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
// Weights are summed as fixed-point integers (scale_log_weights:
// R/particle_filter.R:570-587; resample_weight: SURVEY.md R2b).  Integer
// addition is associative, so the tiled parallel scan gives the same
// cumulative sums as a serial loop and the chosen ancestors do not depend on
// tile size, SM count or scheduling.
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

struct SetInfo {   // per parameter set, rewritten every interval by K2b
  double uu0, du;  // threshold_i = uu0 + i * du   (fixed-point units)
};

struct View {
  double *y;          // [n_state][ld]   live state
  double *y_tmp;      // [n_state][ld]   post-update, pre-resample
  uint64_t *rng;      // [4][ld_rng]     stream n_total is the filter's
  const double *pars; // [n_sets][kMaxPar]
  double *derived;    // [n_sets][kMaxDerived]  parameter-only constants (Model::derive)
  double2 *tab_inf;   // [n_sets][n_tab]                     infection entries (Model::inf_entry)
  double *tab_walk;   // [n_sets][n_const][n_wrows][kWalkK]  walk tables (Model::walk_row)
  const double *data; // [n_data][data_sets][2] = {value, lgamma(value+1)}
  double *logw;       // [n_total]
  unsigned long long *cw;        // [n_total] inclusive fixed-point cumsum inside a tile
  unsigned long long *tile_incl; // [n_sets][tiles_per_set] tile sums, then inclusive prefix
  unsigned long long *wmax;      // [n_data][n_sets] order-preserving key of max logw
  const double *u_draws;         // [n_data][n_sets] resampling uniforms (filter stream)
  SetInfo *setinfo;   // [n_sets]
  double *ll;         // [n_sets] accumulated log-likelihood of this filter call
  const double *min_ll; // [n_min]
  int *flags;         // [0] stopped, [1] stop row
  unsigned int *ticket; // [n_sets + 2] K2 counters: tiles done per set, sets done, any set dead
  const int *index;   // [n_index] rows returned by run()/saved in trajectories
  size_t n_particles, n_sets, n_total, ld, ld_rng;
  int n_state, n_index, n_min, data_sets, tiles_per_set, shift, deterministic, n_tab, n_wrows;
  int pow_groups;   // groups of three bits beyond bit 0 in a count below n_tab (pow_int_lean)
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

// ---------------------------------------------------------------------------
// Systematic resampling (dust resample_weight; the R twin is
// particle_resample, R/particle_filter.R:515-523): output slot `il` of `set`
// takes the first particle whose inclusive cumulative weight reaches
// ceil(uu0 + il du).  Compared exactly, as integers: binary search over the
// set's tile prefix (8 KB, L1 resident), then inside the tile.  Thresholds grow
// with il, so neighbouring threads probe neighbouring addresses and the
// chosen sources are monotone (near-coalesced gather).
// ---------------------------------------------------------------------------
__device__ __forceinline__ size_t resample_source(const View &v, size_t set, size_t il) {
  const SetInfo si = v.setinfo[set];
  const unsigned long long thr =
      (unsigned long long)__double2ll_ru(si.uu0 + (double)il * si.du);
  const unsigned long long *ts = v.tile_incl + set * v.tiles_per_set;
  int lo = 0, hi = v.tiles_per_set - 1;  // first tile whose inclusive prefix reaches thr
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(ts + mid) >= thr) hi = mid; else lo = mid + 1;
  }
  const unsigned long long before = lo ? __ldg(ts + lo - 1) : 0ULL;
  const unsigned long long rem = thr > before ? thr - before : 0ULL;
  const size_t tile_base = (size_t)lo * kTile;
  const size_t in_tile = min((size_t)kTile, v.n_particles - tile_base);
  const unsigned long long *cw = v.cw + set * v.n_particles + tile_base;
  int a = 0, b = (int)in_tile - 1;
  while (a < b) {
    const int mid = (a + b) >> 1;
    if (__ldg(cw + mid) >= rem) b = mid; else a = mid + 1;
  }
  return set * v.n_particles + tile_base + a;
}

// ---------------------------------------------------------------------------
// The same search done by a whole warp for 32 consecutive output slots of one
// set (the usual case in K1).  Thresholds grow with the slot, so the warp first
// finds lane 0's position with 32-way probes (2 rounds over the tile prefix, 3
// inside a 2048-particle tile, every probe round one coalesced or strided load
// instead of a chain of 20 dependent ones), then each lane finishes inside a
// 64-entry window staged in shared memory.  Lanes whose threshold lies beyond
// the window, and warps whose slots span two tiles, fall back to the per-thread
// search: the result is identical by construction.
// ---------------------------------------------------------------------------
// first index in [0, n) with arr[idx] >= key, else n - 1; uniform across the warp
__device__ __forceinline__ int warp_lower_bound(const unsigned long long *__restrict__ arr, int n,
                                                unsigned long long key, int lane) {
  int lo = 0, cnt = n;
  while (cnt > 1) {
    const int step = (cnt + 31) >> 5;
    const int last = lo + cnt - 1;
    const int idx = min(lo + (lane + 1) * step - 1, last);
    const unsigned b = __ballot_sync(0xffffffffu, __ldg(arr + idx) >= key);
    if (b == 0) return last;
    const int j = __ffs(b) - 1;
    lo += j * step;
    cnt = min(lo + step - 1, last) - lo + 1;
  }
  return lo;
}

constexpr int kWindow = 64;
__device__ __forceinline__ size_t resample_source_warp(const View &v, size_t set, size_t il,
                                                       unsigned long long *s_win, int lane) {
  const SetInfo si = v.setinfo[set];
  const unsigned long long thr =
      (unsigned long long)__double2ll_ru(si.uu0 + (double)il * si.du);
  const unsigned long long thr0 = shfl_u64(thr, 0), thr31 = shfl_u64(thr, 31);
  const unsigned long long *ts = v.tile_incl + set * v.tiles_per_set;
  const int T = v.tiles_per_set;
  const int t0 = warp_lower_bound(ts, T, thr0, lane);
  if (t0 < T - 1 && thr31 > __ldg(ts + t0)) return resample_source(v, set, il);  // two tiles
  const unsigned long long before = t0 ? __ldg(ts + t0 - 1) : 0ULL;
  const unsigned long long rem = thr > before ? thr - before : 0ULL;
  const size_t tile_base = (size_t)t0 * kTile;
  const int in_tile = (int)min((size_t)kTile, v.n_particles - tile_base);
  const unsigned long long *cw = v.cw + set * v.n_particles + tile_base;
  const int a0 = warp_lower_bound(cw, in_tile, shfl_u64(rem, 0), lane);
  // window cw[a0 .. a0 + 63] (clamped to the tile) -> shared memory
  s_win[lane] = __ldg(cw + min(a0 + lane, in_tile - 1));
  s_win[lane + 32] = __ldg(cw + min(a0 + 32 + lane, in_tile - 1));
  __syncwarp();
  int lo = 0, hi = kWindow - 1;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const int mid = (lo + hi) >> 1;
    if (s_win[mid] >= rem) hi = mid; else lo = mid + 1;
  }
  int a = min(a0 + lo, in_tile - 1);
  if (lo == kWindow - 1 && s_win[kWindow - 1] < rem && a0 + kWindow < in_tile) {
    // beyond the window: the plain search over what is left of the tile
    int x = a0 + kWindow, y = in_tile - 1;
    while (x < y) {
      const int mid = (x + y) >> 1;
      if (__ldg(cw + mid) >= rem) y = mid; else x = mid + 1;
    }
    a = x;
  }
  __syncwarp();
  return set * v.n_particles + tile_base + a;
}

// ---------------------------------------------------------------------------
// K1: one thread per particle: [resample of the previous interval, fused into
// the load] -> `n_steps` model updates -> compare.  State (kState doubles) and
// the generator (4 words) stay in registers throughout.
//   prev_row >= 0: slot i starts from the particle systematic resampling picks
//                  for it (reads y_in[src]; y_in != y_out), so the resampled
//                  state is never written to HBM and read back;
//   prev_row <  0: slot i continues its own particle (y_in may equal y_out);
//   row >= 0: weigh against data row `row`; row < 0: no compare (model$run).
// kappa_out (trajectories only) records src, the order slice of prev_row.
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
// compartments -- its one form, inline.  s_inv: the block's {k, 1/k} table.
template <class M, int SCR>
__device__ __forceinline__ void run_model_steps(const View &v, const SetData &sd,
                                                const typename M::Ctx &ctx, size_t set, double *y,
                                                Xoshiro &rng, uint64_t t0, uint32_t n_steps,
                                                bool det, const double2 *s_inv) {
  if (n_steps == 0) return;
  uint32_t done = 0;
  if constexpr (M::kHasLean) {
    int yi[M::kLeanState];
    typename M::Lean lc;
    lc.set = (uint32_t)set;
    const LeanTabs tabs{v.tab_inf, v.tab_walk, (uint32_t)v.n_tab, (uint32_t)v.n_wrows,
                        v.pow_groups};
    if (!det && M::lean_load(sd, y, yi, lc)) {
      // the usual case: integer compartments, setup from the memo tables, no calls
      const uint32_t s_inv_addr = (uint32_t)__cvta_generic_to_shared(s_inv);
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
          ok = M::template lean_update<SCR>((resets & 1u) != 0, yi, rng, lc, tabs, s_inv_addr);
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
              uint64_t t0, uint32_t n_steps, int prev_row, int row,
              double *__restrict__ hist, int *__restrict__ kappa_out) {
  __shared__ unsigned long long s_key[kRunThreads / 32];
  __shared__ unsigned long long s_win[kRunThreads / 32][kWindow];
  __shared__ double2 s_inv[kInvSmall + 1];  // one spare entry (walk_lean looks two ahead)
  // programmatic dependent launch: let the next kernel be scheduled as this one
  // drains, and do not read anything an earlier kernel wrote before the wait
  grid_launch_dependents();
  if (threadIdx.x <= kInvSmall) s_inv[threadIdx.x] = __ldg(g_inv_small + threadIdx.x);
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
  // which particle this slot continues: systematic resampling of prev_row
  size_t src = i;
  {
    const bool want = live && prev_row >= 0 && !datum_is_na(v, prev_row, set);
    const int lane = threadIdx.x & 31;
    const unsigned set0 = __shfl_sync(0xffffffffu, (unsigned)set, 0);  // every lane takes part
    const bool together = want && set0 == (unsigned)set;
    if (__all_sync(0xffffffffu, together))
      src = resample_source_warp(v, set, i - set * v.n_particles, s_win[threadIdx.x >> 5], lane);
    else if (want)
      src = resample_source(v, set, i - set * v.n_particles);
  }
  if (live) {
    if (HIST && kappa_out) kappa_out[i] = (int)src;
#pragma unroll
    // state and generator stream through once per launch: keep them out of L1's
    // way (evict-first), the memo tables and the search arrays live there
    for (int s = 0; s < M::kState; ++s) y[s] = __ldcs(y_in + (size_t)s * v.ld + src);
    rng.s0 = __ldcs(v.rng + 0 * v.ld_rng + i);
    rng.s1 = __ldcs(v.rng + 1 * v.ld_rng + i);
    rng.s2 = __ldcs(v.rng + 2 * v.ld_rng + i);
    rng.s3 = __ldcs(v.rng + 3 * v.ld_rng + i);

    run_model_steps<M, SCR>(v, sd, ctx, set, y, rng, t0, n_steps, det, s_inv);

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
__device__ __forceinline__ bool finish_set(const View &v, int row, size_t set,
                                           unsigned long long tot, double u, double mx,
                                           double ll_old) {
  const double n = (double)v.n_particles;
  const double tot_d = (double)tot;
  SetInfo si;
  si.uu0 = tot_d * u / n;
  si.du = tot_d / n;
  v.setinfo[set] = si;
  if (row < 0) return false;
  if (!(mx > __longlong_as_double(0xfff0000000000000LL)) || tot == 0) {
    v.ll[set] = __longlong_as_double(0xfff0000000000000LL);
    return true;
  }
  const double mean = (tot_d * bits_to_double((uint64_t)(1023 - v.shift) << 52)) / n;
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
  __shared__ unsigned long long s_tot[kTileThreads / 32];
  __shared__ unsigned long long s_carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int T = v.tiles_per_set;
  unsigned long long *ts = v.tile_incl + set * T;
  bool dead = false;
  const bool weighed = row < 0 || !datum_is_na(v, row, set);
  if (weighed && T <= 32) {
    // a small set: one warp
    if (warp == 0) {
      unsigned long long incl = lane < T ? __ldcg(ts + lane) : 0ULL;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long o = shfl_up_u64(incl, d);
        if (lane >= d) incl += o;
      }
      if (lane < T) ts[lane] = incl;
      const unsigned long long tot = shfl_u64(incl, 31);
      if (lane == 0) dead = finish_set(v, row, set, tot, pre_u, pre_mx, pre_ll);
    }
  } else if (weighed) {
    // a big set: the whole block scans its tiles; each thread owns `per`
    // consecutive tiles so a set of up to 256 * 8 tiles is one pass
    constexpr int kPer = 8;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < T; base += kTileThreads * kPer) {
      const int span = min(T - base, kTileThreads * kPer);
      const int per = (span + kTileThreads - 1) / kTileThreads;
      const int lo = base + tid * per, hi = min(base + span, lo + per);
      unsigned long long x[kPer];
      unsigned long long local = 0;
#pragma unroll
      for (int j = 0; j < kPer; ++j) {
        x[j] = (j < per && lo + j < hi) ? __ldcg(ts + lo + j) : 0ULL;
        local += x[j];
        x[j] = local;
      }
      unsigned long long incl = local;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long o = shfl_up_u64(incl, d);
        if (lane >= d) incl += o;
      }
      if (lane == 31) s_tot[warp] = incl;
      __syncthreads();
      unsigned long long offset = s_carry + incl - local;
#pragma unroll
      for (int w = 0; w < kTileThreads / 32; ++w)
        if (w < warp) offset += s_tot[w];
#pragma unroll
      for (int j = 0; j < kPer; ++j)
        if (j < per && lo + j < hi) ts[lo + j] = x[j] + offset;
      __syncthreads();
      if (tid == kTileThreads - 1) s_carry = offset + local;
      __syncthreads();
    }
    if (tid == 0) dead = finish_set(v, row, set, s_carry, pre_u, pre_mx, pre_ll);
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
  bool stop = any_dead;
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

// ---------------------------------------------------------------------------
// K2: one block per tile of kTile particles (tiles never straddle a set):
// w = exp(logw - max) -> fixed point, inclusive scan inside the tile, tile sum.
// The block that takes its set's last ticket then runs finalize_set, so the
// whole weights stage is one launch.
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
  if (GIVEN || !datum_is_na(v, row, set)) {
    if (!GIVEN) mx = key_to_double(v.wmax[(size_t)row * v.n_sets + set]);
    const bool dead = !GIVEN && !(mx > __longlong_as_double(0xfff0000000000000LL));
    const double scale = bits_to_double((uint64_t)(1023 + v.shift) << 52);
    const size_t in_set = (size_t)tile * kTile + (size_t)threadIdx.x * kTileItems;
    const size_t base = set * v.n_particles + in_set;
    const bool full = in_set + kTileItems <= v.n_particles && (base & 1) == 0;
    double lw[kTileItems];
    if (full) {  // 16-byte vector loads: the 8 values of a thread are contiguous
      const double2 *src = reinterpret_cast<const double2 *>(v.logw + base);
#pragma unroll
      for (int j = 0; j < kTileItems / 2; ++j) {
        const double2 t = src[j];
        lw[2 * j] = t.x;
        lw[2 * j + 1] = t.y;
      }
    } else {
#pragma unroll
      for (int j = 0; j < kTileItems; ++j)
        lw[j] = in_set + j < v.n_particles ? v.logw[base + j] : __longlong_as_double(0xfff0000000000000LL);
    }
    unsigned long long q[kTileItems];
    unsigned long long run = 0;
#pragma unroll
    for (int j = 0; j < kTileItems; ++j) {
      unsigned long long x = 0;
      if (in_set + j < v.n_particles && !dead) {
        const double w = GIVEN ? lw[j] : det_exp(lw[j] - mx);
        x = (unsigned long long)__double2ll_rn(w * scale);
      }
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
    unsigned long long offset = incl - run;
    unsigned long long total = 0;
#pragma unroll
    for (int w = 0; w < kTileThreads / 32; ++w) {
      if (w < (int)(threadIdx.x >> 5)) offset += s_warp[w];
      total += s_warp[w];
    }
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
    if (threadIdx.x == 0) v.tile_incl[blockIdx.x] = total;
  }
  // ---- the block that takes the set's last ticket finishes the set.  Only the
  // tile sum (written by thread 0) is read by that block; cw is for the next launch.
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = atomicAdd(v.ticket + set, 1u) == (unsigned)v.tiles_per_set - 1;
  }
  __syncthreads();
  if (!s_last) return;
  if (threadIdx.x == 0) v.ticket[set] = 0;
  __threadfence();
  finalize_set(v, row, set, pre_u, mx, pre_ll);
}

// ---------------------------------------------------------------------------
// Stand-alone resample + gather (src -> dst), used where the resampled state
// must exist in memory: snapshots and model$resample.
// row < 0: every set is resampled; else sets whose datum is NA copy through.
// ---------------------------------------------------------------------------
template <int NS>
__global__ void __launch_bounds__(kGatherThreads)
k_resample_gather(View v, int row, const double *__restrict__ src_buf,
                  double *__restrict__ dst_buf, int *__restrict__ kappa_out) {
  if (v.flags[0]) return;
  const size_t i = (size_t)blockIdx.x * kGatherThreads + threadIdx.x;
  if (i >= v.n_total) return;
  const size_t set = set_of(v, i);
  size_t src = i;
  if (row < 0 || !datum_is_na(v, row, set)) src = resample_source(v, set, i - set * v.n_particles);
#pragma unroll
  for (int s = 0; s < NS; ++s) dst_buf[(size_t)s * v.ld + i] = src_buf[(size_t)s * v.ld + src];
  if (kappa_out) kappa_out[i] = (int)src;
}

// End of a filter call over rows [first, last): bring the live state back into
// v.y.  The last row always wrote v.y_tmp.  Running: resample row last-1 (if it
// was weighed) on the way.  Stopped at row rs: the post-update state of rs, as
// it is (R breaks before resampling), from whichever buffer rs wrote.
template <int NS>
__global__ void __launch_bounds__(kGatherThreads)
k_finish_state(View v, int last, int last_weighted, int *__restrict__ kappa_out) {
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
  const size_t set = set_of(v, i);
  size_t src = i;
  if (last_weighted && !datum_is_na(v, last - 1, set))
    src = resample_source(v, set, i - set * v.n_particles);
#pragma unroll
  for (int s = 0; s < NS; ++s) v.y[(size_t)s * v.ld + i] = v.y_tmp[(size_t)s * v.ld + src];
  if (kappa_out) kappa_out[i] = (int)src;
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
