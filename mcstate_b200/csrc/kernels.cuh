// The per-interval kernels of the compiled particle filter
// (model$filter, R/particle_filter_state.R:151; per-interval order as
// R/particle_filter_state.R:63-117).
//
// Three launches per data interval, chained by programmatic dependent launch:
//   K1 k_run_compare   [gather of the previous interval's resampling fused into the
//                      load: slot i starts from particle kappa[i]] -> `rate` model
//                      steps -> compare -> state', rng', logw, max
//   K2 k_weights_scan  w = exp(logw - max) as fixed point, per-tile inclusive scan
//                      and tile sum; no block waits for another
//   K3 k_ancestors     every block adds up its set's tile sums (log-mean-exp -> ll and
//                      the early-exit rule by the set's first block), turns the
//                      cumulative weight of each particle into the range of output
//                      slots it owns -- a division, not a search -- and writes the
//                      particle's index at the head of the range, so K1 and every
//                      other state pass read their ancestor with one coalesced load
//                      and a running maximum over the warp
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
// K2 (weights scan) works on tiles of kTile particles, 4 per thread: the tile is the unit of
// the cumulative weights and of the tile sums.  K3 (ancestors) takes 8 particles per thread
// -- a block covers kAncTiles = 2 weight tiles -- because its per-block prologue (the sum of
// the set's tile sums) is worth spreading over more particles.  Measured, K2 / K3 us per
// launch at 2^20 particles (profiles/ab_r02_v23.txt): 4 and 4: 11.0 / 15.7, 8 and 8:
// 12.3 / 14.4, 16 and 16: 14.3 / 15.9.
#ifndef MCB_TILE_ITEMS
#define MCB_TILE_ITEMS 4
#endif
#ifndef MCB_ANC_ITEMS
#define MCB_ANC_ITEMS 8
#endif
constexpr int kTileItems = MCB_TILE_ITEMS;  // consecutive particles per thread in K2
constexpr int kAncItems = MCB_ANC_ITEMS;    // consecutive particles per thread in K3
#ifndef MCB_TILE_MINBLOCKS
#define MCB_TILE_MINBLOCKS 4
#endif
constexpr int kTile = kTileThreads * kTileItems;     // particles per weight tile
constexpr int kAncTile = kTileThreads * kAncItems;   // particles per block of K3
constexpr int kAncTiles = kAncTile / kTile;          // weight tiles per block of K3
static_assert(kAncTile % kTile == 0 && kAncTiles >= 1 && kTileItems % 4 == 0 && kAncItems % 4 == 0 &&
                  kTile % kAncItems == 0,
              "a block of K3 covers whole weight tiles, a thread of K3 stays inside one");
constexpr int kGatherThreads = 256;

typedef unsigned __int128 u128;
constexpr int kShift = 52;   // fractional bits of a fixed-point weight

struct SetInfo {   // the resampling thresholds of a parameter set in one interval
  double uu0, du;  // threshold_i = ceil(uu0 + i * du)   (fixed-point units)
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
  int *heads;                    // [ld] ancestor slots: emptied by K2, filled by K3 (ancestor_of)
  unsigned long long *wmax;      // [n_data][n_sets] order-preserving key of max logw
  const double *u_draws;         // [n_data][n_sets] resampling uniforms (filter stream)
  double *ll;         // [n_sets] accumulated log-likelihood of this filter call
  const double *min_ll; // [n_min]
  int *flags;         // [0] stopped, [1] stop row
  unsigned int *ticket; // [n_sets + 2] K3 counters: [n_sets] sets done, [n_sets + 1] any set dead
  const int *index;   // [n_index] rows returned by run()/saved in trajectories
  size_t n_particles, n_sets, n_total, ld, ld_rng;
  // tiles_per_set: weight tiles (kTile particles) per set; anc_blocks_per_set: blocks of K3
  int n_state, n_index, n_min, data_sets, tiles_per_set, anc_blocks_per_set, deterministic, n_tab,
      n_wrows;
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

// ---- loads of data that streams through once per launch (state, generator, ancestor
// slots): they must not push the memo tables -- read at every model step -- or the
// kernel's few spilled registers out of L1.  MCB_STREAM_LD: 0 evict-first (ld.cs),
// 1 no L1 allocation.
#ifndef MCB_STREAM_LD
#define MCB_STREAM_LD 1
#endif
__device__ __forceinline__ double ld_stream(const double *p) {
#if MCB_STREAM_LD == 1
  double x;
  asm volatile("ld.global.L1::no_allocate.f64 %0, [%1];" : "=d"(x) : "l"(p));
  return x;
#else
  return __ldcs(p);
#endif
}
__device__ __forceinline__ uint64_t ld_stream(const uint64_t *p) {
#if MCB_STREAM_LD == 1
  uint64_t x;
  asm volatile("ld.global.L1::no_allocate.u64 %0, [%1];" : "=l"(x) : "l"(p));
  return x;
#else
  return __ldcs(reinterpret_cast<const unsigned long long *>(p));
#endif
}
__device__ __forceinline__ int ld_stream(const int *p) {
#if MCB_STREAM_LD == 1
  int x;
  asm volatile("ld.global.L1::no_allocate.s32 %0, [%1];" : "=r"(x) : "l"(p));
  return x;
#else
  return __ldcs(p);
#endif
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

// The particle slot i continues, from the slots K3 filled: the running maximum over
// the warp (a slot that is a multiple of 32 is never empty, ancestors grow with the
// slot, an empty slot holds -1).  Every lane of the warp calls this; lanes beyond
// the particles pass live = false.
__device__ __forceinline__ size_t ancestor_of(const int *__restrict__ heads, size_t i, bool live) {
  int h = live ? ld_stream(heads + i) : -1;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int o = __shfl_up_sync(0xffffffffu, h, d);
    if ((int)(threadIdx.x & 31) >= d) h = max(h, o);
  }
  return (size_t)h;
}

// ---------------------------------------------------------------------------
// K1: one thread per particle: [gather of the previous interval's resampling,
// fused into the load] -> `n_steps` model updates -> compare.  State (kState
// doubles) and the generator (4 words) stay in registers throughout.
//   kappa != nullptr: the ancestor slots K3 filled: slot i starts from the particle
//                  ancestor_of gives it (reads y_in[that]; y_in != y_out), so the
//                  resampled state is never written to HBM and read back; with
//                  trajectories the ancestor is also recorded in `order_out`;
//   kappa == nullptr: slot i continues its own particle (y_in may equal y_out);
//   row >= 0: weigh against data row `row`; row < 0: no compare (model$run).
// ---------------------------------------------------------------------------
// The general form of the model steps (state as doubles, every special case of
// the draws), out of line so the run kernel's registers are sized for the lean
// path: deterministic mode, non-integer or out-of-table states.
// State and generator travel BY VALUE: a pointer to the caller's state would pin it to
// local memory for the whole kernel (every thread's frame in L1, next to the memo tables),
// whereas a value is only spilled around this call, which hardly ever happens.
template <class M>
struct ParticleState {
  double y[M::kState];
  uint64_t s0, s1, s2, s3;
};
template <class M, int SCR>
__device__ __noinline__ ParticleState<M> run_general(ParticleState<M> p, uint64_t t0,
                                                     uint32_t n_steps, const double *derived,
                                                     int n_tab, bool det) {
  const SetData sd{derived, n_tab};
  const typename M::Ctx ctx = M::context(sd);
  Xoshiro rng{p.s0, p.s1, p.s2, p.s3};
  for (uint32_t k = 0; k < n_steps; ++k) M::template update<SCR>(t0 + k, p.y, rng, ctx, det);
  p.s0 = rng.s0; p.s1 = rng.s1; p.s2 = rng.s2; p.s3 = rng.s3;
  return p;
}

// `n_steps` model updates of one particle, state and generator in registers: the
// lean path where the model has one and the state allows it, the general code for
// whatever is left (run_general, out of line), or -- for a model without integer
// compartments -- its one form, inline.
template <class M, int SCR>
__device__ __forceinline__ void run_model_steps(const View &v, const SetData &sd,
                                                const typename M::Ctx &ctx, size_t set, double *y,
                                                Xoshiro &rng, uint64_t t0, uint32_t n_steps,
                                                bool det) {
  if (n_steps == 0) return;
  uint32_t done = 0;
  if constexpr (M::kHasLean) {
    int yi[M::kLeanState];
    typename M::Lean lc;
    lc.set = (uint32_t)set;
    const LeanTabs tabs{v.tab_inf, v.tab_walk, (uint32_t)v.n_tab, (uint32_t)v.n_wrows};
    if (!det && M::lean_load(sd, y, yi, lc)) {
      // the usual case: integer compartments, setup from the memo tables, no calls
      // one counter runs the loop (registers are what this kernel is short of); bit 0 of
      // `resets`: the next step starts a new incidence interval.  The mask comes from the
      // host when it could tell (at most 32 steps, one freq for every set), else it is
      // worked out here, 32 steps at a time.
      uint32_t left = n_steps, resets = v.reset_mask;
      do {
        if (!v.reset_mask_valid && ((n_steps - left) & 31u) == 0) {
          const uint32_t freq = (uint32_t)__ldg(sd.derived + M::D_FREQ);
          const uint64_t t = t0 + (n_steps - left);
          uint32_t phase = t <= 0xffffffffULL ? (uint32_t)t % freq : (uint32_t)(t % freq);
          resets = 0;
          const uint32_t nk = min(left, 32u);   // (only as many bits as steps are left)
          for (uint32_t k = 0; k < nk; ++k) {
            resets |= (phase == 0 ? 1u : 0u) << k;
            phase = phase + 1 == freq ? 0 : phase + 1;
          }
        }
        if (!M::template lean_update<SCR>((resets & 1u) != 0, yi, rng, lc, tabs)) break;
        resets >>= 1;
      } while (--left);
      done = n_steps - left;
      M::lean_store(sd, yi, y);
    }
  }
  if constexpr (!M::kHasLean) {
    // a model without integer compartments has one form only: inline
    for (uint32_t k = 0; k < n_steps; ++k) M::template update<SCR>(t0 + k, y, rng, ctx, det);
  } else if (done < n_steps) {  // whatever the lean path did not take
    ParticleState<M> p;
#pragma unroll
    for (int s = 0; s < M::kState; ++s) p.y[s] = y[s];
    p.s0 = rng.s0; p.s1 = rng.s1; p.s2 = rng.s2; p.s3 = rng.s3;
    p = run_general<M, SCR>(p, t0 + done, n_steps - done, sd.derived, sd.n_tab, det);
#pragma unroll
    for (int s = 0; s < M::kState; ++s) y[s] = p.y[s];
    rng.s0 = p.s0; rng.s1 = p.s1; rng.s2 = p.s2; rng.s3 = p.s3;
  }
}

#ifndef MCB_K1_MINBLOCKS
#define MCB_K1_MINBLOCKS 12
#endif
// the warp's maximum goes to the atomic only if it would raise the maximum as it stands (1),
// or always, as a fire-and-forget reduction (0: a same-address atomic per warp, which a
// short launch cannot absorb)
#ifndef MCB_MAX_FILTER
#define MCB_MAX_FILTER 1
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

// HIST: trajectories are recorded.  SINGLE: one parameter set (n_sets == 1): everything
// that depends on the set is a constant, not a register.
template <class M, int SCR, bool HIST, bool SINGLE>
__global__ void __launch_bounds__(kRunThreads, k1_min_blocks<M>())
k_run_compare(View v, const double *__restrict__ y_in, double *__restrict__ y_out,
              uint64_t t0, uint32_t n_steps, const int *__restrict__ kappa, int row,
              double *__restrict__ hist, int *__restrict__ order_out) {
  // (no shared memory and no block barrier anywhere in this kernel: the warps of a block
  // never wait for one another)
  // programmatic dependent launch: let the next kernel be scheduled as this one
  // drains, and do not read anything an earlier kernel wrote before the wait
  grid_launch_dependents();
  grid_dependency_wait();
  if (v.flags[0]) return;
  const size_t i = (size_t)blockIdx.x * kRunThreads + threadIdx.x;
  const bool live = i < v.n_total;
  const size_t ii = live ? i : v.n_total - 1;
  const size_t set = SINGLE ? 0 : set_of(v, ii);
  const SetData sd{v.derived + set * kMaxDerived, v.n_tab};
  const typename M::Ctx ctx = M::context(sd);
  const bool det = v.deterministic != 0;

  double y[M::kState];
  Xoshiro rng;
  unsigned long long key = 0;
  bool weighted = false;
  // which particle this slot continues: the ancestor systematic resampling chose
  const size_t src = kappa ? ancestor_of(kappa, i, live) : i;
  if (live) {
    if (HIST && order_out) order_out[i] = (int)src;
#pragma unroll
    // state and generator stream through once per launch: keep them out of L1's
    // way (evict-first), the memo tables live there
    for (int s = 0; s < M::kState; ++s) y[s] = ld_stream(y_in + (size_t)s * v.ld + src);
    rng.s0 = ld_stream(v.rng + 0 * v.ld_rng + i);
    rng.s1 = ld_stream(v.rng + 1 * v.ld_rng + i);
    rng.s2 = ld_stream(v.rng + 2 * v.ld_rng + i);
    rng.s3 = ld_stream(v.rng + 3 * v.ld_rng + i);

    run_model_steps<M, SCR>(v, sd, ctx, set, y, rng, t0, n_steps, det);

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

  // max of logw per parameter set: one reduction per warp when the warp sits inside one
  // set (the common case), and an atomic only if it would raise the maximum as it stands
  // (the maximum only grows: after the first blocks hardly any warp has to).  Without the
  // look 32768 atomics on one address bound a short launch: the SIR kernel gains 0.3 us
  // (profiles/ab_r02_v14.txt) but the volatility kernel goes from 24 to 34 us
  // (profiles/weights_r02_v25.txt).  Else one atomic per weighted thread
  const size_t first = i - (threadIdx.x & 31);
  size_t last = first + 31;
  if (last >= v.n_total) last = v.n_total - 1;
  const size_t set_first = SINGLE ? 0 : set_of(v, first < v.n_total ? first : v.n_total - 1);
  const bool one_set = SINGLE || set_first == set_of(v, last);
  unsigned long long *slot = v.wmax + (size_t)row * v.n_sets;
  if (one_set) {
    // warp max of a 64-bit key with two 32-bit REDUX: high words, then the low
    // words of the lanes that hold the maximal high word
    const unsigned hi = (unsigned)(key >> 32);
    const unsigned m_hi = __reduce_max_sync(0xffffffffu, hi);
    const unsigned m_lo = __reduce_max_sync(0xffffffffu, hi == m_hi ? (unsigned)key : 0u);
    key = ((unsigned long long)m_hi << 32) | m_lo;
#if MCB_MAX_FILTER
    if ((threadIdx.x & 31) == 0 && key && key > __ldcg(slot + set_first))
      atomicMax(slot + set_first, key);
#else
    if ((threadIdx.x & 31) == 0 && key) atomicMax(slot + set_first, key);
#endif
  } else if (weighted) {
    atomicMax(slot + set, key);
  }
}

// ---------------------------------------------------------------------------
// Per parameter set, once its total weight is known: log-mean-exp into ll
// (scale_log_weights, R/particle_filter.R:570-587).  One thread.  Returns true if
// the set is dead (all -Inf).
// ---------------------------------------------------------------------------
__device__ __forceinline__ bool finish_set(const View &v, int row, size_t set, u128 tot, double tot_d,
                                           double mx) {
  const double neg_inf = __longlong_as_double(0xfff0000000000000LL);
  if (row < 0) return false;
  if (!(mx > neg_inf) || tot == 0) {
    v.ll[set] = neg_inf;
    return true;
  }
  const double mean = (tot_d * 0x1p-52) / (double)v.n_particles;
  v.ll[set] = v.ll[set] + (det_log(mean) + mx);
  return false;
}

// The early-exit rule of R/particle_filter_state.R:463-474 over all the sets, by the
// thread that finished the last set of the interval (a counter) -- the only thread of
// the launch that runs this.
//   ticket [n_sets] the counter of finished sets, [n_sets + 1] "some set is dead"
__device__ __forceinline__ void apply_stop_rule(const View &v, int row, bool dead) {
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
// K2: one block per tile of kTile particles, kTileItems per thread (tiles never straddle a set):
// w = exp(logw - max) -> fixed point, inclusive scan inside the tile -> the
// tile's cumulative weights and its sum.  Nothing here waits for another block:
// the sums are put together by every block of K3 for itself.  The block also
// empties the ancestor slots of its particles' positions for K3 to fill.
// GIVEN = true: `logw` already holds weights in [0,1] (model$resample(weights)).
// ---------------------------------------------------------------------------
template <bool GIVEN>
__global__ void __launch_bounds__(kTileThreads, MCB_TILE_MINBLOCKS) k_weights_scan(View v, int row) {
  __shared__ unsigned long long s_warp[kTileThreads / 32];
  grid_launch_dependents();
  const size_t set = v.n_sets == 1 ? 0 : blockIdx.x / (unsigned)v.tiles_per_set;
  const int tile = (int)(blockIdx.x - (unsigned)set * (unsigned)v.tiles_per_set);
  const size_t in_set = (size_t)tile * kTile + (size_t)threadIdx.x * kTileItems;
  const size_t base = set * v.n_particles + in_set;
  const bool full = in_set + kTileItems <= v.n_particles && (base & 3) == 0;
  grid_dependency_wait();
  // everything the block reads is asked for at once (the stop flag, the set's maximum,
  // the log-weights): one trip to memory, not three in a row
  const int stopped = __ldcg(v.flags);
  double mx = 0.0;
  if (!GIVEN) mx = key_to_double(__ldcg(v.wmax + (size_t)row * v.n_sets + set));
  const bool have = GIVEN || !datum_is_na(v, row, set);   // block-uniform
  const bool dead = !GIVEN && !(mx > __longlong_as_double(0xfff0000000000000LL));
  double lw[kTileItems];
  if (full) {  // 16-byte vector loads: a thread's values are contiguous
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
      lw[j] = in_set + j < v.n_particles ? v.logw[base + j]
                                         : __longlong_as_double(0xfff0000000000000LL);
  }
  // a stopped filter: nothing more is written.  A set without a datum in this row: K3
  // leaves it as it is and reads nothing of this tile
  if (stopped || !have) return;
  if (full) {
    int4 *h = reinterpret_cast<int4 *>(v.heads + base);
#pragma unroll
    for (int j = 0; j < kTileItems / 4; ++j) h[j] = make_int4(-1, -1, -1, -1);
  } else {
#pragma unroll
    for (int j = 0; j < kTileItems; ++j)
      if (in_set + j < v.n_particles) v.heads[base + j] = -1;
  }
  unsigned long long q[kTileItems];
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
  unsigned long long offset = incl - run, total = 0;
#pragma unroll
  for (int w = 0; w < kTileThreads / 32; ++w) {
    if (w < (int)(threadIdx.x >> 5)) offset += s_warp[w];
    total += s_warp[w];
  }
  if (threadIdx.x == 0) v.tile_sum[blockIdx.x] = total;
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

// ---------------------------------------------------------------------------
// K3: the ancestors.  Output slot i of a set continues the first particle whose
// cumulative weight reaches ceil(uu0 + i du); turned round, particle j owns the
// slots [c(C_{j-1}), c(C_j)) where C_j is its inclusive cumulative weight and
// c(C) = #{i : ceil(uu0 + i du) <= C} -- a count that (C - uu0)/du gives directly.
// So nothing is searched: a block (two weight tiles, 8 particles per thread) puts the
// set's tile sums together for itself (what lies before its tiles, and the total: the
// thresholds, and for the set's first block the log-mean-exp -> ll and the early-exit
// rule), every thread turns its 8 cumulative weights into counts, and a particle with
// offspring writes its index at its FIRST slot and at every multiple of 32 its
// slots cover.  Whoever loads by ancestor (ancestor_of) reads its slot and takes the
// running maximum over its warp: one coalesced 4-byte load, no dependent probes, and
// a cost that does not depend on how the weights are spread.
//
// The count is exact.  x = (C - uu0)/du is evaluated in double with an error below
// n 2^-49 slots (C, the tile offset, their difference and the product each round
// once, 2^-53 relative to the total apiece, and n / total stands in for 1 / du to
// within two ulps); a threshold ceil(fl(uu0 + fl(i du)))
// lies within n 2^-52 slots of the line uu0 + i du.  So when x is further than
// kCountEps = 2^-12 from an integer, floor(x) + 1 IS the count (n < 2^31); otherwise
// (one particle in 2000) the thresholds next to the estimate are evaluated as the
// definition has them and compared as integers.
//   heads [n_total]: emptied by K2; global index of the particle at the slots above
// ---------------------------------------------------------------------------
// The sum over a warp of values below 2^110 (a thread's share of the tile sums: each at
// most 2^63, fewer than 2^20 tiles), in 22-bit limbs: a 32-lane integer reduction (REDUX)
// of limbs cannot overflow 32 bits, and five of them replace twenty 64-bit shuffles.
__device__ __forceinline__ u128 warp_sum_u128(u128 x) {
  u128 sum = 0;
#pragma unroll
  for (int l = 0; l < 5; ++l) {
    const unsigned limb = (unsigned)(x >> (22 * l)) & 0x3fffffu;
    sum += (u128)__reduce_add_sync(0xffffffffu, limb) << (22 * l);
  }
  return sum;
}

constexpr double kCountEps = 0x1p-12;
constexpr uint32_t kHeavySlots = 2048;   // a particle with this many slots: the block writes them
constexpr int kHeavyCap = 32;

struct CountCtx {
  double d0, inv_du;   // x = (q + d0) inv_du
  SetInfo si;
  u128 offset;         // cumulative weight before the tile
  uint32_t n;          // particles (= slots) of the set
};

__device__ __noinline__ uint32_t count_exact(const CountCtx &c, unsigned long long q, uint32_t k) {
  const u128 C = c.offset + q;
  while (k < c.n && slot_threshold(c.si, k) <= C) ++k;
  while (k > 0 && slot_threshold(c.si, k - 1) > C) --k;
  return k;
}

// RN(q) for a 64-bit q from its two halves: hi 2^32 is exact, the fma rounds hi 2^32 + lo
// once -- the correctly rounded conversion in three native instructions (the 64-bit
// conversion is a routine)
__device__ __forceinline__ double u64_to_double(unsigned long long q) {
  return fma((double)(uint32_t)(q >> 32), 4294967296.0, (double)(uint32_t)q);
}

// #{slots of the set whose threshold is at or below offset + q}
__device__ __forceinline__ uint32_t slots_reached(const CountCtx &c, unsigned long long q) {
  const double x = (u64_to_double(q) + c.d0) * c.inv_du;
  const double xc = fmin(fmax(x, -1.0), (double)c.n);
  const double fl = floor(xc);
  const uint32_t k = min((uint32_t)(__double2int_rd(xc) + 1), c.n);
  const double fr = xc - fl;
  if (fr < kCountEps || fr > 1.0 - kCountEps) return count_exact(c, q, k);
  return k;
}

__global__ void __launch_bounds__(kTileThreads, MCB_TILE_MINBLOCKS) k_ancestors(View v, int row) {
  __shared__ u128 s_red[2][kTileThreads / 32];
  __shared__ uint32_t s_cnt[kTileThreads / 32];
  __shared__ uint32_t s_heavy[kHeavyCap][3];
  __shared__ int s_heavy_n;
  grid_launch_dependents();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int T = v.tiles_per_set;          // weight tiles of a set (K2's)
  const int NB = v.anc_blocks_per_set;    // blocks of this kernel per set
  const size_t set = v.n_sets == 1 ? 0 : blockIdx.x / (unsigned)NB;
  const int blk = (int)(blockIdx.x - (unsigned)set * (unsigned)NB);
  const int tile = blk * kAncTiles;       // the block's first weight tile
  const size_t gbase = set * v.n_particles;
  const uint32_t n_p = (uint32_t)v.n_particles;
  const uint32_t in_set = (uint32_t)blk * kAncTile + tid * kAncItems;   // the thread's first particle
  const int n_mine = in_set >= n_p ? 0 : (int)min((uint32_t)kAncItems, n_p - in_set);
  const int my_tile = tile + (tid * kAncItems) / kTile;   // the weight tile of the thread's particles
  if (tid == 0) s_heavy_n = 0;
  const size_t base = gbase + in_set;
  const bool vec = n_mine == kAncItems && (base & 1) == 0;
  const unsigned long long *sums = v.tile_sum + set * T;
  grid_dependency_wait();
  // everything the block reads is asked for at once -- the stop flag, the first round of
  // tile sums, the uniform, the maximum, the thread's cumulative weights -- so the block
  // makes one trip to memory, not five in a row (stale values of a set without a datum
  // are read and never used)
  const int stopped = __ldcg(v.flags);
  const unsigned long long sum0 = tid < T ? __ldcg(sums + tid) : 0ULL;
  const double u = __ldcg(v.u_draws + (size_t)(row < 0 ? 0 : row) * v.n_sets + set);
  const unsigned long long mx_key = row >= 0 ? __ldcg(v.wmax + (size_t)row * v.n_sets + set) : 0ULL;
  unsigned long long q[kAncItems];
  if (vec) {
    const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(v.cw + base);
#pragma unroll
    for (int j = 0; j < kAncItems / 2; ++j) {
      const ulonglong2 t = __ldcs(src + j);
      q[2 * j] = t.x;
      q[2 * j + 1] = t.y;
    }
  } else {
#pragma unroll
    for (int j = 0; j < kAncItems; ++j) q[j] = j < n_mine ? __ldcs(v.cw + base + j) : 0ULL;
  }
  if (stopped) return;
  int *heads = v.heads;
  const bool leader = blk == 0 && tid == 0;   // finishes the set
  if (row >= 0 && datum_is_na(v, row, set)) {   // block-uniform: nothing weighed, identity
    for (int k = 0; k < n_mine; ++k) heads[gbase + in_set + k] = (int)(gbase + in_set + k);
    if (leader) apply_stop_rule(v, row, false);
    return;
  }
  // ---- the set's tile sums: what lies before this tile, and the total.  Every thread
  // adds up its share, the warps' sums meet in shared memory
  u128 before = 0, total = 0;
  {
    total = sum0;
    if (tid < tile) before = sum0;
    for (int t = tid + kTileThreads; t < T; t += kTileThreads) {
      const unsigned long long x = __ldcg(sums + t);
      total += x;
      if (t < tile) before += x;
    }
    total = warp_sum_u128(total);
    before = warp_sum_u128(before);
    if (lane == 0) { s_red[0][warp] = total; s_red[1][warp] = before; }
    __syncthreads();
    total = 0; before = 0;
#pragma unroll
    for (int w = 0; w < kTileThreads / 32; ++w) { total += s_red[0][w]; before += s_red[1][w]; }
  }
  const double tot_d = u128_to_double(total);
  const double n_d = (double)v.n_particles;
  CountCtx c;
  // three independent divisions: the two the thresholds are defined by, and n / total for
  // the estimate (1 / du to within an ulp or two, which its error bound allows for)
  c.si.uu0 = tot_d * u / n_d;
  c.si.du = tot_d / n_d;
  c.inv_du = n_d / tot_d;
  c.offset = before;
  c.n = n_p;
  bool dead = false;
  if (row >= 0) {
    const double mx = key_to_double(mx_key);
    dead = !(mx > __longlong_as_double(0xfff0000000000000LL)) || total == 0;
    if (leader) {
      finish_set(v, row, set, total, tot_d, mx);
      apply_stop_rule(v, row, dead);
    }
  } else if (leader) {
    apply_stop_rule(v, row, false);
  }
  if (dead || total == 0) {   // every weight vanished: the set keeps its particles
    for (int k = 0; k < n_mine; ++k) heads[gbase + in_set + k] = (int)(gbase + in_set + k);
    return;
  }
  // the cumulative weights are per weight tile: a thread beyond the block's first tile adds
  // the sums of the tiles in between (one, with two tiles per block)
  for (int t = tile; t < my_tile && t < T; ++t) before += __ldcg(sums + t);
  c.offset = before;
  c.d0 = u128_to_double(before) - c.si.uu0;

  // ---- counts of the thread's 8 particles
  uint32_t cnt[kAncItems];
  {
#pragma unroll
    for (int j = 0; j < kAncItems; ++j) {
      cnt[j] = j < n_mine ? slots_reached(c, q[j]) : n_p;
      if (in_set + j >= n_p - 1) cnt[j] = n_p;   // the last particle takes whatever is left
    }
  }
  // the count before the thread's first particle: the previous thread's last, the
  // count at the tile's offset for the first thread (nothing for the set's first)
  uint32_t prev = __shfl_up_sync(0xffffffffu, cnt[kAncItems - 1], 1);
  if (lane == 31) s_cnt[warp] = cnt[kAncItems - 1];
  __syncthreads();
  if (lane == 0) prev = warp ? s_cnt[warp - 1] : (tile ? slots_reached(c, 0ULL) : 0u);

  // ---- a particle with offspring: its index at its first slot (global slot numbers) ...
  const uint32_t g0 = (uint32_t)gbase + in_set;   // the thread's first particle
  const uint32_t lo = (uint32_t)gbase + prev, hi = (uint32_t)gbase + cnt[kAncItems - 1];
  uint32_t heavy = 0;   // bit j: particle j's slots are written by the whole block
  {
    uint32_t a = prev;
#pragma unroll
    for (int j = 0; j < kAncItems; ++j) {
      const uint32_t b = cnt[j];
      if (b > a) heads[(uint32_t)gbase + a] = (int)(g0 + j);
      if (b - a >= kHeavySlots) {
        const int e = atomicAdd(&s_heavy_n, 1);
        if (e < kHeavyCap) {
          s_heavy[e][0] = (((uint32_t)gbase + a) | 31u) + 1u;
          s_heavy[e][1] = (uint32_t)gbase + b;
          s_heavy[e][2] = g0 + j;
          heavy |= 1u << j;
        }
      }
      a = b;
    }
  }
  // ... and at every multiple of 32 its slots cover.  The thread's 8 particles own the
  // slots [lo, hi) -- about 8 of them, so mostly one multiple of 32 or none: the
  // multiples are walked once per thread, each given to the particle whose range holds it
  for (uint32_t m = (lo | 31u) + 1u; m < hi; m += 32) {
    const uint32_t rel = m - (uint32_t)gbase;
    int j = 0;
#pragma unroll
    for (int k = 0; k < kAncItems - 1; ++k) j += cnt[k] <= rel ? 1 : 0;
    if ((heavy >> j) & 1u) {   // the block writes that particle's: on to the end of its range
      uint32_t end = cnt[0];
#pragma unroll
      for (int k = 1; k < kAncItems; ++k) end = k == j ? cnt[k] : end;
      m = ((((uint32_t)gbase + end - 1u) | 31u) + 1u) - 32u;
      continue;
    }
    heads[m] = (int)(g0 + j);
  }
  __syncthreads();
  const int n_heavy = min(s_heavy_n, kHeavyCap);
  for (int e = 0; e < n_heavy; ++e) {
    const uint32_t B = s_heavy[e][1];
    const int g = (int)s_heavy[e][2];
    for (uint32_t m = s_heavy[e][0] + 32u * tid; m < B; m += 32u * kTileThreads) heads[m] = g;
  }
}

// the slots in place: heads -> the ancestor of every slot (model$resample returns it)
__global__ void __launch_bounds__(kGatherThreads) k_fill_ancestors(View v, int *heads) {
  const size_t i = (size_t)blockIdx.x * kGatherThreads + threadIdx.x;
  const bool live = i < v.n_total;
  const size_t src = ancestor_of(heads, i, live);
  if (live) heads[i] = (int)src;
}

// ---------------------------------------------------------------------------
// Stand-alone gather (src -> dst) by the ancestor slots K3 filled, used where the
// resampled state must exist in memory: snapshots and model$resample.
// ---------------------------------------------------------------------------
template <int NS>
__global__ void __launch_bounds__(kGatherThreads)
k_resample_gather(View v, const int *__restrict__ kappa, const double *__restrict__ src_buf,
                  double *__restrict__ dst_buf) {
  if (v.flags[0]) return;
  const size_t i = (size_t)blockIdx.x * kGatherThreads + threadIdx.x;
  const bool live = i < v.n_total;
  const size_t src = ancestor_of(kappa, i, live);
  if (!live) return;
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
k_finish_state(View v, int last, const int *__restrict__ kappa, int *__restrict__ order_out) {
  const size_t i = (size_t)blockIdx.x * kGatherThreads + threadIdx.x;
  const bool live = i < v.n_total;
  if (v.flags[0]) {
    if (!live) return;
    const int rs = v.flags[1];
    if (((last - 1 - rs) & 1) == 0) {
#pragma unroll
      for (int s = 0; s < NS; ++s) v.y[(size_t)s * v.ld + i] = v.y_tmp[(size_t)s * v.ld + i];
    }
    return;
  }
  const size_t src = kappa ? ancestor_of(kappa, i, live) : i;
  if (!live) return;
  if (order_out) order_out[i] = (int)src;
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
