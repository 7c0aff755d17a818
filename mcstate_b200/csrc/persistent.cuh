// The whole filter call as ONE cooperative kernel, for filters that fit the GPU in
// a single wave (one particle per thread, every block resident): BASELINE config
// C3 -- a pMCMC chain of 2^16 particles per GPU -- is this case.
//
// A filter of 2^16 particles is latency bound under the per-interval launches of
// kernels.cuh: its kernels run for a few microseconds each and the three launch
// boundaries per interval and the HBM round trip of the generator state dominate.
// Here a thread owns output slot i for all the data rows: the generator never leaves
// its registers, the log-weight goes from the compare straight into the scan, and
// the intervals are separated by two grid-wide barriers instead of kernel boundaries:
//
//   per data row:  [gather: slot i takes the particle systematic resampling picks]
//                  -> n_steps model updates -> compare -> state to HBM, block max
//                  == grid barrier ==
//                  w = exp(logw - max) as fixed point -> block-wide inclusive scan
//                  -> block sum and in-block cumulative weights to HBM
//                  == grid barrier ==
//                  every block: prefix over the block sums (shared memory),
//                  log-mean-exp, thresholds and the early-exit rule, redundantly
//                  and identically in every block
//
// Arithmetic, draw order and the integer weight sums are those of kernels.cuh --
// integer addition is associative, so the tiling (256 here, 2048 there) does not
// change a cumulative sum -- and the ancestors (searched here, counted there) are the
// same: the results are bit-identical to the per-interval path and to the oracle
// (tests/test_engine_gpu.py).
//
// Used for one parameter set without snapshots; everything else takes the general
// path.  Reference semantics: R/particle_filter_state.R:63-117 (loop),
// R/particle_filter.R:570-587 (log-mean-exp), :515-523 / SURVEY.md R2b
// (resampling), R/particle_filter_state.R:463-474 (early exit).
#pragma once
#include <cooperative_groups.h>

#include "kernels.cuh"

namespace mcb {

// 256-thread blocks: half as many block sums to prefix in every block and wider tiles for
// the search than 128 (measured per interval: 15.3 vs 16.6 us at 2^16 particles, 15.6 vs
// 24.5 at 75 000, 22.8 vs 31.6 at 150 000; 13.0 vs 12.3 at 4096)
#ifndef MCB_PERSIST_THREADS
#define MCB_PERSIST_THREADS 256
#endif
constexpr int kPersistThreads = MCB_PERSIST_THREADS;
constexpr int kPersistMaxBlocks = 1184 * 128 / kPersistThreads;   // 151 552 particles (8 x 128 threads per SM on 148 SMs)
// the single launch is used up to this many blocks (all that can be co-resident: it beats
// the two-launch path over the whole range, 22.8 vs 28.3 us per interval at 150 000)
#ifndef MCB_PERSIST_USE_BLOCKS
#define MCB_PERSIST_USE_BLOCKS (1184 * 128 / MCB_PERSIST_THREADS)
#endif
constexpr int kPersistUseBlocks = MCB_PERSIST_USE_BLOCKS;
#ifndef MCB_PERSIST_MINBLOCKS
#define MCB_PERSIST_MINBLOCKS (8 * 128 / MCB_PERSIST_THREADS)
#endif

struct PersistArgs {
  const uint64_t *data_time;     // [n_data] model step at the end of each row
  unsigned long long *blk_sum;   // [n_blocks] fixed-point weight of each block
  int first, last;               // rows [first, last)
  uint64_t t_start;              // model step the state is at
  int first_out_tmp;             // the first weighed row writes y_tmp (else y)
  // trajectories (or nullptr): values [rows + 1][n_index][ld] of the post-update states,
  // order [rows + 1][n_total] = the particle each slot continued (kernels.cuh K1, HIST)
  double *hist_value;
  int *hist_order;
};

// first index in [lo, hi] (hi = last valid) with arr[idx] >= key, arr in shared memory
template <class K>
__device__ __forceinline__ int smem_lower_bound(const K *arr, int lo, int hi, K key) {
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (arr[mid] >= key) hi = mid; else lo = mid + 1;
  }
  return lo;
}

template <class M, int SCR>
__global__ void __launch_bounds__(kPersistThreads, MCB_PERSIST_MINBLOCKS)
k_filter_persistent(View v, PersistArgs a) {
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  __shared__ u128 s_pre[kPersistMaxBlocks];   // inclusive prefix of block sums (128-bit: kernels.cuh)
  __shared__ unsigned long long s_win[kPersistThreads / 32][2 * kPersistThreads];
  __shared__ unsigned long long s_warp[kPersistThreads / 32];
  __shared__ u128 s_warp2[kPersistThreads / 32];
  __shared__ unsigned long long s_key[kPersistThreads / 32];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_blocks = gridDim.x;
  const size_t i = (size_t)blockIdx.x * kPersistThreads + tid;
  const size_t N = v.n_total;
  const bool live = i < N;

  const SetData sd{v.derived, v.n_tab};
  const typename M::Ctx ctx = M::context(sd);
  const bool det = v.deterministic != 0;
  const double n_d = (double)v.n_particles;
  const double neg_inf = __longlong_as_double(0xfff0000000000000LL);

  double y[M::kState];
  Xoshiro rng{0, 0, 0, 0};
  if (live) {
#pragma unroll
    for (int s = 0; s < M::kState; ++s) y[s] = v.y[(size_t)s * v.ld + i];
    rng.s0 = v.rng[0 * v.ld_rng + i];
    rng.s1 = v.rng[1 * v.ld_rng + i];
    rng.s2 = v.rng[2 * v.ld_rng + i];
    rng.s3 = v.rng[3 * v.ld_rng + i];
  }

  const bool hist = a.hist_value != nullptr;
  auto save_slice = [&](int slice, size_t src) {   // slice 0 = the state before the first row
    if (!live) return;
    for (int k = 0; k < v.n_index; ++k) {
      const int s = v.index[k];
      double val = y[0];
#pragma unroll
      for (int q = 1; q < M::kState; ++q) val = (s == q) ? y[q] : val;
      a.hist_value[((size_t)slice * v.n_index + k) * v.ld + i] = val;
    }
    (void)src;
  };
  if (hist) {
    save_slice(0, i);
    if (live) a.hist_order[i] = (int)i;
  }

  double ll = 0.0;            // the same value in every thread
  bool stopped = false;
  int stop_row = 0;
  bool owe_resample = false;  // the previous weighed row's resampling is still to be applied
  double uu0 = 0.0, du = 0.0;
  const double *buf_prev = nullptr;          // where the weighed row's post-update state went
  bool out_tmp = a.first_out_tmp != 0;
  uint64_t t = a.t_start;
  bool touched = false;       // the generator moved

  // slot i <- the particle systematic resampling picks: first j whose inclusive
  // cumulative weight reaches ceil(uu0 + i du).  Level 1: the block, in the shared
  // prefix.  Level 2: inside the block, in its 128 in-block cumulative weights --
  // staged in shared memory by the warp when its 32 slots fall into at most two
  // neighbouring blocks (thresholds grow with the slot), else searched in place.
  auto resample_source_p = [&](size_t slot) -> size_t {
    const u128 thr = ceil_to_u128(uu0 + (double)(uint32_t)slot * du);
    const int nb = n_blocks;
    const int b = smem_lower_bound<u128>(s_pre, 0, nb - 1, thr);
    const u128 before = b ? s_pre[b - 1] : (u128)0;
    const u128 dd = thr > before ? thr - before : (u128)0;
    const unsigned long long rem = (unsigned long long)(dd >> 64) ? ~0ULL : (unsigned long long)dd;
    const int b0 = __shfl_sync(0xffffffffu, b, 0);
    const int b31 = __shfl_sync(0xffffffffu, b, 31);
    const size_t base = (size_t)b * kPersistThreads;
    const int in_blk = (int)min((size_t)kPersistThreads, N - base);
    int pos;
    if (b31 - b0 <= 1) {
      unsigned long long *win = s_win[warp];
      const size_t wbase = (size_t)b0 * kPersistThreads;
#pragma unroll
      for (int k = 0; k < 2 * kPersistThreads / 32; ++k) {
        const size_t idx = wbase + (size_t)k * 32 + lane;
        win[k * 32 + lane] = idx < N ? __ldcg(v.cw + idx) : ~0ULL;
      }
      __syncwarp();
      const int off = (b - b0) * kPersistThreads;
      pos = smem_lower_bound<unsigned long long>(win + off, 0, in_blk - 1, rem);
      __syncwarp();
    } else {
      const unsigned long long *cw = v.cw + base;
      int lo = 0, hi = in_blk - 1;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldcg(cw + mid) >= rem) hi = mid; else lo = mid + 1;
      }
      pos = lo;
    }
    return base + pos;
  };

  for (int row = a.first; row < a.last; ++row) {
    const uint64_t t_end = __ldg(a.data_time + row);
    const uint32_t n_steps = (uint32_t)(t_end - t);
    const bool weigh = !datum_missing(v, row, 0);

    // ---- gather (resampling of the previous weighed row), update, compare
    const int rel = row - a.first;
    if (owe_resample) {
      const size_t src = resample_source_p(live ? i : N - 1);
      if (live) {
#pragma unroll
        for (int s = 0; s < M::kState; ++s) y[s] = __ldcg(buf_prev + (size_t)s * v.ld + src);
        if (hist) a.hist_order[(size_t)rel * N + i] = (int)src;
      }
      owe_resample = false;
    } else if (hist && rel > 0 && live) {
      a.hist_order[(size_t)rel * N + i] = (int)i;
    }
    double lw = neg_inf;
    if (live) {
      run_model_steps<M, SCR>(v, sd, ctx, 0, y, rng, t, n_steps, det);
      touched = touched || n_steps > 0;
      if (hist) save_slice(rel + 1, i);
      if (weigh) {
        const double *d = v.data + (size_t)row * v.data_sets * 2;
        lw = M::template compare<SCR>(y, __ldg(d), __ldg(d + 1), rng, ctx, det);
        if (lw != lw) lw = neg_inf;  // NaN -> -Inf
        touched = true;
      }
    }
    t = t_end;
    if (!weigh) continue;   // a row without data: no weights, no resampling

    double *out = out_tmp ? v.y_tmp : v.y;
    out_tmp = !out_tmp;
    if (live) {
#pragma unroll
      for (int s = 0; s < M::kState; ++s) __stcg(out + (size_t)s * v.ld + i, y[s]);
    }
    buf_prev = out;
    {
      unsigned long long key = live ? key_of(lw) : 0ULL;
      const unsigned hi = (unsigned)(key >> 32);
      const unsigned m_hi = __reduce_max_sync(0xffffffffu, hi);
      const unsigned m_lo = __reduce_max_sync(0xffffffffu, hi == m_hi ? (unsigned)key : 0u);
      key = ((unsigned long long)m_hi << 32) | m_lo;
      if (lane == 0) s_key[warp] = key;
      __syncthreads();
      if (tid == 0) {
#pragma unroll
        for (int w = 1; w < kPersistThreads / 32; ++w) key = s_key[w] > key ? s_key[w] : key;
        if (key) atomicMax(v.wmax + row, key);
      }
    }
    grid.sync();

    // ---- weights: fixed point, block-wide inclusive scan
    const double mx = key_to_double(__ldcg(v.wmax + row));
    const bool dead = !(mx > neg_inf);
    unsigned long long x = 0;
    if (live && !dead) {
      x = weight_to_fixed(det_exp_nonpos(lw - mx));
    }
    unsigned long long incl = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned long long o = shfl_up_u64(incl, d);
      if (lane >= d) incl += o;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    unsigned long long total = 0;
#pragma unroll
    for (int w = 0; w < kPersistThreads / 32; ++w) {
      if (w < warp) incl += s_warp[w];
      total += s_warp[w];
    }
    if (live) __stcg(v.cw + i, incl);
    if (tid == 0) __stcg(a.blk_sum + blockIdx.x, total);
    grid.sync();

    // ---- every block: inclusive prefix of the block sums into shared memory
    {
      constexpr int kPer = (kPersistMaxBlocks + kPersistThreads - 1) / kPersistThreads;  // 10
      const int per = (n_blocks + kPersistThreads - 1) / kPersistThreads;
      const int lo = tid * per;
      // a block sum is at most 2^60 (256 weights of at most 2^52): a thread's kPer <= 10
      // of them fit 64 bits, the sums across threads are 128-bit
      static_assert(kPer <= 15, "a thread's block sums must fit 64 bits");
      unsigned long long xs[kPer];
      unsigned long long local = 0;
#pragma unroll
      for (int j = 0; j < kPer; ++j) {
        if (j < per && lo + j < n_blocks) local += __ldcg(a.blk_sum + lo + j);
        xs[j] = local;
      }
      u128 inc2 = local;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const u128 o = shfl_up_u128(inc2, d);
        if (lane >= d) inc2 += o;
      }
      if (lane == 31) s_warp2[warp] = inc2;
      __syncthreads();
      u128 offset = inc2 - local;
#pragma unroll
      for (int w = 0; w < kPersistThreads / 32; ++w)
        if (w < warp) offset += s_warp2[w];
#pragma unroll
      for (int j = 0; j < kPer; ++j)
        if (j < per && lo + j < n_blocks) s_pre[lo + j] = xs[j] + offset;
      __syncthreads();
    }
    // ---- finish the interval (finish_set + the early-exit rule), in every thread alike
    const u128 tot = s_pre[n_blocks - 1];
    const double tot_d = u128_to_double(tot);
    const double u = __ldg(v.u_draws + row);
    uu0 = tot_d * u / n_d;
    du = tot_d / n_d;
    bool stop;
    if (dead || tot == 0) {
      stop = true;
    } else {
      const double mean = (tot_d * 0x1p-52) / n_d;
      ll += det_log(mean) + mx;
      stop = v.n_min > 0 && ll < __ldg(v.min_ll);
    }
    if (stop) {
      stopped = true;
      stop_row = row;
      break;
    }
    owe_resample = true;
  }

  // ---- end of the call: the live state into v.y (the last weighed row wrote y_tmp,
  // so the gather never reads what it writes), generator and results back
  if (!stopped && owe_resample) {
    const size_t src = resample_source_p(live ? i : N - 1);
    if (live) {
#pragma unroll
      for (int s = 0; s < M::kState; ++s) y[s] = __ldcg(buf_prev + (size_t)s * v.ld + src);
      if (hist) a.hist_order[(size_t)(a.last - a.first) * N + i] = (int)src;
    }
  } else if (!stopped && hist && live) {
    a.hist_order[(size_t)(a.last - a.first) * N + i] = (int)i;
  }
  if (live) {
#pragma unroll
    for (int s = 0; s < M::kState; ++s) v.y[(size_t)s * v.ld + i] = y[s];
    if (touched) {
      v.rng[0 * v.ld_rng + i] = rng.s0;
      v.rng[1 * v.ld_rng + i] = rng.s1;
      v.rng[2 * v.ld_rng + i] = rng.s2;
      v.rng[3 * v.ld_rng + i] = rng.s3;
    }
  }
  if (blockIdx.x == 0 && tid == 0) {
    v.ll[0] = stopped ? neg_inf : ll;
    v.flags[1] = stopped ? stop_row : 0;
    v.flags[0] = stopped ? 1 : 0;
  }
}

}  // namespace mcb
