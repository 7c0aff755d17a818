// C ABI of the B200 particle-filter engine (include/mcstate_b200.h).
//
// One mcb_model = one dust model object: particle state, one
// RNG stream per particle (+ the filter stream), parameters, observations and
// the scratch the filter needs, all resident in HBM for the life of the
// handle.  model$filter() is recorded once per (row range, options) as a CUDA
// graph and replayed: ~4 kernels per data interval, no host round trip until
// the log-likelihood is read.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <set>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/mcstate_b200.h"
#include "kernels.cuh"
#include "persistent.cuh"

namespace {

thread_local std::string g_error;

int fail(int code, const std::string &msg) {
  g_error = msg;
  return code;
}

#define MCB_CUDA(expr)                                                             \
  do {                                                                             \
    cudaError_t err__ = (expr);                                                    \
    if (err__ != cudaSuccess) {                                                    \
      return fail(MCB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(err__)); \
    }                                                                              \
  } while (0)

#define MCB_TRY(expr)              \
  do {                             \
    int rc__ = (expr);             \
    if (rc__ != MCB_OK) return rc__; \
  } while (0)

struct ModelDesc {
  const char *name;
  int n_state, n_par, n_data;
  const char *par_names[mcb::kMaxPar];
  double par_defaults[mcb::kMaxPar];
  const char *state_names[mcb::kMaxState];
  const char *data_names[2];
};

// parameter slots of the closed population (S0, I0, R0) of a compartmental model, or
// none; draws with a parameter-only probability (walk tables)
struct ModelTables {
  int pop[3];
  int n_const;
  int freq;      // parameter slot of the steps-per-incidence-interval count, or -1
};

// ---- which models this build of the library holds ---------------------------------
// The stock library holds the four dust_example models the reference uses.  A build
// with -DMCB_USER_MODEL="<header>" holds ONE model instead, the `mcb::UserModel` that
// header defines by the contract at the top of models.cuh (what dust::dust(<file>) /
// odin.dust::odin_dust do for the reference: a model source compiled into its own
// shared object behind the same generator interface -- tests/testthat/
// test-pmcmc-parallel.R:197, helper-mcstate.R:478-492 with compare_variable.cpp).  The
// header also defines MCB_USER_MODEL_DESC, the ModelDesc initialiser (name, counts,
// parameter names and defaults, state names, data field).
#ifdef MCB_USER_MODEL
}  // namespace
#include MCB_USER_MODEL
namespace {
static_assert(mcb::UserModel::kHasLean || mcb::UserModel::kConst == 0,
              "memo tables (kConst > 0) are read by a lean path only");
static_assert(mcb::UserModel::kConst <= mcb::kMaxConst, "at most kMaxConst fixed-probability draws");
static_assert(mcb::UserModel::kState >= 1 && mcb::UserModel::kState <= mcb::kMaxState &&
              mcb::UserModel::kPar >= 1 && mcb::UserModel::kPar <= mcb::kMaxPar &&
              mcb::UserModel::kData >= 0 && mcb::UserModel::kData <= 1,
              "user model: 1..8 state variables, 1..8 parameters, at most one data field");
#define MCB_MODELS(X) X(0, mcb::UserModel)
const ModelDesc kModels[] = {MCB_USER_MODEL_DESC};
#ifdef MCB_USER_MODEL_TABLES   // a model with a lean path says what its memo tables cover
const ModelTables kModelTables[] = {MCB_USER_MODEL_TABLES};
#else
const ModelTables kModelTables[] = {{{-1, -1, -1}, 0, -1}};   // no memo tables: general form only
#endif
#else
#define MCB_MODELS(X) \
  X(0, mcb::SirModel) X(1, mcb::SirsModel) X(2, mcb::VolatilityModel) X(3, mcb::WalkModel)
// defaults: scripts/generate_vignette_data.R:11-12, tests/testthat/helper-mcstate.R:11-15
const ModelDesc kModels[] = {
    {"sir", 5, 7, 1,
     {"beta", "gamma", "I0", "exp_noise", "S0", "R0", "freq"},
     {0.2, 0.1, 10.0, 1e6, 1000.0, 0.0, 4.0},
     {"S", "I", "R", "cases_cumul", "cases_inc"},
     {"incidence"}},
    {"sirs", 4, 8, 1,
     {"alpha", "beta", "gamma", "I0", "exp_noise", "S0", "R0", "freq"},
     {0.1, 0.2, 0.1, 10.0, 1e6, 1000.0, 0.0, 4.0},
     {"S", "I", "R", "cases_inc"},
     {"incidence"}},
    // tests/testthat/helper-mcstate.R:99
    {"volatility", 1, 5, 1,
     {"alpha", "sigma", "gamma", "tau", "x0"},
     {0.91, 1.0, 1.0, 1.0, 0.0},
     {"x"},
     {"observed"}},
    // no compiled compare (n_data = 0): tests/testthat/test-particle-filter.R:579-586
    {"walk", 1, 1, 0,
     {"sd"},
     {1.0},
     {"x"},
     {nullptr}},
};
const ModelTables kModelTables[] = {{{4, 2, 5}, mcb::SirModel::kConst, 6},
                                    {{5, 3, 6}, mcb::SirsModel::kConst, 7},
                                    {{-1, -1, -1}, 0, -1},
                                    {{-1, -1, -1}, 0, -1}};
#endif
constexpr int kNumModels = sizeof(kModels) / sizeof(kModels[0]);
constexpr size_t kMaxWalkRows = 2048;  // counts covered by a walk table (samplers.cuh)

size_t round_up(size_t x, size_t m) { return (x + m - 1) / m * m; }

// per-device table of jump-matrix powers, built once
struct JumpCache {
  std::mutex mu;
  std::map<int, uint64_t *> dev;
  mcb::JumpTable *host = nullptr;
};
JumpCache g_jump;

// per-device constants (reciprocals of small integers for the inversion sampler)
int device_constants(int device) {
  static std::mutex mu;
  static std::map<int, bool> done;
  std::lock_guard<std::mutex> lock(mu);
  if (done[device]) return MCB_OK;
  double2 inv[mcb::kInvSmall];
  inv[0] = make_double2(0.0, 0.0);
  for (int k = 1; k < mcb::kInvSmall; ++k) inv[k] = make_double2((double)k, 1.0 / (double)k);
  MCB_CUDA(cudaMemcpyToSymbol(mcb::c_inv_small, inv, sizeof inv));
  // the fast walk's tables (samplers.cuh): 1/k and 2^(j/32); approximations are enough
  double recip[mcb::kInvSmall];
  recip[0] = 0.0;
  for (int k = 1; k < mcb::kInvSmall; ++k) recip[k] = 1.0 / (double)k;
  MCB_CUDA(cudaMemcpyToSymbol(mcb::c_recip, recip, sizeof recip));
  double exp2_tab[mcb::kExp2Tab];
  for (int j = 0; j < mcb::kExp2Tab; ++j) exp2_tab[j] = std::exp2((double)j / mcb::kExp2Tab);
  MCB_CUDA(cudaMemcpyToSymbol(mcb::g_exp2_tab, exp2_tab, sizeof exp2_tab));
  done[device] = true;
  return MCB_OK;
}

int jump_table_device(int device, uint64_t **out) {
  std::lock_guard<std::mutex> lock(g_jump.mu);
  auto it = g_jump.dev.find(device);
  if (it != g_jump.dev.end()) { *out = it->second; return MCB_OK; }
  if (!g_jump.host) {
    g_jump.host = new mcb::JumpTable;
    mcb::jump_table_build(*g_jump.host);
  }
  uint64_t *d = nullptr;
  MCB_CUDA(cudaMalloc(&d, sizeof(mcb::JumpTable)));
  MCB_CUDA(cudaMemcpy(d, g_jump.host, sizeof(mcb::JumpTable), cudaMemcpyHostToDevice));
  g_jump.dev[device] = d;
  *out = d;
  return MCB_OK;
}

// What a captured filter graph depends on besides the handle's buffers: the row range and
// options, and -- baked into the kernel arguments at capture -- the model step the call
// starts at and the incidence-reset schedule of every launch (from the parameters' freq).
struct GraphKey {
  int row_first, row_last, hist, n_snap, n_min;
  uint64_t snap_hash, time, sched_hash;
  bool operator<(const GraphKey &o) const {
    return std::tie(row_first, row_last, hist, n_snap, n_min, snap_hash, time, sched_hash) <
           std::tie(o.row_first, o.row_last, o.hist, o.n_snap, o.n_min, o.snap_hash, o.time,
                    o.sched_hash);
  }
};

struct GraphEntry {
  cudaGraphExec_t exec = nullptr;
  uint64_t n_kernels = 0;
};

}  // namespace

struct mcb_model {
  int model_id = 0, scrambler = 0, device = 0, n_const = 0;
  bool chain = false;   // K1/K2 launched with programmatic dependent launch
  // the single-launch filter (persistent.cuh): blocks that can be co-resident, or 0
  int persist_blocks = 0;
  uint64_t *d_data_time = nullptr;
  unsigned long long *d_blk_sum = nullptr;
  const ModelDesc *desc = nullptr;
  mcb::View v{};
  uint64_t time = 0;
  std::vector<double> pars_host;

  // observations
  size_t n_data = 0, data_cursor = 0;
  std::vector<uint64_t> data_time;
  std::vector<char> row_all_na;  // per row: every set NA
  double *d_data = nullptr;
  double *d_u = nullptr;
  uint64_t *d_fs_after = nullptr;
  unsigned long long *d_wmax = nullptr;
  unsigned char *d_na_global = nullptr;

  // state
  double *d_pars = nullptr;
  int *d_index = nullptr;
  int *d_flags = nullptr;
  unsigned int *d_ticket = nullptr;
  double *d_min_ll = nullptr;
  double *d_ll = nullptr;
  int *d_kappa = nullptr;      // scratch for reorder/resample
  double *d_stage = nullptr;   // staging for packed state in/out
  size_t stage_bytes = 0;
  std::vector<int> index_host;

  // history / snapshots of the last filter call
  double *d_hist_value = nullptr;
  int *d_hist_order = nullptr;
  double *d_traj = nullptr;
  int hist_rows = -1, hist_first = 0;  // rows of the history held on the device, or -1
  double *d_snap = nullptr;
  size_t hist_rows_cap = 0, hist_index_cap = 0, snap_cap = 0;

  // pinned result block: ll[n_sets], flags[2]
  double *h_ll = nullptr;
  int *h_flags = nullptr;

  cudaStream_t own_stream = nullptr, stream = nullptr;
  std::map<GraphKey, GraphEntry> graphs;
  std::set<GraphKey> graph_seen;   // keys launched directly once (a graph from the second use)
  uint64_t launches = 0;
  cudaError_t launch_error = cudaSuccess;   // first failed chained launch since the last check

  // optional per-kernel timing (bench.py roofline): 5 events per data row
  bool timing = false;
  std::vector<cudaEvent_t> events;
  double kernel_ms[4] = {0, 0, 0, 0};
  uint64_t kernel_n[4] = {0, 0, 0, 0};

  // pending enqueue
  bool pending = false;
  int pend_first = 0, pend_last = 0;
};

namespace {

using mcb::View;

// a chained launch (launch_chained) that failed since the last check
int check_launches(mcb_model *m) {
  if (m->launch_error == cudaSuccess) return MCB_OK;
  const cudaError_t le = m->launch_error;
  m->launch_error = cudaSuccess;
  return fail(MCB_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(le));
}

void clear_graphs(mcb_model *m) {
  for (auto &g : m->graphs)
    if (g.second.exec) cudaGraphExecDestroy(g.second.exec);
  m->graphs.clear();
  m->graph_seen.clear();
}

int ensure_stage(mcb_model *m, size_t bytes) {
  if (bytes <= m->stage_bytes) return MCB_OK;
  if (m->d_stage) cudaFree(m->d_stage);
  m->d_stage = nullptr;
  m->stage_bytes = 0;
  MCB_CUDA(cudaMalloc(&m->d_stage, bytes));
  m->stage_bytes = bytes;
  return MCB_OK;
}

unsigned blocks_for(size_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

// Bit k of the result: step t0 + k starts a new incidence interval (step % freq == 0).
// Known on the host when every parameter set has the same freq and the launch covers at
// most 32 steps (the usual case); else the kernel works it out per thread.
uint32_t reset_mask(const mcb_model *m, uint64_t t0, uint32_t n_steps, int *valid) {
  *valid = 0;
  const int slot = kModelTables[m->model_id].freq;
  if (slot < 0 || n_steps > 32 || m->pars_host.empty()) return 0;
  const double f0 = m->pars_host[slot];
  for (size_t p = 1; p < m->v.n_sets; ++p)
    if (m->pars_host[p * mcb::kMaxPar + slot] != f0) return 0;
  if (!(f0 >= 1.0) || f0 > 4294967295.0 || f0 != std::floor(f0)) return 0;
  const uint64_t freq = (uint64_t)f0;
  uint32_t mask = 0;
  for (uint32_t k = 0; k < n_steps; ++k)
    if ((t0 + k) % freq == 0) mask |= 1u << k;
  *valid = 1;
  return mask;
}

// Launch with (chained = true) programmatic dependent launch: the kernel may be
// scheduled while its predecessor in the stream drains -- its blocks do their
// prologue and then wait (griddepcontrol.wait) until the predecessor has completed
// and its writes are visible.  Inside a captured graph this becomes a programmatic
// edge.  K1 and K2 of the filter loop are chained this way; both kernels wait
// before they touch anything another kernel wrote.
template <class... P, class... A>
void launch_chained(mcb_model *m, bool chained, void (*kernel)(P...), unsigned grid, int threads,
                    A... args) {
  if (m->launch_error != cudaSuccess) return;   // reported by the caller (check_launches)
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3((unsigned)threads);
  cfg.stream = m->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = chained ? 1 : 0;
  const cudaError_t err = cudaLaunchKernelEx(&cfg, kernel, static_cast<P>(args)...);
  if (err != cudaSuccess) m->launch_error = err;
  ++m->launches;
}

template <class M, int SCR>
void launch_run_compare(mcb_model *m, const double *y_in, double *y_out, uint64_t t0,
                        uint32_t n_steps, const int *kappa, int row, double *hist, int *order) {
  const unsigned grid = blocks_for(m->v.n_total, mcb::kRunThreads);
  // the run kernel uses no shared memory: ask for all of the L1 / shared-memory array as
  // cache, for the walk tables (0 and 20 % measure the same, profiles/ab_r02_v18.txt;
  // MCB_CARVEOUT overrides).  The attribute is per device: set once for every device this
  // instantiation runs on.
  static const int carve = getenv("MCB_CARVEOUT") ? atoi(getenv("MCB_CARVEOUT")) : 0;
  static std::mutex mu;
  static std::map<int, bool> carved;
  {
    std::lock_guard<std::mutex> lock(mu);
    if (!carved[m->device]) {
      cudaFuncSetAttribute(mcb::k_run_compare<M, SCR, false, false>,
                           cudaFuncAttributePreferredSharedMemoryCarveout, carve);
      cudaFuncSetAttribute(mcb::k_run_compare<M, SCR, true, false>,
                           cudaFuncAttributePreferredSharedMemoryCarveout, carve);
      cudaFuncSetAttribute(mcb::k_run_compare<M, SCR, false, true>,
                           cudaFuncAttributePreferredSharedMemoryCarveout, carve);
      cudaFuncSetAttribute(mcb::k_run_compare<M, SCR, true, true>,
                           cudaFuncAttributePreferredSharedMemoryCarveout, carve);
      carved[m->device] = true;
    }
  }
  m->v.reset_mask = reset_mask(m, t0, n_steps, &m->v.reset_mask_valid);
  const bool single = m->v.n_sets == 1;
  if (hist && single)
    launch_chained(m, m->chain, mcb::k_run_compare<M, SCR, true, true>, grid, mcb::kRunThreads,
                   m->v, y_in, y_out, t0, n_steps, kappa, row, hist, order);
  else if (hist)
    launch_chained(m, m->chain, mcb::k_run_compare<M, SCR, true, false>, grid, mcb::kRunThreads,
                   m->v, y_in, y_out, t0, n_steps, kappa, row, hist, order);
  else if (single)
    launch_chained(m, m->chain, mcb::k_run_compare<M, SCR, false, true>, grid, mcb::kRunThreads,
                   m->v, y_in, y_out, t0, n_steps, kappa, row, (double *)nullptr, (int *)nullptr);
  else
    launch_chained(m, m->chain, mcb::k_run_compare<M, SCR, false, false>, grid, mcb::kRunThreads,
                   m->v, y_in, y_out, t0, n_steps, kappa, row, (double *)nullptr, (int *)nullptr);
}

// K1: [gather by kappa on load] -> n_steps updates -> compare against `row`
void run_compare(mcb_model *m, const double *y_in, double *y_out, uint64_t t0, uint32_t n_steps,
                 const int *kappa, int row, double *hist, int *order = nullptr) {
  const int key = m->model_id * 2 + m->scrambler;
  switch (key) {
#define X(ID, M)                                                                                      \
    case 2 * ID: launch_run_compare<M, 0>(m, y_in, y_out, t0, n_steps, kappa, row, hist, order); break; \
    case 2 * ID + 1: launch_run_compare<M, 1>(m, y_in, y_out, t0, n_steps, kappa, row, hist, order); break;
    MCB_MODELS(X)
#undef X
    default: break;
  }
}

// K2: weights -> fixed point -> tile scan and tile sums; empties the ancestor slots
void weights_scan(mcb_model *m, int row) {
  const unsigned grid = (unsigned)(m->v.n_sets * m->v.tiles_per_set);
  if (row < 0) launch_chained(m, m->chain, mcb::k_weights_scan<true>, grid, mcb::kTileThreads, m->v, row);
  else launch_chained(m, m->chain, mcb::k_weights_scan<false>, grid, mcb::kTileThreads, m->v, row);
}

// K3: per set the total weight -> ll, thresholds, early-exit rule; every particle with
// offspring into the ancestor slots (v.heads), read back by mcb::ancestor_of
void ancestors(mcb_model *m, int row) {
  const unsigned grid = (unsigned)(m->v.n_sets * m->v.anc_blocks_per_set);
  launch_chained(m, m->chain, mcb::k_ancestors, grid, mcb::kTileThreads, m->v, row);
}

// one instantiation per state count a model may have (models.cuh: kMaxState)
#define MCB_STATE_COUNTS(X) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8)
static_assert(mcb::kMaxState == 8, "MCB_STATE_COUNTS covers 1..kMaxState");

void resample_gather(mcb_model *m, const int *kappa, const double *src, double *dst) {
  const unsigned grid = blocks_for(m->v.n_total, mcb::kGatherThreads);
  switch (m->v.n_state) {
#define X(N) case N: mcb::k_resample_gather<N><<<grid, mcb::kGatherThreads, 0, m->stream>>>(m->v, kappa, src, dst); break;
    MCB_STATE_COUNTS(X)
#undef X
    default: break;
  }
  ++m->launches;
}

void finish_state(mcb_model *m, int last, const int *kappa, int *order) {
  const unsigned grid = blocks_for(m->v.n_total, mcb::kGatherThreads);
  switch (m->v.n_state) {
#define X(N) case N: mcb::k_finish_state<N><<<grid, mcb::kGatherThreads, 0, m->stream>>>(m->v, last, kappa, order); break;
    MCB_STATE_COUNTS(X)
#undef X
    default: break;
  }
  ++m->launches;
}

void prepare_sets(mcb_model *m) {
  const int n_const = kModelTables[m->model_id].n_const;
  const size_t rows = n_const ? (size_t)m->v.n_tab + (size_t)n_const * m->v.n_wrows : 1;
  const dim3 grid(blocks_for(rows, 128), (unsigned)m->v.n_sets);
  switch (m->model_id) {
#define X(ID, M) case ID: mcb::k_prepare_sets<M><<<grid, 128, 0, m->stream>>>(m->v); break;
    MCB_MODELS(X)
#undef X
    default: break;
  }
  ++m->launches;
}

void set_initial(mcb_model *m) {
  const unsigned grid = blocks_for(m->v.n_total, 256);
  switch (m->model_id) {
#define X(ID, M) case ID: mcb::k_set_initial<M><<<grid, 256, 0, m->stream>>>(m->v); break;
    MCB_MODELS(X)
#undef X
    default: break;
  }
  ++m->launches;
}

void draw_uniforms(mcb_model *m, int row_first, int row_last) {
  if (m->v.n_fs == 1) {   // one stream shared by every set: a warp, look-ups batched
    if (m->scrambler == 0)
      mcb::k_draw_uniforms_shared<0><<<1, 32, 0, m->stream>>>(m->v, row_first, row_last, m->d_u, m->d_fs_after);
    else
      mcb::k_draw_uniforms_shared<1><<<1, 32, 0, m->stream>>>(m->v, row_first, row_last, m->d_u, m->d_fs_after);
    ++m->launches;
    return;
  }
  const unsigned grid = blocks_for((size_t)m->v.n_fs, 128);
  if (m->scrambler == 0)
    mcb::k_draw_uniforms<0><<<grid, 128, 0, m->stream>>>(m->v, row_first, row_last, m->d_u, m->d_fs_after);
  else
    mcb::k_draw_uniforms<1><<<grid, 128, 0, m->stream>>>(m->v, row_first, row_last, m->d_u, m->d_fs_after);
  ++m->launches;
}

// boundary k of a data row: 0 before K1, 1 after K1, 2 after K2, 3 after K3; row n_data
// holds the two boundaries around k_finish_state
constexpr size_t kMarks = 4;
void mark(mcb_model *m, int row, int k) {
  if (!m->timing) return;
  const size_t need = (size_t)(row + 1) * kMarks;
  while (m->events.size() < need) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    m->events.push_back(e);
  }
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(m->stream, &st);
  cudaEventRecordWithFlags(m->events[(size_t)row * kMarks + k], m->stream,
                           st == cudaStreamCaptureStatusActive ? cudaEventRecordExternal
                                                               : cudaEventRecordDefault);
}

void reset_cursor(mcb_model *m) {
  m->data_cursor = 0;
  while (m->data_cursor < m->n_data && m->data_time[m->data_cursor] <= m->time) ++m->data_cursor;
}

int pack_to_host(mcb_model *m, const double *y, const int *d_index, int n_rows, double *out) {
  const size_t bytes = m->v.n_total * (size_t)n_rows * sizeof(double);
  MCB_TRY(ensure_stage(m, bytes));
  mcb::k_pack_rows<<<blocks_for(m->v.n_total, 256), 256, 0, m->stream>>>(
      y, m->v.ld, m->v.n_total, d_index, n_rows, m->d_stage);
  ++m->launches;
  MCB_CUDA(cudaMemcpyAsync(out, m->d_stage, bytes, cudaMemcpyDeviceToHost, m->stream));
  MCB_CUDA(cudaStreamSynchronize(m->stream));
  return MCB_OK;
}

// Enqueue rows [first, last) of the filter on m->stream (also used under capture).
// Returns the number of kernels enqueued.
uint64_t enqueue_filter_rows(mcb_model *m, int first, int last, bool hist,
                             const std::vector<int> &snap_rows) {
  const uint64_t before = m->launches;
  View &v = m->v;
  const size_t n_sets = v.n_sets;
  cudaMemsetAsync(m->d_flags, 0, 2 * sizeof(int), m->stream);
  cudaMemsetAsync(m->d_ll, 0, n_sets * sizeof(double), m->stream);
  cudaMemsetAsync(m->d_wmax + (size_t)first * n_sets, 0,
                  (size_t)(last - first) * n_sets * sizeof(unsigned long long), m->stream);
  draw_uniforms(m, first, last);
  auto iota = [&](int *dst) {
    mcb::k_iota<<<blocks_for(v.n_total, 256), 256, 0, m->stream>>>(dst, v.n_total);
    ++m->launches;
  };
  if (hist) {
    // slice 0: the state before the first interval, identity order
    run_compare(m, v.y, v.y, 0, 0, nullptr, -1, m->d_hist_value);
    iota(m->d_hist_order);
  }
  uint64_t t = m->time;
  size_t snap_i = 0;
  // the ancestor slots K3 filled for the last weighed row (its resampling is applied by
  // the next launch that loads the state), or nullptr.  With trajectories that launch
  // also records the ancestors in the order slice of the history.
  const int *kappa_prev = nullptr;
  const double *in = v.y;
  // the loop's K1, K2 and K3 are chained by programmatic dependent launch (not when
  // events sit between them for per-kernel timing); MCB_PDL=0 turns it off
  static const bool pdl = !(getenv("MCB_PDL") && atoi(getenv("MCB_PDL")) == 0);
  m->chain = pdl && !m->timing;
  for (int row = first; row < last; ++row) {
    const uint64_t t_end = m->data_time[row];
    const uint32_t n_steps = (uint32_t)(t_end - t);
    const int rel = row - first;
    // buffers alternate so that the last row always writes y_tmp
    double *out = ((last - 1 - row) & 1) == 0 ? v.y_tmp : v.y;
    double *hist_slice = hist ? m->d_hist_value + (size_t)(rel + 1) * v.n_index * v.ld : nullptr;
    const int weigh = m->row_all_na[row] ? -1 : row;
    mark(m, row, 0);
    run_compare(m, in, out, t, n_steps, kappa_prev, weigh, hist_slice,
                hist && kappa_prev ? m->d_hist_order + (size_t)rel * v.n_total : nullptr);
    if (hist && rel > 0 && !kappa_prev) iota(m->d_hist_order + (size_t)rel * v.n_total);
    mark(m, row, 1);
    kappa_prev = nullptr;
    if (weigh >= 0) {
      weights_scan(m, row);
      mark(m, row, 2);
      ancestors(m, row);
      kappa_prev = v.heads;
    } else {
      mark(m, row, 2);
    }
    mark(m, row, 3);
    while (snap_i < snap_rows.size() && snap_rows[snap_i] < row) ++snap_i;
    if (snap_i < snap_rows.size() && snap_rows[snap_i] == row) {
      double *dst = m->d_snap + snap_i * (size_t)v.n_state * v.ld;
      if (weigh >= 0) {
        resample_gather(m, kappa_prev, out, dst);
      } else {
        mcb::k_copy_state_if_running<<<blocks_for(v.n_total, 256), 256, 0, m->stream>>>(v, out, dst);
        ++m->launches;
      }
      ++snap_i;
    }
    in = out;
    t = t_end;
  }
  m->chain = false;
  mark(m, (int)m->n_data, 0);
  finish_state(m, last, kappa_prev,
               hist && kappa_prev ? m->d_hist_order + (size_t)(last - first) * v.n_total : nullptr);
  if (hist && !kappa_prev) iota(m->d_hist_order + (size_t)(last - first) * v.n_total);
  mark(m, (int)m->n_data, 1);
  mcb::k_filter_finish<<<blocks_for((size_t)v.n_fs, 128), 128, 0, m->stream>>>(v, first, last,
                                                                                m->d_fs_after);
  ++m->launches;
  return m->launches - before;
}

int filter_rows_range(const mcb_model *m, uint64_t time_end, int *first, int *last) {
  size_t a = m->data_cursor, b = a;
  while (b < m->n_data && m->data_time[b] <= time_end) ++b;
  *first = (int)a;
  *last = (int)b;
  return MCB_OK;
}

int prepare_history(mcb_model *m, int rows) {
  View &v = m->v;
  if ((size_t)rows + 1 > m->hist_rows_cap || (size_t)v.n_index > m->hist_index_cap) {
    if (m->d_hist_value) cudaFree(m->d_hist_value);
    if (m->d_hist_order) cudaFree(m->d_hist_order);
    if (m->d_traj) cudaFree(m->d_traj);
    m->d_hist_value = nullptr; m->d_hist_order = nullptr; m->d_traj = nullptr;
    m->hist_rows_cap = 0;
    clear_graphs(m);  // baked pointers are stale
    const size_t T1 = (size_t)rows + 1;
    MCB_CUDA(cudaMalloc(&m->d_hist_value, T1 * v.n_index * v.ld * sizeof(double)));
    MCB_CUDA(cudaMalloc(&m->d_hist_order, T1 * v.n_total * sizeof(int)));
    m->hist_rows_cap = T1;
    m->hist_index_cap = v.n_index;
  }
  return MCB_OK;
}

// ---- the single-launch filter (persistent.cuh) -----------------------------------
template <class M, int SCR>
const void *persistent_kernel() {
  return reinterpret_cast<const void *>(&mcb::k_filter_persistent<M, SCR>);
}
const void *persistent_kernel_for(const mcb_model *m) {
  switch (m->model_id * 2 + m->scrambler) {
#define X(ID, M)                                          \
    case 2 * ID: return persistent_kernel<M, 0>();        \
    case 2 * ID + 1: return persistent_kernel<M, 1>();
    MCB_MODELS(X)
#undef X
    default: return nullptr;
  }
}

// how many blocks of the persistent kernel fit the device at once (0: not usable)
int persistent_capacity(const mcb_model *m) {
  static const bool on = !(getenv("MCB_PERSISTENT") && atoi(getenv("MCB_PERSISTENT")) == 0);
  if (!on) return 0;
  int coop = 0, sms = 0, per_sm = 0;
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, m->device);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, m->device);
  if (!coop) return 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, persistent_kernel_for(m),
                                                    mcb::kPersistThreads, 0) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  const int cap = per_sm * sms;
  return cap < mcb::kPersistUseBlocks ? cap : mcb::kPersistUseBlocks;
}

// One parameter set, no snapshots, every particle resident at once: the whole call is
// one cooperative launch between the uniforms and the stream update.
bool use_persistent(const mcb_model *m, size_t n_snap) {
  const View &v = m->v;
  return m->persist_blocks > 0 && n_snap == 0 && !m->timing && v.n_sets == 1 &&
         v.n_fs == 1 && blocks_for(v.n_total, mcb::kPersistThreads) <= (unsigned)m->persist_blocks;
}

int launch_persistent(mcb_model *m, int first, int last, bool hist) {
  View &v = m->v;
  cudaMemsetAsync(m->d_flags, 0, 2 * sizeof(int), m->stream);
  cudaMemsetAsync(m->d_ll, 0, sizeof(double), m->stream);
  cudaMemsetAsync(m->d_wmax + (size_t)first, 0, (size_t)(last - first) * sizeof(unsigned long long),
                  m->stream);
  draw_uniforms(m, first, last);
  int weighed = 0;
  for (int row = first; row < last; ++row) weighed += m->row_all_na[row] ? 0 : 1;
  mcb::PersistArgs a;
  a.data_time = m->d_data_time;
  a.blk_sum = m->d_blk_sum;
  a.first = first;
  a.last = last;
  a.t_start = m->time;
  a.first_out_tmp = (weighed & 1) ? 1 : 0;   // so that the last weighed row writes y_tmp
  a.hist_value = hist ? m->d_hist_value : nullptr;
  a.hist_order = hist ? m->d_hist_order : nullptr;
  v.reset_mask_valid = 0;   // one launch covers rows with different start steps
  void *args[] = {&v, &a};
  const unsigned grid = blocks_for(v.n_total, mcb::kPersistThreads);
  MCB_CUDA(cudaLaunchCooperativeKernel(persistent_kernel_for(m), dim3(grid),
                                       dim3(mcb::kPersistThreads), args, 0, m->stream));
  ++m->launches;
  mcb::k_filter_finish<<<blocks_for((size_t)v.n_fs, 128), 128, 0, m->stream>>>(v, first, last,
                                                                                m->d_fs_after);
  ++m->launches;
  return MCB_OK;
}

int run_filter(mcb_model *m, uint64_t time_end, bool hist, const uint64_t *snapshot_times,
               size_t n_snap, const double *min_ll, size_t n_min, bool wait,
               double *ll_out, double *traj_out, double *snap_out) {
  if (m->pending) return fail(MCB_ERR_STATE, "a filter is already enqueued; collect it first");
  if (m->n_data == 0) return fail(MCB_ERR_STATE, "filter: no data set (call set_data first)");
  View &v = m->v;
  if (n_min != 0 && n_min != 1 && n_min != v.n_sets)
    return fail(MCB_ERR_ARG, "min_log_likelihood must have length 0, 1 or n_pars");
  int first, last;
  filter_rows_range(m, time_end, &first, &last);
  const int rows = last - first;

  std::vector<int> snap_rows;
  uint64_t snap_hash = 1469598103934665603ULL;
  for (size_t k = 0; k < n_snap; ++k) {
    int found = -1;
    for (int r = first; r < last; ++r)
      if (m->data_time[r] == snapshot_times[k]) found = r;
    if (found < 0) return fail(MCB_ERR_ARG, "snapshot time is not a data time of this call");
    if (!snap_rows.empty() && found <= snap_rows.back())
      return fail(MCB_ERR_ARG, "snapshot times must be strictly increasing");
    snap_rows.push_back(found);
    snap_hash = (snap_hash ^ (uint64_t)found) * 1099511628211ULL;
  }
  if (n_snap > m->snap_cap) {
    if (m->d_snap) cudaFree(m->d_snap);
    m->d_snap = nullptr;
    clear_graphs(m);
    MCB_CUDA(cudaMalloc(&m->d_snap, n_snap * (size_t)v.n_state * v.ld * sizeof(double)));
    m->snap_cap = n_snap;
  }
  if (hist) MCB_TRY(prepare_history(m, rows));
  if (n_snap)
    MCB_CUDA(cudaMemsetAsync(m->d_snap, 0xff, n_snap * (size_t)v.n_state * v.ld * sizeof(double),
                             m->stream));
  v.n_min = (int)n_min;
  if (n_min)
    MCB_CUDA(cudaMemcpyAsync(m->d_min_ll, min_ll, n_min * sizeof(double), cudaMemcpyHostToDevice,
                             m->stream));

  if (rows > 0 && use_persistent(m, n_snap)) {
    MCB_TRY(launch_persistent(m, first, last, hist));
  } else if (rows > 0) {
    // the reset schedule of every launch, as enqueue_filter_rows will bake it in
    uint64_t sched_hash = 1469598103934665603ULL;
    {
      uint64_t t = m->time;
      for (int r = first; r < last; ++r) {
        int valid = 0;
        const uint32_t mask = reset_mask(m, t, (uint32_t)(m->data_time[r] - t), &valid);
        sched_hash = (sched_hash ^ (((uint64_t)mask << 1) | (uint64_t)valid)) * 1099511628211ULL;
        t = m->data_time[r];
      }
    }
    const GraphKey key{first, last, (hist ? 1 : 0) | (m->timing ? 2 : 0), (int)n_snap, (int)n_min,
                       snap_hash, m->time, sched_hash};
    auto it = m->graphs.find(key);
    // A graph pays for itself from the second use on: the first call with a key launches
    // the kernels directly (SMC^2's proposals run a row range once -- capturing and
    // instantiating a graph of hundreds of nodes for it costs more than it saves).
    if (it == m->graphs.end() && !m->graph_seen.count(key) && !m->timing) {
      m->graph_seen.insert(key);
      enqueue_filter_rows(m, first, last, hist, snap_rows);
      MCB_TRY(check_launches(m));
    } else if (it == m->graphs.end()) {
      cudaGraph_t graph = nullptr;
      MCB_CUDA(cudaStreamBeginCapture(m->stream, cudaStreamCaptureModeThreadLocal));
      const uint64_t n_k = enqueue_filter_rows(m, first, last, hist, snap_rows);
      m->launches -= n_k;  // counted per replay below
      cudaError_t err = cudaStreamEndCapture(m->stream, &graph);
      if (m->launch_error != cudaSuccess) {
        const cudaError_t le = m->launch_error;
        m->launch_error = cudaSuccess;
        if (graph) cudaGraphDestroy(graph);
        return fail(MCB_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(le));
      }
      if (err != cudaSuccess)
        return fail(MCB_ERR_CUDA, std::string("graph capture: ") + cudaGetErrorString(err));
      GraphEntry e;
      err = cudaGraphInstantiate(&e.exec, graph, 0);
      cudaGraphDestroy(graph);
      if (err != cudaSuccess)
        return fail(MCB_ERR_CUDA, std::string("graph instantiate: ") + cudaGetErrorString(err));
      e.n_kernels = n_k;
      it = m->graphs.emplace(key, e).first;
    }
    if (it != m->graphs.end()) {
      MCB_CUDA(cudaGraphLaunch(it->second.exec, m->stream));
      m->launches += it->second.n_kernels;
    }
  } else {
    MCB_CUDA(cudaMemsetAsync(m->d_flags, 0, 2 * sizeof(int), m->stream));
    MCB_CUDA(cudaMemsetAsync(m->d_ll, 0, v.n_sets * sizeof(double), m->stream));
  }
  MCB_CUDA(cudaMemcpyAsync(m->h_ll, m->d_ll, v.n_sets * sizeof(double), cudaMemcpyDeviceToHost,
                           m->stream));
  MCB_CUDA(cudaMemcpyAsync(m->h_flags, m->d_flags, 2 * sizeof(int), cudaMemcpyDeviceToHost,
                           m->stream));
  m->hist_rows = hist ? rows : -1;   // what mcb_filter_history may trace afterwards
  m->hist_first = first;
  m->pending = true;
  m->pend_first = first;
  m->pend_last = last;
  if (!wait) return MCB_OK;

  if (hist && traj_out) {
    if (!m->d_traj)
      MCB_CUDA(cudaMalloc(&m->d_traj, m->hist_rows_cap * v.n_total * m->hist_index_cap * sizeof(double)));
    mcb::k_trace_history<<<blocks_for(v.n_total, 128), 128, 0, m->stream>>>(
        v, m->d_hist_value, m->d_hist_order, rows, first, m->d_traj);
    ++m->launches;
    MCB_CUDA(cudaMemcpyAsync(traj_out, m->d_traj,
                             (size_t)(rows + 1) * v.n_total * v.n_index * sizeof(double),
                             cudaMemcpyDeviceToHost, m->stream));
  }
  if (n_snap && snap_out) {
    for (size_t k = 0; k < n_snap; ++k) {
      MCB_TRY(pack_to_host(m, m->d_snap + k * (size_t)v.n_state * v.ld, nullptr, v.n_state,
                           snap_out + k * v.n_total * (size_t)v.n_state));
    }
  }
  return mcb_filter_collect(m, ll_out);
}

}  // namespace

// ===========================================================================
extern "C" {

int mcb_version(void) { return 100; }
const char *mcb_last_error(void) { return g_error.c_str(); }

int mcb_device_count(int *n_devices) {
  if (!n_devices) return fail(MCB_ERR_ARG, "n_devices is NULL");
  int n = 0;
  cudaError_t err = cudaGetDeviceCount(&n);
  if (err != cudaSuccess) {
    *n_devices = 0;
    return fail(MCB_ERR_CUDA, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(err));
  }
  *n_devices = n;
  return MCB_OK;
}

int mcb_device_info(int device_id, char *name, size_t name_len, int *cc_major, int *cc_minor,
                    int *n_sm, size_t *mem_bytes) {
  cudaDeviceProp p;
  MCB_CUDA(cudaGetDeviceProperties(&p, device_id));
  if (name && name_len) { std::strncpy(name, p.name, name_len - 1); name[name_len - 1] = 0; }
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  if (n_sm) *n_sm = p.multiProcessorCount;
  if (mem_bytes) *mem_bytes = p.totalGlobalMem;
  return MCB_OK;
}

int mcb_model_lookup(const char *name, int *model_id, size_t *n_state, size_t *n_par,
                     size_t *n_data_fields) {
  if (!name) return fail(MCB_ERR_ARG, "model name is NULL");
  for (int i = 0; i < kNumModels; ++i)
    if (std::strcmp(name, kModels[i].name) == 0) {
      if (model_id) *model_id = i;
      if (n_state) *n_state = kModels[i].n_state;
      if (n_par) *n_par = kModels[i].n_par;
      if (n_data_fields) *n_data_fields = kModels[i].n_data;
      return MCB_OK;
    }
  return fail(MCB_ERR_ARG, std::string("unknown model '") + name + "'");
}

const char *mcb_model_par_name(int model_id, size_t i) {
  if (model_id < 0 || model_id >= kNumModels || i >= (size_t)kModels[model_id].n_par) return nullptr;
  return kModels[model_id].par_names[i];
}
double mcb_model_par_default(int model_id, size_t i) {
  if (model_id < 0 || model_id >= kNumModels || i >= (size_t)kModels[model_id].n_par) return NAN;
  return kModels[model_id].par_defaults[i];
}
const char *mcb_model_data_name(int model_id, size_t i) {
  if (model_id < 0 || model_id >= kNumModels || i >= (size_t)kModels[model_id].n_data) return nullptr;
  return kModels[model_id].data_names[i];
}
const char *mcb_model_state_name(int model_id, size_t i) {
  if (model_id < 0 || model_id >= kNumModels || i >= (size_t)kModels[model_id].n_state) return nullptr;
  return kModels[model_id].state_names[i];
}

// ---- RNG plumbing on the host ---------------------------------------------
int mcb_rng_seed(uint64_t seed, uint64_t state[4]) {
  if (!state) return fail(MCB_ERR_ARG, "state is NULL");
  mcb::xo_seed(seed, state);
  return MCB_OK;
}
int mcb_rng_jump(uint64_t state[4]) {
  if (!state) return fail(MCB_ERR_ARG, "state is NULL");
  mcb::xo_jump(state);
  return MCB_OK;
}
int mcb_rng_long_jump(uint64_t state[4]) {
  if (!state) return fail(MCB_ERR_ARG, "state is NULL");
  mcb::xo_long_jump(state);
  return MCB_OK;
}
int mcb_rng_random_real(uint64_t state[4], int scrambler, size_t n, double *out) {
  if (!state || !out) return fail(MCB_ERR_ARG, "NULL argument");
  mcb::Xoshiro r{state[0], state[1], state[2], state[3]};
  for (size_t i = 0; i < n; ++i)
    out[i] = scrambler == MCB_SCRAMBLER_PLUS ? mcb::xo_real<0>(r) : mcb::xo_real<1>(r);
  state[0] = r.s0; state[1] = r.s1; state[2] = r.s2; state[3] = r.s3;
  return MCB_OK;
}
int mcb_rng_distributed_state(const uint64_t seed_state[4], size_t n_streams, size_t n_nodes,
                              uint64_t *out) {
  if (!seed_state || !out) return fail(MCB_ERR_ARG, "NULL argument");
  uint64_t base[4] = {seed_state[0], seed_state[1], seed_state[2], seed_state[3]};
  for (size_t k = 0; k < n_nodes; ++k) {
    uint64_t cur[4] = {base[0], base[1], base[2], base[3]};
    for (size_t s = 0; s < n_streams; ++s) {
      std::memcpy(out + (k * n_streams + s) * 4, cur, sizeof cur);
      mcb::xo_jump(cur);
    }
    mcb::xo_long_jump(base);
  }
  return MCB_OK;
}

// ---- model object ------------------------------------------------------------
int mcb_alloc(int model_id, const double *pars, size_t n_particles, size_t n_pars, uint64_t time,
              const uint64_t *seed_state, size_t n_seed_streams, int scrambler,
              int deterministic, int device_id, mcb_model **out) {
  return mcb_alloc_ex(model_id, pars, n_particles, n_pars, time, seed_state, n_seed_streams,
                      MCB_SEED_SHARED, scrambler, deterministic, device_id, out);
}

int mcb_alloc_ex(int model_id, const double *pars, size_t n_particles, size_t n_pars,
                 uint64_t time, const uint64_t *seed_state, size_t n_seed_streams, int seed_mode,
                 int scrambler, int deterministic, int device_id, mcb_model **out) {
  if (!out) return fail(MCB_ERR_ARG, "out is NULL");
  *out = nullptr;
  if (model_id < 0 || model_id >= kNumModels) return fail(MCB_ERR_ARG, "unknown model id");
  if (n_particles == 0) return fail(MCB_ERR_ARG, "'n_particles' must be at least 1");
  if (!pars) return fail(MCB_ERR_ARG, "pars is NULL");
  if (!seed_state || n_seed_streams == 0) return fail(MCB_ERR_ARG, "seed state is empty");
  if (scrambler != MCB_SCRAMBLER_PLUS && scrambler != MCB_SCRAMBLER_STARSTAR)
    return fail(MCB_ERR_ARG, "unknown scrambler");
  if (seed_mode != MCB_SEED_SHARED && seed_mode != MCB_SEED_PER_SET)
    return fail(MCB_ERR_ARG, "unknown seed mode");
  if (seed_mode == MCB_SEED_PER_SET && n_seed_streams != (n_pars == 0 ? 1 : n_pars))
    return fail(MCB_ERR_ARG, "per-set seeding needs exactly one seed stream per parameter set");
  const size_t n_sets = n_pars == 0 ? 1 : n_pars;
  const size_t n_total = n_particles * n_sets;
  if (n_total >= ((size_t)1 << 31)) return fail(MCB_ERR_ARG, "too many particles (int32 kappa)");
  int n_dev = 0;
  MCB_TRY(mcb_device_count(&n_dev));
  if (n_dev == 0) return fail(MCB_ERR_CUDA, "no CUDA device: this engine has no CPU fallback");
  if (device_id < 0 || device_id >= n_dev) return fail(MCB_ERR_ARG, "device_id out of range");
  MCB_CUDA(cudaSetDevice(device_id));
  MCB_TRY(device_constants(device_id));

  mcb_model *m = new mcb_model;
  m->model_id = model_id;
  m->desc = &kModels[model_id];
  m->scrambler = scrambler;
  m->device = device_id;
  m->time = time;
  View &v = m->v;
  v.n_particles = n_particles;
  v.n_sets = n_sets;
  v.n_total = n_total;
  v.ld = round_up(n_total, 32);
  v.n_fs = seed_mode == MCB_SEED_PER_SET ? (int)n_sets : 1;
  v.n_sets_global = (int)n_sets;
  v.set_offset = 0;
  v.na_global = nullptr;
  v.ld_rng = round_up(n_total + v.n_fs, 32);
  v.n_state = m->desc->n_state;
  v.n_index = v.n_state;
  v.n_min = 0;
  v.data_sets = 1;
  v.tiles_per_set = (int)((n_particles + mcb::kTile - 1) / mcb::kTile);
  v.anc_blocks_per_set = (int)((n_particles + mcb::kAncTile - 1) / mcb::kAncTile);
  v.deterministic = deterministic;
  v.independent = (seed_mode == MCB_SEED_PER_SET && n_sets > 1) ? 1 : 0;

  auto cleanup = [&](int rc) { mcb_free(m); return rc; };
#define ALLOC(ptr, bytes)                                                         \
  do {                                                                            \
    cudaError_t e__ = cudaMalloc(&(ptr), (bytes));                                \
    if (e__ != cudaSuccess)                                                       \
      return cleanup(fail(MCB_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e__))); \
  } while (0)
  ALLOC(v.y, (size_t)v.n_state * v.ld * sizeof(double));
  ALLOC(v.y_tmp, (size_t)v.n_state * v.ld * sizeof(double));
  ALLOC(v.rng, 4 * v.ld_rng * sizeof(uint64_t));
  ALLOC(m->d_pars, n_sets * mcb::kMaxPar * sizeof(double));
  {
    // memo tables cover compartment counts up to the largest initial population,
    // within 64 Ki rows and 1 GB in total; larger counts take the direct path.  The
    // walk tables stop at kMaxWalkRows counts (beyond: the direct path).
    double max_pop = 0.0;
    const int np = m->desc->n_par;
    for (size_t p = 0; p < n_sets; ++p) {
      const double *pp = pars + p * np;
      const int *ix = kModelTables[model_id].pop;
      const double pop = ix[0] < 0 ? 0.0 : pp[ix[0]] + pp[ix[1]] + pp[ix[2]];
      if (pop > max_pop) max_pop = pop;
    }
    size_t n_tab = max_pop < 65535.0 ? (size_t)max_pop + 1 : 65536;
    n_tab = round_up(n_tab < 256 ? 256 : n_tab, 256);
    const size_t n_const = (size_t)kModelTables[model_id].n_const;
    const size_t row_bytes = (4 + n_const * mcb::kWalkK) * sizeof(double);
    const size_t budget = ((size_t)1 << 30) / (n_sets * row_bytes);
    if (n_tab > budget) n_tab = budget / 256 * 256;
    if (n_tab < 256) n_tab = 256;
    v.n_tab = (int)n_tab;
    v.n_wrows = (int)(n_tab < kMaxWalkRows ? n_tab : kMaxWalkRows);
    m->n_const = (int)n_const;
  }
  ALLOC(v.derived, n_sets * mcb::kMaxDerived * sizeof(double));
  ALLOC(v.tab_inf, n_sets * (size_t)v.n_tab * sizeof(double4));
  ALLOC(v.tab_walk, n_sets * (size_t)(m->n_const ? m->n_const : 1) * v.n_wrows * mcb::kWalkK * sizeof(double));
  ALLOC(v.logw, v.ld * sizeof(double));
  ALLOC(v.cw, v.ld * sizeof(unsigned long long));
  ALLOC(v.tile_sum, n_sets * v.tiles_per_set * sizeof(unsigned long long));
  ALLOC(m->d_ll, n_sets * sizeof(double));
  ALLOC(m->d_min_ll, n_sets * sizeof(double));
  ALLOC(m->d_flags, 2 * sizeof(int));
  ALLOC(m->d_ticket, (n_sets + 2) * sizeof(unsigned int));
  ALLOC(m->d_blk_sum, mcb::kPersistMaxBlocks * sizeof(unsigned long long));
  ALLOC(m->d_index, mcb::kMaxState * sizeof(int));
  ALLOC(m->d_kappa, v.ld * sizeof(int));
  ALLOC(m->d_u, n_sets * sizeof(double));
  ALLOC(m->d_wmax, n_sets * sizeof(unsigned long long));
  ALLOC(m->d_fs_after, 4 * (size_t)v.n_fs * sizeof(uint64_t));
#undef ALLOC
  v.pars = m->d_pars;
  v.ll = m->d_ll;
  v.min_ll = m->d_min_ll;
  v.flags = m->d_flags;
  v.ticket = m->d_ticket;
  v.heads = m->d_kappa;
  v.index = m->d_index;
  v.u_draws = m->d_u;
  v.wmax = m->d_wmax;
  if (cudaMallocHost(&m->h_ll, n_sets * sizeof(double)) != cudaSuccess ||
      cudaMallocHost(&m->h_flags, 2 * sizeof(int)) != cudaSuccess ||
      cudaStreamCreateWithFlags(&m->own_stream, cudaStreamNonBlocking) != cudaSuccess)
    return cleanup(fail(MCB_ERR_CUDA, "pinned memory / stream creation failed"));
  m->stream = m->own_stream;
  cudaMemsetAsync(m->d_flags, 0, 2 * sizeof(int), m->stream);
  cudaMemsetAsync(m->d_ticket, 0, (n_sets + 2) * sizeof(unsigned int), m->stream);
  cudaMemsetAsync(v.y, 0, (size_t)v.n_state * v.ld * sizeof(double), m->stream);
  cudaMemsetAsync(v.y_tmp, 0, (size_t)v.n_state * v.ld * sizeof(double), m->stream);
  m->index_host.resize(v.n_state);
  for (int s = 0; s < v.n_state; ++s) m->index_host[s] = s;
  cudaMemcpyAsync(m->d_index, m->index_host.data(), v.n_state * sizeof(int),
                  cudaMemcpyHostToDevice, m->stream);

  // RNG streams: supplied ones verbatim, the rest by parallel jump-ahead
  if (seed_mode == MCB_SEED_PER_SET) {
    uint64_t *d_bases = nullptr, *d_jump = nullptr;
    int rc = jump_table_device(device_id, &d_jump);
    if (rc != MCB_OK) return cleanup(rc);
    if (cudaMalloc(&d_bases, n_sets * 4 * sizeof(uint64_t)) != cudaSuccess)
      return cleanup(fail(MCB_ERR_CUDA, "cudaMalloc: seed bases"));
    cudaMemcpyAsync(d_bases, seed_state, n_sets * 4 * sizeof(uint64_t), cudaMemcpyHostToDevice,
                    m->stream);
    mcb::k_derive_streams_per_set<<<blocks_for((n_particles + 1) * n_sets, 128), 128, 0,
                                    m->stream>>>(v.rng, v.ld_rng, d_bases, n_particles, n_sets,
                                                 d_jump);
    ++m->launches;
    cudaStreamSynchronize(m->stream);
    cudaFree(d_bases);
  } else {
    const size_t n_streams = n_total + 1;
    const size_t have = n_seed_streams < n_streams ? n_seed_streams : n_streams;
    std::vector<uint64_t> soa(4 * have);
    for (size_t i = 0; i < have; ++i)
      for (int w = 0; w < 4; ++w) soa[w * have + i] = seed_state[4 * i + w];
    for (int w = 0; w < 4; ++w)
      cudaMemcpyAsync(v.rng + w * v.ld_rng, soa.data() + w * have, have * sizeof(uint64_t),
                      cudaMemcpyHostToDevice, m->stream);
    cudaStreamSynchronize(m->stream);
    if (have < n_streams) {
      uint64_t *d_jump = nullptr;
      int rc = jump_table_device(device_id, &d_jump);
      if (rc != MCB_OK) return cleanup(rc);
      mcb::k_derive_streams<<<blocks_for(n_streams - have, 128), 128, 0, m->stream>>>(
          v.rng, v.ld_rng, have, n_streams, d_jump);
      ++m->launches;
    }
  }
  int rc = mcb_update_state(m, pars, nullptr, 0, 1, time, 1);
  if (rc != MCB_OK) return cleanup(rc);
  cudaError_t err = cudaStreamSynchronize(m->stream);
  if (err != cudaSuccess)
    return cleanup(fail(MCB_ERR_CUDA, std::string("alloc: ") + cudaGetErrorString(err)));
  m->persist_blocks = persistent_capacity(m);
  *out = m;
  return MCB_OK;
}

int mcb_free(mcb_model *m) {
  if (!m) return MCB_OK;
  cudaSetDevice(m->device);
  if (m->stream) cudaStreamSynchronize(m->stream);
  clear_graphs(m);
  void *ptrs[] = {m->v.y, m->v.y_tmp, m->v.rng, m->v.derived, m->v.tab_inf, m->v.tab_walk, m->d_pars, m->v.logw, m->v.cw, m->v.tile_sum,
                  m->d_ll, m->d_min_ll, m->d_flags, m->d_ticket, m->d_index, m->d_kappa, m->d_u,
                  m->d_wmax, m->d_fs_after, m->d_na_global, m->d_data, m->d_stage, m->d_hist_value,
                  m->d_hist_order, m->d_traj, m->d_snap, m->d_data_time, m->d_blk_sum};
  for (void *p : ptrs)
    if (p) cudaFree(p);
  for (cudaEvent_t e : m->events) cudaEventDestroy(e);
  if (m->h_ll) cudaFreeHost(m->h_ll);
  if (m->h_flags) cudaFreeHost(m->h_flags);
  if (m->own_stream) cudaStreamDestroy(m->own_stream);
  delete m;
  return MCB_OK;
}

int mcb_set_stream(mcb_model *m, void *cuda_stream) {
  if (!m) return fail(MCB_ERR_ARG, "model is NULL");
  cudaSetDevice(m->device);
  cudaStreamSynchronize(m->stream);
  m->stream = cuda_stream ? (cudaStream_t)cuda_stream : m->own_stream;
  return MCB_OK;
}

int mcb_synchronize(mcb_model *m) {
  if (!m) return fail(MCB_ERR_ARG, "model is NULL");
  MCB_CUDA(cudaSetDevice(m->device));
  MCB_CUDA(cudaStreamSynchronize(m->stream));
  return MCB_OK;
}

int mcb_shape(const mcb_model *m, size_t *n_state, size_t *n_particles, size_t *n_pars,
              size_t *n_index) {
  if (!m) return fail(MCB_ERR_ARG, "model is NULL");
  if (n_state) *n_state = m->v.n_state;
  if (n_particles) *n_particles = m->v.n_particles;
  if (n_pars) *n_pars = m->v.n_sets;
  if (n_index) *n_index = m->v.n_index;
  return MCB_OK;
}

int mcb_time(const mcb_model *m, uint64_t *time) {
  if (!m || !time) return fail(MCB_ERR_ARG, "NULL argument");
  *time = m->time;
  return MCB_OK;
}

int mcb_pars(const mcb_model *m, double *out) {
  if (!m || !out) return fail(MCB_ERR_ARG, "NULL argument");
  for (size_t p = 0; p < m->v.n_sets; ++p)
    std::memcpy(out + p * m->desc->n_par, m->pars_host.data() + p * mcb::kMaxPar,
                m->desc->n_par * sizeof(double));
  return MCB_OK;
}

int mcb_set_index(mcb_model *m, const size_t *index, size_t n_index) {
  if (!m) return fail(MCB_ERR_ARG, "model is NULL");
  MCB_CUDA(cudaSetDevice(m->device));
  if (!index) {
    n_index = m->v.n_state;
    m->index_host.resize(n_index);
    for (size_t s = 0; s < n_index; ++s) m->index_host[s] = (int)s;
  } else {
    if (n_index == 0 || n_index > (size_t)mcb::kMaxState)
      return fail(MCB_ERR_ARG, "index length must be between 1 and 8");
    for (size_t k = 0; k < n_index; ++k)
      if (index[k] >= (size_t)m->v.n_state) return fail(MCB_ERR_ARG, "index out of range");
    m->index_host.assign(n_index, 0);
    for (size_t k = 0; k < n_index; ++k) m->index_host[k] = (int)index[k];
  }
  MCB_CUDA(cudaStreamSynchronize(m->stream));
  MCB_CUDA(cudaMemcpy(m->d_index, m->index_host.data(), n_index * sizeof(int),
                      cudaMemcpyHostToDevice));
  if ((int)n_index != m->v.n_index) {
    m->v.n_index = (int)n_index;
    clear_graphs(m);
  }
  return MCB_OK;
}

int mcb_run(mcb_model *m, uint64_t time_end, double *out) {
  if (!m) return fail(MCB_ERR_ARG, "model is NULL");
  MCB_CUDA(cudaSetDevice(m->device));
  if (time_end < m->time) return fail(MCB_ERR_ARG, "'time_end' must be at least the current time");
  if (time_end > m->time) {
    run_compare(m, m->v.y, m->v.y, m->time, (uint32_t)(time_end - m->time), nullptr, -1, nullptr);
    m->time = time_end;
    reset_cursor(m);
    MCB_TRY(check_launches(m));
  }
  if (out) return pack_to_host(m, m->v.y, m->d_index, m->v.n_index, out);
  MCB_CUDA(cudaGetLastError());
  return MCB_OK;
}

// model$simulate(times) (R/predict.R:79): every launch of the run kernel also writes
// the selected rows of its end state into a slice buffer in HBM; slices are packed
// to R's layout and copied out once per chunk (256 MB of slices), not once per time.
int mcb_simulate(mcb_model *m, const uint64_t *times, size_t n_times, double *out) {
  if (!m || !times || !out) return fail(MCB_ERR_ARG, "NULL argument");
  MCB_CUDA(cudaSetDevice(m->device));
  View &v = m->v;
  for (size_t k = 0; k < n_times; ++k) {
    if (k > 0 && times[k] <= times[k - 1]) return fail(MCB_ERR_ARG, "'times' must be increasing");
    if (times[k] < m->time) return fail(MCB_ERR_ARG, "'time_end' must be at least the current time");
  }
  if (n_times == 0) return MCB_OK;
  const size_t slice = (size_t)v.n_index * v.ld;                 // doubles per saved slice
  const size_t packed = (size_t)v.n_index * v.n_total;           // doubles per packed slice
  size_t chunk = ((size_t)256 << 20) / (slice * sizeof(double));
  if (chunk < 1) chunk = 1;
  if (chunk > n_times) chunk = n_times;
  const size_t slice_bytes = round_up(chunk * slice * sizeof(double), 256);
  MCB_TRY(ensure_stage(m, slice_bytes + chunk * packed * sizeof(double)));
  double *d_slices = m->d_stage;
  double *d_packed = reinterpret_cast<double *>(reinterpret_cast<char *>(m->d_stage) + slice_bytes);
  for (size_t c0 = 0; c0 < n_times; c0 += chunk) {
    const size_t nc = n_times - c0 < chunk ? n_times - c0 : chunk;
    for (size_t k = 0; k < nc; ++k) {
      const uint64_t t_end = times[c0 + k];
      // n_steps = 0 (t_end == current time) still records the slice
      run_compare(m, v.y, v.y, m->time, (uint32_t)(t_end - m->time), nullptr, -1,
                  d_slices + k * slice);
      m->time = t_end;
    }
    MCB_TRY(check_launches(m));
    const dim3 grid(blocks_for(v.n_total, 256), (unsigned)nc);
    mcb::k_pack_slices<<<grid, 256, 0, m->stream>>>(d_slices, v.ld, v.n_total, v.n_index, nc, d_packed);
    ++m->launches;
    MCB_CUDA(cudaMemcpyAsync(out + c0 * packed, d_packed, nc * packed * sizeof(double),
                             cudaMemcpyDeviceToHost, m->stream));
    MCB_CUDA(cudaStreamSynchronize(m->stream));
  }
  reset_cursor(m);
  return MCB_OK;
}

int mcb_state(mcb_model *m, const size_t *index, size_t n_index, double *out) {
  if (!m || !out) return fail(MCB_ERR_ARG, "NULL argument");
  MCB_CUDA(cudaSetDevice(m->device));
  if (!index) return pack_to_host(m, m->v.y, nullptr, m->v.n_state, out);
  if (n_index == 0) return MCB_OK;
  std::vector<int> idx(n_index);
  for (size_t k = 0; k < n_index; ++k) {
    if (index[k] >= (size_t)m->v.n_state) return fail(MCB_ERR_ARG, "index out of range");
    idx[k] = (int)index[k];
  }
  int *d_idx = nullptr;
  MCB_CUDA(cudaMalloc(&d_idx, n_index * sizeof(int)));
  cudaMemcpy(d_idx, idx.data(), n_index * sizeof(int), cudaMemcpyHostToDevice);
  int rc = pack_to_host(m, m->v.y, d_idx, (int)n_index, out);
  cudaFree(d_idx);
  return rc;
}

int mcb_state_particles(mcb_model *m, const size_t *particles, size_t n, double *out) {
  if (!m || !particles || !out) return fail(MCB_ERR_ARG, "NULL argument");
  if (n == 0) return MCB_OK;
  MCB_CUDA(cudaSetDevice(m->device));
  View &v = m->v;
  std::vector<unsigned long long> idx(n);
  for (size_t j = 0; j < n; ++j) {
    if (particles[j] >= v.n_total) return fail(MCB_ERR_ARG, "particle index out of range");
    idx[j] = particles[j];
  }
  const size_t idx_bytes = round_up(n * sizeof(unsigned long long), 256);
  const size_t out_bytes = n * (size_t)v.n_state * sizeof(double);
  MCB_TRY(ensure_stage(m, idx_bytes + out_bytes));
  unsigned long long *d_idx = reinterpret_cast<unsigned long long *>(m->d_stage);
  double *d_out = reinterpret_cast<double *>(reinterpret_cast<char *>(m->d_stage) + idx_bytes);
  MCB_CUDA(cudaMemcpyAsync(d_idx, idx.data(), n * sizeof(unsigned long long),
                           cudaMemcpyHostToDevice, m->stream));
  mcb::k_state_particles<<<blocks_for(n, 64), 64, 0, m->stream>>>(v.y, v.ld, v.n_state, d_idx, n,
                                                                    d_out);
  ++m->launches;
  MCB_CUDA(cudaMemcpyAsync(out, d_out, out_bytes, cudaMemcpyDeviceToHost, m->stream));
  MCB_CUDA(cudaStreamSynchronize(m->stream));
  return MCB_OK;
}

int mcb_update_state(mcb_model *m, const double *pars, const double *state, size_t state_len,
                     int has_time, uint64_t time, int set_initial_state) {
  if (!m) return fail(MCB_ERR_ARG, "model is NULL");
  MCB_CUDA(cudaSetDevice(m->device));
  View &v = m->v;
  if (state && state_len != (size_t)v.n_state && state_len != (size_t)v.n_state * v.n_total)
    return fail(MCB_ERR_ARG, "'state' must have length n_state or n_state * n_particles");
  if (has_time) {
    m->time = time;
    reset_cursor(m);
  }
  if (pars) {
    const int np = m->desc->n_par;
    m->pars_host.assign(v.n_sets * mcb::kMaxPar, 0.0);
    for (size_t p = 0; p < v.n_sets; ++p)
      for (int k = 0; k < np; ++k) m->pars_host[p * mcb::kMaxPar + k] = pars[p * np + k];
    // pageable source: the runtime stages it before returning, so pars_host may change
    MCB_CUDA(cudaMemcpyAsync(m->d_pars, m->pars_host.data(),
                             m->pars_host.size() * sizeof(double), cudaMemcpyHostToDevice,
                             m->stream));
    prepare_sets(m);
    if (set_initial_state && !state) set_initial(m);
  }
  if (state) {
    MCB_TRY(ensure_stage(m, state_len * sizeof(double)));
    MCB_CUDA(cudaMemcpyAsync(m->d_stage, state, state_len * sizeof(double),
                             cudaMemcpyHostToDevice, m->stream));
    mcb::k_unpack_state<<<blocks_for(v.n_total, 256), 256, 0, m->stream>>>(
        m->d_stage, state_len == (size_t)v.n_state, v.n_state, v.ld, v.n_total, v.y);
    ++m->launches;
    MCB_CUDA(cudaStreamSynchronize(m->stream));
  }
  MCB_CUDA(cudaGetLastError());
  return MCB_OK;
}

int mcb_reorder(mcb_model *m, const int *kappa) {
  if (!m || !kappa) return fail(MCB_ERR_ARG, "NULL argument");
  MCB_CUDA(cudaSetDevice(m->device));
  View &v = m->v;
  for (size_t i = 0; i < v.n_total; ++i)
    if (kappa[i] < 1 || (size_t)kappa[i] > v.n_particles)
      return fail(MCB_ERR_ARG, "'index' must be in [1, n_particles]");
  MCB_CUDA(cudaMemcpyAsync(m->d_kappa, kappa, v.n_total * sizeof(int), cudaMemcpyHostToDevice,
                           m->stream));
  mcb::k_gather_by_kappa<<<blocks_for(v.n_total, 256), 256, 0, m->stream>>>(v, m->d_kappa);
  ++m->launches;
  MCB_CUDA(cudaMemcpyAsync(v.y, v.y_tmp, (size_t)v.n_state * v.ld * sizeof(double),
                           cudaMemcpyDeviceToDevice, m->stream));
  MCB_CUDA(cudaStreamSynchronize(m->stream));
  return MCB_OK;
}

int mcb_rng_state(mcb_model *m, int first_only, uint64_t *out) {
  if (!m || !out) return fail(MCB_ERR_ARG, "NULL argument");
  MCB_CUDA(cudaSetDevice(m->device));
  const View &v = m->v;
  const size_t n = first_only ? 1 : v.n_total + (size_t)v.n_fs;
  std::vector<uint64_t> soa(4 * n);
  MCB_CUDA(cudaStreamSynchronize(m->stream));
  for (int w = 0; w < 4; ++w)
    MCB_CUDA(cudaMemcpy(soa.data() + w * n, v.rng + w * v.ld_rng, n * sizeof(uint64_t),
                        cudaMemcpyDeviceToHost));
  for (size_t i = 0; i < n; ++i)
    for (int w = 0; w < 4; ++w) out[4 * i + w] = soa[w * n + i];
  return MCB_OK;
}

int mcb_set_rng_state(mcb_model *m, const uint64_t *state, size_t n_words) {
  if (!m || !state) return fail(MCB_ERR_ARG, "NULL argument");
  MCB_CUDA(cudaSetDevice(m->device));
  const View &v = m->v;
  const size_t n = v.n_total + (size_t)v.n_fs;
  if (n_words != 4 * n)
    return fail(MCB_ERR_ARG, "'rng_state' must hold 32 bytes for each particle and filter stream");
  std::vector<uint64_t> soa(4 * n);
  for (size_t i = 0; i < n; ++i)
    for (int w = 0; w < 4; ++w) soa[w * n + i] = state[4 * i + w];
  MCB_CUDA(cudaStreamSynchronize(m->stream));
  for (int w = 0; w < 4; ++w)
    MCB_CUDA(cudaMemcpy(v.rng + w * v.ld_rng, soa.data() + w * n, n * sizeof(uint64_t),
                        cudaMemcpyHostToDevice));
  return MCB_OK;
}

int mcb_copy_sets(mcb_model *dst, mcb_model *src, const unsigned char *take, int copy_state,
                  int copy_rng) {
  if (!dst || !src || !take) return fail(MCB_ERR_ARG, "NULL argument");
  const View &a = dst->v, &b = src->v;
  if (a.n_particles != b.n_particles || a.n_sets != b.n_sets || a.n_state != b.n_state ||
      a.n_fs != b.n_fs || dst->device != src->device)
    return fail(MCB_ERR_ARG, "copy_sets needs two handles of the same shape on one device");
  if (copy_rng && a.n_fs == 1 && a.n_sets > 1)
    return fail(MCB_ERR_STATE, "per-set RNG copies need per-set filter streams");
  MCB_CUDA(cudaSetDevice(dst->device));
  MCB_CUDA(cudaStreamSynchronize(src->stream));
  // the flags travel through the handle's staging buffer (no allocation per call: SMC^2
  // copies sets at every replenishment)
  MCB_TRY(ensure_stage(dst, a.n_sets));
  unsigned char *d_take = reinterpret_cast<unsigned char *>(dst->d_stage);
  MCB_CUDA(cudaMemcpyAsync(d_take, take, a.n_sets, cudaMemcpyHostToDevice, dst->stream));
  mcb::k_copy_sets<<<blocks_for(a.n_total, 256), 256, 0, dst->stream>>>(a, b, d_take, copy_state,
                                                                        copy_rng);
  ++dst->launches;
  cudaError_t err = cudaStreamSynchronize(dst->stream);
  if (err != cudaSuccess) return fail(MCB_ERR_CUDA, std::string("copy_sets: ") + cudaGetErrorString(err));
  return MCB_OK;
}

int mcb_rng_jump_n(uint64_t state[4], uint64_t n) {
  if (!state) return fail(MCB_ERR_ARG, "state is NULL");
  if (n >> mcb::kJumpLevels) return fail(MCB_ERR_ARG, "jump count too large");
  {
    std::lock_guard<std::mutex> lock(g_jump.mu);
    if (!g_jump.host) {
      g_jump.host = new mcb::JumpTable;
      mcb::jump_table_build(*g_jump.host);
    }
  }
  for (int l = 0; l < mcb::kJumpLevels; ++l) {
    if (!((n >> l) & 1ULL)) continue;
    uint64_t acc[4] = {0, 0, 0, 0};
    for (int b = 0; b < 256; ++b)
      if ((state[b >> 6] >> (b & 63)) & 1ULL)
        for (int w = 0; w < 4; ++w) acc[w] ^= g_jump.host->col[l][b][w];
    for (int w = 0; w < 4; ++w) state[w] = acc[w];
  }
  return MCB_OK;
}

int mcb_set_filter_stream(mcb_model *m, const uint64_t state[4]) {
  if (!m || !state) return fail(MCB_ERR_ARG, "NULL argument");
  if (m->v.n_fs != 1) return fail(MCB_ERR_STATE, "the model has one filter stream per set");
  MCB_CUDA(cudaSetDevice(m->device));
  MCB_CUDA(cudaStreamSynchronize(m->stream));
  for (int w = 0; w < 4; ++w)
    MCB_CUDA(cudaMemcpy(m->v.rng + w * m->v.ld_rng + m->v.n_total, state + w, sizeof(uint64_t),
                        cudaMemcpyHostToDevice));
  return MCB_OK;
}

int mcb_set_stream_sharing(mcb_model *m, size_t n_sets_global, size_t set_offset,
                           const unsigned char *na_global) {
  if (!m) return fail(MCB_ERR_ARG, "model is NULL");
  View &v = m->v;
  if (v.n_fs != 1) return fail(MCB_ERR_STATE, "the model has one filter stream per set");
  if (m->n_data == 0) return fail(MCB_ERR_STATE, "set the data before the stream sharing");
  if (set_offset + v.n_sets > n_sets_global)
    return fail(MCB_ERR_ARG, "this handle's sets do not fit in the global set range");
  if (!na_global) return fail(MCB_ERR_ARG, "na_global is NULL");
  MCB_CUDA(cudaSetDevice(m->device));
  MCB_CUDA(cudaStreamSynchronize(m->stream));
  clear_graphs(m);
  if (m->d_na_global) cudaFree(m->d_na_global);
  m->d_na_global = nullptr;
  MCB_CUDA(cudaMalloc(&m->d_na_global, m->n_data * n_sets_global));
  MCB_CUDA(cudaMemcpy(m->d_na_global, na_global, m->n_data * n_sets_global,
                      cudaMemcpyHostToDevice));
  v.na_global = m->d_na_global;
  v.n_sets_global = (int)n_sets_global;
  v.set_offset = (int)set_offset;
  return MCB_OK;
}

int mcb_set_data(mcb_model *m, size_t n_data, const uint64_t *time_end, const double *values,
                 int shared) {
  if (!m || !time_end || !values) return fail(MCB_ERR_ARG, "NULL argument");
  if (n_data == 0) return fail(MCB_ERR_ARG, "data must have at least one row");
  if (m->desc->n_data == 0)
    return fail(MCB_ERR_STATE, "Your model does not have a built-in 'compare' function");
  MCB_CUDA(cudaSetDevice(m->device));
  View &v = m->v;
  for (size_t i = 1; i < n_data; ++i)
    if (time_end[i] <= time_end[i - 1]) return fail(MCB_ERR_ARG, "data times must be increasing");
  MCB_CUDA(cudaStreamSynchronize(m->stream));
  clear_graphs(m);
  const size_t sets = shared ? 1 : v.n_sets;
  std::vector<double> packed(n_data * sets * 2);
  m->row_all_na.assign(n_data, 1);
  for (size_t r = 0; r < n_data; ++r)
    for (size_t s = 0; s < sets; ++s) {
      const double x = values[r * sets + s];
      packed[(r * sets + s) * 2] = x;
      // lgamma(x + 1) depends on the datum only: once here, never per particle
      packed[(r * sets + s) * 2 + 1] = std::isnan(x) ? NAN : std::lgamma(x + 1.0);
      if (!std::isnan(x)) m->row_all_na[r] = 0;
    }
  void *old[] = {m->d_data, m->d_u, m->d_fs_after, m->d_wmax, m->d_data_time};
  for (void *p : old)
    if (p) cudaFree(p);
  m->d_data = nullptr; m->d_u = nullptr; m->d_fs_after = nullptr; m->d_wmax = nullptr;
  m->d_data_time = nullptr;
  MCB_CUDA(cudaMalloc(&m->d_data_time, n_data * sizeof(uint64_t)));
  MCB_CUDA(cudaMemcpy(m->d_data_time, time_end, n_data * sizeof(uint64_t), cudaMemcpyHostToDevice));
  MCB_CUDA(cudaMalloc(&m->d_data, packed.size() * sizeof(double)));
  MCB_CUDA(cudaMalloc(&m->d_u, n_data * v.n_sets * sizeof(double)));
  MCB_CUDA(cudaMalloc(&m->d_fs_after, n_data * (size_t)v.n_fs * 4 * sizeof(uint64_t)));
  MCB_CUDA(cudaMalloc(&m->d_wmax, n_data * v.n_sets * sizeof(unsigned long long)));
  MCB_CUDA(cudaMemcpy(m->d_data, packed.data(), packed.size() * sizeof(double),
                      cudaMemcpyHostToDevice));
  if (m->d_na_global) cudaFree(m->d_na_global);
  m->d_na_global = nullptr;
  v.na_global = nullptr;   // a new series: mcb_set_stream_sharing must be called again
  v.n_sets_global = (int)v.n_sets;
  v.set_offset = 0;
  v.data = m->d_data;
  v.data_sets = (int)sets;
  v.u_draws = m->d_u;
  v.wmax = m->d_wmax;
  m->n_data = n_data;
  m->data_time.assign(time_end, time_end + n_data);
  reset_cursor(m);
  return MCB_OK;
}

int mcb_compare_data(mcb_model *m, double *logw, int *has_data) {
  if (!m || !logw || !has_data) return fail(MCB_ERR_ARG, "NULL argument");
  MCB_CUDA(cudaSetDevice(m->device));
  View &v = m->v;
  *has_data = 0;
  int row = -1;
  for (size_t r = 0; r < m->n_data; ++r)
    if (m->data_time[r] == m->time) row = (int)r;
  if (row < 0 || m->row_all_na[row]) return MCB_OK;
  MCB_CUDA(cudaMemsetAsync(v.logw, 0, v.n_total * sizeof(double), m->stream));
  MCB_CUDA(cudaMemsetAsync(m->d_flags, 0, 2 * sizeof(int), m->stream));
  run_compare(m, v.y, v.y, m->time, 0, nullptr, row, nullptr);
  MCB_TRY(check_launches(m));
  MCB_CUDA(cudaMemcpyAsync(logw, v.logw, v.n_total * sizeof(double), cudaMemcpyDeviceToHost,
                           m->stream));
  MCB_CUDA(cudaStreamSynchronize(m->stream));
  *has_data = 1;
  return MCB_OK;
}

int mcb_resample(mcb_model *m, const double *weights, int *kappa) {
  if (!m || !weights || !kappa) return fail(MCB_ERR_ARG, "NULL argument");
  MCB_CUDA(cudaSetDevice(m->device));
  View &v = m->v;
  for (size_t i = 0; i < v.n_total; ++i)
    if (!(weights[i] >= 0.0 && weights[i] <= 1.0))
      return fail(MCB_ERR_ARG, "'weights' must lie in [0, 1] (exp(log_weight - max))");
  MCB_CUDA(cudaMemsetAsync(m->d_flags, 0, 2 * sizeof(int), m->stream));
  MCB_CUDA(cudaMemcpyAsync(v.logw, weights, v.n_total * sizeof(double), cudaMemcpyHostToDevice,
                           m->stream));
  draw_uniforms(m, -1, 0);
  weights_scan(m, -1);
  ancestors(m, -1);
  MCB_TRY(check_launches(m));
  MCB_CUDA(cudaMemcpyAsync(v.y_tmp, v.y, (size_t)v.n_state * v.ld * sizeof(double),
                           cudaMemcpyDeviceToDevice, m->stream));
  resample_gather(m, m->d_kappa, v.y_tmp, v.y);
  mcb::k_fill_ancestors<<<blocks_for(v.n_total, mcb::kGatherThreads), mcb::kGatherThreads, 0,
                          m->stream>>>(v, m->d_kappa);
  ++m->launches;
  std::vector<int> k0(v.n_total);
  MCB_CUDA(cudaMemcpyAsync(k0.data(), m->d_kappa, v.n_total * sizeof(int), cudaMemcpyDeviceToHost,
                           m->stream));
  MCB_CUDA(cudaStreamSynchronize(m->stream));
  for (size_t i = 0; i < v.n_total; ++i)
    kappa[i] = k0[i] - (int)((i / v.n_particles) * v.n_particles) + 1;
  return MCB_OK;
}

int mcb_filter_rows(const mcb_model *m, uint64_t time_end, size_t *n_rows) {
  if (!m || !n_rows) return fail(MCB_ERR_ARG, "NULL argument");
  int first, last;
  filter_rows_range(m, time_end, &first, &last);
  *n_rows = (size_t)(last - first);
  return MCB_OK;
}

int mcb_filter(mcb_model *m, uint64_t time_end, int save_trajectories,
               const uint64_t *snapshot_times, size_t n_snapshots,
               const double *min_log_likelihood, size_t n_min, double *log_likelihood,
               double *trajectories, double *snapshots) {
  if (!m || !log_likelihood) return fail(MCB_ERR_ARG, "NULL argument");
  if (save_trajectories == 1 && !trajectories) return fail(MCB_ERR_ARG, "trajectories is NULL");
  if (save_trajectories == MCB_TRAJECTORIES_ON_DEVICE) trajectories = nullptr;
  if (n_snapshots && (!snapshot_times || !snapshots)) return fail(MCB_ERR_ARG, "snapshots is NULL");
  MCB_CUDA(cudaSetDevice(m->device));
  return run_filter(m, time_end, save_trajectories != 0, snapshot_times, n_snapshots,
                    min_log_likelihood, n_min, true, log_likelihood, trajectories, snapshots);
}

int mcb_filter_history(mcb_model *m, const size_t *particles, size_t n, double *out) {
  if (!m || !particles || !out) return fail(MCB_ERR_ARG, "NULL argument");
  if (m->hist_rows < 0)
    return fail(MCB_ERR_STATE, "no history on the device: run filter with save_trajectories first");
  if (n == 0) return MCB_OK;
  MCB_CUDA(cudaSetDevice(m->device));
  View &v = m->v;
  std::vector<unsigned long long> idx(n);
  for (size_t j = 0; j < n; ++j) {
    if (particles[j] >= v.n_total) return fail(MCB_ERR_ARG, "particle index out of range");
    idx[j] = particles[j];
  }
  const size_t T1 = (size_t)m->hist_rows + 1;
  const size_t out_bytes = T1 * n * v.n_index * sizeof(double);
  const size_t idx_bytes = round_up(n * sizeof(unsigned long long), 256);
  MCB_TRY(ensure_stage(m, idx_bytes + out_bytes));
  unsigned long long *d_idx = reinterpret_cast<unsigned long long *>(m->d_stage);
  double *d_out = reinterpret_cast<double *>(reinterpret_cast<char *>(m->d_stage) + idx_bytes);
  MCB_CUDA(cudaMemcpyAsync(d_idx, idx.data(), n * sizeof(unsigned long long),
                           cudaMemcpyHostToDevice, m->stream));
  mcb::k_trace_particles<<<blocks_for(n, 64), 64, 0, m->stream>>>(
      v, m->d_hist_value, m->d_hist_order, m->hist_rows, m->hist_first, d_idx, n, d_out);
  ++m->launches;
  MCB_CUDA(cudaMemcpyAsync(out, d_out, out_bytes, cudaMemcpyDeviceToHost, m->stream));
  MCB_CUDA(cudaStreamSynchronize(m->stream));
  return MCB_OK;
}

int mcb_set_persistent(mcb_model *m, int enable) {
  if (!m) return fail(MCB_ERR_ARG, "model is NULL");
  MCB_CUDA(cudaSetDevice(m->device));
  m->persist_blocks = enable ? persistent_capacity(m) : 0;
  return MCB_OK;
}

int mcb_filter_enqueue(mcb_model *m, uint64_t time_end) {
  if (!m) return fail(MCB_ERR_ARG, "model is NULL");
  MCB_CUDA(cudaSetDevice(m->device));
  return run_filter(m, time_end, false, nullptr, 0, nullptr, 0, false, nullptr, nullptr, nullptr);
}

int mcb_filter_collect(mcb_model *m, double *log_likelihood) {
  if (!m) return fail(MCB_ERR_ARG, "model is NULL");
  if (!m->pending) return fail(MCB_ERR_STATE, "no filter enqueued");
  MCB_CUDA(cudaSetDevice(m->device));
  m->pending = false;
  MCB_CUDA(cudaStreamSynchronize(m->stream));
  MCB_CUDA(cudaGetLastError());
  const bool stopped = m->h_flags[0] != 0;
  const int last_row = stopped ? m->h_flags[1] : m->pend_last - 1;
  if (m->timing && !stopped) {
    auto add = [&](int slot, size_t e0, size_t e1) {
      float ms = 0.f;
      if (e1 < m->events.size() &&
          cudaEventElapsedTime(&ms, m->events[e0], m->events[e1]) == cudaSuccess) {
        m->kernel_ms[slot] += ms;
        ++m->kernel_n[slot];
      }
    };
    for (int row = m->pend_first; row < m->pend_last; ++row) {
      add(0, (size_t)row * kMarks, (size_t)row * kMarks + 1);
      if (!m->row_all_na[row]) {
        add(1, (size_t)row * kMarks + 1, (size_t)row * kMarks + 2);
        add(2, (size_t)row * kMarks + 2, (size_t)row * kMarks + 3);
      }
    }
    if (m->pend_last > m->pend_first) add(3, m->n_data * kMarks, m->n_data * kMarks + 1);
  }
  if (last_row >= m->pend_first) {
    m->time = m->data_time[last_row];
    m->data_cursor = (size_t)last_row + 1;
  }
  if (log_likelihood)
    std::memcpy(log_likelihood, m->h_ll, m->v.n_sets * sizeof(double));
  return MCB_OK;
}

int mcb_log_likelihood_device(mcb_model *m, double **device_ptr) {
  if (!m || !device_ptr) return fail(MCB_ERR_ARG, "NULL argument");
  *device_ptr = m->d_ll;
  return MCB_OK;
}

int mcb_set_kernel_timing(mcb_model *m, int enable) {
  if (!m) return fail(MCB_ERR_ARG, "model is NULL");
  m->timing = enable != 0;
  for (int k = 0; k < 4; ++k) { m->kernel_ms[k] = 0; m->kernel_n[k] = 0; }
  return MCB_OK;
}

int mcb_kernel_timing(mcb_model *m, double ms[4], uint64_t launches[4]) {
  if (!m || !ms || !launches) return fail(MCB_ERR_ARG, "NULL argument");
  for (int k = 0; k < 4; ++k) {
    ms[k] = m->kernel_ms[k];
    launches[k] = m->kernel_n[k];
    m->kernel_ms[k] = 0;
    m->kernel_n[k] = 0;
  }
  return MCB_OK;
}

int mcb_launch_count(const mcb_model *m, uint64_t *n) {
  if (!m || !n) return fail(MCB_ERR_ARG, "NULL argument");
  *n = m->launches;
  return MCB_OK;
}

}  // extern "C"
