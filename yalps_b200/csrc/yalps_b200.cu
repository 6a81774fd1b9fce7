// yalps_b200.cu -- C ABI (include/yalps_b200.h) over the sm_100a kernels.
//
// Host runtime: context, device buffer pool, double-buffered chunked H2D -> kernel -> D2H pipeline,
// kernel-path selection, and the branch-and-cut driver (speculative node waves replayed in the
// reference's pop order, src/branchAndCut.ts:89-176).
#include "../../include/yalps_b200.h"

#include <cuda_runtime.h>

#include <dlfcn.h>
#include <nccl.h>  // types and prototypes only: libnccl.so.2 is opened at run time (multi.inl)

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <condition_variable>
#include <functional>
#include <limits>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "kernel_table.h"
#include "aux_kernels.cuh"
#include "cluster_kernel.cuh"
#include "grid_kernel.cuh"
#include "multigrid_kernel.cuh"
#include "tmem_launch.h"
#include "bnb_launch.h"

using namespace yalps;

namespace {

thread_local std::string g_create_error;

// cudaFuncAttributeMaxDynamicSharedMemorySize is a property of the function on the device, shared by every ctx of
// the process: it is only ever raised, under a lock (several contexts may run on different host threads).
std::mutex g_attr_mutex;
std::unordered_map<std::string, int> g_smem_attr;

cudaError_t raise_smem_limit(int device, const void *fn, int bytes) {
  std::lock_guard<std::mutex> lock(g_attr_mutex);
  int &cur = g_smem_attr[std::to_string(device) + ":" + std::to_string((size_t)fn)];
  if (bytes <= cur) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) cur = bytes;
  return e;
}

struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
};

struct Root {
  bool valid = false;
  int H = 0, W = 0, max_extra = 0;
  double density = 1.0;  // fraction of non-zero cells of (a sample of) the root-optimal tableau
  DevBuf m, pos, var;
  std::vector<double> h_m;  // host copy of column 0 is enough for the driver, but keep rhs + pos + var
  std::vector<double> h_rhs;
  std::vector<int32_t> h_pos, h_var;
};

}  // namespace

struct yalps_ctx {
  int device = 0;
  cudaDeviceProp prop{};
  int smem_optin = 0;
  std::string error;
  cudaStream_t streams[2]{};
  cudaEvent_t events[2]{};
  int64_t launches = 0;
  int tune_path = 0, tune_threads = 0, tune_rows = 0;
  std::vector<cudaStream_t> aux_streams;  // concurrent K4 launches of one node wave
  std::vector<cudaEvent_t> aux_events;
  cudaEvent_t fork_event = nullptr;
  bool keep_final = false;         // solve_host (n == 1): leave the final tableau on the device, no D2H copy of it
  double *kept_final = nullptr;    // ... and where it is (valid until the next batch call on this ctx)
  // yalps_solve_sparse: solve_host (n == 1) builds its device tableau from these (cell, value) pairs instead of copying
  // a dense host matrix (coo_nnz < 0: not in use); coo_dup reports that duplicates with different values were seen
  const int32_t *coo_cell = nullptr;
  const double *coo_val = nullptr;
  int64_t coo_nnz = -1;
  int coo_dup = 0;
  // YALPS_BNB_DEBUG: where the host spends a node wave (ns): before the first CUDA call, enqueueing, waiting, copying out
  int64_t wave_ns[4] = {0, 0, 0, 0};
  int wave = 256;  // upper bound of the adaptive look-ahead of the branch-and-cut driver
  int kc_tma = 0;    // KC: staging of the winner's pivot row (YALPS_KC_TMA: 0 ld.global.cg, 1 cp.async.bulk, 2 multicast)
  int bnb_mode = 0;  // 0: device-resident search when it fits, else host waves; 1: host waves only; 2: device only
  int bnb_workers = 0;  // device-resident search: worker CTAs (0 = one per SM beside the scheduler; yalps_multi_solve_many
                        // shares the SMs between its concurrent searches)
  int64_t replica_forks = -1;  // last yalps_solve_replicas call: replicas that left the shared path (-1: sharing not used)
  int replica_sharing = 1;  // yalps_solve_replicas: follow the base tableau's pivot path while a replica makes the same choices
  unsigned long long *d_rows = nullptr;  // roofline diagnostics: device counter(s) of rewritten rows (yalps_set_row_counter)
  int rows_per_lp = 0;
  // pooled device buffers (index = purpose * 2 + pipeline slot)
  std::unordered_map<std::string, DevBuf> pool;
  std::unordered_map<std::string, DevBuf> pinned;
  std::unordered_map<std::string, int> occ_cache;
  std::unordered_map<std::string, double> density_cache;
  Root root;
};

namespace {

int fail(yalps_ctx *ctx, int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (ctx)
    ctx->error = buf;
  else
    g_create_error = buf;
  return code;
}

#define CU(ctx, call)                                                                                   \
  do {                                                                                                  \
    cudaError_t e_ = (call);                                                                            \
    if (e_ != cudaSuccess)                                                                              \
      return fail(ctx, YALPS_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

int dev_ensure(yalps_ctx *ctx, const std::string &name, size_t bytes, void **out) {
  DevBuf &b = ctx->pool[name];
  if (b.cap < bytes) {
    if (b.p) CU(ctx, cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
    size_t want = std::max(bytes, (size_t)256);
    CU(ctx, cudaMalloc(&b.p, want));
    b.cap = want;
  }
  *out = b.p;
  return 0;
}

int pin_ensure(yalps_ctx *ctx, const std::string &name, size_t bytes, void **out) {
  DevBuf &b = ctx->pinned[name];
  if (b.cap < bytes) {
    if (b.p) CU(ctx, cudaFreeHost(b.p));
    b.p = nullptr;
    b.cap = 0;
    size_t want = std::max(bytes * 2, (size_t)4096);
    CU(ctx, cudaHostAlloc(&b.p, want, cudaHostAllocMapped | cudaHostAllocPortable));
    b.cap = want;
  }
  *out = b.p;
  return 0;
}

// Device-side alias of a buffer from pin_ensure (zero-copy access over PCIe for latency-bound small launches).
int pin_device_ptr(yalps_ctx *ctx, void *host, void **dev) {
  CU(ctx, cudaHostGetDevicePointer(dev, host, 0));
  return 0;
}

// ---- kernel table -------------------------------------------------------------------------------------
// All k_simplex instantiations (kernel_table.h), gathered once.
const std::vector<KernelEntry> &all_kernels() {
  static const std::vector<KernelEntry> table = [] {
    std::vector<KernelEntry> v;
    int n = 0;
    const KernelEntry *t;
    t = kernel_table_base(&n);
    v.insert(v.end(), t, t + n);
    t = kernel_table_split_a(&n);
    v.insert(v.end(), t, t + n);
    t = kernel_table_split_b(&n);
    v.insert(v.end(), t, t + n);
    t = kernel_table_split_c(&n);
    v.insert(v.end(), t, t + n);
    return v;
  }();
  return table;
}

// Thread (row group, t) keeps the pivot-row cells of vector-columns t, t+NTC, ... in registers:
// 32*NWC*KC*VW >= W-1 is required.  Among the kernels with `nwr` row groups: smallest NWC >= nwc_want whose
// widest KC covers the row (else the largest NWC), then the smallest covering KC.
const KernelEntry *pick_kernel(int nwc_want, int nwr, int W, bool resident) {
  const int vw = resident ? 2 : 1;
  const int wm1 = std::max(W - 1, 1);
  const KernelEntry *best = nullptr;
  for (const auto &k : all_kernels()) {
    if (k.nwr != nwr) continue;
    if ((long long)k.nw * 32 * k.kc * vw < wm1) continue;
    if (!best) {
      best = &k;
      continue;
    }
    const bool k_ok = k.nw >= nwc_want, b_ok = best->nw >= nwc_want;
    if (k_ok != b_ok) {
      if (k_ok) best = &k;
      continue;
    }
    if (k_ok) {  // both at least as wide as wanted: prefer the narrower CTA, then the smaller KC
      if (k.nw < best->nw || (k.nw == best->nw && k.kc < best->kc)) best = &k;
    } else {     // both narrower than wanted: prefer the wider CTA, then the smaller KC
      if (k.nw > best->nw || (k.nw == best->nw && k.kc < best->kc)) best = &k;
    }
  }
  return best;
}

// CTA width by tableau size, from scripts/sweep_paths.py on B200 (profiles/r01_sweep_paths.jsonl).
int default_warps(long long cells, bool resident) {
  if (resident) {
    if (cells < 3000) return 1;
    if (cells < 12000) return 2;
    if (cells < 40000) return 4;
    if (cells < 120000) return 8;
    return 16;
  }
  if (cells < 3000) return 1;
  if (cells < 6000) return 2;
  if (cells < 25000) return 4;
  if (cells < 50000) return 8;
  if (cells < 400000) return 16;
  return 32;
}

struct LaunchPlan {
  bool tmem = false;  // K1t: tableau in tensor memory, one LP per warp (tmem_kernel.cuh)
  bool small_for_grid = false;  // few LPs outside shared memory, yet small enough that one row-split CTA beats K4
  bool resident;
  const KernelEntry *k;
  size_t smem;
  int grid;
};

// density: fraction of non-zero cells of a sample of the input (< 0 = unknown).  Sparse tableaus skip most of
// the rank-1 update, their pivots are latency-bound, and the HBM/L2-resident kernel wins because it needs no
// shared memory for the tableau and therefore runs many more LPs per SM (profiles/r01_sweep_paths.jsonl).
int plan_launch(yalps_ctx *ctx, long long n, int Hcap, int Wcap, bool check_cycles, LaunchPlan *plan,
                double density = -1.0, bool allow_tmem = false) {
  // a forced cluster path is resolved by maybe_cluster(); what it cannot take is planned as in automatic mode
  const int tune_path = (ctx->tune_path == YALPS_PATH_CLUSTER || ctx->tune_path == YALPS_PATH_GRID_RESIDENT) ? (int)YALPS_PATH_AUTO : ctx->tune_path;
  SmemLayout Lr(Hcap, Wcap, true, 32), Lg(Hcap, Wcap, false, 32);
  bool resident = Lr.total <= (size_t)ctx->smem_optin;
  // (throughput mode only: with at most two LPs per SM the shared-memory row-split kernels are 3x faster than K2 on
  // sparse Netlib models -- AFIRO, 148 LPs: 46 vs 155 us)
  if (resident && tune_path == YALPS_PATH_AUTO && density >= 0.0 && density < 0.35 &&
      n > std::max(64LL, 2LL * ctx->prop.multiProcessorCount))
    resident = false;
  if (tune_path == YALPS_PATH_SMEM) {
    if (!resident) return fail(ctx, YALPS_ERR_TOO_LARGE, "tableau %dx%d does not fit in shared memory", Hcap, Wcap);
  } else if (tune_path == YALPS_PATH_GMEM) {
    resident = false;
  }
  const bool grid_ok = tune_path == YALPS_PATH_GRID || (tune_path == YALPS_PATH_AUTO && n <= 16);
  plan->tmem = false;
  // K1t (tensor-memory resident, one LP per warp, no CTA barrier in the pivot loop): every batch whose tableaus fit
  // it, whatever its size and density.  Measured against the previous choices on B200: 1.47-1.68x K1 on big dense
  // batches of every shape from 9x17 to 33x65 (scripts/tmem_vs_smem.py); 1.5-2.3x K2 on big batches with 70-90 % zeros
  // (scripts/tmem_sparse_batches.py); 8-33 % faster than the row-split latency kernels for 1..296 LPs, dense or sparse
  // (scripts/tmem_small_batches.py).
  // Tableaus of 34..65 rows take the 256-column shape (8 LPs per SM, scripts/tmem_tall_crossover.py): 1.4-1.85x K1 on
  // big batches, dense or sparse; 10-20 % faster than K1s for 1..296 DENSE LPs, but 26-40 % slower on sparse Netlib
  // models (AFIRO, KLEIN1: the row-split kernels compact the few active rows, one warp walks all blocks) -- so in
  // latency mode the tall shape needs a density probe that says "dense".
  const bool latency_mode = n <= 2LL * ctx->prop.multiProcessorCount;
  // The short shape wins or ties everywhere, very sparse tableaus included (its sparse row pass touches only the
  // active rows: 90 % zeros, 33x65, one LP: 18.8 vs 18.9 us for K1s; scripts/tmem_small_batches.py).
  const bool tmem_auto = tune_path == YALPS_PATH_AUTO && ctx->tune_threads <= 0 && ctx->tune_rows <= 0 &&
                         (!tmem_kernel_is_tall(Hcap) ||
                          (latency_mode ? density >= 0.5
                                        // throughput mode (scripts/policy_mid_batches.py): KLEIN1 55x55 (density 0.24) 2.2-2.6x
                                        // the HBM-resident K2; very sparse AFIRO 36x33 (0.1) only while one wave of
                                        // 8 LPs per SM covers the batch, beyond that K2 is 20-30 % faster
                                        : (density < 0.0 || density >= 0.15 || n <= 8LL * ctx->prop.multiProcessorCount)));
  if (allow_tmem && tmem_kernel_fits(Hcap, Wcap) && (tune_path == YALPS_PATH_TMEM || tmem_auto)) {
    plan->tmem = true;
    plan->resident = true;
    plan->k = nullptr;
    plan->smem = tmem_kernel_dynamic_smem(Hcap);
    CU(ctx, raise_smem_limit(ctx->device, tmem_kernel_fn(Hcap, false, check_cycles), (int)plan->smem));
    if (ctx->d_rows) CU(ctx, raise_smem_limit(ctx->device, tmem_kernel_fn(Hcap, true, check_cycles), (int)plan->smem));
    // checkCycles: one history buffer per warp, so one CTA per SM bounds the memory (4 warps x 2 x hist_cap ints each)
    const long long ctas = (long long)(check_cycles ? 1 : tmem_kernel_ctas_per_sm(Hcap)) * ctx->prop.multiProcessorCount;
    plan->grid = (int)std::max(1LL, std::min(ctas, n));  // few LPs: one per CTA (the kernel deals LP i to CTA i % grid)
    return 0;
  }
  if (tune_path == YALPS_PATH_TMEM)
    return fail(ctx, YALPS_ERR_TOO_LARGE, "tableau %dx%d does not fit the tensor-memory kernel (max %dx%d, no node mode)",
                Hcap, Wcap, 65, 65);
  plan->resident = resident;
  plan->k = nullptr;
  plan->smem = 0;
  plan->grid = 0;
  if (!resident && Lg.total > (size_t)ctx->smem_optin) {
    if (grid_ok) return 0;  // only the grid kernel (K4) can take it
    return fail(ctx, YALPS_ERR_TOO_LARGE, "tableau %dx%d: pivot row/column staging exceeds shared memory", Hcap, Wcap);
  }
  // CTA shape.  Explicit tuning wins; otherwise (scripts/single_lp_latency.py, scripts/sweep_config3.py on B200):
  //  * latency mode (every LP can have an SM to itself): row-split CTAs, sized by the tableau;
  //  * throughput mode: one row group and many CTAs per SM for tableaus in shared memory, a small row split for
  //    the HBM/L2-resident kernel on mid-size tableaus.
  const long long cells = (long long)Hcap * Wcap;
  int want_threads = ctx->tune_threads, want_rows = ctx->tune_rows;
  if (want_threads <= 0 && want_rows <= 0) {
    if (n <= 2LL * ctx->prop.multiProcessorCount) {
      // latency mode wants the row-split layout (compacted row list next to the tableau): a tableau that only fits
      // shared memory without it is better off on the cluster / HBM-resident split kernels than on one-row-group K1
      if (resident && tune_path == YALPS_PATH_AUTO && SmemLayout(Hcap, Wcap, true, 16, true).total > (size_t)ctx->smem_optin)
        resident = false;
      if (!resident) {
        want_threads = 512, want_rows = 4;
      } else if (cells < 2500) {
        want_threads = 128, want_rows = 4;
      } else if (cells < 4000) {
        want_threads = 256, want_rows = 8;
      } else if (cells < 12000) {
        want_threads = 256, want_rows = 4;
      } else {
        want_threads = 512, want_rows = 8;
      }
    } else if (!resident) {
      // (with the RHS column and the basis in shared memory a pivot of the split kernel waits on L2 less often, and one
      // column warp per row group -- twice the cells per thread and row -- beats two: SC105 78.5 -> 81.8, ADLITTLE
      // 139.8 -> 151.1 M pivots/s, profiles/r02u_sweep_config3.jsonl)
      if (cells >= 5000) want_threads = Wcap <= 129 ? 64 : 128, want_rows = 2;
    }
  }
  plan->small_for_grid = !resident && cells < 40000;
  const int nw = want_threads > 0 ? std::max(1, want_threads / 32) : default_warps(cells, resident);
  int nwr = want_rows > 0 ? want_rows : 1;
  const KernelEntry *k = nullptr;
  if (nwr > 1) {
    k = pick_kernel(std::max(1, nw / nwr), nwr, Wcap, resident);
    if (k && SmemLayout(Hcap, Wcap, resident, k->nw * k->nwr, true).total > (size_t)ctx->smem_optin) k = nullptr;
  }
  if (!k) {
    nwr = 1;
    k = pick_kernel(nw, 1, Wcap, resident);
  }
  if (k) {
    Lr = SmemLayout(Hcap, Wcap, true, k->nw * k->nwr, k->nwr > 1);
    Lg = SmemLayout(Hcap, Wcap, false, k->nw * k->nwr, k->nwr > 1);
  }
  if (k) {  // the attribute / occupancy queries cost microseconds: remember them per (kernel, shared memory size)
    const size_t smem_c = resident ? Lr.total : Lg.total;
    const std::string key = std::to_string((size_t)(resident ? (void *)k->resident : (void *)k->global)) + ":" + std::to_string(smem_c);
    auto it = ctx->occ_cache.find(key);
    if (it != ctx->occ_cache.end()) {
      CU(ctx, raise_smem_limit(ctx->device, (const void *)(resident ? k->resident : k->global), (int)smem_c));
      int occ_c = it->second;
      if (check_cycles) occ_c = std::min(occ_c, 4);
      if (const char *env = getenv("YALPS_CTAS_PER_SM")) occ_c = std::max(1, std::min(occ_c, atoi(env)));  // experiments
      long long grid_c = std::max(1LL, std::min((long long)occ_c * ctx->prop.multiProcessorCount, n));
      plan->resident = resident;
      plan->k = k;
      plan->smem = smem_c;
      plan->grid = (int)grid_c;
      return 0;
    }
  }
  if (!k) {
    if (grid_ok) return 0;
    return fail(ctx, YALPS_ERR_TOO_LARGE, "tableau width %d exceeds the widest kernel", Wcap);
  }
  SimplexKernel fn = resident ? k->resident : k->global;
  const size_t smem = resident ? Lr.total : Lg.total;
  CU(ctx, raise_smem_limit(ctx->device, (const void *)fn, (int)smem));
  int occ = 0;
  CU(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, k->nw * k->nwr * 32, smem));
  if (occ < 1) return fail(ctx, YALPS_ERR_TOO_LARGE, "kernel does not fit on an SM (smem %zu)", smem);
  ctx->occ_cache[std::to_string((size_t)(void *)fn) + ":" + std::to_string(smem)] = occ;
  if (check_cycles) occ = std::min(occ, 4);  // bounds the history buffer
  if (const char *env = getenv("YALPS_CTAS_PER_SM")) occ = std::max(1, std::min(occ, atoi(env)));  // experiments
  long long grid = (long long)occ * ctx->prop.multiProcessorCount;
  grid = std::max(1LL, std::min(grid, n));
  plan->resident = resident;
  plan->k = k;
  plan->smem = smem;
  plan->grid = (int)grid;
  return 0;
}

// K4 (whole grid per LP, LPs one after another) beats K2 (one CTA per LP) when the tableaus do not fit in
// shared memory and there are too few of them to give every SM its own LP.
bool use_grid_path(const yalps_ctx *ctx, long long n, const LaunchPlan &plan) {
  if (plan.tmem) return false;
  if (ctx->tune_path == YALPS_PATH_GRID || plan.k == nullptr) return true;
  if (ctx->tune_path != YALPS_PATH_AUTO && ctx->tune_path != YALPS_PATH_CLUSTER) return false;
  return !plan.resident && n <= 16 && !plan.small_for_grid;
}

int hist_capacity(const yalps_options *opt) {
  if (!opt->check_cycles) return 0;
  double cap = opt->max_pivots;
  if (!(cap < 262144.0)) cap = 262144.0;
  if (cap < 1.0) cap = 1.0;
  return (int)cap;
}

// Enqueue one kernel launch for `args` (device pointers filled in by the caller).
int launch_simplex(yalps_ctx *ctx, const LaunchPlan &plan, BatchArgs &args, const std::string &slot,
                   cudaStream_t stream) {
  if (!args.rows_out && ctx->d_rows && !ctx->rows_per_lp) args.rows_out = ctx->d_rows;  // total mode: every launch of the ctx
  args.counter = nullptr;
  const long long solvers = plan.tmem ? (long long)plan.grid * tmem_kernel_warps() : plan.grid;  // K1t: one LP per warp
  if (args.n > solvers) {  // more LPs than CTAs: dynamic queue (pivot counts vary per LP)
    void *counter = nullptr;
    if (int rc = dev_ensure(ctx, "counter" + slot, 8, &counter)) return rc;
    CU(ctx, cudaMemsetAsync(counter, 0, 8, stream));
    args.counter = (unsigned long long *)counter;
  }
  if (args.check_cycles) {
    void *hist = nullptr;
    const size_t per_cta = (size_t)(plan.tmem ? tmem_kernel_warps() : 1) * 2 * args.hist_cap * sizeof(int);  // K1t: per warp
    if (int rc = dev_ensure(ctx, "hist" + slot, (size_t)plan.grid * per_cta, &hist)) return rc;
    args.hist = (int *)hist;
  } else {
    args.hist = nullptr;
  }
  if (plan.tmem) {
    CU(ctx, launch_simplex_tmem(args, plan.grid, stream));
    ctx->launches++;
    return 0;
  }
  SimplexKernel fn = plan.resident ? plan.k->resident : plan.k->global;
  fn<<<plan.grid, plan.k->nw * plan.k->nwr * 32, plan.smem, stream>>>(args);
  CU(ctx, cudaGetLastError());
  ctx->launches++;
  return 0;
}

// ---- KC: one LP per thread-block cluster (cluster_kernel.cuh) -------------------------------------------
struct ClusterPlan {
  const KernelEntry *k = nullptr;
  int C = 0;          // CTAs per cluster
  size_t smem = 0;
  int clusters = 0;   // co-resident clusters
};

// Smallest cluster (2, 4, 8, 16 CTAs) whose shared memory holds the tableau, and the kernel covering its width.
// plan->k stays null when the tableau does not fit or the device cannot co-schedule such a cluster.
int plan_cluster(yalps_ctx *ctx, int Hcap, int Wcap, ClusterPlan *plan) {
  plan->k = nullptr;
  int count = 0;
  const KernelEntry *tab = kernel_table_cluster(&count);
  const KernelEntry *k = nullptr;
  for (int i = 0; i < count && !k; i++)
    if (tab[i].resident && (long long)tab[i].nw * 32 * tab[i].kc * 2 >= std::max(Wcap - 1, 1)) k = &tab[i];
  if (!k) return 0;
  for (int C = 2; C <= kMaxCluster; C *= 2) {
    const ClusterSmem L(Hcap, Wcap, C);
    if (L.total > (size_t)ctx->smem_optin) continue;
    // more CTAs than the capacity needs when a CTA would otherwise own many rows (dense updates are row-parallel)
    if (L.Hloc > 40 && C < kMaxCluster) continue;
    const std::string key = "cl:" + std::to_string((size_t)(void *)k->resident) + ":" + std::to_string(C) + ":" + std::to_string(L.total);
    int ncl = 0;
    auto it = ctx->occ_cache.find(key);
    if (it != ctx->occ_cache.end()) {
      ncl = it->second;
    } else {
      CU(ctx, raise_smem_limit(ctx->device, (const void *)k->resident, (int)L.total));
      if (C > 8) CU(ctx, cudaFuncSetAttribute((const void *)k->resident, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3((unsigned)C * 64);
      cfg.blockDim = dim3((unsigned)(k->nw * k->nwr * 32));
      cfg.dynamicSmemBytes = L.total;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = (unsigned)C;
      attr[0].val.clusterDim.y = attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      if (cudaOccupancyMaxActiveClusters(&ncl, (const void *)k->resident, &cfg) != cudaSuccess) {
        cudaGetLastError();
        ncl = 0;
      }
      ctx->occ_cache[key] = ncl;
    }
    if (ncl < 1) continue;
    CU(ctx, raise_smem_limit(ctx->device, (const void *)k->resident, (int)L.total));
    plan->k = k;
    plan->C = C;
    plan->smem = L.total;
    plan->clusters = ncl;
    return 0;
  }
  return 0;
}

int launch_cluster(yalps_ctx *ctx, const ClusterPlan &plan, BatchArgs &args, const std::string &slot, cudaStream_t stream) {
  const int ncl = (int)std::max<long long>(1, std::min<long long>(plan.clusters, args.n));
  if (!args.rows_out && ctx->d_rows && !ctx->rows_per_lp) args.rows_out = ctx->d_rows;
  args.counter = nullptr;
  args.hist = nullptr;
  if (args.check_cycles) {
    void *hist = nullptr;
    if (int rc = dev_ensure(ctx, "hist_cl" + slot, (size_t)ncl * plan.C * 2 * args.hist_cap * sizeof(int), &hist)) return rc;
    args.hist = (int *)hist;
  }
  {
    void *scr = nullptr;
    const size_t per_cluster = (size_t)2 * plan.C * SmemLayout::ld_for(args.Wcap) * sizeof(double);
    if (int rc = dev_ensure(ctx, "cl_scratch" + slot, per_cluster * ncl, &scr)) return rc;
    args.cl_scratch = (double *)scr;
  }
  args.tma_mode = ctx->kc_tma;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(ncl * plan.C));
  cfg.blockDim = dim3((unsigned)(plan.k->nw * plan.k->nwr * 32));
  cfg.dynamicSmemBytes = plan.smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)plan.C;
  attr[0].val.clusterDim.y = attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CU(ctx, cudaLaunchKernelEx(&cfg, plan.k->resident, (const BatchArgs)args));
  ctx->launches++;
  return 0;
}

// ---- KG: the cluster kernel with the whole cooperative grid as its "cluster" (tableau resident in the SMs' shared memory)
struct GridResPlan {
  const KernelEntry *k = nullptr;
  int C = 0;  // CTAs (one per SM at most)
  size_t smem = 0;
};

int plan_gridres(yalps_ctx *ctx, int Hcap, int Wcap, GridResPlan *plan) {
  plan->k = nullptr;
  int count = 0, coop = 0;
  const KernelEntry *tab = kernel_table_cluster(&count);
  const KernelEntry *k = nullptr;
  // narrowest kernel that covers the width; among those the one with the most threads (YALPS_KG_THREADS caps them)
  int max_threads = 512;
  if (const char *env = getenv("YALPS_KG_THREADS")) max_threads = atoi(env);
  for (int i = 0; i < count; i++) {
    const long long cap = (long long)tab[i].nw * 32 * tab[i].kc * 2;
    const int threads = tab[i].nw * tab[i].nwr * 32;
    if (!tab[i].global || cap < std::max(Wcap - 1, 1) || threads > max_threads) continue;
    const long long kcap = k ? (long long)k->nw * 32 * k->kc * 2 : 0;
    if (!k || cap < kcap || (cap == kcap && threads > k->nw * k->nwr * 32)) k = &tab[i];
  }
  if (!k) return 0;
  CU(ctx, cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device));
  if (!coop) return 0;
  int C = std::max(1, std::min(std::min(ctx->prop.multiProcessorCount, 160), Hcap - 1));  // (grid_select: 5 slots per lane)
  if (const char *env = getenv("YALPS_KG_CTAS")) C = std::max(1, std::min(C, atoi(env)));
  const ClusterSmem L(Hcap, Wcap, C);
  if (L.total > (size_t)ctx->smem_optin) return 0;
  CU(ctx, raise_smem_limit(ctx->device, (const void *)k->global, (int)L.total));
  const std::string key = "kg:" + std::to_string((size_t)(void *)k->global) + ":" + std::to_string(L.total);
  int occ = 0;
  auto it = ctx->occ_cache.find(key);
  if (it != ctx->occ_cache.end()) {
    occ = it->second;
  } else {
    CU(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void *)k->global, k->nw * k->nwr * 32, L.total));
    ctx->occ_cache[key] = occ;
  }
  if ((long long)occ * ctx->prop.multiProcessorCount < C) return 0;
  plan->k = k;
  plan->C = C;
  plan->smem = L.total;
  return 0;
}

int launch_gridres(yalps_ctx *ctx, const GridResPlan &plan, BatchArgs &args, const std::string &slot, cudaStream_t stream) {
  if (!args.rows_out && ctx->d_rows && !ctx->rows_per_lp) args.rows_out = ctx->d_rows;
  args.hist = nullptr;
  if (args.check_cycles) {
    void *hist = nullptr;
    if (int rc = dev_ensure(ctx, "hist_kg" + slot, (size_t)plan.C * 2 * args.hist_cap * sizeof(int), &hist)) return rc;
    args.hist = (int *)hist;
  }
  void *p = nullptr;
  const size_t ldA = (size_t)SmemLayout::ld_for(args.Wcap);
  // candidate rows as 16-byte flag-in-data slots {low word, sequence, high word, sequence}: [2][C][ldA]
  const size_t pub_bytes = (size_t)2 * plan.C * ldA * sizeof(uint4);
  if (int rc = dev_ensure(ctx, "kg_scratch" + slot, pub_bytes, &p)) return rc;
  args.cl_scratch = (double *)p;
  CU(ctx, cudaMemsetAsync(p, 0, pub_bytes, stream));  // sequence numbers of an earlier launch must not look current
  const size_t inbox_bytes = (size_t)2 * plan.C * plan.C * sizeof(uint4);  // [2][receiver][sender] selection records
  if (int rc = dev_ensure(ctx, "kg_slots" + slot, inbox_bytes + 64, &p)) return rc;
  args.gx_slots = (uint4 *)p;
  args.counter = (unsigned long long *)((char *)p + inbox_bytes);
  // slots: sequence numbers of an earlier launch must not look current; counter: the grid barrier's arrival count
  CU(ctx, cudaMemsetAsync(p, 0, inbox_bytes + 64, stream));
  args.tma_mode = 0;
  void *params[] = {&args};
  CU(ctx, cudaLaunchCooperativeKernel((const void *)plan.k->global, dim3((unsigned)plan.C),
                                      dim3((unsigned)(plan.k->nw * plan.k->nwr * 32)), params, plan.smem, stream));
  ctx->launches++;
  return 0;
}

// The batches K4 would take (few LPs beyond one SM's -- and one cluster's -- shared memory) go to KG when the tableau
// fits the shared memory of the whole grid.  1: launched, 0: not applicable, < 0: error.
int maybe_gridres(yalps_ctx *ctx, long long n, int Hcap, int Wcap, const LaunchPlan *plan, BatchArgs &a, const std::string &slot,
                  cudaStream_t stream) {
  const bool forced = ctx->tune_path == YALPS_PATH_GRID_RESIDENT;
  if (!forced) {
    if (!plan || ctx->tune_path != YALPS_PATH_AUTO || !use_grid_path(ctx, n, *plan)) return 0;
    if (getenv("YALPS_NO_KG")) return 0;
  }
  GridResPlan gp;
  if (int rc = plan_gridres(ctx, Hcap, Wcap, &gp)) return rc;
  if (!gp.k) {
    if (forced)
      return fail(ctx, YALPS_ERR_TOO_LARGE, "tableau %dx%d does not fit the shared memory of the grid (or is wider than 2049)", Hcap, Wcap);
    return 0;
  }
  if (int rc = launch_gridres(ctx, gp, a, slot, stream)) return rc;
  return 1;
}

// Few LPs that do not fit one SM's shared memory: KC when they fit a cluster's (tune_path AUTO or CLUSTER).
bool want_cluster(const yalps_ctx *ctx, long long n, const LaunchPlan &plan, const ClusterPlan &cp) {
  if (!cp.k) return false;
  if (ctx->tune_path == YALPS_PATH_CLUSTER) return true;
  if (ctx->tune_path != YALPS_PATH_AUTO || plan.tmem) return false;
  return !plan.resident && n <= 2LL * std::max(1, cp.clusters);
}

// 1 = launched on the cluster path, 0 = not applicable (caller goes on with its plan), < 0 = error.
// plan == nullptr: before plan_launch, only the forced YALPS_PATH_CLUSTER is considered.
int maybe_cluster(yalps_ctx *ctx, long long n, int Hcap, int Wcap, const LaunchPlan *plan, BatchArgs &a,
                  const std::string &slot, cudaStream_t stream) {
  const bool forced = ctx->tune_path == YALPS_PATH_CLUSTER;
  if (ctx->tune_path == YALPS_PATH_GRID_RESIDENT) return plan ? 0 : maybe_gridres(ctx, n, Hcap, Wcap, nullptr, a, slot, stream);
  if (!plan && !forced) return 0;
  if (plan && (forced || ctx->tune_path != YALPS_PATH_AUTO || plan->resident || plan->tmem)) return 0;
  ClusterPlan cp;
  if (int rc = plan_cluster(ctx, Hcap, Wcap, &cp)) return rc;
  if (forced) {
    if (!cp.k)
      return fail(ctx, YALPS_ERR_TOO_LARGE, "tableau %dx%d does not fit the shared memory of a %d-CTA cluster", Hcap, Wcap, kMaxCluster);
  } else if (!want_cluster(ctx, n, *plan, cp)) {
    return maybe_gridres(ctx, n, Hcap, Wcap, plan, a, slot, stream);  // beyond a cluster: the whole grid's shared memory (KG)
  }
  if (int rc = launch_cluster(ctx, cp, a, slot, stream)) return rc;
  return 1;
}

void fill_options(BatchArgs &a, const yalps_options *opt) {
  a.rows_out = nullptr;
  a.rows_per_lp = 0;
  a.precision = opt->precision;
  a.max_pivots = opt->max_pivots;
  a.check_cycles = opt->check_cycles ? 1 : 0;
  a.hist_cap = hist_capacity(opt);
}

int check_device_status(yalps_ctx *ctx, const int32_t *status, long long n) {
  if (!status) return 0;
  for (long long i = 0; i < n; i++)
    if (status[i] == ST_ERR_HISTORY)
      return fail(ctx, YALPS_ERR_HISTORY, "checkCycles history exhausted for LP %lld (more than 262144 pivots in a phase)", i);
    else if (status[i] == ST_ERR_PEER)
      return fail(ctx, YALPS_ERR_CUDA, "LP %lld: an in-kernel exchange between the CTAs of the grid timed out", i);
  return 0;
}

}  // namespace

// =========================================================================================================
extern "C" {

void yalps_default_options(yalps_options *opt) {
  opt->precision = 1e-8;
  opt->max_pivots = 8192;
  opt->tolerance = 0;
  opt->timeout_ms = std::numeric_limits<double>::infinity();
  opt->max_iterations = 32768;
  opt->check_cycles = 0;
  opt->reserved = 0;
}

int yalps_create(int device, yalps_ctx **out) {
  if (!out) return fail(nullptr, YALPS_ERR_ARGUMENT, "out is null");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(nullptr, YALPS_ERR_CUDA, "no CUDA device available (%s); yalps_b200 has no CPU fallback",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
  if (device < 0 || device >= count) return fail(nullptr, YALPS_ERR_ARGUMENT, "device %d out of range [0,%d)", device, count);
  std::unique_ptr<yalps_ctx> ctx(new yalps_ctx);
  ctx->device = device;
  if ((e = cudaSetDevice(device)) != cudaSuccess) return fail(nullptr, YALPS_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
  if ((e = cudaGetDeviceProperties(&ctx->prop, device)) != cudaSuccess)
    return fail(nullptr, YALPS_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (ctx->prop.major < 10)
    return fail(nullptr, YALPS_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                ctx->prop.major, ctx->prop.minor);
  ctx->smem_optin = (int)ctx->prop.sharedMemPerBlockOptin;
  if (const char *env = getenv("YALPS_KC_TMA")) ctx->kc_tma = std::max(0, std::min(2, atoi(env)));
  for (int i = 0; i < 2; i++) {
    if ((e = cudaStreamCreateWithFlags(&ctx->streams[i], cudaStreamNonBlocking)) != cudaSuccess)
      return fail(nullptr, YALPS_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
    if ((e = cudaEventCreateWithFlags(&ctx->events[i], cudaEventDisableTiming)) != cudaSuccess)
      return fail(nullptr, YALPS_ERR_CUDA, "cudaEventCreate: %s", cudaGetErrorString(e));
  }
  *out = ctx.release();
  return 0;
}

void yalps_destroy(yalps_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  for (auto &kv : ctx->pool)
    if (kv.second.p) cudaFree(kv.second.p);
  for (auto &kv : ctx->pinned)
    if (kv.second.p) cudaFreeHost(kv.second.p);
  for (DevBuf *b : {&ctx->root.m, &ctx->root.pos, &ctx->root.var})
    if (b->p) cudaFree(b->p);
  for (int i = 0; i < 2; i++) {
    if (ctx->streams[i]) cudaStreamDestroy(ctx->streams[i]);
    if (ctx->events[i]) cudaEventDestroy(ctx->events[i]);
  }
  for (cudaStream_t st : ctx->aux_streams) cudaStreamDestroy(st);
  for (cudaEvent_t ev : ctx->aux_events) cudaEventDestroy(ev);
  if (ctx->fork_event) cudaEventDestroy(ctx->fork_event);
  delete ctx;
}

const char *yalps_last_error(const yalps_ctx *ctx) { return ctx ? ctx->error.c_str() : g_create_error.c_str(); }

int yalps_device_info(const yalps_ctx *ctx, int32_t *sm_count, int32_t *smem_per_block_optin, int32_t *cc_major,
                      int32_t *cc_minor) {
  if (!ctx) return YALPS_ERR_ARGUMENT;
  if (sm_count) *sm_count = ctx->prop.multiProcessorCount;
  if (smem_per_block_optin) *smem_per_block_optin = ctx->smem_optin;
  if (cc_major) *cc_major = ctx->prop.major;
  if (cc_minor) *cc_minor = ctx->prop.minor;
  return 0;
}

int yalps_set_tuning(yalps_ctx *ctx, int32_t path, int32_t threads_per_lp) {
  if (!ctx) return YALPS_ERR_ARGUMENT;
  if (path < 0 || path > YALPS_PATH_GRID_RESIDENT || path == 4) return fail(ctx, YALPS_ERR_ARGUMENT, "bad path %d", path);
  ctx->tune_path = path;
  ctx->tune_threads = threads_per_lp;
  return 0;
}

int yalps_set_row_groups(yalps_ctx *ctx, int32_t row_groups) {
  if (!ctx) return YALPS_ERR_ARGUMENT;
  if (row_groups < 0 || row_groups > 16 || (row_groups & (row_groups - 1)))
    return fail(ctx, YALPS_ERR_ARGUMENT, "row_groups must be 0 (automatic), 1, 2, 4, 8 or 16");
  ctx->tune_rows = row_groups;
  return 0;
}

int64_t yalps_launch_count(const yalps_ctx *ctx) { return ctx ? ctx->launches : 0; }

int yalps_set_row_counter(yalps_ctx *ctx, uint64_t *d_rows, int32_t per_lp) {
  if (!ctx) return YALPS_ERR_ARGUMENT;
  ctx->d_rows = (unsigned long long *)d_rows;
  ctx->rows_per_lp = per_lp ? 1 : 0;
  return 0;
}

int yalps_host_alloc(yalps_ctx *ctx, uint64_t bytes, void **out) {
  if (!ctx || !out) return YALPS_ERR_ARGUMENT;
  CU(ctx, cudaSetDevice(ctx->device));
  CU(ctx, cudaMallocHost(out, bytes ? bytes : 1));
  return 0;
}

int yalps_host_free(yalps_ctx *ctx, void *ptr) {
  if (!ctx) return YALPS_ERR_ARGUMENT;
  if (ptr) CU(ctx, cudaFreeHost(ptr));
  return 0;
}

// K4: one LP over the whole grid.  d_M (H*W doubles, reference layout) is solved in place.
static int launch_grid(yalps_ctx *ctx, int H, int W, double *d_M, const yalps_options *opt, int *d_status,
                       double *d_value, long long *d_pivots, double *d_rhs, int *d_pos, int *d_var,
                       cudaStream_t stream, const int *d_init_var = nullptr, int init_n = 0, int max_ctas = 0,
                       const std::string &slot = "", unsigned long long *d_rows = nullptr) {
  const GridSmem L(H, W);
  if (L.total > (size_t)ctx->smem_optin)
    return fail(ctx, YALPS_ERR_TOO_LARGE, "tableau %dx%d: pivot row/column staging exceeds shared memory", H, W);
  int coop = 0;
  CU(ctx, cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device));
  if (!coop) return fail(ctx, YALPS_ERR_CUDA, "device does not support cooperative launches");
  CU(ctx, raise_smem_limit(ctx->device, (const void *)k_simplex_grid, (int)L.total));
  int occ = 0;
  CU(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_simplex_grid, kGridThreads, L.total));
  if (occ < 1) return fail(ctx, YALPS_ERR_TOO_LARGE, "grid kernel does not fit on an SM");
  // at most ~2 update items (row x 256-column segment) per warp, never more CTAs than can be co-resident
  int grid = ctx->prop.multiProcessorCount;
  {
    const long long items = (long long)H * ((W + kSegCols - 1) / kSegCols);
    grid = (int)std::min<long long>(grid, std::max(1LL, (items + 2 * kGridWarps - 1) / (2 * kGridWarps)));
  }
  if (const char *env = getenv("YALPS_GRID_CTAS")) grid = std::max(1, std::min(atoi(env), ctx->prop.multiProcessorCount * occ));
  if (max_ctas > 0) grid = std::max(1, std::min(grid, max_ctas));  // several K4 launches sharing the GPU (node waves)
  GridArgs a{};
  a.M = d_M;
  a.H = H;
  a.W = W;
  a.status = d_status;
  a.value = d_value;
  a.pivots = d_pivots;
  a.rhs_out = d_rhs;
  a.pos_out = d_pos;
  a.precision = opt->precision;
  a.max_pivots = opt->max_pivots;
  a.check_cycles = opt->check_cycles ? 1 : 0;
  a.hist_cap = hist_capacity(opt);
  void *p = nullptr;
  if (!d_var) {
    if (int rc = dev_ensure(ctx, "grid_var" + slot, (size_t)(W + H) * 4, &p)) return rc;
    d_var = (int *)p;
  }
  a.var = d_var;
  a.init_var = d_init_var;
  a.init_n = init_n;
  a.rows_out = d_rows ? d_rows : ((ctx->d_rows && !ctx->rows_per_lp) ? ctx->d_rows : nullptr);
  if (int rc = dev_ensure(ctx, "grid_flags" + slot, 64, &p)) return rc;
  a.flags = (int *)p;
  a.barrier = (unsigned long long *)((char *)p + 32);
  CU(ctx, cudaMemsetAsync(p, 0, 64, stream));
  if (a.check_cycles) {
    if (int rc = dev_ensure(ctx, "grid_hist" + slot, (size_t)2 * a.hist_cap * sizeof(int), &p)) return rc;
    a.hist = (int *)p;
  }
  void *params[] = {&a};
  CU(ctx, cudaLaunchCooperativeKernel((void *)k_simplex_grid, dim3(grid), dim3(kGridThreads), params, L.total, stream));
  ctx->launches++;
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// `slot` names the pooled scratch (LP queue counter, cycle history, cluster scratch) of this launch: calls that may be in
// flight at the same time on different streams of one ctx must use different slots.
static int solve_batch_device_impl(yalps_ctx *ctx, int64_t n, int32_t height, int32_t width, const double *d_matrices,
                                   double *d_work, const yalps_options *opt, int32_t *d_status, double *d_value,
                                   int64_t *d_pivots, double *d_rhs_out, int32_t *d_pos_out, int32_t *d_var_out,
                                   double *d_matrices_out, void *stream, const std::string &slot, int64_t rows_base = 0) {
  if (!ctx) return YALPS_ERR_ARGUMENT;
  if (n < 0 || height < 1 || width < 1 || !opt || (n > 0 && !d_matrices))
    return fail(ctx, YALPS_ERR_ARGUMENT, "bad arguments (n=%lld, %dx%d)", (long long)n, height, width);
  if ((long long)height * width >= (1LL << 31)) return fail(ctx, YALPS_ERR_TOO_LARGE, "height*width must be < 2^31");
  if (n == 0) return 0;
  CU(ctx, cudaSetDevice(ctx->device));
  // density probe for the path policy, once per (buffer, shape): synchronises `stream` the first time only
  double density = -1.0;
  if (ctx->tune_path == YALPS_PATH_AUTO && n > 64 && d_work) {
    const std::string key = std::to_string((size_t)d_matrices) + ":" + std::to_string(height) + "x" + std::to_string(width);
    auto it = ctx->density_cache.find(key);
    if (it == ctx->density_cache.end()) {
      void *hp, *dp;
      if (int rc = pin_ensure(ctx, "density_probe", 64, &hp)) return rc;
      if (int rc = pin_device_ptr(ctx, hp, &dp)) return rc;
      ((int *)hp)[0] = ((int *)hp)[1] = 0;
      const long long cells = (long long)height * width;
      k_sample_density<<<1, 256, 0, (cudaStream_t)stream>>>(d_matrices, cells, std::max(1LL, cells / 4096), (int *)dp);
      CU(ctx, cudaGetLastError());
      CU(ctx, cudaStreamSynchronize((cudaStream_t)stream));
      ctx->launches++;
      const int seen = ((int *)hp)[0], nz = ((int *)hp)[1];
      density = seen ? (double)nz / seen : 1.0;
      if (ctx->density_cache.size() > 64) ctx->density_cache.clear();
      ctx->density_cache[key] = density;
    } else {
      density = it->second;
    }
  }
  BatchArgs a{};
  a.n = n;
  a.mode = kModeBatch;
  a.H = a.Hcap = height;
  a.W = a.Wcap = width;
  a.in = d_matrices;
  a.work = d_work;
  a.mat_out = d_matrices_out;
  a.status = d_status;
  a.value = d_value;
  a.pivots = (long long *)d_pivots;
  a.rhs_out = d_rhs_out;
  a.pos_out = d_pos_out;
  a.var_out = d_var_out;
  fill_options(a, opt);
  a.rows_out = ctx->d_rows ? ctx->d_rows + (ctx->rows_per_lp ? rows_base : 0) : nullptr;
  a.rows_per_lp = ctx->rows_per_lp;
  if (int rc = maybe_cluster(ctx, n, height, width, nullptr, a, slot, (cudaStream_t)stream)) return rc < 0 ? rc : 0;
  LaunchPlan plan;
  if (int rc = plan_launch(ctx, n, height, width, opt->check_cycles != 0, &plan, density, true)) return rc;
  if (int rc = maybe_cluster(ctx, n, height, width, &plan, a, slot, (cudaStream_t)stream)) return rc < 0 ? rc : 0;
  if (use_grid_path(ctx, n, plan)) {
    if (!d_work) return fail(ctx, YALPS_ERR_ARGUMENT, "d_work is required for the grid path");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t cells = (size_t)height * width;
    if (d_work != d_matrices)
      CU(ctx, cudaMemcpyAsync(d_work, d_matrices, (size_t)n * cells * 8, cudaMemcpyDeviceToDevice, st));
    for (int64_t i = 0; i < n; i++) {
      if (int rc = launch_grid(ctx, height, width, d_work + i * cells, opt, d_status ? d_status + i : nullptr,
                               d_value ? d_value + i : nullptr, d_pivots ? (long long *)d_pivots + 2 * i : nullptr,
                               d_rhs_out ? d_rhs_out + i * height : nullptr,
                               d_pos_out ? d_pos_out + i * (width + height) : nullptr,
                               d_var_out ? d_var_out + i * (width + height) : nullptr, st, nullptr, 0, 0, slot,
                               ctx->d_rows ? ctx->d_rows + (ctx->rows_per_lp ? rows_base + i : 0) : nullptr))
        return rc;
    }
    if (d_matrices_out && d_matrices_out != d_work)
      CU(ctx, cudaMemcpyAsync(d_matrices_out, d_work, (size_t)n * cells * 8, cudaMemcpyDeviceToDevice, st));
    return 0;
  }
  if (!plan.resident) {
    if (!d_work) return fail(ctx, YALPS_ERR_ARGUMENT, "d_work is required for the HBM-resident path");
    if (!d_pos_out || !d_var_out) {
      void *p = nullptr;
      if (int rc = dev_ensure(ctx, "posvar_scratch" + slot, (size_t)n * (width + height) * 2 * sizeof(int), &p)) return rc;
      if (!a.pos_out) a.pos_out = (int *)p;
      if (!a.var_out) a.var_out = (int *)p + (size_t)n * (width + height);
    }
  }
  return launch_simplex(ctx, plan, a, slot, (cudaStream_t)stream);
}

int yalps_solve_batch_device(yalps_ctx *ctx, int64_t n, int32_t height, int32_t width, const double *d_matrices,
                             double *d_work, const yalps_options *opt, int32_t *d_status, double *d_value,
                             int64_t *d_pivots, double *d_rhs_out, int32_t *d_pos_out, int32_t *d_var_out,
                             double *d_matrices_out, void *stream) {
  return solve_batch_device_impl(ctx, n, height, width, d_matrices, d_work, opt, d_status, d_value, d_pivots, d_rhs_out,
                                 d_pos_out, d_var_out, d_matrices_out, stream, "dev");
}

// Shared host-side pipeline for uniform and ragged batches.
static int solve_host(yalps_ctx *ctx, int64_t n, int32_t height, int32_t width, const int32_t *heights,
                      const int32_t *widths, const int64_t *mat_offsets, const double *matrices,
                      const yalps_options *opt, int32_t *status, double *value, int64_t *pivots, double *rhs_out,
                      int32_t *pos_out, int32_t *var_out, double *matrices_out, const int32_t *pos_in = nullptr,
                      const int32_t *var_in = nullptr) {
  if (!ctx) return YALPS_ERR_ARGUMENT;
  if (n < 0 || !opt || (n > 0 && !matrices)) return fail(ctx, YALPS_ERR_ARGUMENT, "bad arguments");
  if ((pos_in == nullptr) != (var_in == nullptr))
    return fail(ctx, YALPS_ERR_ARGUMENT, "pos_in and var_in must be given together (or both null for the identity)");
  if (n == 0) return 0;
  CU(ctx, cudaSetDevice(ctx->device));
  const bool ragged = heights != nullptr;

  // per-LP prefix offsets (ragged) and caps
  std::vector<long long> moff, roff, poff;
  int Hcap = height, Wcap = width;
  if (ragged) {
    if (!widths || !mat_offsets) return fail(ctx, YALPS_ERR_ARGUMENT, "ragged batch needs widths and mat_offsets");
    moff.resize(n + 1);
    roff.resize(n + 1);
    poff.resize(n + 1);
    long long r = 0, p = 0, m = 0;
    Hcap = Wcap = 1;
    for (int64_t i = 0; i < n; i++) {
      if (heights[i] < 1 || widths[i] < 1) return fail(ctx, YALPS_ERR_ARGUMENT, "LP %lld has empty shape", (long long)i);
      if ((long long)heights[i] * widths[i] >= (1LL << 31)) return fail(ctx, YALPS_ERR_TOO_LARGE, "height*width must be < 2^31");
      moff[i] = m;
      roff[i] = r;
      poff[i] = p;
      m += (long long)heights[i] * widths[i];
      r += heights[i];
      p += heights[i] + widths[i];
      Hcap = std::max(Hcap, heights[i]);
      Wcap = std::max(Wcap, widths[i]);
    }
    moff[n] = m;
    roff[n] = r;
    poff[n] = p;
  } else {
    if (height < 1 || width < 1) return fail(ctx, YALPS_ERR_ARGUMENT, "bad shape %dx%d", height, width);
    if ((long long)height * width >= (1LL << 31)) return fail(ctx, YALPS_ERR_TOO_LARGE, "height*width must be < 2^31");
  }

  auto cells_upto = [&](int64_t i) -> long long { return ragged ? moff[i] : (long long)i * height * width; };
  auto rows_upto = [&](int64_t i) -> long long { return ragged ? roff[i] : (long long)i * height; };
  auto pv_upto = [&](int64_t i) -> long long { return ragged ? poff[i] : (long long)i * (height + width); };

  // caller-supplied basis: the two arrays must be inverse permutations of 0..W+H-1 (src/tableau.ts:9-15); the kernels
  // carry variableAtPosition and rebuild positionOfVariable on output, so an inconsistent pair would silently
  // change meaning instead of behaving like the reference
  if (var_in) {
    for (int64_t i = 0; i < n; i++) {
      const long long o = pv_upto(i), len = pv_upto(i + 1) - o;
      for (long long k = 0; k < len; k++) {
        const int32_t v = var_in[o + k];
        if (v < 0 || v >= len || pos_in[o + v] != (int32_t)k)
          return fail(ctx, YALPS_ERR_ARGUMENT, "LP %lld: pos_in/var_in are not inverse permutations at position %lld",
                      (long long)i, k);
      }
    }
  }

  // ---- small calls (a single LP, a handful of models): latency-bound.  Stage through one mapped pinned buffer,
  // let the kernel read the tableaus and write the results zero-copy: launch + synchronise, no memcpy calls.
  {
    const size_t in_b = (size_t)cells_upto(n) * 8, rows_b = (size_t)rows_upto(n) * 8, pv_b = (size_t)pv_upto(n) * 4;
    const size_t out_b = (size_t)n * 32 + rows_b + 2 * pv_b + 64 + (matrices_out ? in_b : 0);
    const size_t desc_b = (ragged ? (size_t)n * 32 + 64 : 0) + (var_in ? pv_b + 16 : 0);
    LaunchPlan plan;
    // density of (a sample of) the first tableau: the latency-mode policy distinguishes dense from very sparse LPs
    double small_density = -1.0;
    if (ctx->coo_nnz < 0 && n > 0 && in_b + out_b + desc_b <= ((size_t)768 << 10)) {
      const double *m0 = matrices + (ragged ? mat_offsets[0] : 0);
      const long long c0 = ragged ? (long long)heights[0] * widths[0] : (long long)height * width;
      const long long step = std::max(1LL, c0 / 4096);
      long long seen = 0, nz = 0;
      for (long long k = 0; k < c0; k += step, seen++) nz += m0[k] != 0.0;
      small_density = seen ? (double)nz / (double)seen : 1.0;
    }
    if (!ctx->keep_final && ctx->coo_nnz < 0 && in_b + out_b + desc_b <= ((size_t)768 << 10) && ctx->tune_path != YALPS_PATH_GRID && ctx->tune_path != YALPS_PATH_CLUSTER && ctx->tune_path != YALPS_PATH_GRID_RESIDENT &&
        plan_launch(ctx, n, Hcap, Wcap, opt->check_cycles != 0, &plan, small_density, true) == 0 && (plan.k || plan.tmem) && plan.resident) {
      auto up16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
      size_t o = 0;
      const size_t o_in = o; o += up16(in_b);
      const size_t o_h = o; o += ragged ? up16((size_t)n * 4) : 0;
      const size_t o_w = o; o += ragged ? up16((size_t)n * 4) : 0;
      const size_t o_mo = o; o += ragged ? up16((size_t)n * 8) : 0;
      const size_t o_ro = o; o += ragged ? up16((size_t)n * 8) : 0;
      const size_t o_po = o; o += ragged ? up16((size_t)n * 8) : 0;
      const size_t o_vin = o; o += var_in ? up16(pv_b) : 0;
      const size_t o_st = o; o += up16((size_t)n * 4);
      const size_t o_val = o; o += up16((size_t)n * 8);
      const size_t o_piv = o; o += up16((size_t)n * 16);
      const size_t o_rhs = o; o += up16(rows_b);
      const size_t o_pos = o; o += up16(pv_b);
      const size_t o_var = o; o += up16(pv_b);
      const size_t o_mat = o; o += matrices_out ? up16(in_b) : 0;
      void *hb, *db;
      int rc;
      if ((rc = pin_ensure(ctx, "small_io", o + 16, &hb))) return rc;
      if ((rc = pin_device_ptr(ctx, hb, &db))) return rc;
      char *h = (char *)hb, *d = (char *)db;
      if (ragged) {
        for (int64_t i = 0; i < n; i++)
          std::memcpy(h + o_in + (size_t)moff[i] * 8, matrices + mat_offsets[i], (size_t)heights[i] * widths[i] * 8);
        std::memcpy(h + o_h, heights, (size_t)n * 4);
        std::memcpy(h + o_w, widths, (size_t)n * 4);
        std::memcpy(h + o_mo, moff.data(), (size_t)n * 8);
        std::memcpy(h + o_ro, roff.data(), (size_t)n * 8);
        std::memcpy(h + o_po, poff.data(), (size_t)n * 8);
      } else {
        std::memcpy(h + o_in, matrices, in_b);
      }
      if (var_in) std::memcpy(h + o_vin, var_in, pv_b);
      BatchArgs a{};
      a.n = n;
      a.mode = kModeBatch;
      a.H = height;
      a.W = width;
      a.Hcap = Hcap;
      a.Wcap = Wcap;
      a.in = (const double *)(d + o_in);
      a.work = nullptr;
      a.mat_out = matrices_out ? (double *)(d + o_mat) : nullptr;
      a.status = (int *)(d + o_st);
      a.value = (double *)(d + o_val);
      a.pivots = (long long *)(d + o_piv);
      a.rhs_out = (double *)(d + o_rhs);
      a.pos_out = (int *)(d + o_pos);
      a.var_out = (int *)(d + o_var);
      a.var_in = var_in ? (const int *)(d + o_vin) : nullptr;
      if (ragged) {
        a.heights = (const int *)(d + o_h);
        a.widths = (const int *)(d + o_w);
        a.mat_off = (const long long *)(d + o_mo);
        a.rhs_off = (const long long *)(d + o_ro);
        a.pos_off = (const long long *)(d + o_po);
      }
      fill_options(a, opt);
      cudaStream_t st = ctx->streams[0];
      if ((rc = launch_simplex(ctx, plan, a, "small", st))) return rc;
      CU(ctx, cudaStreamSynchronize(st));
      if (status) std::memcpy(status, h + o_st, (size_t)n * 4);
      if (value) std::memcpy(value, h + o_val, (size_t)n * 8);
      if (pivots) std::memcpy(pivots, h + o_piv, (size_t)n * 16);
      if (rhs_out) std::memcpy(rhs_out, h + o_rhs, rows_b);
      if (pos_out) std::memcpy(pos_out, h + o_pos, pv_b);
      if (var_out) std::memcpy(var_out, h + o_var, pv_b);
      if (matrices_out) {
        if (ragged)
          for (int64_t i = 0; i < n; i++)
            std::memcpy(matrices_out + mat_offsets[i], h + o_mat + (size_t)moff[i] * 8, (size_t)heights[i] * widths[i] * 8);
        else
          std::memcpy(matrices_out, h + o_mat, in_b);
      }
      return check_device_status(ctx, status, n);
    }
  }

  // chunking: two pipeline slots; a large batch is cut into ~8 chunks (64 MiB .. 1.5 GiB each) so that the
  // H2D copy of chunk k+1 overlaps the kernel and the D2H copies of chunk k
  const size_t total_bytes = (size_t)cells_upto(n) * 8;
  const size_t kSlotBytes = std::min((size_t)1536 << 20, std::max((size_t)64 << 20, total_bytes / 8 + 4096));

  int64_t begin = 0;
  int slot = 0;
  int rc = 0;
  double density = -1.0;
  bool used[2] = {false, false};
  while (begin < n) {
    int64_t end = begin + 1;
    while (end < n && (size_t)(cells_upto(end + 1) - cells_upto(begin)) * 8 <= kSlotBytes) end++;
    const int64_t cn = end - begin;
    const size_t ccells = (size_t)(cells_upto(end) - cells_upto(begin));
    const size_t crows = (size_t)(rows_upto(end) - rows_upto(begin));
    const size_t cpv = (size_t)(pv_upto(end) - pv_upto(begin));
    const std::string s = std::to_string(slot);
    cudaStream_t st = ctx->streams[slot];
    if (used[slot]) CU(ctx, cudaEventSynchronize(ctx->events[slot]));

    int chcap = Hcap, cwcap = Wcap;
    if (ragged) {
      chcap = cwcap = 1;
      for (int64_t i = begin; i < end; i++) {
        chcap = std::max(chcap, heights[i]);
        cwcap = std::max(cwcap, widths[i]);
      }
    }
    LaunchPlan plan;
    const bool from_cells = ctx->coo_nnz >= 0 && n == 1 && !ragged;
    if (from_cells) density = (double)ctx->coo_nnz / (double)((long long)height * width);
    if (density < 0.0) {  // sample the first tableau once
      const double *m0 = matrices + (ragged ? mat_offsets[0] : 0);
      const long long c0 = ragged ? (long long)heights[0] * widths[0] : (long long)height * width;
      const long long step = std::max(1LL, c0 / 4096);
      long long seen = 0, nz = 0;
      for (long long k = 0; k < c0; k += step, seen++) nz += m0[k] != 0.0;
      density = seen ? (double)nz / (double)seen : 1.0;
    }
    if ((rc = plan_launch(ctx, cn, chcap, cwcap, opt->check_cycles != 0, &plan, density, true))) return rc;

    void *d_in, *d_status, *d_value, *d_piv, *d_rhs, *d_pos, *d_var;
    if ((rc = dev_ensure(ctx, "in" + s, ccells * 8, &d_in))) return rc;
    if ((rc = dev_ensure(ctx, "status" + s, (size_t)cn * 4, &d_status))) return rc;
    if ((rc = dev_ensure(ctx, "value" + s, (size_t)cn * 8, &d_value))) return rc;
    if ((rc = dev_ensure(ctx, "pivots" + s, (size_t)cn * 16, &d_piv))) return rc;
    if ((rc = dev_ensure(ctx, "rhs" + s, crows * 8, &d_rhs))) return rc;
    if ((rc = dev_ensure(ctx, "pos" + s, cpv * 4, &d_pos))) return rc;
    if ((rc = dev_ensure(ctx, "var" + s, cpv * 4, &d_var))) return rc;
    void *d_vin = nullptr;
    if (var_in) {
      if ((rc = dev_ensure(ctx, "varin" + s, cpv * 4, &d_vin))) return rc;
      CU(ctx, cudaMemcpyAsync(d_vin, var_in + pv_upto(begin), cpv * 4, cudaMemcpyHostToDevice, st));
    }
    void *d_out = nullptr;
    const bool want_mat = matrices_out != nullptr || (ctx->keep_final && n == 1 && !ragged);
    if (want_mat && (plan.resident || ragged))
      if ((rc = dev_ensure(ctx, "matout" + s, ccells * 8, &d_out))) return rc;

    const double *src = matrices + (ragged ? mat_offsets[begin] : cells_upto(begin));
    if (ragged) {
      // caller offsets may be non-contiguous: copy LP by LP when they are
      bool contiguous = true;
      for (int64_t i = begin; i < end && contiguous; i++)
        contiguous = (mat_offsets[i] - mat_offsets[begin]) == (moff[i] - moff[begin]);
      if (contiguous) {
        CU(ctx, cudaMemcpyAsync(d_in, src, ccells * 8, cudaMemcpyHostToDevice, st));
      } else {
        for (int64_t i = begin; i < end; i++)
          CU(ctx, cudaMemcpyAsync((double *)d_in + (moff[i] - moff[begin]), matrices + mat_offsets[i],
                                  (size_t)heights[i] * widths[i] * 8, cudaMemcpyHostToDevice, st));
      }
    } else if (from_cells) {
      // the (cell, value) pairs cross PCIe, the zeros are written by the device
      const size_t nnz = (size_t)ctx->coo_nnz, val_off = (nnz * 4 + 15) & ~(size_t)15;
      void *h_coo, *d_coo, *d_flag;
      if ((rc = pin_ensure(ctx, "coo", val_off + nnz * 8 + 16, &h_coo))) return rc;
      if ((rc = dev_ensure(ctx, "coo", val_off + nnz * 8 + 16, &d_coo))) return rc;
      if ((rc = dev_ensure(ctx, "coo_flag", 4, &d_flag))) return rc;
      if (nnz) {
        std::memcpy(h_coo, ctx->coo_cell, nnz * 4);
        std::memcpy((char *)h_coo + val_off, ctx->coo_val, nnz * 8);
        CU(ctx, cudaMemcpyAsync(d_coo, h_coo, val_off + nnz * 8, cudaMemcpyHostToDevice, st));
      }
      CU(ctx, cudaMemsetAsync(d_in, 0, ccells * 8, st));
      CU(ctx, cudaMemsetAsync(d_flag, 0, 4, st));
      if (nnz) {
        const int blocks = (int)std::min<size_t>((nnz + 255) / 256, (size_t)ctx->prop.multiProcessorCount * 8);
        const int *d_cell = (const int *)d_coo;
        const double *d_val = (const double *)((char *)d_coo + val_off);
        k_scatter_cells<<<blocks, 256, 0, st>>>((long long)nnz, d_cell, d_val, (double *)d_in);
        k_verify_cells<<<blocks, 256, 0, st>>>((long long)nnz, d_cell, d_val, (const double *)d_in, (int *)d_flag);
        ctx->launches += 2;
        CU(ctx, cudaGetLastError());
      }
      CU(ctx, cudaMemcpyAsync(&ctx->coo_dup, d_flag, 4, cudaMemcpyDeviceToHost, st));
    } else {
      CU(ctx, cudaMemcpyAsync(d_in, src, ccells * 8, cudaMemcpyHostToDevice, st));
    }

    BatchArgs a{};
    a.n = cn;
    a.mode = kModeBatch;
    a.H = height;
    a.W = width;
    a.Hcap = chcap;
    a.Wcap = cwcap;
    a.in = (const double *)d_in;
    a.work = (double *)d_in;  // K2 works in place on the device copy
    a.mat_out = want_mat ? (plan.resident ? (double *)d_out : (double *)d_in) : nullptr;
    a.status = (int *)d_status;
    a.value = (double *)d_value;
    a.pivots = (long long *)d_piv;
    a.rhs_out = (double *)d_rhs;
    a.pos_out = (int *)d_pos;
    a.var_out = (int *)d_var;
    a.var_in = (const int *)d_vin;
    if (ragged) {
      void *d_h, *d_w, *d_mo, *d_ro, *d_po;
      std::vector<long long> lm(cn), lr(cn), lpv(cn);
      for (int64_t i = 0; i < cn; i++) {
        lm[i] = moff[begin + i] - moff[begin];
        lr[i] = roff[begin + i] - roff[begin];
        lpv[i] = poff[begin + i] - poff[begin];
      }
      if ((rc = dev_ensure(ctx, "rg_h" + s, (size_t)cn * 4, &d_h))) return rc;
      if ((rc = dev_ensure(ctx, "rg_w" + s, (size_t)cn * 4, &d_w))) return rc;
      if ((rc = dev_ensure(ctx, "rg_mo" + s, (size_t)cn * 8, &d_mo))) return rc;
      if ((rc = dev_ensure(ctx, "rg_ro" + s, (size_t)cn * 8, &d_ro))) return rc;
      if ((rc = dev_ensure(ctx, "rg_po" + s, (size_t)cn * 8, &d_po))) return rc;
      CU(ctx, cudaMemcpyAsync(d_h, heights + begin, (size_t)cn * 4, cudaMemcpyHostToDevice, st));
      CU(ctx, cudaMemcpyAsync(d_w, widths + begin, (size_t)cn * 4, cudaMemcpyHostToDevice, st));
      CU(ctx, cudaMemcpyAsync(d_mo, lm.data(), (size_t)cn * 8, cudaMemcpyHostToDevice, st));
      CU(ctx, cudaMemcpyAsync(d_ro, lr.data(), (size_t)cn * 8, cudaMemcpyHostToDevice, st));
      CU(ctx, cudaMemcpyAsync(d_po, lpv.data(), (size_t)cn * 8, cudaMemcpyHostToDevice, st));
      CU(ctx, cudaStreamSynchronize(st));  // the staging vectors die at the end of this scope
      a.heights = (const int *)d_h;
      a.widths = (const int *)d_w;
      a.mat_off = (const long long *)d_mo;
      a.rhs_off = (const long long *)d_ro;
      a.pos_off = (const long long *)d_po;
    }
    fill_options(a, opt);
    int on_cluster = 0;
    if (!ragged) {
      if ((on_cluster = maybe_cluster(ctx, cn, chcap, cwcap, nullptr, a, s, st)) < 0) return on_cluster;
      if (!on_cluster && (on_cluster = maybe_cluster(ctx, cn, chcap, cwcap, &plan, a, s, st)) < 0) return on_cluster;
    }
    if (on_cluster) {
      // launched: one LP per thread-block cluster
    } else if (!ragged) {
      if (use_grid_path(ctx, cn, plan)) {
        for (int64_t i = 0; i < cn; i++) {
          const long long mo = (long long)i * height * width;
          if ((rc = launch_grid(ctx, height, width, (double *)d_in + mo, opt, (int *)d_status + i, (double *)d_value + i,
                                (long long *)d_piv + 2 * i, (double *)d_rhs + (size_t)i * height,
                                (int *)d_pos + (size_t)i * (width + height), (int *)d_var + (size_t)i * (width + height),
                                st, d_vin ? (const int *)d_vin + (size_t)i * (width + height) : nullptr,
                                d_vin ? width + height : 0)))
            return rc;
        }
        a.mat_out = want_mat ? (double *)d_in : nullptr;
      } else if ((rc = launch_simplex(ctx, plan, a, s, st))) {
        return rc;
      }
    } else {
      // Ragged chunk: LPs of very different sizes must not share one launch configuration (one large model would
      // push a thousand small ones onto the HBM-resident kernel).  Group by (fits in shared memory, CTA width) and
      // launch each group with its own plan through an index indirection.
      if (matrices_out) a.mat_out = (double *)d_out;
      std::vector<std::vector<int>> groups;
      std::vector<std::pair<int, int>> caps;  // (Hcap, Wcap) per group
      std::unordered_map<int, int> group_of;
      for (int64_t i = 0; i < cn; i++) {
        const int hi = heights[begin + i], wi = widths[begin + i];
        const bool res = SmemLayout(hi, wi, true, 32).total <= (size_t)ctx->smem_optin;
        const int key = (res ? 0 : 64) + default_warps((long long)hi * wi, res);
        auto it = group_of.find(key);
        if (it == group_of.end()) {
          it = group_of.emplace(key, (int)groups.size()).first;
          groups.emplace_back();
          caps.emplace_back(1, 1);
        }
        groups[it->second].push_back((int)i);
        caps[it->second].first = std::max(caps[it->second].first, hi);
        caps[it->second].second = std::max(caps[it->second].second, wi);
      }
      std::vector<int> flat;
      for (auto &g : groups) flat.insert(flat.end(), g.begin(), g.end());
      void *d_index;
      if ((rc = dev_ensure(ctx, "rg_index" + s, flat.size() * 4, &d_index))) return rc;
      CU(ctx, cudaMemcpyAsync(d_index, flat.data(), flat.size() * 4, cudaMemcpyHostToDevice, st));
      CU(ctx, cudaStreamSynchronize(st));
      size_t at = 0;
      for (size_t g = 0; g < groups.size(); g++) {
        const long long gn = (long long)groups[g].size();
        LaunchPlan gp;
        if ((rc = plan_launch(ctx, gn, caps[g].first, caps[g].second, opt->check_cycles != 0, &gp,
                              groups.size() == 1 ? density : -1.0, true)))
          return rc;
        BatchArgs ca = a;
        ca.n = gn;
        ca.Hcap = caps[g].first;
        ca.Wcap = caps[g].second;
        ca.index = (const int *)d_index + at;
        int g_cluster = maybe_cluster(ctx, gn, caps[g].first, caps[g].second, nullptr, ca, s, st);
        if (g_cluster == 0) g_cluster = maybe_cluster(ctx, gn, caps[g].first, caps[g].second, &gp, ca, s, st);
        if (g_cluster < 0) return g_cluster;
        if (g_cluster) {
          // launched on the cluster path
        } else if (use_grid_path(ctx, gn, gp)) {
          for (int id : groups[g]) {
            const int hi = heights[begin + id], wi = widths[begin + id];
            const long long mo = moff[begin + id] - moff[begin], ro = roff[begin + id] - roff[begin],
                            po = poff[begin + id] - poff[begin];
            if ((rc = launch_grid(ctx, hi, wi, (double *)d_in + mo, opt, (int *)d_status + id, (double *)d_value + id,
                                  (long long *)d_piv + 2 * id, (double *)d_rhs + ro, (int *)d_pos + po, (int *)d_var + po,
                                  st, d_vin ? (const int *)d_vin + po : nullptr, d_vin ? wi + hi : 0)))
              return rc;
            if (matrices_out)
              CU(ctx, cudaMemcpyAsync((double *)d_out + mo, (double *)d_in + mo, (size_t)hi * wi * 8,
                                      cudaMemcpyDeviceToDevice, st));
          }
        } else {
          BatchArgs ga = a;
          ga.n = gn;
          ga.Hcap = caps[g].first;
          ga.Wcap = caps[g].second;
          ga.index = (const int *)d_index + at;
          if ((rc = launch_simplex(ctx, gp, ga, s, st))) return rc;
        }
        at += groups[g].size();
      }
    }

    if (status) CU(ctx, cudaMemcpyAsync(status + begin, d_status, (size_t)cn * 4, cudaMemcpyDeviceToHost, st));
    if (value) CU(ctx, cudaMemcpyAsync(value + begin, d_value, (size_t)cn * 8, cudaMemcpyDeviceToHost, st));
    if (pivots) CU(ctx, cudaMemcpyAsync(pivots + 2 * begin, d_piv, (size_t)cn * 16, cudaMemcpyDeviceToHost, st));
    if (rhs_out) CU(ctx, cudaMemcpyAsync(rhs_out + rows_upto(begin), d_rhs, crows * 8, cudaMemcpyDeviceToHost, st));
    if (pos_out) CU(ctx, cudaMemcpyAsync(pos_out + pv_upto(begin), d_pos, cpv * 4, cudaMemcpyDeviceToHost, st));
    if (var_out) CU(ctx, cudaMemcpyAsync(var_out + pv_upto(begin), d_var, cpv * 4, cudaMemcpyDeviceToHost, st));
    if (ctx->keep_final && n == 1 && !ragged) ctx->kept_final = a.mat_out;
    if (matrices_out) {
      if (ragged) {
        for (int64_t i = begin; i < end; i++)
          CU(ctx, cudaMemcpyAsync(matrices_out + mat_offsets[i], (double *)a.mat_out + (moff[i] - moff[begin]),
                                  (size_t)heights[i] * widths[i] * 8, cudaMemcpyDeviceToHost, st));
      } else {
        CU(ctx, cudaMemcpyAsync(matrices_out + cells_upto(begin), a.mat_out, ccells * 8, cudaMemcpyDeviceToHost, st));
      }
    }
    CU(ctx, cudaEventRecord(ctx->events[slot], st));
    used[slot] = true;
    slot ^= 1;
    begin = end;
  }
  for (int i = 0; i < 2; i++)
    if (used[i]) CU(ctx, cudaStreamSynchronize(ctx->streams[i]));
  return check_device_status(ctx, status, n);
}

int yalps_solve_batch(yalps_ctx *ctx, int64_t n, int32_t height, int32_t width, const double *matrices,
                      const yalps_options *opt, int32_t *status, double *value, int64_t *pivots, double *rhs_out,
                      int32_t *pos_out, int32_t *var_out, double *matrices_out) {
  return solve_host(ctx, n, height, width, nullptr, nullptr, nullptr, matrices, opt, status, value, pivots, rhs_out,
                    pos_out, var_out, matrices_out);
}

int yalps_solve_ragged(yalps_ctx *ctx, int64_t n, const int32_t *heights, const int32_t *widths,
                       const int64_t *mat_offsets, const double *matrices, const yalps_options *opt, int32_t *status,
                       double *value, int64_t *pivots, double *rhs_out, int32_t *pos_out, int32_t *var_out,
                       double *matrices_out) {
  if (!heights) return fail(ctx, YALPS_ERR_ARGUMENT, "heights is null");
  return solve_host(ctx, n, 0, 0, heights, widths, mat_offsets, matrices, opt, status, value, pivots, rhs_out,
                    pos_out, var_out, matrices_out);
}

int yalps_solve_batch_basis(yalps_ctx *ctx, int64_t n, int32_t height, int32_t width, const double *matrices,
                            const int32_t *pos_in, const int32_t *var_in, const yalps_options *opt, int32_t *status,
                            double *value, int64_t *pivots, double *rhs_out, int32_t *pos_out, int32_t *var_out,
                            double *matrices_out) {
  return solve_host(ctx, n, height, width, nullptr, nullptr, nullptr, matrices, opt, status, value, pivots, rhs_out,
                    pos_out, var_out, matrices_out, pos_in, var_in);
}

int yalps_solve_ragged_basis(yalps_ctx *ctx, int64_t n, const int32_t *heights, const int32_t *widths,
                             const int64_t *mat_offsets, const double *matrices, const int32_t *pos_in,
                             const int32_t *var_in, const yalps_options *opt, int32_t *status, double *value,
                             int64_t *pivots, double *rhs_out, int32_t *pos_out, int32_t *var_out,
                             double *matrices_out) {
  if (!heights) return fail(ctx, YALPS_ERR_ARGUMENT, "heights is null");
  return solve_host(ctx, n, 0, 0, heights, widths, mat_offsets, matrices, opt, status, value, pivots, rhs_out,
                    pos_out, var_out, matrices_out, pos_in, var_in);
}

// ---- replica path sharing (aux_kernels.cuh: k_replica_follow) ------------------------------------------------------
// The leader (the base tableau itself) is solved once by a row-split kernel that records its trace; see the comment
// above FollowArgs for why following it gives every replica its own result, bit for bit.
struct ReplicaTrace {
  bool ok = false;
  int K = 0, Hp = 0, cap = 0, leader_done = 0, leader_status = 0;
  int *steps = nullptr, *var = nullptr;
  double *q = nullptr, *col = nullptr, *snap = nullptr;
  double density = -1.0;
};

static int replica_leader_trace(yalps_ctx *ctx, int H, int W, const double *base_host, const double *d_base,
                                const yalps_options *opt, cudaStream_t st, ReplicaTrace *tr) {
  tr->ok = false;
  if (!ctx->replica_sharing || opt->check_cycles || getenv("YALPS_NO_REPLICA_SHARING")) return 0;
  if (ctx->tune_path != YALPS_PATH_AUTO) return 0;  // a forced kernel path means "measure that kernel"
  const size_t cells = (size_t)H * W, pv = (size_t)H + W;
  const SmemLayout Lg(H, W, false, 32);
  if (Lg.total > (size_t)ctx->smem_optin) return 0;  // the forks need a one-CTA-per-LP kernel
  // trace capacity: both phases of the budget, bounded by 512 MiB of snapshots
  const double budget = !(opt->max_pivots > 0.0) ? 0.0 : std::min(opt->max_pivots, 1.0e6);
  long long cap = (long long)std::min(2.0 * std::ceil(budget), 4096.0);
  cap = std::min<long long>(cap, (long long)(((size_t)512 << 20) / (cells * 8)) - 1);
  if (cap < 8) return 0;
  LaunchPlan plan;
  if (int rc = plan_launch(ctx, 1, H, W, false, &plan, -1.0, false)) return rc;
  if (plan.tmem || !plan.k || plan.k->nwr < 2) return 0;  // only the row-split kernels record a trace
  size_t nz = 0;
  for (size_t i = 0; i < cells; i++) nz += base_host[i] != 0.0;
  tr->density = (double)nz / (double)cells;
  tr->Hp = (H + 1) & ~1;
  tr->cap = (int)cap;
  void *p;
  int rc;
  if ((rc = dev_ensure(ctx, "rs_steps", (size_t)(cap + 1) * 4 * sizeof(int), &p))) return rc;
  tr->steps = (int *)p;
  if ((rc = dev_ensure(ctx, "rs_q", (size_t)cap * 8, &p))) return rc;
  tr->q = (double *)p;
  if ((rc = dev_ensure(ctx, "rs_col", (size_t)cap * tr->Hp * 8, &p))) return rc;
  tr->col = (double *)p;
  if ((rc = dev_ensure(ctx, "rs_snap", (size_t)(cap + 1) * cells * 8, &p))) return rc;
  tr->snap = (double *)p;
  if ((rc = dev_ensure(ctx, "rs_var", (size_t)(cap + 1) * pv * sizeof(int), &p))) return rc;
  tr->var = (int *)p;
  void *d_lwork, *d_lres, *d_lpv;
  if ((rc = dev_ensure(ctx, "rs_lwork", cells * 8, &d_lwork))) return rc;
  if ((rc = dev_ensure(ctx, "rs_lres", 64, &d_lres))) return rc;
  if ((rc = dev_ensure(ctx, "rs_lpv", 2 * pv * sizeof(int), &d_lpv))) return rc;
  BatchArgs a{};
  a.n = 1;
  a.mode = kModeBatch;
  a.H = a.Hcap = H;
  a.W = a.Wcap = W;
  a.in = d_base;
  a.work = (double *)d_lwork;
  a.status = (int *)d_lres;
  a.value = (double *)((char *)d_lres + 8);
  a.pivots = (long long *)((char *)d_lres + 16);
  a.pos_out = (int *)d_lpv;
  a.var_out = (int *)d_lpv + pv;
  fill_options(a, opt);
  a.tr_steps = tr->steps;
  a.tr_q = tr->q;
  a.tr_col = tr->col;
  a.tr_snap = tr->snap;
  a.tr_var = tr->var;
  a.tr_hp = tr->Hp;
  a.tr_cap = tr->cap;
  if ((rc = launch_simplex(ctx, plan, a, "rs_lead", st))) return rc;
  struct {
    int status, pad;
    double value;
    long long p1, p2;
  } res;
  CU(ctx, cudaMemcpyAsync(&res, d_lres, sizeof res, cudaMemcpyDeviceToHost, st));
  CU(ctx, cudaStreamSynchronize(st));
  if (res.status == ST_ERR_HISTORY) return 0;
  const long long K = res.p1 + res.p2;
  // a trace that did not fit ends one step early: snapshots exist for the states BEFORE pivots 0 .. cap - 1 only
  tr->leader_done = K <= cap;
  tr->K = (int)(K <= cap ? K : cap - 1);
  tr->leader_status = res.status;
  tr->ok = true;
  return 0;
}

// One chunk of replicas whose right-hand sides are in d_rin: followers first, then the replicas that left the path
// (tableau = the leader's snapshot of the fork step + their own column 0) through the ordinary batch kernels.
static int replica_chunk_shared(yalps_ctx *ctx, const ReplicaTrace &tr, int64_t cn, int H, int W, const double *d_rin,
                                double *d_work, const yalps_options *opt, int *d_status, double *d_value, long long *d_piv,
                                double *d_rhs, int *d_pos, int *d_var, cudaStream_t st, const std::string &s,
                                int64_t *forks_out) {
  const size_t cells = (size_t)H * W, pv = (size_t)H + W;
  void *p, *hp;
  int rc;
  if ((rc = dev_ensure(ctx, "rs_fcount" + s, 64, &p))) return rc;
  int *d_fcount = (int *)p;
  if ((rc = dev_ensure(ctx, "rs_fids" + s, (size_t)cn * 4, &p))) return rc;
  int *d_fids = (int *)p;
  if ((rc = dev_ensure(ctx, "rs_fstep" + s, (size_t)cn * 4, &p))) return rc;
  int *d_fstep = (int *)p;
  if ((rc = dev_ensure(ctx, "rs_fres" + s, (size_t)cn * 24, &p))) return rc;
  long long *d_fres = (long long *)p;
  if ((rc = pin_ensure(ctx, "rs_fcount_h" + s, 64, &hp))) return rc;
  CU(ctx, cudaMemsetAsync(d_fcount, 0, 4, st));
  FollowArgs fa{};
  fa.n = cn;
  fa.H = H;
  fa.W = W;
  fa.Hp = tr.Hp;
  fa.rhs_in = d_rin;
  fa.steps = tr.steps;
  fa.q = tr.q;
  fa.colraw = tr.col;
  fa.leader_var = tr.var + (size_t)tr.K * pv;
  fa.K = tr.K;
  fa.leader_done = tr.leader_done;
  fa.leader_status = tr.leader_status;
  fa.precision = opt->precision;
  fa.max_pivots = opt->max_pivots;
  fa.status = d_status;
  fa.value = d_value;
  fa.pivots = d_piv;
  fa.rhs_out = d_rhs;
  fa.pos_out = d_pos;
  fa.var_out = d_var;
  fa.fork_count = d_fcount;
  fa.fork_ids = d_fids;
  fa.fork_step = d_fstep;
  fa.fork_resume = d_fres;
  const int grid = (int)std::min<int64_t>((cn + kFollowWarps - 1) / kFollowWarps, (int64_t)ctx->prop.multiProcessorCount * 8);
  k_replica_follow<<<grid, kFollowWarps * 32, (size_t)kFollowWarps * tr.Hp * 8, st>>>(fa);
  CU(ctx, cudaGetLastError());
  ctx->launches++;
  CU(ctx, cudaMemcpyAsync(hp, d_fcount, 4, cudaMemcpyDeviceToHost, st));
  CU(ctx, cudaStreamSynchronize(st));
  const int nf = *(int *)hp;
  if (forks_out) *forks_out += nf;
  if (nf == 0) return 0;
  // compact buffers of the forked replicas
  if ((rc = dev_ensure(ctx, "rs_cvin" + s, (size_t)nf * pv * 4, &p))) return rc;
  int *c_vin = (int *)p;
  if ((rc = dev_ensure(ctx, "rs_cres" + s, (size_t)nf * 24, &p))) return rc;
  long long *c_resume = (long long *)p;
  if ((rc = dev_ensure(ctx, "rs_cstatus" + s, (size_t)nf * 4, &p))) return rc;
  int *c_status = (int *)p;
  if ((rc = dev_ensure(ctx, "rs_cvalue" + s, (size_t)nf * 8, &p))) return rc;
  double *c_value = (double *)p;
  if ((rc = dev_ensure(ctx, "rs_cpiv" + s, (size_t)nf * 16, &p))) return rc;
  long long *c_piv = (long long *)p;
  if ((rc = dev_ensure(ctx, "rs_crhs" + s, (size_t)nf * H * 8, &p))) return rc;
  double *c_rhs = (double *)p;
  if ((rc = dev_ensure(ctx, "rs_cpos" + s, (size_t)nf * pv * 4, &p))) return rc;
  int *c_pos = (int *)p;
  if ((rc = dev_ensure(ctx, "rs_cvar" + s, (size_t)nf * pv * 4, &p))) return rc;
  int *c_var = (int *)p;
  {
    const dim3 g((unsigned)std::min<size_t>((cells + 255) / 256, 64), (unsigned)std::min(nf, 32768));
    k_replica_materialise<<<g, 256, 0, st>>>(0, nf, H, W, d_fids, d_fstep, d_fres, d_rhs, tr.snap, tr.var, d_work, c_vin, c_resume);
    CU(ctx, cudaGetLastError());
    ctx->launches++;
  }
  LaunchPlan plan;
  if ((rc = plan_launch(ctx, nf, H, W, false, &plan, tr.density, false))) return rc;
  if (use_grid_path(ctx, nf, plan) || plan.tmem || !plan.k) {  // few large forks: one CTA per LP on the HBM-resident kernel
    const int keep = ctx->tune_path;
    ctx->tune_path = YALPS_PATH_GMEM;
    rc = plan_launch(ctx, nf, H, W, false, &plan, tr.density, false);
    ctx->tune_path = keep;
    if (rc) return rc;
  }
  BatchArgs a{};
  a.n = nf;
  a.mode = kModeBatch;
  a.H = a.Hcap = H;
  a.W = a.Wcap = W;
  a.in = d_work;
  a.work = d_work;
  a.status = c_status;
  a.value = c_value;
  a.pivots = c_piv;
  a.rhs_out = c_rhs;
  a.pos_out = c_pos;
  a.var_out = c_var;
  a.var_in = c_vin;
  a.resume = c_resume;
  fill_options(a, opt);
  if ((rc = launch_simplex(ctx, plan, a, "rs_fork" + s, st))) return rc;
  k_replica_scatter<<<(unsigned)std::min(nf, ctx->prop.multiProcessorCount * 8), 128, 0, st>>>(
      0, nf, H, W, d_fids, c_status, c_value, c_piv, c_rhs, c_pos, c_var, d_status, d_value, d_piv, d_rhs, d_pos, d_var);
  CU(ctx, cudaGetLastError());
  ctx->launches++;
  return 0;
}

int yalps_solve_replicas(yalps_ctx *ctx, int64_t n, int32_t height, int32_t width, const double *base,
                         const double *rhs, const yalps_options *opt, int32_t *status, double *value, int64_t *pivots,
                         double *rhs_out, int32_t *pos_out, int32_t *var_out) {
  if (!ctx) return YALPS_ERR_ARGUMENT;
  if (n < 0 || height < 1 || width < 1 || !opt || (n > 0 && (!base || !rhs)))
    return fail(ctx, YALPS_ERR_ARGUMENT, "bad arguments (n=%lld, %dx%d)", (long long)n, height, width);
  if ((long long)height * width >= (1LL << 31)) return fail(ctx, YALPS_ERR_TOO_LARGE, "height*width must be < 2^31");
  if (n == 0) return 0;
  CU(ctx, cudaSetDevice(ctx->device));
  const size_t cells = (size_t)height * width, pv = (size_t)height + width;
  void *d_base;
  int rc;
  if ((rc = dev_ensure(ctx, "rep_base", cells * 8, &d_base))) return rc;
  CU(ctx, cudaMemcpy(d_base, base, cells * 8, cudaMemcpyHostToDevice));
  ReplicaTrace trace;
  if ((rc = replica_leader_trace(ctx, height, width, base, (const double *)d_base, opt, ctx->streams[0], &trace))) return rc;
  ctx->replica_forks = trace.ok ? 0 : -1;
  // two pipeline slots of at most ~1.5 GiB of working copies each: the H2D of chunk k+1's right-hand sides and the
  // D2H of chunk k-1's results overlap the solve of chunk k
  const int64_t per_chunk = std::max<int64_t>(1, std::min<int64_t>((n + 1) / 2 > 4096 ? (n + 7) / 8 : n,
                                                                  (int64_t)(((size_t)1536 << 20) / (cells * 8))));
  bool used[2] = {false, false};
  int slot = 0;
  for (int64_t begin = 0; begin < n; begin += per_chunk, slot ^= 1) {
    const int64_t cn = std::min(per_chunk, n - begin);
    const std::string s = "rp" + std::to_string(slot);
    cudaStream_t st = ctx->streams[slot];
    if (used[slot]) CU(ctx, cudaEventSynchronize(ctx->events[slot]));
    void *d_rin, *d_work, *d_status, *d_value, *d_piv, *d_rhs, *d_pos, *d_var;
    if ((rc = dev_ensure(ctx, "rin" + s, (size_t)cn * height * 8, &d_rin))) return rc;
    if ((rc = dev_ensure(ctx, "work" + s, (size_t)cn * cells * 8, &d_work))) return rc;
    if ((rc = dev_ensure(ctx, "status" + s, (size_t)cn * 4, &d_status))) return rc;
    if ((rc = dev_ensure(ctx, "value" + s, (size_t)cn * 8, &d_value))) return rc;
    if ((rc = dev_ensure(ctx, "pivots" + s, (size_t)cn * 16, &d_piv))) return rc;
    if ((rc = dev_ensure(ctx, "rhs" + s, (size_t)cn * height * 8, &d_rhs))) return rc;
    if ((rc = dev_ensure(ctx, "pos" + s, (size_t)cn * pv * 4, &d_pos))) return rc;
    if ((rc = dev_ensure(ctx, "var" + s, (size_t)cn * pv * 4, &d_var))) return rc;
    CU(ctx, cudaMemcpyAsync(d_rin, rhs + (size_t)begin * height, (size_t)cn * height * 8, cudaMemcpyHostToDevice, st));
    if (trace.ok) {
      if ((rc = replica_chunk_shared(ctx, trace, cn, height, width, (const double *)d_rin, (double *)d_work, opt, (int *)d_status,
                                     (double *)d_value, (long long *)d_piv, (double *)d_rhs, (int *)d_pos, (int *)d_var, st, s,
                                     &ctx->replica_forks)))
        return rc;
    } else {
      const size_t total = (size_t)cn * cells;
      const int grid = (int)std::min<size_t>((total + 255) / 256, (size_t)ctx->prop.multiProcessorCount * 16);
      k_expand_replicas<<<grid, 256, 0, st>>>(cn, height, width, (const double *)d_base, (const double *)d_rin, (double *)d_work);
      CU(ctx, cudaGetLastError());
      ctx->launches++;
      if ((rc = solve_batch_device_impl(ctx, cn, height, width, (const double *)d_work, (double *)d_work, opt,
                                        (int32_t *)d_status, (double *)d_value, (int64_t *)d_piv, (double *)d_rhs,
                                        (int32_t *)d_pos, (int32_t *)d_var, nullptr, st, s, begin)))
        return rc;
    }
    if (status) CU(ctx, cudaMemcpyAsync(status + begin, d_status, (size_t)cn * 4, cudaMemcpyDeviceToHost, st));
    if (value) CU(ctx, cudaMemcpyAsync(value + begin, d_value, (size_t)cn * 8, cudaMemcpyDeviceToHost, st));
    if (pivots) CU(ctx, cudaMemcpyAsync(pivots + 2 * begin, d_piv, (size_t)cn * 16, cudaMemcpyDeviceToHost, st));
    if (rhs_out) CU(ctx, cudaMemcpyAsync(rhs_out + (size_t)begin * height, d_rhs, (size_t)cn * height * 8, cudaMemcpyDeviceToHost, st));
    if (pos_out) CU(ctx, cudaMemcpyAsync(pos_out + (size_t)begin * pv, d_pos, (size_t)cn * pv * 4, cudaMemcpyDeviceToHost, st));
    if (var_out) CU(ctx, cudaMemcpyAsync(var_out + (size_t)begin * pv, d_var, (size_t)cn * pv * 4, cudaMemcpyDeviceToHost, st));
    CU(ctx, cudaEventRecord(ctx->events[slot], st));
    used[slot] = true;
  }
  for (int i = 0; i < 2; i++)
    if (used[i]) CU(ctx, cudaStreamSynchronize(ctx->streams[i]));
  return check_device_status(ctx, status, n);
}

int yalps_set_replica_sharing(yalps_ctx *ctx, int32_t on) {
  if (!ctx) return YALPS_ERR_ARGUMENT;
  ctx->replica_sharing = on ? 1 : 0;
  return 0;
}

int64_t yalps_replica_forks(const yalps_ctx *ctx) { return ctx ? ctx->replica_forks : -1; }

int yalps_generate_synthetic_device(yalps_ctx *ctx, int64_t first, int64_t n, int32_t m, int32_t nvars,
                                    int32_t neg_rows, uint32_t salt, double *d_out, void *stream) {
  if (!ctx) return YALPS_ERR_ARGUMENT;
  if (n < 0 || m < 0 || nvars < 0 || !d_out) return fail(ctx, YALPS_ERR_ARGUMENT, "bad arguments");
  if (n == 0) return 0;
  CU(ctx, cudaSetDevice(ctx->device));
  const size_t total = (size_t)n * (m + 1) * (nvars + 1);
  const int grid = (int)std::min<size_t>((total + 255) / 256, (size_t)ctx->prop.multiProcessorCount * 16);
  k_generate_synthetic<<<grid, 256, 0, (cudaStream_t)stream>>>(first, n, m, nvars, neg_rows, salt, d_out);
  CU(ctx, cudaGetLastError());
  ctx->launches++;
  return 0;
}

int yalps_generate_replicas_device(yalps_ctx *ctx, int64_t first, int64_t n, int32_t height, int32_t width,
                                   const double *base_host, const int32_t *group_host, int32_t ngroups, double eps,
                                   uint32_t salt, double *d_out, void *stream) {
  if (!ctx) return YALPS_ERR_ARGUMENT;
  if (n < 0 || height < 1 || width < 1 || !base_host || !group_host || !d_out)
    return fail(ctx, YALPS_ERR_ARGUMENT, "bad arguments");
  (void)ngroups;
  if (n == 0) return 0;
  CU(ctx, cudaSetDevice(ctx->device));
  void *d_base, *d_group;
  const size_t cells = (size_t)height * width;
  if (int rc = dev_ensure(ctx, "rep_base", cells * 8, &d_base)) return rc;
  if (int rc = dev_ensure(ctx, "rep_group", (size_t)height * 4, &d_group)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  CU(ctx, cudaMemcpyAsync(d_base, base_host, cells * 8, cudaMemcpyHostToDevice, st));
  CU(ctx, cudaMemcpyAsync(d_group, group_host, (size_t)height * 4, cudaMemcpyHostToDevice, st));
  const size_t total = (size_t)n * cells;
  const int grid = (int)std::min<size_t>((total + 255) / 256, (size_t)ctx->prop.multiProcessorCount * 16);
  k_generate_replicas<<<grid, 256, 0, st>>>(first, n, height, width, (const double *)d_base, (const int *)d_group, eps,
                                            salt, d_out);
  CU(ctx, cudaGetLastError());
  CU(ctx, cudaStreamSynchronize(st));  // base_host / group_host may be pageable
  ctx->launches++;
  return 0;
}

int yalps_round_to_precision(yalps_ctx *ctx, int64_t n, const double *x, double precision, double *out) {
  if (!ctx) return YALPS_ERR_ARGUMENT;
  if (n < 0 || (n > 0 && (!x || !out))) return fail(ctx, YALPS_ERR_ARGUMENT, "bad arguments");
  if (n == 0) return 0;
  CU(ctx, cudaSetDevice(ctx->device));
  void *d_x, *d_o;
  if (int rc = dev_ensure(ctx, "round_x", (size_t)n * 8, &d_x)) return rc;
  if (int rc = dev_ensure(ctx, "round_o", (size_t)n * 8, &d_o)) return rc;
  cudaStream_t st = ctx->streams[0];
  CU(ctx, cudaMemcpyAsync(d_x, x, (size_t)n * 8, cudaMemcpyHostToDevice, st));
  k_round_to_precision<<<(int)((n + 255) / 256), 256, 0, st>>>(n, (const double *)d_x, precision, (double *)d_o);
  CU(ctx, cudaGetLastError());
  ctx->launches++;
  CU(ctx, cudaMemcpyAsync(out, d_o, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
  CU(ctx, cudaStreamSynchronize(st));
  return 0;
}

int yalps_probe_division(yalps_ctx *ctx, int64_t n, uint64_t seed, int32_t mode, uint64_t *mismatches,
                         uint64_t *first_bad_bits) {
  if (!ctx || !mismatches || n < 0 || mode < 0 || mode > 4) return YALPS_ERR_ARGUMENT;
  CU(ctx, cudaSetDevice(ctx->device));
  void *d;
  if (int rc = dev_ensure(ctx, "probe_div", 32, &d)) return rc;
  cudaStream_t st = ctx->streams[0];
  CU(ctx, cudaMemsetAsync(d, 0, 32, st));
  k_probe_division<<<ctx->prop.multiProcessorCount * 8, 256, 0, st>>>(n, seed, mode, (unsigned long long *)d);
  CU(ctx, cudaGetLastError());
  ctx->launches++;
  unsigned long long h[4];
  CU(ctx, cudaMemcpyAsync(h, d, 32, cudaMemcpyDeviceToHost, st));
  CU(ctx, cudaStreamSynchronize(st));
  *mismatches = h[0];
  if (first_bad_bits) {
    first_bad_bits[0] = h[2];
    first_bad_bits[1] = h[3];
  }
  return 0;
}

int yalps_measure_smem_bandwidth(yalps_ctx *ctx, double *gbs, double *sm_clock_mhz) {
  if (!ctx || !gbs) return YALPS_ERR_ARGUMENT;
  CU(ctx, cudaSetDevice(ctx->device));
  const int words = 24 * 1024 / 8 * 4;  // 96 KiB per CTA, 2 CTAs per SM
  const int threads = 512, iters = 2000;
  const size_t smem = (size_t)words * 8;
  CU(ctx, raise_smem_limit(ctx->device, (const void *)k_smem_stream, (int)smem));
  void *sink;
  if (int rc = dev_ensure(ctx, "sink", 64, &sink)) return rc;
  const int grid = ctx->prop.multiProcessorCount * 2;
  cudaStream_t st = ctx->streams[0];
  cudaEvent_t e0, e1;
  CU(ctx, cudaEventCreate(&e0));
  CU(ctx, cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 5; rep++) {
    CU(ctx, cudaEventRecord(e0, st));
    k_smem_stream<<<grid, threads, smem, st>>>(words, iters, (double *)sink);
    CU(ctx, cudaEventRecord(e1, st));
    CU(ctx, cudaEventSynchronize(e1));
    CU(ctx, cudaGetLastError());
    ctx->launches++;
    float ms = 0;
    CU(ctx, cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0) best = std::min(best, ms);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  const double bytes = (double)grid * iters * words * 16.0;
  *gbs = bytes / (best * 1e-3) / 1e9;
  if (sm_clock_mhz) {
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, ctx->device);
    *sm_clock_mhz = khz / 1000.0;
  }
  return 0;
}

int yalps_measure_l2_bandwidth(yalps_ctx *ctx, uint64_t bytes, double *gbs) {
  if (!ctx || !gbs || bytes < 4096) return YALPS_ERR_ARGUMENT;
  CU(ctx, cudaSetDevice(ctx->device));
  void *d;
  if (int rc = dev_ensure(ctx, "l2_probe", bytes, &d)) return rc;
  cudaStream_t st = ctx->streams[0];
  CU(ctx, cudaMemsetAsync(d, 0, bytes, st));
  const int iters = 50, grid = ctx->prop.multiProcessorCount * 8;
  cudaEvent_t e0, e1;
  CU(ctx, cudaEventCreate(&e0));
  CU(ctx, cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    CU(ctx, cudaEventRecord(e0, st));
    k_l2_stream<<<grid, 256, 0, st>>>((double2 *)d, bytes / 16, iters, 0.5);
    CU(ctx, cudaEventRecord(e1, st));
    CU(ctx, cudaEventSynchronize(e1));
    CU(ctx, cudaGetLastError());
    ctx->launches++;
    float ms = 0;
    CU(ctx, cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0) best = std::min(best, ms);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *gbs = (double)(bytes / 16 * 16) * iters * 2.0 / (best * 1e-3) / 1e9;
  return 0;
}

int yalps_measure_h2d_bandwidth(yalps_ctx *ctx, const void *pinned_host, uint64_t bytes, int32_t reps, int32_t nstreams,
                                double *seconds) {
  if (!ctx || !pinned_host || !seconds || bytes == 0 || reps < 1 || nstreams < 1 || nstreams > 2)
    return ctx ? fail(ctx, YALPS_ERR_ARGUMENT, "bad arguments") : YALPS_ERR_ARGUMENT;
  CU(ctx, cudaSetDevice(ctx->device));
  void *d;
  if (int rc = dev_ensure(ctx, "h2d_probe", bytes, &d)) return rc;
  const size_t part = ((bytes / nstreams) + 255) & ~(size_t)255;
  CU(ctx, cudaDeviceSynchronize());
  const auto t0 = std::chrono::steady_clock::now();
  for (int r = 0; r < reps; r++)
    for (int k = 0; k < nstreams; k++) {
      const size_t off = (size_t)k * part;
      if (off >= bytes) break;
      const size_t len = std::min(part, (size_t)bytes - off);
      CU(ctx, cudaMemcpyAsync((char *)d + off, (const char *)pinned_host + off, len, cudaMemcpyHostToDevice, ctx->streams[k]));
    }
  for (int k = 0; k < nstreams; k++) CU(ctx, cudaStreamSynchronize(ctx->streams[k]));
  *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / reps;
  return 0;
}

int yalps_measure_tmem_bandwidth(yalps_ctx *ctx, double *gbs, double *sm_clock_mhz) {
  if (!ctx || !gbs) return YALPS_ERR_ARGUMENT;
  CU(ctx, cudaSetDevice(ctx->device));
  CU(ctx, raise_smem_limit(ctx->device, tmem_stream_fn(), (int)tmem_stream_dynamic_smem()));
  void *sink;
  if (int rc = dev_ensure(ctx, "sink", 64, &sink)) return rc;
  const int grid = ctx->prop.multiProcessorCount * tmem_stream_ctas_per_sm();
  cudaStream_t st = ctx->streams[0];
  cudaEvent_t e0, e1;
  CU(ctx, cudaEventCreate(&e0));
  CU(ctx, cudaEventCreate(&e1));
  float best = 1e30f;
  double bytes = 0.0;
  for (int rep = 0; rep < 5; rep++) {
    CU(ctx, cudaEventRecord(e0, st));
    CU(ctx, launch_tmem_stream(grid, 4000, (double *)sink, st, &bytes));
    CU(ctx, cudaEventRecord(e1, st));
    CU(ctx, cudaEventSynchronize(e1));
    ctx->launches++;
    float ms = 0;
    CU(ctx, cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0) best = std::min(best, ms);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *gbs = bytes / (best * 1e-3) / 1e9;
  if (sm_clock_mhz) {
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, ctx->device);
    *sm_clock_mhz = khz / 1000.0;
  }
  return 0;
}

}  // extern "C"

#include "bnb.inl"
#include "multi.inl"
