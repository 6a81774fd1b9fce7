// multi.inl -- one process, several GPUs (included by yalps_b200.cu; API in include/yalps_b200.h).
//
// The reference is a synchronous single-process library (src/YALPS.ts:73-92); its N-API addon cannot be launched by
// torchrun.  A yalps_multi therefore owns the ranks itself: one yalps_ctx and one host thread per rank (a ctx is
// single-threaded by contract), extra worker ctxs per rank for concurrent branch-and-cut searches.
//   * independent LPs (batch / ragged / replicas): contiguous shards, LP i -> rank floor(i*ndev/n), no collective on the
//     data path -- every rank runs the single-GPU pipeline on its slice of the caller's arrays;
//   * branch and bound: root replicated, the nodes of a speculative wave dealt over the ranks, the replay is the
//     reference's sequential loop (bnb.inl), the incumbent goes through ncclAllReduce(min) every k waves;
//   * many MILPs: searches dealt to ndev * searches_per_device worker contexts.
// NCCL is opened at run time (libnccl.so.2; a process that already loaded one, e.g. through PyTorch, shares it).

namespace {

thread_local std::string g_multi_create_error;

struct MultiWorker {
  yalps_ctx *ctx = nullptr;
  std::thread th;
  std::mutex mu;
  std::condition_variable cv;
  std::function<int()> job;
  bool pending = false, quit = false;
  int rc = 0;

  void loop() {
    std::unique_lock<std::mutex> lk(mu);
    for (;;) {
      cv.wait(lk, [&] { return pending || quit; });
      if (quit) return;
      std::function<int()> j = std::move(job);
      lk.unlock();
      const int r = j();
      lk.lock();
      rc = r;
      pending = false;
      cv.notify_all();
    }
  }
  void submit(std::function<int()> j) {
    std::lock_guard<std::mutex> lk(mu);
    job = std::move(j);
    pending = true;
    cv.notify_all();
  }
  int wait() {
    std::unique_lock<std::mutex> lk(mu);
    cv.wait(lk, [&] { return !pending; });
    return rc;
  }
  void stop() {
    {
      std::lock_guard<std::mutex> lk(mu);
      quit = true;
      cv.notify_all();
    }
    if (th.joinable()) th.join();
  }
};

struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  std::string error;

  bool load() {
    if (handle) return true;
    handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // already in the process (PyTorch's)?
    if (!handle) handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!handle) handle = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
    if (!handle) {
      error = std::string("cannot open libnccl.so.2: ") + dlerror();
      return false;
    }
    auto sym = [&](const char *name) -> void * {
      void *p = dlsym(handle, name);
      if (!p && error.empty()) error = std::string("libnccl.so.2 lacks ") + name;
      return p;
    };
    CommInitAll = (decltype(CommInitAll))sym("ncclCommInitAll");
    CommDestroy = (decltype(CommDestroy))sym("ncclCommDestroy");
    AllReduce = (decltype(AllReduce))sym("ncclAllReduce");
    GroupStart = (decltype(GroupStart))sym("ncclGroupStart");
    GroupEnd = (decltype(GroupEnd))sym("ncclGroupEnd");
    GetErrorString = (decltype(GetErrorString))sym("ncclGetErrorString");
    if (!error.empty()) {
      handle = nullptr;
      return false;
    }
    return true;
  }
};

}  // namespace

struct yalps_multi {
  std::vector<int> devices;  // rank -> GPU (an entry may repeat: several logical ranks on one GPU)
  std::vector<std::vector<std::unique_ptr<MultiWorker>>> workers;  // [rank][w]; w == 0 is the rank's own ctx
  std::string error;
  // incumbent allreduce
  std::vector<int> gpus;                 // distinct GPUs in order of first appearance
  std::vector<std::vector<int>> gpu_ranks;  // ranks on each of them
  std::vector<double *> d_slots;         // per GPU: (ranks on it + 1) doubles
  std::vector<cudaStream_t> ar_streams;
  std::vector<ncclComm_t> comms;
  NcclApi nccl;
  int64_t allreduces = 0;
  bool peers_enabled = false;  // cudaDeviceEnablePeerAccess between all distinct GPUs (yalps_multi_solve_large)
};

namespace {

int mfail(yalps_multi *m, int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (m)
    m->error = buf;
  else
    g_multi_create_error = buf;
  return code;
}

#define MCU(m, call)                                                                                            \
  do {                                                                                                          \
    cudaError_t e_ = (call);                                                                                    \
    if (e_ != cudaSuccess)                                                                                      \
      return mfail(m, YALPS_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

inline int64_t shard_begin(int64_t n, int rank, int world) {
  return (int64_t)(((__int128)rank * n + world - 1) / world);  // ceil(rank*n/world): LP i -> rank floor(i*world/n)
}

MultiWorker *ensure_worker(yalps_multi *m, int rank, int w) {
  auto &row = m->workers[rank];
  while ((int)row.size() <= w) {
    std::unique_ptr<MultiWorker> mw(new MultiWorker);
    if (yalps_create(m->devices[rank], &mw->ctx) != 0) {
      m->error = yalps_last_error(nullptr);
      return nullptr;
    }
    MultiWorker *raw = mw.get();
    mw->th = std::thread([raw] { raw->loop(); });
    row.push_back(std::move(mw));
  }
  return row[w].get();
}

// Runs job(rank) on every rank's own thread and waits; the first failure's message becomes the multi's error.
int run_on_ranks(yalps_multi *m, const std::function<int(int, yalps_ctx *)> &job) {
  const int world = (int)m->devices.size();
  for (int r = 0; r < world; r++) {
    MultiWorker *w = m->workers[r][0].get();
    w->submit([&job, r, w] { return job(r, w->ctx); });
  }
  int rc = 0;
  for (int r = 0; r < world; r++) {
    const int rr = m->workers[r][0]->wait();
    if (rr != 0 && rc == 0) {
      rc = rr;
      m->error = "rank " + std::to_string(r) + " (GPU " + std::to_string(m->devices[r]) + "): " +
                 yalps_last_error(m->workers[r][0]->ctx);
    }
  }
  return rc;
}

int ensure_allreduce(yalps_multi *m) {
  if (!m->comms.empty()) return 0;
  if (!m->nccl.load()) return mfail(m, YALPS_ERR_CUDA, "incumbent allreduce needs NCCL: %s", m->nccl.error.c_str());
  const int ng = (int)m->gpus.size();
  m->d_slots.assign(ng, nullptr);
  m->ar_streams.assign(ng, nullptr);
  for (int g = 0; g < ng; g++) {
    MCU(m, cudaSetDevice(m->gpus[g]));
    MCU(m, cudaMalloc((void **)&m->d_slots[g], (m->gpu_ranks[g].size() + 1) * sizeof(double)));
    MCU(m, cudaStreamCreateWithFlags(&m->ar_streams[g], cudaStreamNonBlocking));
  }
  std::vector<ncclComm_t> comms(ng);
  const ncclResult_t r = m->nccl.CommInitAll(comms.data(), ng, m->gpus.data());
  if (r != ncclSuccess) return mfail(m, YALPS_ERR_CUDA, "ncclCommInitAll over %d GPUs failed: %s", ng, m->nccl.GetErrorString(r));
  m->comms = comms;
  return 0;
}

}  // namespace

extern "C" {

int yalps_create_multi(const int32_t *devices, int32_t ndev, yalps_multi **out) {
  if (!out) return mfail(nullptr, YALPS_ERR_ARGUMENT, "out is null");
  *out = nullptr;
  if (!devices || ndev < 1 || ndev > 64) return mfail(nullptr, YALPS_ERR_ARGUMENT, "need 1..64 devices");
  std::unique_ptr<yalps_multi> m(new yalps_multi);
  m->devices.assign(devices, devices + ndev);
  m->workers.resize(ndev);
  for (int r = 0; r < ndev; r++) {
    if (!ensure_worker(m.get(), r, 0)) {
      g_multi_create_error = "rank " + std::to_string(r) + ": " + m->error;
      yalps_destroy_multi(m.release());
      return YALPS_ERR_CUDA;
    }
    size_t g = 0;
    while (g < m->gpus.size() && m->gpus[g] != devices[r]) g++;
    if (g == m->gpus.size()) {
      m->gpus.push_back(devices[r]);
      m->gpu_ranks.emplace_back();
    }
    m->gpu_ranks[g].push_back(r);
  }
  *out = m.release();
  return 0;
}

void yalps_destroy_multi(yalps_multi *m) {
  if (!m) return;
  for (auto &row : m->workers)
    for (auto &w : row) {
      w->stop();
      if (w->ctx) yalps_destroy(w->ctx);
    }
  for (size_t g = 0; g < m->comms.size(); g++)
    if (m->comms[g]) m->nccl.CommDestroy(m->comms[g]);
  for (size_t g = 0; g < m->d_slots.size(); g++) {
    cudaSetDevice(m->gpus[g]);
    if (m->d_slots[g]) cudaFree(m->d_slots[g]);
    if (m->ar_streams[g]) cudaStreamDestroy(m->ar_streams[g]);
  }
  delete m;
}

const char *yalps_multi_last_error(const yalps_multi *m) { return m ? m->error.c_str() : g_multi_create_error.c_str(); }

int32_t yalps_multi_size(const yalps_multi *m) { return m ? (int32_t)m->devices.size() : 0; }

yalps_ctx *yalps_multi_ctx(yalps_multi *m, int32_t rank) {
  if (!m || rank < 0 || rank >= (int32_t)m->devices.size()) return nullptr;
  return m->workers[rank][0]->ctx;
}

int64_t yalps_multi_launch_count(const yalps_multi *m) {
  if (!m) return 0;
  int64_t total = m->allreduces * 2;  // the reduce and broadcast kernels of every incumbent allreduce
  for (const auto &row : m->workers)
    for (const auto &w : row) total += yalps_launch_count(w->ctx);
  return total;
}

int yalps_multi_solve_batch(yalps_multi *m, int64_t n, int32_t height, int32_t width, const double *matrices,
                            const int32_t *pos_in, const int32_t *var_in, const yalps_options *opt, int32_t *status,
                            double *value, int64_t *pivots, double *rhs_out, int32_t *pos_out, int32_t *var_out,
                            double *matrices_out) {
  if (!m) return YALPS_ERR_ARGUMENT;
  if (n < 0 || height < 1 || width < 1 || !opt || (n > 0 && !matrices)) return mfail(m, YALPS_ERR_ARGUMENT, "bad arguments");
  const int world = (int)m->devices.size();
  const size_t cells = (size_t)height * width, pv = (size_t)height + width;
  return run_on_ranks(m, [&](int r, yalps_ctx *ctx) -> int {
    const int64_t lo = shard_begin(n, r, world), hi = shard_begin(n, r + 1, world);
    if (hi <= lo) return 0;
    return yalps_solve_batch_basis(ctx, hi - lo, height, width, matrices + lo * cells, pos_in ? pos_in + lo * pv : nullptr,
                                   var_in ? var_in + lo * pv : nullptr, opt, status ? status + lo : nullptr,
                                   value ? value + lo : nullptr, pivots ? pivots + 2 * lo : nullptr,
                                   rhs_out ? rhs_out + lo * height : nullptr, pos_out ? pos_out + lo * pv : nullptr,
                                   var_out ? var_out + lo * pv : nullptr, matrices_out ? matrices_out + lo * cells : nullptr);
  });
}

int yalps_multi_solve_ragged(yalps_multi *m, int64_t n, const int32_t *heights, const int32_t *widths,
                             const int64_t *mat_offsets, const double *matrices, const yalps_options *opt,
                             int32_t *status, double *value, int64_t *pivots, double *rhs_out, int32_t *pos_out,
                             int32_t *var_out, double *matrices_out) {
  if (!m) return YALPS_ERR_ARGUMENT;
  if (n < 0 || !opt || (n > 0 && (!heights || !widths || !mat_offsets || !matrices)))
    return mfail(m, YALPS_ERR_ARGUMENT, "bad arguments");
  const int world = (int)m->devices.size();
  // ranges balanced by cells (LPs of a ragged batch differ in size), still contiguous: rank r takes the LPs whose
  // cumulative cell count falls into its 1/world share
  std::vector<int64_t> roff(n + 1, 0), poff(n + 1, 0);
  std::vector<double> cum(n + 1, 0.0);
  for (int64_t i = 0; i < n; i++) {
    if (heights[i] < 1 || widths[i] < 1) return mfail(m, YALPS_ERR_ARGUMENT, "LP %lld has empty shape", (long long)i);
    roff[i + 1] = roff[i] + heights[i];
    poff[i + 1] = poff[i] + heights[i] + widths[i];
    cum[i + 1] = cum[i] + (double)heights[i] * widths[i];
  }
  std::vector<int64_t> cut(world + 1, n);
  cut[0] = 0;
  for (int r = 1; r < world; r++) {
    const double want = cum[n] * r / world;
    cut[r] = std::lower_bound(cum.begin(), cum.end(), want) - cum.begin();
    cut[r] = std::max(cut[r - 1], std::min<int64_t>(cut[r], n));
  }
  return run_on_ranks(m, [&](int r, yalps_ctx *ctx) -> int {
    const int64_t lo = cut[r], hi = cut[r + 1];
    if (hi <= lo) return 0;
    return yalps_solve_ragged(ctx, hi - lo, heights + lo, widths + lo, mat_offsets + lo, matrices, opt,
                              status ? status + lo : nullptr, value ? value + lo : nullptr,
                              pivots ? pivots + 2 * lo : nullptr, rhs_out ? rhs_out + roff[lo] : nullptr,
                              pos_out ? pos_out + poff[lo] : nullptr, var_out ? var_out + poff[lo] : nullptr,
                              matrices_out);
  });
}

int yalps_multi_solve_replicas(yalps_multi *m, int64_t n, int32_t height, int32_t width, const double *base,
                               const double *rhs, const yalps_options *opt, int32_t *status, double *value,
                               int64_t *pivots, double *rhs_out, int32_t *pos_out, int32_t *var_out) {
  if (!m) return YALPS_ERR_ARGUMENT;
  if (n < 0 || height < 1 || width < 1 || !opt || (n > 0 && (!base || !rhs))) return mfail(m, YALPS_ERR_ARGUMENT, "bad arguments");
  const int world = (int)m->devices.size();
  const size_t pv = (size_t)height + width;
  return run_on_ranks(m, [&](int r, yalps_ctx *ctx) -> int {
    const int64_t lo = shard_begin(n, r, world), hi = shard_begin(n, r + 1, world);
    if (hi <= lo) return 0;
    return yalps_solve_replicas(ctx, hi - lo, height, width, base, rhs + lo * height, opt, status ? status + lo : nullptr,
                                value ? value + lo : nullptr, pivots ? pivots + 2 * lo : nullptr,
                                rhs_out ? rhs_out + lo * height : nullptr, pos_out ? pos_out + lo * pv : nullptr,
                                var_out ? var_out + lo * pv : nullptr);
  });
}

int yalps_incumbent_allreduce(yalps_multi *m, const double *local, double *agreed) {
  if (!m) return YALPS_ERR_ARGUMENT;
  if (!local || !agreed) return mfail(m, YALPS_ERR_ARGUMENT, "bad arguments");
  if (int rc = ensure_allreduce(m)) return rc;
  const int ng = (int)m->gpus.size();
  std::vector<std::vector<double>> h(ng);
  for (int g = 0; g < ng; g++) {
    const int k = (int)m->gpu_ranks[g].size();
    h[g].resize(k);
    for (int i = 0; i < k; i++) {
      const double v = local[m->gpu_ranks[g][i]];
      h[g][i] = v == v ? v : std::numeric_limits<double>::infinity();  // NaN = no incumbent
    }
    MCU(m, cudaSetDevice(m->gpus[g]));
    MCU(m, cudaMemcpyAsync(m->d_slots[g], h[g].data(), k * sizeof(double), cudaMemcpyHostToDevice, m->ar_streams[g]));
    k_incumbent_reduce<<<1, 1, 0, m->ar_streams[g]>>>(m->d_slots[g], k);
    MCU(m, cudaGetLastError());
  }
  ncclResult_t nr = m->nccl.GroupStart();
  for (int g = 0; g < ng && nr == ncclSuccess; g++) {
    const int k = (int)m->gpu_ranks[g].size();
    nr = m->nccl.AllReduce(m->d_slots[g] + k, m->d_slots[g] + k, 1, ncclDouble, ncclMin, m->comms[g], m->ar_streams[g]);
  }
  const ncclResult_t ne = m->nccl.GroupEnd();
  if (nr == ncclSuccess) nr = ne;
  if (nr != ncclSuccess) return mfail(m, YALPS_ERR_CUDA, "ncclAllReduce(min) failed: %s", m->nccl.GetErrorString(nr));
  for (int g = 0; g < ng; g++) {
    const int k = (int)m->gpu_ranks[g].size();
    MCU(m, cudaSetDevice(m->gpus[g]));
    k_incumbent_broadcast<<<1, 1, 0, m->ar_streams[g]>>>(m->d_slots[g], k);
    MCU(m, cudaGetLastError());
    MCU(m, cudaMemcpyAsync(h[g].data(), m->d_slots[g], k * sizeof(double), cudaMemcpyDeviceToHost, m->ar_streams[g]));
  }
  for (int g = 0; g < ng; g++) {
    MCU(m, cudaSetDevice(m->gpus[g]));
    MCU(m, cudaStreamSynchronize(m->ar_streams[g]));
    for (size_t i = 0; i < m->gpu_ranks[g].size(); i++) agreed[m->gpu_ranks[g][i]] = h[g][i];
  }
  m->allreduces++;
  return 0;
}

int yalps_multi_solve(yalps_multi *m, int32_t height, int32_t width, const double *matrix, const int32_t *ints,
                      int32_t nints, double sign, const yalps_options *opt, int32_t allreduce_every, int32_t *status,
                      double *result, int32_t *out_height, double *rhs_out, int32_t *pos_out, int32_t *var_out,
                      int32_t *root_status, double *root_value, int64_t *root_pivots, int64_t *stats) {
  if (!m) return YALPS_ERR_ARGUMENT;
  if (height < 1 || width < 1 || !matrix || !opt || !status || !result || !out_height || !rhs_out || !pos_out ||
      !var_out || nints < 0 || (nints && !ints))
    return mfail(m, YALPS_ERR_ARGUMENT, "bad arguments");
  const int world = (int)m->devices.size();
  yalps_ctx *ctx0 = m->workers[0][0]->ctx;
  m->error.clear();
  if (stats) std::memset(stats, 0, sizeof(int64_t) * 10);
  auto ctx_fail = [&](int rc) {
    m->error = std::string("rank 0: ") + yalps_last_error(ctx0);
    return rc;
  };
  if (world == 1 || nints == 0) {  // nothing to shard: the single-GPU path
    const int rc = yalps_solve(ctx0, height, width, matrix, ints, nints, sign, opt, status, result, out_height, rhs_out,
                               pos_out, var_out, root_status, root_value, root_pivots, stats);
    return rc ? ctx_fail(rc) : 0;
  }
  // root LP on rank 0 (src/YALPS.ts:79), final tableau back to the host once, then replicated on every rank
  const size_t cells = (size_t)height * width;
  std::vector<double> final_m(cells);
  int32_t st = 0;
  double val = 0;
  int64_t piv[2] = {0, 0};
  if (int rc = yalps_solve_batch(ctx0, 1, height, width, matrix, opt, &st, &val, piv, rhs_out, pos_out, var_out, final_m.data()))
    return ctx_fail(rc);
  if (root_status) *root_status = st;
  if (root_value) *root_value = val;
  if (root_pivots) root_pivots[0] = piv[0], root_pivots[1] = piv[1];
  *out_height = height;
  if (st != YALPS_OPTIMAL) {  // src/YALPS.ts:81-86
    *status = st;
    *result = val;
    return 0;
  }
  if (int rc = run_on_ranks(m, [&](int, yalps_ctx *ctx) {
        return yalps_bnb_set_root(ctx, height, width, final_m.data(), pos_out, var_out, 2 * nints);
      }))
    return rc;

  // A wave is worth dealing over the ranks when its nodes are too big for one SM's shared memory (each then takes a
  // cluster or the whole grid: a handful of nodes already fills a GPU) or when there are many small ones.
  const bool big_nodes = SmemLayout(height + 2, width, true, 32).total > (size_t)ctx0->smem_optin;
  int64_t sharded_waves = 0, allreduces = 0;
  std::vector<std::vector<int32_t>> loc_off(world);
  WaveEval eval = [&](int64_t n, const int32_t *off, const double *csign, const int32_t *cvar, const double *cval,
                      int maxcuts, int32_t *w_status, double *w_value, int64_t *w_piv, double *w_rhs, int32_t *w_pos,
                      int32_t *w_var) -> int {
    const int Hcap = height + maxcuts;
    const bool shard = n >= 2 && (big_nodes || n >= 8LL * world);
    if (!shard) {
      const int rc = bnb_solve_nodes_impl(ctx0, n, off, csign, cvar, cval, opt, w_status, w_value, w_piv, w_rhs, w_pos,
                                          w_var, nullptr, maxcuts);
      return rc ? ctx_fail(rc) : 0;
    }
    sharded_waves++;
    return run_on_ranks(m, [&](int r, yalps_ctx *ctx) -> int {
      const int64_t lo = shard_begin(n, r, world), hi = shard_begin(n, r + 1, world);
      if (hi <= lo) return 0;
      std::vector<int32_t> &lo_off = loc_off[r];
      lo_off.resize(hi - lo + 1);
      for (int64_t j = lo; j <= hi; j++) lo_off[j - lo] = off[j] - off[lo];
      const size_t c0 = (size_t)off[lo];
      return bnb_solve_nodes_impl(ctx, hi - lo, lo_off.data(), csign + c0, cvar + c0, cval + c0, opt, w_status + lo,
                                  w_value + lo, w_piv + 2 * lo, w_rhs + (size_t)lo * Hcap,
                                  w_pos + (size_t)lo * (width + Hcap), w_var + (size_t)lo * (width + Hcap), nullptr,
                                  maxcuts);
    });
  };
  // north_star: "an NCCL min-allreduce of the incumbent objective every k node batches".  The replay is sequential and
  // replicated in this one process, so every rank contributes the same incumbent and the collective doubles as a
  // consistency check of the communicator path.
  std::vector<double> loc(world), agreed(world);
  WaveHook hook = [&](int64_t wave_index, double best_eval) -> int {
    if (allreduce_every <= 0 || wave_index % allreduce_every != 0) return 0;
    std::fill(loc.begin(), loc.end(), best_eval);
    if (int rc = yalps_incumbent_allreduce(m, loc.data(), agreed.data())) return rc;
    allreduces++;
    for (int r = 0; r < world; r++)
      if (agreed[r] != best_eval) return mfail(m, YALPS_ERR_CUDA, "incumbent allreduce returned %g on rank %d, expected %g", agreed[r], r, best_eval);
    return 0;
  };
  const int rc = branch_and_cut_impl(ctx0, &eval, &hook, ints, nints, sign, val, opt, status, result, out_height, rhs_out,
                                     pos_out, var_out, stats);
  if (stats) {
    stats[8] = sharded_waves;
    stats[9] = allreduces;
  }
  if (rc && m->error.empty()) return ctx_fail(rc);
  return rc;
}

int yalps_multi_solve_many(yalps_multi *m, int64_t n_models, const int32_t *heights, const int32_t *widths,
                           const int64_t *mat_offsets, const double *matrices, const int64_t *ints_offsets,
                           const int32_t *ints, const double *signs, const yalps_options *opt,
                           int32_t searches_per_device, int32_t *status, double *result, int32_t *out_height,
                           double *rhs_out, int32_t *pos_out, int32_t *var_out) {
  if (!m) return YALPS_ERR_ARGUMENT;
  if (n_models < 0 || !opt || (n_models > 0 && (!heights || !widths || !mat_offsets || !matrices || !ints_offsets ||
                                                !signs || !status || !result || !out_height || !rhs_out || !pos_out || !var_out)))
    return mfail(m, YALPS_ERR_ARGUMENT, "bad arguments");
  if (n_models == 0) return 0;
  const int world = (int)m->devices.size();
  const int64_t n = n_models;
  // packed offsets: inputs by height / width+height, outputs by the same plus 2*nints_i extra rows
  std::vector<int64_t> r_in(n + 1, 0), p_in(n + 1, 0), r_out(n + 1, 0), p_out(n + 1, 0);
  bool any_int = false;
  for (int64_t i = 0; i < n; i++) {
    const int64_t ni = ints_offsets[i + 1] - ints_offsets[i];
    if (ni < 0 || heights[i] < 1 || widths[i] < 1) return mfail(m, YALPS_ERR_ARGUMENT, "model %lld: bad shape or integer list", (long long)i);
    if (ni > 0 && !ints) return mfail(m, YALPS_ERR_ARGUMENT, "ints is null");
    any_int |= ni > 0;
    r_in[i + 1] = r_in[i] + heights[i];
    p_in[i + 1] = p_in[i] + heights[i] + widths[i];
    r_out[i + 1] = r_out[i] + heights[i] + 2 * ni;
    p_out[i + 1] = p_out[i] + heights[i] + widths[i] + 2 * ni;
  }
  // ---- all root LPs as one ragged batch over the ranks (src/YALPS.ts:79 per model)
  std::vector<int32_t> r_status(n);
  std::vector<double> r_value(n), r_rhs(r_in[n]);
  std::vector<int32_t> r_pos(p_in[n]), r_var(p_in[n]);
  std::vector<int64_t> packed_off(n + 1, 0);
  for (int64_t i = 0; i < n; i++) packed_off[i + 1] = packed_off[i] + (int64_t)heights[i] * widths[i];
  std::vector<double> finals;  // final root tableaus (applyCuts needs the whole matrix, src/branchAndCut.ts:28,38-42)
  if (any_int) finals.resize((size_t)packed_off[n]);
  {
    // yalps_solve_ragged writes matrices_out at the caller's mat_offsets: solve from a packed view so that the final
    // tableaus land packed as well
    std::vector<double> packed;
    const double *src = matrices;
    const int64_t *offs = mat_offsets;
    bool is_packed = true;
    for (int64_t i = 0; i < n && is_packed; i++) is_packed = mat_offsets[i] == packed_off[i];
    if (!is_packed) {
      packed.resize((size_t)packed_off[n]);
      for (int64_t i = 0; i < n; i++)
        std::memcpy(packed.data() + packed_off[i], matrices + mat_offsets[i], (size_t)heights[i] * widths[i] * 8);
      src = packed.data();
      offs = packed_off.data();
    }
    if (int rc = yalps_multi_solve_ragged(m, n, heights, widths, offs, src, opt, r_status.data(), r_value.data(), nullptr,
                                          r_rhs.data(), r_pos.data(), r_var.data(), any_int ? finals.data() : nullptr))
      return rc;
  }
  std::vector<int64_t> searches;
  for (int64_t i = 0; i < n; i++) {
    const int h = heights[i], w = widths[i];
    status[i] = r_status[i];
    result[i] = r_value[i];
    out_height[i] = h;
    std::memcpy(rhs_out + r_out[i], r_rhs.data() + r_in[i], (size_t)h * 8);
    std::memcpy(pos_out + p_out[i], r_pos.data() + p_in[i], (size_t)(h + w) * 4);
    std::memcpy(var_out + p_out[i], r_var.data() + p_in[i], (size_t)(h + w) * 4);
    if (ints_offsets[i + 1] > ints_offsets[i] && r_status[i] == YALPS_OPTIMAL) searches.push_back(i);  // src/YALPS.ts:81-89
  }
  if (searches.empty()) return 0;
  // ---- branch and cut: searches dealt dynamically to world * per_dev worker contexts
  const int per_dev = std::max(1, std::min<int>(searches_per_device > 0 ? searches_per_device : 4,
                                                (int)((searches.size() + world - 1) / world)));
  std::vector<MultiWorker *> pool;
  for (int w = 0; w < per_dev; w++)
    for (int r = 0; r < world; r++) {
      MultiWorker *mw = ensure_worker(m, r, w);
      if (!mw) return YALPS_ERR_CUDA;
      // the per_dev searches of one GPU run at the same time: each gets its share of the SMs (scheduler CTA included)
      mw->ctx->bnb_workers = std::max(8, mw->ctx->prop.multiProcessorCount / per_dev - 1);
      pool.push_back(mw);
    }
  std::mutex qmu;
  size_t next = 0;
  std::string first_error;
  auto take = [&]() -> int64_t {
    std::lock_guard<std::mutex> lk(qmu);
    return next < searches.size() ? searches[next++] : -1;
  };
  for (MultiWorker *mw : pool) {
    mw->submit([&, mw]() -> int {
      for (int64_t i = take(); i >= 0; i = take()) {
        const int h = heights[i], w = widths[i];
        const int32_t ni = (int32_t)(ints_offsets[i + 1] - ints_offsets[i]);
        int rc = yalps_bnb_set_root(mw->ctx, h, w, finals.data() + packed_off[i], r_pos.data() + p_in[i],
                                    r_var.data() + p_in[i], 2 * ni);
        if (!rc)
          rc = yalps_branch_and_cut(mw->ctx, ints + ints_offsets[i], ni, signs[i], r_value[i], opt, status + i, result + i,
                                    out_height + i, rhs_out + r_out[i], pos_out + p_out[i], var_out + p_out[i], nullptr);
        if (rc) {
          std::lock_guard<std::mutex> lk(qmu);
          if (first_error.empty()) first_error = "model " + std::to_string(i) + ": " + yalps_last_error(mw->ctx);
          return rc;
        }
      }
      return 0;
    });
  }
  int rc = 0;
  for (MultiWorker *mw : pool) {
    const int r = mw->wait();
    if (r && !rc) rc = r;
  }
  if (rc) m->error = first_error;
  return rc;
}

// ---- one large LP over all ranks: multigrid_kernel.cuh ---------------------------------------------
namespace {

struct LargeRank {
  double *M = nullptr, *value = nullptr, *rhs = nullptr;
  uint4 *colx = nullptr, *rowx = nullptr;  // exchange buffers of 16-byte {data, sequence} slots
  unsigned long long *sync = nullptr;  // grid barrier counter, 2 verdict ints, give-up record
  int *var = nullptr, *pos = nullptr, *status = nullptr, *hist = nullptr;
  long long *piv = nullptr;
  int Hl = 0, grid = 0;
  float ms = 0.f;
  int32_t h_status = 0;
  long long giveup[6] = {0, 0, 0, 0, 0, 0};
};

int enable_peers(yalps_multi *m) {
  if (m->peers_enabled) return 0;
  for (int a : m->gpus)
    for (int b : m->gpus) {
      if (a == b) continue;
      int can = 0;
      MCU(m, cudaDeviceCanAccessPeer(&can, a, b));
      if (!can) return mfail(m, YALPS_ERR_CUDA, "GPU %d cannot access the memory of GPU %d (no peer path)", a, b);
      MCU(m, cudaSetDevice(a));
      const cudaError_t e = cudaDeviceEnablePeerAccess(b, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
        return mfail(m, YALPS_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d -> %d) failed: %s", a, b, cudaGetErrorString(e));
      (void)cudaGetLastError();
    }
  m->peers_enabled = true;
  return 0;
}

}  // namespace

int yalps_multi_solve_large(yalps_multi *m, int32_t height, int32_t width, const double *matrix, const yalps_options *opt,
                            int32_t *status, double *value, int64_t *pivots, double *rhs_out, int32_t *pos_out,
                            int32_t *var_out, double *matrix_out, double *kernel_ms) {
  if (!m) return YALPS_ERR_ARGUMENT;
  if (height < 1 || width < 1 || !opt || !matrix) return mfail(m, YALPS_ERR_ARGUMENT, "bad arguments");
  const int G = (int)m->devices.size();
  const int H = height, W = width;
  if (G > kMaxGridRanks) return mfail(m, YALPS_ERR_ARGUMENT, "at most %d ranks share one LP", kMaxGridRanks);
  if ((long long)H + W >= (1LL << 31)) return mfail(m, YALPS_ERR_TOO_LARGE, "height + width must be < 2^31");
  if (int rc = enable_peers(m)) return rc;
  const GridSmem L(H, W);
  const int Hlmax = (H + G - 1) / G, Wpad = (W + 1) & ~1;
  std::vector<LargeRank> R(G);
  std::vector<int> ranks_on_gpu(G, 1);
  for (int r = 0; r < G; r++) {
    int cnt = 0;
    for (int q = 0; q < G; q++) cnt += m->devices[q] == m->devices[r];
    ranks_on_gpu[r] = cnt;
  }
  // ---- phase A: every rank allocates, clears its flags and takes its rows (global row r -> rank r % G, local r / G)
  int rc = run_on_ranks(m, [&](int r, yalps_ctx *ctx) -> int {
    LargeRank &k = R[r];
    CU(ctx, cudaSetDevice(ctx->device));
    if (L.total > (size_t)ctx->smem_optin)
      return fail(ctx, YALPS_ERR_TOO_LARGE, "tableau %dx%d: pivot row/column staging exceeds shared memory", H, W);
    int coop = 0;
    CU(ctx, cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device));
    if (!coop) return fail(ctx, YALPS_ERR_CUDA, "device does not support cooperative launches");
    CU(ctx, raise_smem_limit(ctx->device, (const void *)k_simplex_grid_multi, (int)L.total));
    int occ = 0;
    CU(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_simplex_grid_multi, kGridThreads, L.total));
    if (occ < 1) return fail(ctx, YALPS_ERR_TOO_LARGE, "grid kernel does not fit on an SM");
    k.Hl = r < H ? (H - r + G - 1) / G : 0;
    // the ranks that share a GPU share its SMs (all their kernels must be resident at once: they wait for one another)
    int grid = ctx->prop.multiProcessorCount * occ / ranks_on_gpu[r];
    {
      const long long items = (long long)std::max(k.Hl, 1) * ((W + kSegCols - 1) / kSegCols);
      grid = (int)std::min<long long>(grid, std::max(1LL, (items + 2 * kGridWarps - 1) / (2 * kGridWarps)));
    }
    if (const char *env = getenv("YALPS_GRID_CTAS")) grid = std::min(grid, std::max(1, atoi(env)));
    grid = std::max(grid, G);  // CTA d of a rank is its sender to rank d
    if (grid > ctx->prop.multiProcessorCount * occ / ranks_on_gpu[r])
      return fail(ctx, YALPS_ERR_TOO_LARGE, "%d ranks on GPU %d leave fewer than %d co-resident CTAs per rank", ranks_on_gpu[r],
                  ctx->device, G);
    k.grid = grid;
    void *p;
    int e;
    if ((e = dev_ensure(ctx, "lg_M", std::max<size_t>(1, (size_t)k.Hl) * W * 8, &p))) return e;
    k.M = (double *)p;
    const size_t colx_bytes = (size_t)2 * G * Hlmax * 16, rowx_bytes = (size_t)2 * Wpad * 16;
    if ((e = dev_ensure(ctx, "lg_colx", colx_bytes, &p))) return e;
    k.colx = (uint4 *)p;
    if ((e = dev_ensure(ctx, "lg_rowx", rowx_bytes, &p))) return e;
    k.rowx = (uint4 *)p;
    if ((e = dev_ensure(ctx, "lg_sync", (size_t)(kMaxGridRanks + kMaxRowParts + 12) * 8, &p))) return e;
    k.sync = (unsigned long long *)p;
    if ((e = dev_ensure(ctx, "lg_var", (size_t)(W + H) * 4, &p))) return e;
    k.var = (int *)p;
    if ((e = dev_ensure(ctx, "lg_pos", (size_t)(W + H) * 4, &p))) return e;
    k.pos = (int *)p;
    if ((e = dev_ensure(ctx, "lg_rhs", (size_t)H * 8, &p))) return e;
    k.rhs = (double *)p;
    if ((e = dev_ensure(ctx, "lg_res", 64, &p))) return e;
    k.status = (int *)p;
    k.value = (double *)((char *)p + 8);
    k.piv = (long long *)((char *)p + 16);
    if (opt->check_cycles) {
      if ((e = dev_ensure(ctx, "lg_hist", (size_t)2 * hist_capacity(opt) * sizeof(int), &p))) return e;
      k.hist = (int *)p;
    }
    cudaStream_t st = ctx->streams[0];
    CU(ctx, cudaMemsetAsync(k.sync, 0, (size_t)(kMaxGridRanks + kMaxRowParts + 12) * 8, st));
    CU(ctx, cudaMemsetAsync(k.colx, 0, colx_bytes, st));  // sequence numbers of an earlier call must not look current
    CU(ctx, cudaMemsetAsync(k.rowx, 0, rowx_bytes, st));
    if (k.Hl > 0)
      CU(ctx, cudaMemcpy2DAsync(k.M, (size_t)W * 8, matrix + (size_t)r * W, (size_t)G * W * 8, (size_t)W * 8, k.Hl,
                                cudaMemcpyHostToDevice, st));
    CU(ctx, cudaStreamSynchronize(st));
    return 0;
  });
  if (rc) return rc;
  // ---- phase B: one cooperative kernel per rank; the kernels exchange pivot rows and columns through peer memory
  int row_parts = kMaxRowParts;
  for (int r = 0; r < G; r++) row_parts = std::min(row_parts, std::max(1, R[r].grid / G));
  if (const char *env = getenv("YALPS_LARGE_ROW_PARTS")) row_parts = std::max(1, std::min(row_parts, atoi(env)));
  // Every rank launches before any rank issues a call that waits for its own kernel (a device-to-host copy into
  // pageable memory blocks inside the driver; a rank whose launch queued up behind such a call would never start, and
  // the kernels wait for one another).
  std::atomic<int> launched{0};
  std::vector<cudaEvent_t> ev0(G), ev1(G);
  std::vector<long long *> d_giveup(G, nullptr);
  auto launch = [&](int r, yalps_ctx *ctx) -> int {
    LargeRank &k = R[r];
    CU(ctx, cudaSetDevice(ctx->device));
    MultiGridArgs a{};
    a.M = k.M;
    a.H = H;
    a.W = W;
    a.G = G;
    a.g = r;
    a.Hl = k.Hl;
    a.Hlmax = Hlmax;
    a.Wpad = Wpad;
    a.row_parts = row_parts;
    for (int q = 0; q < G; q++) {
      a.colx[q] = R[q].colx;
      a.rowx[q] = R[q].rowx;
    }
    a.barrier = k.sync + kMaxGridRanks + kMaxRowParts;
    a.flags = (int *)(k.sync + kMaxGridRanks + kMaxRowParts + 2);
    a.var = k.var;
    a.pos_out = k.pos;
    a.rhs_out = k.rhs;
    a.status = k.status;
    a.value = k.value;
    a.pivots = k.piv;
    a.precision = opt->precision;
    a.max_pivots = opt->max_pivots;
    a.check_cycles = opt->check_cycles ? 1 : 0;
    a.hist = k.hist;
    a.hist_cap = hist_capacity(opt);
    a.rows_out = (r == 0 && ctx->d_rows && !ctx->rows_per_lp) ? ctx->d_rows : nullptr;
    a.spin_limit = (long long)(2.0 * ctx->prop.clockRate * 1e3);  // ~2 s of SM clock
    a.giveup = (long long *)(k.sync + kMaxGridRanks + kMaxRowParts + 4);
    cudaStream_t st = ctx->streams[0];
    d_giveup[r] = a.giveup;
    CU(ctx, cudaEventCreate(&ev0[r]));  // (the ctx's own events are created without timing)
    CU(ctx, cudaEventCreate(&ev1[r]));
    CU(ctx, cudaEventRecord(ev0[r], st));
    void *params[] = {&a};
    CU(ctx, cudaLaunchCooperativeKernel((void *)k_simplex_grid_multi, dim3(k.grid), dim3(kGridThreads), params, L.total, st));
    ctx->launches++;
    CU(ctx, cudaEventRecord(ev1[r], st));
    return 0;
  };
  rc = run_on_ranks(m, [&](int r, yalps_ctx *ctx) -> int {
    LargeRank &k = R[r];
    const int lrc = launch(r, ctx);
    launched.fetch_add(1);
    while (launched.load() < G) std::this_thread::yield();
    if (lrc) return lrc;
    cudaStream_t st = ctx->streams[0];
    CU(ctx, cudaMemcpyAsync(&k.h_status, k.status, 4, cudaMemcpyDeviceToHost, st));
    CU(ctx, cudaMemcpyAsync(k.giveup, d_giveup[r], sizeof k.giveup, cudaMemcpyDeviceToHost, st));
    if (r == 0) {
      if (status) CU(ctx, cudaMemcpyAsync(status, k.status, 4, cudaMemcpyDeviceToHost, st));
      if (value) CU(ctx, cudaMemcpyAsync(value, k.value, 8, cudaMemcpyDeviceToHost, st));
      if (pivots) CU(ctx, cudaMemcpyAsync(pivots, k.piv, 16, cudaMemcpyDeviceToHost, st));
      if (rhs_out) CU(ctx, cudaMemcpyAsync(rhs_out, k.rhs, (size_t)H * 8, cudaMemcpyDeviceToHost, st));
      if (pos_out) CU(ctx, cudaMemcpyAsync(pos_out, k.pos, (size_t)(W + H) * 4, cudaMemcpyDeviceToHost, st));
      if (var_out) CU(ctx, cudaMemcpyAsync(var_out, k.var, (size_t)(W + H) * 4, cudaMemcpyDeviceToHost, st));
    }
    if (matrix_out && k.Hl > 0)
      CU(ctx, cudaMemcpy2DAsync(matrix_out + (size_t)r * W, (size_t)G * W * 8, k.M, (size_t)W * 8, (size_t)W * 8, k.Hl,
                                cudaMemcpyDeviceToHost, st));
    CU(ctx, cudaStreamSynchronize(st));
    CU(ctx, cudaEventElapsedTime(&k.ms, ev0[r], ev1[r]));
    cudaEventDestroy(ev0[r]);
    cudaEventDestroy(ev1[r]);
    return 0;
  });
  if (rc) return rc;
  float ms = 0.f;
  for (int r = 0; r < G; r++) {
    ms = std::max(ms, R[r].ms);
    if (R[r].h_status == ST_ERR_PEER)
      return mfail(m, YALPS_ERR_CUDA,
                   "rank %d (GPU %d) gave up waiting for a peer (wait %lld, exchange %lld, phase %lld, CTA %lld, after %lld + %lld "
                   "pivots): the %d kernels were not running at the same time",
                   r, m->devices[r], R[r].giveup[0], R[r].giveup[1], R[r].giveup[2], R[r].giveup[3], R[r].giveup[4],
                   R[r].giveup[5], G);
    if (R[r].h_status != R[0].h_status)
      return mfail(m, YALPS_ERR_CUDA, "ranks disagree on the outcome (rank 0: %d, rank %d: %d)", R[0].h_status, r, R[r].h_status);
  }
  if (R[0].h_status == ST_ERR_HISTORY)
    return mfail(m, YALPS_ERR_HISTORY, "checkCycles history exhausted (more than 262144 pivots in a phase)");
  if (kernel_ms) *kernel_ms = ms;
  return 0;
}

}  // extern "C"
