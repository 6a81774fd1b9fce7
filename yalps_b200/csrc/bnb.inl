// bnb.inl -- branch and cut on top of the node-wave kernel (included by yalps_b200.cu).
//
// Control flow follows src/branchAndCut.ts:89-176 statement by statement; what changes is *when* the
// node LPs are evaluated.  Every node is root-optimal-tableau + its cut list (applyCuts always starts
// from the root, :126), so nodes are independent given their cuts.  The driver therefore peeks at the
// next `wave` branches the reference heap would pop, evaluates the ones it has not seen yet in one
// device launch (K3 assembly + K1/K2 simplex), caches the results by branch id, and then replays the
// reference loop consuming cached results in true pop order.  A child pushed during the replay that
// outranks the speculated branches simply triggers the next wave.  The incumbent, pruning, tolerance
// early-exit and iteration counting are therefore exactly the reference's.

namespace {

struct Cut {
  double sign;
  int32_t variable;
  double value;
};

struct Branch {
  double eval;
  std::vector<Cut> cuts;
  int64_t id;
};

// npm heap@0.2.7 (package.json:157) == CPython heapq; comparator x[0]-y[0] (src/branchAndCut.ts:100).
struct BranchHeap {
  std::vector<std::shared_ptr<Branch>> a;
  static bool lt(const Branch &x, const Branch &y) { return x.eval - y.eval < 0; }
  void sift_toward_root(size_t start, size_t pos) {
    auto item = a[pos];
    while (pos > start) {
      const size_t parent = (pos - 1) >> 1;
      if (!lt(*item, *a[parent])) break;
      a[pos] = a[parent];
      pos = parent;
    }
    a[pos] = item;
  }
  void sift_to_leaf(size_t pos) {
    const size_t end = a.size(), start = pos;
    auto item = a[pos];
    size_t child = 2 * pos + 1;
    while (child < end) {
      const size_t right = child + 1;
      if (right < end && !lt(*a[child], *a[right])) child = right;
      a[pos] = a[child];
      pos = child;
      child = 2 * pos + 1;
    }
    a[pos] = item;
    sift_toward_root(start, pos);
  }
  void push(std::shared_ptr<Branch> b) {
    a.push_back(std::move(b));
    sift_toward_root(0, a.size() - 1);
  }
  std::shared_ptr<Branch> pop() {
    auto last = a.back();
    a.pop_back();
    if (a.empty()) return last;
    auto top = a[0];
    a[0] = last;
    sift_to_leaf(0);
    return top;
  }
  bool empty() const { return a.empty(); }
};

struct NodeResult {
  int32_t status;
  double value;
  int64_t pivots;
  int32_t height;
  std::vector<double> rhs;
  std::vector<int32_t> pos, var;
};

double host_js_round(double x) {
  if (!(std::fabs(x) < 4503599627370496.0)) return x;
  double r = std::floor(x);
  if (x - r >= 0.5) r += 1.0;
  if (r == 0.0 && std::signbit(x)) r = -0.0;
  return r;
}

// mostFractionalVar, src/branchAndCut.ts:64-85, on (rhs column, positionOfVariable)
void most_fractional(const double *rhs, const int32_t *pos, int W, const int32_t *ints, int nints, int32_t *variable,
                     double *value, double *frac) {
  double highest = 0.0, val_out = 0.0;
  int32_t v_out = 0;
  for (int i = 0; i < nints; i++) {
    const int32_t iv = ints[i];
    const int32_t row = pos[iv] - W;
    if (row < 0) continue;
    const double val = rhs[row];
    const double fr = std::fabs(val - host_js_round(val));
    if (fr > highest) {
      highest = fr;
      v_out = iv;
      val_out = val;
    }
  }
  *variable = v_out;
  *value = val_out;
  *frac = highest;
}

double now_ms() {
  using namespace std::chrono;
  return std::floor((double)duration_cast<microseconds>(system_clock::now().time_since_epoch()).count() / 1000.0);
}

int upload_root_host(yalps_ctx *ctx, int32_t height, int32_t width, const double *matrix, const int32_t *pos,
                     const int32_t *var, int32_t max_extra_rows) {
  Root &R = ctx->root;
  const size_t cells = (size_t)height * width;
  const size_t nv = (size_t)height + width;
  auto ensure = [&](DevBuf &b, size_t bytes) -> int {
    if (b.cap < bytes) {
      if (b.p) CU(ctx, cudaFree(b.p));
      b.p = nullptr;
      b.cap = 0;
      CU(ctx, cudaMalloc(&b.p, bytes));
      b.cap = bytes;
    }
    return 0;
  };
  if (int rc = ensure(R.m, cells * 8)) return rc;
  if (int rc = ensure(R.pos, nv * 4)) return rc;
  if (int rc = ensure(R.var, nv * 4)) return rc;
  cudaStream_t st = ctx->streams[0];
  CU(ctx, cudaMemcpyAsync(R.m.p, matrix, cells * 8, cudaMemcpyHostToDevice, st));
  CU(ctx, cudaMemcpyAsync(R.pos.p, pos, nv * 4, cudaMemcpyHostToDevice, st));
  CU(ctx, cudaMemcpyAsync(R.var.p, var, nv * 4, cudaMemcpyHostToDevice, st));
  CU(ctx, cudaStreamSynchronize(st));
  R.H = height;
  R.W = width;
  R.max_extra = max_extra_rows;
  R.h_rhs.resize(height);
  for (int r = 0; r < height; r++) R.h_rhs[r] = matrix[(size_t)r * width];
  R.h_pos.assign(pos, pos + nv);
  R.h_var.assign(var, var + nv);
  R.valid = true;
  return 0;
}

}  // namespace

extern "C" {

int yalps_bnb_set_root(yalps_ctx *ctx, int32_t height, int32_t width, const double *matrix, const int32_t *pos,
                       const int32_t *var, int32_t max_extra_rows) {
  if (!ctx) return YALPS_ERR_ARGUMENT;
  if (height < 1 || width < 1 || !matrix || !pos || !var || max_extra_rows < 0)
    return fail(ctx, YALPS_ERR_ARGUMENT, "bad arguments");
  if (((long long)height + max_extra_rows) * width >= (1LL << 31))
    return fail(ctx, YALPS_ERR_TOO_LARGE, "(height+extra)*width must be < 2^31");
  CU(ctx, cudaSetDevice(ctx->device));
  return upload_root_host(ctx, height, width, matrix, pos, var, max_extra_rows);
}

int yalps_bnb_set_wave(yalps_ctx *ctx, int32_t wave) {
  if (!ctx || wave < 1) return YALPS_ERR_ARGUMENT;
  ctx->wave = wave;
  return 0;
}

int yalps_bnb_solve_nodes(yalps_ctx *ctx, int64_t n, const int32_t *cut_offsets, const double *cut_sign,
                          const int32_t *cut_var, const double *cut_value, const yalps_options *opt, int32_t *status,
                          double *value, int64_t *pivots, double *rhs_out, int32_t *pos_out, int32_t *var_out,
                          double *matrices_out) {
  if (!ctx) return YALPS_ERR_ARGUMENT;
  Root &R = ctx->root;
  if (!R.valid) return fail(ctx, YALPS_ERR_ARGUMENT, "no root tableau: call yalps_bnb_set_root first");
  if (n < 0 || !opt || (n > 0 && !cut_offsets)) return fail(ctx, YALPS_ERR_ARGUMENT, "bad arguments");
  if (n == 0) return 0;
  CU(ctx, cudaSetDevice(ctx->device));
  const int W = R.W;
  int maxcuts = 0;
  for (int64_t j = 0; j < n; j++) {
    const int k = cut_offsets[j + 1] - cut_offsets[j];
    if (k < 0) return fail(ctx, YALPS_ERR_ARGUMENT, "cut_offsets must be non-decreasing");
    maxcuts = std::max(maxcuts, k);
  }
  const int ncut_total = cut_offsets[n];
  for (int i = 0; i < ncut_total; i++)
    if (cut_var[i] < 0 || cut_var[i] >= R.W + R.H) return fail(ctx, YALPS_ERR_ARGUMENT, "cut %d names variable %d", i, cut_var[i]);
  const int Hcap = R.H + maxcuts;
  if ((long long)Hcap * W >= (1LL << 31)) return fail(ctx, YALPS_ERR_TOO_LARGE, "(height+cuts)*width must be < 2^31");
  LaunchPlan plan;
  if (int rc = plan_launch(ctx, n, Hcap, W, opt->check_cycles != 0, &plan)) return rc;

  cudaStream_t st = ctx->streams[0];
  void *d_off, *d_sign, *d_var, *d_val, *d_status, *d_value, *d_piv, *d_rhs, *d_pos, *d_vr;
  int rc;
  const size_t nc = (size_t)std::max(ncut_total, 1);
  if ((rc = dev_ensure(ctx, "nd_off", (size_t)(n + 1) * 4, &d_off))) return rc;
  if ((rc = dev_ensure(ctx, "nd_sign", nc * 8, &d_sign))) return rc;
  if ((rc = dev_ensure(ctx, "nd_var", nc * 4, &d_var))) return rc;
  if ((rc = dev_ensure(ctx, "nd_val", nc * 8, &d_val))) return rc;
  if ((rc = dev_ensure(ctx, "nd_status", (size_t)n * 4, &d_status))) return rc;
  if ((rc = dev_ensure(ctx, "nd_value", (size_t)n * 8, &d_value))) return rc;
  if ((rc = dev_ensure(ctx, "nd_piv", (size_t)n * 16, &d_piv))) return rc;
  if ((rc = dev_ensure(ctx, "nd_rhs", (size_t)n * Hcap * 8, &d_rhs))) return rc;
  if ((rc = dev_ensure(ctx, "nd_pos", (size_t)n * (W + Hcap) * 4, &d_pos))) return rc;
  if ((rc = dev_ensure(ctx, "nd_vr", (size_t)n * (W + Hcap) * 4, &d_vr))) return rc;
  void *d_work = nullptr, *d_out = nullptr;
  const size_t mat_bytes = (size_t)n * Hcap * W * 8;
  if (!plan.resident) {
    if ((rc = dev_ensure(ctx, "nd_work", mat_bytes, &d_work))) return rc;
    d_out = matrices_out ? d_work : nullptr;
  } else if (matrices_out) {
    if ((rc = dev_ensure(ctx, "nd_out", mat_bytes, &d_out))) return rc;
  }
  CU(ctx, cudaMemcpyAsync(d_off, cut_offsets, (size_t)(n + 1) * 4, cudaMemcpyHostToDevice, st));
  if (ncut_total > 0) {
    CU(ctx, cudaMemcpyAsync(d_sign, cut_sign, (size_t)ncut_total * 8, cudaMemcpyHostToDevice, st));
    CU(ctx, cudaMemcpyAsync(d_var, cut_var, (size_t)ncut_total * 4, cudaMemcpyHostToDevice, st));
    CU(ctx, cudaMemcpyAsync(d_val, cut_value, (size_t)ncut_total * 8, cudaMemcpyHostToDevice, st));
  }
  BatchArgs a{};
  a.n = n;
  a.mode = kModeNodes;
  a.H = R.H;
  a.W = W;
  a.Hcap = Hcap;
  a.Wcap = W;
  a.work = (double *)d_work;
  a.mat_out = (double *)d_out;
  a.status = (int *)d_status;
  a.value = (double *)d_value;
  a.pivots = (long long *)d_piv;
  a.rhs_out = (double *)d_rhs;
  a.pos_out = (int *)d_pos;
  a.var_out = (int *)d_vr;
  a.root = (const double *)R.m.p;
  a.root_pos = (const int *)R.pos.p;
  a.root_var = (const int *)R.var.p;
  a.cut_off = (const int *)d_off;
  a.cut_sign = (const double *)d_sign;
  a.cut_var = (const int *)d_var;
  a.cut_val = (const double *)d_val;
  fill_options(a, opt);
  if (use_grid_path(ctx, n, plan)) {
    // few large nodes: assemble them in HBM (K3), then give each node the whole grid (K4)
    if (!d_work) {
      if ((rc = dev_ensure(ctx, "nd_work", mat_bytes, &d_work))) return rc;
      if (matrices_out) d_out = d_work;
    }
    const int gx = std::max(1, std::min(ctx->prop.multiProcessorCount * 4, (int)(((size_t)R.H * W + 255) / 256)));
    const int gy = (int)std::min<int64_t>(n, 65535);
    k_assemble_nodes<<<dim3(gx, gy), 256, 0, st>>>(n, R.H, W, Hcap, (const double *)R.m.p, (const int *)R.pos.p,
                                                   (const int *)d_off, (const double *)d_sign, (const int *)d_var,
                                                   (const double *)d_val, (double *)d_work);
    CU(ctx, cudaGetLastError());
    ctx->launches++;
    for (int64_t j = 0; j < n; j++) {
      const int hj = R.H + (cut_offsets[j + 1] - cut_offsets[j]);
      if ((rc = launch_grid(ctx, hj, W, (double *)d_work + (size_t)j * Hcap * W, opt, (int *)d_status + j,
                            (double *)d_value + j, (long long *)d_piv + 2 * j, (double *)d_rhs + (size_t)j * Hcap,
                            (int *)d_pos + (size_t)j * (W + Hcap), (int *)d_vr + (size_t)j * (W + Hcap), st,
                            (const int *)R.var.p, W + R.H)))
        return rc;
    }
  } else if ((rc = launch_simplex(ctx, plan, a, "nd", st))) {
    return rc;
  }
  if (status) CU(ctx, cudaMemcpyAsync(status, d_status, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
  if (value) CU(ctx, cudaMemcpyAsync(value, d_value, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
  if (pivots) CU(ctx, cudaMemcpyAsync(pivots, d_piv, (size_t)n * 16, cudaMemcpyDeviceToHost, st));
  if (rhs_out) CU(ctx, cudaMemcpyAsync(rhs_out, d_rhs, (size_t)n * Hcap * 8, cudaMemcpyDeviceToHost, st));
  if (pos_out) CU(ctx, cudaMemcpyAsync(pos_out, d_pos, (size_t)n * (W + Hcap) * 4, cudaMemcpyDeviceToHost, st));
  if (var_out) CU(ctx, cudaMemcpyAsync(var_out, d_vr, (size_t)n * (W + Hcap) * 4, cudaMemcpyDeviceToHost, st));
  if (matrices_out) CU(ctx, cudaMemcpyAsync(matrices_out, d_out, mat_bytes, cudaMemcpyDeviceToHost, st));
  CU(ctx, cudaStreamSynchronize(st));
  return check_device_status(ctx, status, n);
}

int yalps_branch_and_cut(yalps_ctx *ctx, const int32_t *ints, int32_t nints, double sign, double init_result,
                         const yalps_options *opt, int32_t *status, double *result, int32_t *out_height,
                         double *rhs_out, int32_t *pos_out, int32_t *var_out, int64_t *stats) {
  if (!ctx) return YALPS_ERR_ARGUMENT;
  Root &R = ctx->root;
  if (!R.valid) return fail(ctx, YALPS_ERR_ARGUMENT, "no root tableau: call yalps_bnb_set_root first");
  if (!opt || !status || !result || !out_height || !rhs_out || !pos_out || !var_out || nints < 0 || (nints && !ints))
    return fail(ctx, YALPS_ERR_ARGUMENT, "bad arguments");
  const int W = R.W, H = R.H;
  const double precision = opt->precision;
  int64_t st_nodes = 0, st_pivots = 0, st_maxcuts = 0, st_maxheap = 0, st_waves = 0, st_devnodes = 0;

  auto write_best = [&](const double *rhs, const int32_t *pos, const int32_t *var, int h) {
    *out_height = h;
    std::memcpy(rhs_out, rhs, sizeof(double) * (size_t)h);
    std::memcpy(pos_out, pos, sizeof(int32_t) * (size_t)(W + h));
    std::memcpy(var_out, var, sizeof(int32_t) * (size_t)(W + h));
  };
  auto write_stats = [&]() {
    if (!stats) return;
    stats[0] = st_nodes;
    stats[1] = st_pivots;
    stats[2] = st_maxcuts;
    stats[3] = st_maxheap;
    stats[4] = st_waves;
    stats[5] = st_devnodes;
    stats[6] = stats[7] = 0;
  };

  int32_t init_var;
  double init_val, init_frac;
  most_fractional(R.h_rhs.data(), R.h_pos.data(), W, ints, nints, &init_var, &init_val, &init_frac);
  if (init_frac <= precision) {  // :98
    *status = YALPS_OPTIMAL;
    *result = init_result;
    write_best(R.h_rhs.data(), R.h_pos.data(), R.h_var.data(), H);
    write_stats();
    return 0;
  }
  if (2 * nints > R.max_extra)
    return fail(ctx, YALPS_ERR_ARGUMENT, "root was set with max_extra_rows=%d < 2*|integers|=%d", R.max_extra, 2 * nints);

  int64_t next_id = 0;
  BranchHeap heap;
  {
    auto b1 = std::make_shared<Branch>();
    b1->eval = init_result;
    b1->cuts = {Cut{-1.0, init_var, std::ceil(init_val)}};
    b1->id = next_id++;
    auto b2 = std::make_shared<Branch>();
    b2->eval = init_result;
    b2->cuts = {Cut{1.0, init_var, std::floor(init_val)}};
    b2->id = next_id++;
    heap.push(b1);
    heap.push(b2);
  }

  std::unordered_map<int64_t, NodeResult> cache;
  const double threshold = init_result * (1.0 - sign * opt->tolerance);
  const double stop_time = opt->timeout_ms + now_ms();
  bool timedout = now_ms() >= stop_time;
  bool found = false;
  double best_eval = std::numeric_limits<double>::infinity();
  NodeResult best;
  double iter = 0;

  // wave staging
  std::vector<int32_t> w_off;
  std::vector<double> w_sign, w_val;
  std::vector<int32_t> w_var;
  std::vector<int32_t> w_status;
  std::vector<double> w_value;
  std::vector<int64_t> w_piv;
  std::vector<double> w_rhs;
  std::vector<int32_t> w_pos, w_vr;

  auto run_wave = [&](const std::shared_ptr<Branch> &needed) -> int {
    // the needed branch first, then the branches the heap would pop next (on a copy of the heap)
    std::vector<std::shared_ptr<Branch>> wave{needed};
    BranchHeap peek = heap;
    while ((int)wave.size() < ctx->wave && !peek.empty()) {
      auto b = peek.pop();
      if (b->eval > best_eval) break;  // would be pruned (:124)
      if (cache.find(b->id) == cache.end()) wave.push_back(b);
    }
    const int64_t n = (int64_t)wave.size();
    w_off.assign(1, 0);
    w_sign.clear();
    w_var.clear();
    w_val.clear();
    int maxcuts = 0;
    for (auto &b : wave) {
      for (auto &c : b->cuts) {
        w_sign.push_back(c.sign);
        w_var.push_back(c.variable);
        w_val.push_back(c.value);
      }
      w_off.push_back((int32_t)w_sign.size());
      maxcuts = std::max(maxcuts, (int)b->cuts.size());
    }
    const int Hcap = H + maxcuts;
    w_status.resize(n);
    w_value.resize(n);
    w_piv.resize(2 * n);
    w_rhs.resize((size_t)n * Hcap);
    w_pos.resize((size_t)n * (W + Hcap));
    w_vr.resize((size_t)n * (W + Hcap));
    if (int rc = yalps_bnb_solve_nodes(ctx, n, w_off.data(), w_sign.data(), w_var.data(), w_val.data(), opt,
                                       w_status.data(), w_value.data(), w_piv.data(), w_rhs.data(), w_pos.data(),
                                       w_vr.data(), nullptr))
      return rc;
    st_waves++;
    st_devnodes += n;
    for (int64_t j = 0; j < n; j++) {
      NodeResult nr;
      nr.status = w_status[j];
      nr.value = w_value[j];
      nr.pivots = w_piv[2 * j] + w_piv[2 * j + 1];
      nr.height = H + (int)wave[j]->cuts.size();
      nr.rhs.assign(w_rhs.begin() + (size_t)j * Hcap, w_rhs.begin() + (size_t)j * Hcap + nr.height);
      nr.pos.assign(w_pos.begin() + (size_t)j * (W + Hcap), w_pos.begin() + (size_t)j * (W + Hcap) + W + nr.height);
      nr.var.assign(w_vr.begin() + (size_t)j * (W + Hcap), w_vr.begin() + (size_t)j * (W + Hcap) + W + nr.height);
      cache.emplace(wave[j]->id, std::move(nr));
    }
    return 0;
  };

  while (iter < opt->max_iterations && !heap.empty() && best_eval >= threshold && !timedout) {  // :122
    st_maxheap = std::max<int64_t>(st_maxheap, (int64_t)heap.a.size());
    auto br = heap.pop();
    if (br->eval > best_eval) break;  // :124

    auto it = cache.find(br->id);
    if (it == cache.end()) {
      if (int rc = run_wave(br)) return rc;
      it = cache.find(br->id);
    }
    NodeResult nr = std::move(it->second);
    cache.erase(it);
    st_nodes++;
    st_pivots += nr.pivots;
    st_maxcuts = std::max<int64_t>(st_maxcuts, (int64_t)br->cuts.size());

    if (nr.status == YALPS_OPTIMAL && nr.value < best_eval) {  // :130
      int32_t variable;
      double value, frac;
      most_fractional(nr.rhs.data(), nr.pos.data(), W, ints, nints, &variable, &value, &frac);
      if (frac <= precision) {  // integer solution, new incumbent (:132-139)
        found = true;
        best_eval = nr.value;
        best = std::move(nr);
      } else {  // branch (:141-156)
        auto upper = std::make_shared<Branch>();
        auto lower = std::make_shared<Branch>();
        for (const Cut &cut : br->cuts) {
          if (cut.variable == variable) {
            if (cut.sign < 0)
              lower->cuts.push_back(cut);
            else
              upper->cuts.push_back(cut);
          } else {
            upper->cuts.push_back(cut);
            lower->cuts.push_back(cut);
          }
        }
        lower->cuts.push_back(Cut{1.0, variable, std::floor(value)});
        upper->cuts.push_back(Cut{-1.0, variable, std::ceil(value)});
        upper->eval = lower->eval = nr.value;
        upper->id = next_id++;
        lower->id = next_id++;
        heap.push(upper);
        heap.push(lower);
      }
    }
    timedout = now_ms() >= stop_time;  // :162
    iter++;
  }

  const bool unfinished = (timedout || iter >= opt->max_iterations) && !heap.empty() && best_eval >= threshold;  // :167
  *status = unfinished ? YALPS_TIMEDOUT : (!found ? YALPS_INFEASIBLE : YALPS_OPTIMAL);
  *result = found ? best_eval : std::numeric_limits<double>::quiet_NaN();
  if (found)
    write_best(best.rhs.data(), best.pos.data(), best.var.data(), best.height);
  else
    write_best(R.h_rhs.data(), R.h_pos.data(), R.h_var.data(), H);  // bestTableau = root (:119)
  write_stats();
  return 0;
}

int yalps_solve(yalps_ctx *ctx, int32_t height, int32_t width, const double *matrix, const int32_t *ints,
                int32_t nints, double sign, const yalps_options *opt, int32_t *status, double *result,
                int32_t *out_height, double *rhs_out, int32_t *pos_out, int32_t *var_out, int32_t *root_status,
                double *root_value, int64_t *root_pivots, int64_t *stats) {
  if (!ctx) return YALPS_ERR_ARGUMENT;
  if (height < 1 || width < 1 || !matrix || !opt || !status || !result || !out_height || !rhs_out || !pos_out ||
      !var_out || nints < 0 || (nints && !ints))
    return fail(ctx, YALPS_ERR_ARGUMENT, "bad arguments");
  if (stats) std::memset(stats, 0, sizeof(int64_t) * 8);
  int32_t st = 0;
  double val = 0;
  int64_t piv[2] = {0, 0};
  const size_t cells = (size_t)height * width;
  const bool milp = nints > 0;
  std::vector<double> final_m;
  if (milp) final_m.resize(cells);
  // root LP (src/YALPS.ts:79)
  if (int rc = yalps_solve_batch(ctx, 1, height, width, matrix, opt, &st, &val, piv, rhs_out, pos_out, var_out,
                                 milp ? final_m.data() : nullptr))
    return rc;
  if (root_status) *root_status = st;
  if (root_value) *root_value = val;
  if (root_pivots) {
    root_pivots[0] = piv[0];
    root_pivots[1] = piv[1];
  }
  *out_height = height;
  if (!milp || st != YALPS_OPTIMAL) {  // src/YALPS.ts:81-86
    *status = st;
    *result = val;
    return 0;
  }
  if (int rc = yalps_bnb_set_root(ctx, height, width, final_m.data(), pos_out, var_out, 2 * nints)) return rc;
  return yalps_branch_and_cut(ctx, ints, nints, sign, val, opt, status, result, out_height, rhs_out, pos_out, var_out,
                              stats);
}

}  // extern "C"
