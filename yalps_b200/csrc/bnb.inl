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
};

// npm heap@0.2.7 (package.json:157) == CPython heapq; comparator x[0]-y[0] (src/branchAndCut.ts:100).
// The heap holds indices into the branch arena (copying it for the wave look-ahead is a memcpy).
struct BranchHeap {
  const std::vector<Branch> *arena;
  std::vector<int32_t> a;
  bool lt(int32_t x, int32_t y) const { return (*arena)[x].eval - (*arena)[y].eval < 0; }
  void sift_toward_root(size_t start, size_t pos) {
    const int32_t item = a[pos];
    while (pos > start) {
      const size_t parent = (pos - 1) >> 1;
      if (!lt(item, a[parent])) break;
      a[pos] = a[parent];
      pos = parent;
    }
    a[pos] = item;
  }
  void sift_to_leaf(size_t pos) {
    const size_t end = a.size(), start = pos;
    const int32_t item = a[pos];
    size_t child = 2 * pos + 1;
    while (child < end) {
      const size_t right = child + 1;
      if (right < end && !lt(a[child], a[right])) child = right;
      a[pos] = a[child];
      pos = child;
      child = 2 * pos + 1;
    }
    a[pos] = item;
    sift_toward_root(start, pos);
  }
  void push(int32_t b) {
    a.push_back(b);
    sift_toward_root(0, a.size() - 1);
  }
  int32_t pop() {
    const int32_t last = a.back();
    a.pop_back();
    if (a.empty()) return last;
    const int32_t top = a[0];
    a[0] = last;
    sift_to_leaf(0);
    return top;
  }
  bool empty() const { return a.empty(); }
};

// Results of one device wave, kept as downloaded; nodes are read in place (no per-node copies).
struct WaveBuf {
  int n = 0, Hcap = 0;
  std::vector<int32_t> ids;
  std::vector<int32_t> status;
  std::vector<double> value;
  std::vector<int64_t> piv;
  std::vector<double> rhs;
  std::vector<int32_t> pos, var;
  size_t bytes() const { return rhs.size() * 8 + (pos.size() + var.size()) * 4; }
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
  {
    const size_t step = std::max<size_t>(1, cells / 65536);
    size_t seen = 0, nz = 0;
    for (size_t k = 0; k < cells; k += step, seen++) nz += matrix[k] != 0.0;
    R.density = seen ? (double)nz / (double)seen : 1.0;
  }
  R.h_pos.assign(pos, pos + nv);
  R.h_var.assign(var, var + nv);
  R.valid = true;
  return 0;
}

// Root tableau already on the device (the root LP was just solved there): device-to-device copy, no PCIe round trip
// of the matrix.  rhs / pos / var are the host copies the driver needs anyway.
int adopt_root_device(yalps_ctx *ctx, int32_t height, int32_t width, const double *d_matrix, const double *rhs,
                      const int32_t *pos, const int32_t *var, int32_t max_extra_rows) {
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
  CU(ctx, cudaMemcpyAsync(R.m.p, d_matrix, cells * 8, cudaMemcpyDeviceToDevice, st));
  CU(ctx, cudaMemcpyAsync(R.pos.p, pos, nv * 4, cudaMemcpyHostToDevice, st));
  CU(ctx, cudaMemcpyAsync(R.var.p, var, nv * 4, cudaMemcpyHostToDevice, st));
  void *hp, *dp;
  if (int rc = pin_ensure(ctx, "density_probe", 64, &hp)) return rc;
  if (int rc = pin_device_ptr(ctx, hp, &dp)) return rc;
  ((int *)hp)[0] = ((int *)hp)[1] = 0;
  k_sample_density<<<1, 256, 0, st>>>((const double *)R.m.p, (long long)cells, std::max<long long>(1, (long long)cells / 65536), (int *)dp);
  CU(ctx, cudaGetLastError());
  ctx->launches++;
  CU(ctx, cudaStreamSynchronize(st));
  R.density = ((int *)hp)[0] ? (double)((int *)hp)[1] / (double)((int *)hp)[0] : 1.0;
  R.H = height;
  R.W = width;
  R.max_extra = max_extra_rows;
  R.h_rhs.assign(rhs, rhs + height);
  R.h_pos.assign(pos, pos + nv);
  R.h_var.assign(var, var + nv);
  R.valid = true;
  return 0;
}

int bnb_solve_nodes_impl(yalps_ctx *ctx, int64_t n, const int32_t *cut_offsets, const double *cut_sign,
                         const int32_t *cut_var, const double *cut_value, const yalps_options *opt, int32_t *status,
                         double *value, int64_t *pivots, double *rhs_out, int32_t *pos_out, int32_t *var_out,
                         double *matrices_out, int min_maxcuts);

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

int yalps_bnb_set_mode(yalps_ctx *ctx, int32_t mode) {
  if (!ctx || mode < 0 || mode > 2) return YALPS_ERR_ARGUMENT;
  ctx->bnb_mode = mode;
  return 0;
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
  return bnb_solve_nodes_impl(ctx, n, cut_offsets, cut_sign, cut_var, cut_value, opt, status, value, pivots, rhs_out,
                              pos_out, var_out, matrices_out, 0);
}

}  // extern "C"

namespace {

// min_maxcuts: lower bound of the output stride (height + max cuts) -- a wave sharded over several GPUs must come
// back with ONE stride although every shard only sees its own nodes.
int bnb_solve_nodes_impl(yalps_ctx *ctx, int64_t n, const int32_t *cut_offsets, const double *cut_sign,
                         const int32_t *cut_var, const double *cut_value, const yalps_options *opt, int32_t *status,
                         double *value, int64_t *pivots, double *rhs_out, int32_t *pos_out, int32_t *var_out,
                         double *matrices_out, int min_maxcuts) {
  if (!ctx) return YALPS_ERR_ARGUMENT;
  Root &R = ctx->root;
  if (!R.valid) return fail(ctx, YALPS_ERR_ARGUMENT, "no root tableau: call yalps_bnb_set_root first");
  if (n < 0 || !opt || (n > 0 && !cut_offsets)) return fail(ctx, YALPS_ERR_ARGUMENT, "bad arguments");
  if (n == 0) return 0;
  const auto tw0 = std::chrono::steady_clock::now();
  CU(ctx, cudaSetDevice(ctx->device));
  const int W = R.W;
  int maxcuts = std::max(0, min_maxcuts);
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
  // Big, very sparse nodes (Monster-class models: ~1 % non-zeros, ~10 rows rewritten per pivot): the row-split HBM/L2
  // kernel with one CTA per node compacts the few active rows and needs no grid barrier -- 5.8 us per pivot for every
  // node of the wave at once, against 9.5 us per pivot and node on K4 (scripts/big_sparse_paths.py).  Denser nodes
  // (Vendor Selection: 28 % non-zeros at the root optimum) stay on K4.
  bool sparse_split = false;
  if (!plan.resident && ctx->tune_path == YALPS_PATH_AUTO && ctx->tune_threads <= 0 && ctx->tune_rows <= 0 &&
      R.density < 0.03 && use_grid_path(ctx, n, plan)) {
    int split_nwc = 8, split_nwr = 2;
    if (const char *env = getenv("YALPS_NODE_SPLIT")) sscanf(env, "%d,%d", &split_nwc, &split_nwr);  // experiments: "column warps,row groups"
    if (const KernelEntry *k = pick_kernel(split_nwc, split_nwr, W, false)) {
      if (k->nwr == split_nwr) {
        const size_t smem_k = SmemLayout(Hcap, W, false, k->nw * k->nwr, true).total;
        if (smem_k <= (size_t)ctx->smem_optin) {
          CU(ctx, raise_smem_limit(ctx->device, (const void *)k->global, (int)smem_k));
          int occ = 0;
          CU(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k->global, k->nw * k->nwr * 32, smem_k));
          if (occ >= 1) {
            plan.k = k;
            plan.smem = smem_k;
            plan.grid = (int)std::max<long long>(1, std::min<long long>((long long)occ * ctx->prop.multiProcessorCount, n));
            sparse_split = true;
          }
        }
      }
    }
  }

  cudaStream_t st = ctx->streams[0];
  int rc;
  // One pinned staging buffer and ONE copy per direction: a wave is latency-bound, and every extra
  // cudaMemcpyAsync from pageable memory costs more than the node kernel itself.
  auto up8 = [](size_t x) { return (x + 7) & ~(size_t)7; };
  const size_t nc = (size_t)std::max(ncut_total, 0);
  const size_t in_off = 0, in_var = up8((size_t)(n + 1) * 4), in_sign = in_var + up8(nc * 4), in_val = in_sign + nc * 8,
               in_bytes = in_val + nc * 8;
  const size_t o_status = 0, o_piv = up8((size_t)n * 4), o_value = o_piv + (size_t)n * 16, o_rhs = o_value + (size_t)n * 8,
               o_pos = o_rhs + (size_t)n * Hcap * 8, o_var = o_pos + up8((size_t)n * (W + Hcap) * 4),
               out_bytes = o_var + up8((size_t)n * (W + Hcap) * 4);
  void *h_in, *h_out, *d_inb, *d_outb;
  if ((rc = pin_ensure(ctx, "nd_h_in", in_bytes + 8, &h_in))) return rc;
  if ((rc = pin_ensure(ctx, "nd_h_out", out_bytes + 8, &h_out))) return rc;
  // small waves: the kernel reads the cut lists and writes its results straight through the mapped pinned
  // buffers (zero-copy), so a wave is launch + synchronise; big waves use one explicit copy per direction
  // (only for shared-memory resident nodes: the HBM/L2-resident kernels and K4 keep variableAtPosition in the output
  // buffer while they pivot, and that must not be host memory)
  const bool zero_copy = plan.resident && in_bytes + out_bytes <= ((size_t)1 << 20);
  if (zero_copy) {
    if ((rc = pin_device_ptr(ctx, h_in, &d_inb))) return rc;
    if ((rc = pin_device_ptr(ctx, h_out, &d_outb))) return rc;
  } else {
    if ((rc = dev_ensure(ctx, "nd_d_in", in_bytes + 8, &d_inb))) return rc;
    if ((rc = dev_ensure(ctx, "nd_d_out", out_bytes + 8, &d_outb))) return rc;
  }
  std::memcpy((char *)h_in + in_off, cut_offsets, (size_t)(n + 1) * 4);
  if (nc) {
    std::memcpy((char *)h_in + in_var, cut_var, nc * 4);
    std::memcpy((char *)h_in + in_sign, cut_sign, nc * 8);
    std::memcpy((char *)h_in + in_val, cut_value, nc * 8);
  }
  const auto tw1 = std::chrono::steady_clock::now();
  if (!zero_copy) CU(ctx, cudaMemcpyAsync(d_inb, h_in, in_bytes, cudaMemcpyHostToDevice, st));
  void *d_off = (char *)d_inb + in_off, *d_var = (char *)d_inb + in_var, *d_sign = (char *)d_inb + in_sign,
       *d_val = (char *)d_inb + in_val;
  void *d_status = (char *)d_outb + o_status, *d_piv = (char *)d_outb + o_piv, *d_value = (char *)d_outb + o_value,
       *d_rhs = (char *)d_outb + o_rhs, *d_pos = (char *)d_outb + o_pos, *d_vr = (char *)d_outb + o_var;
  void *d_work = nullptr, *d_out = nullptr;
  const size_t mat_bytes = (size_t)n * Hcap * W * 8;
  if (!plan.resident) {
    if ((rc = dev_ensure(ctx, "nd_work", mat_bytes, &d_work))) return rc;
    d_out = matrices_out ? d_work : nullptr;
  } else if (matrices_out) {
    if ((rc = dev_ensure(ctx, "nd_out", mat_bytes, &d_out))) return rc;
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
  a.rows_out = ctx->rows_per_lp ? nullptr : ctx->d_rows;  // node waves only feed a total (their LP indices are wave-local)
  if (sparse_split) {
    // assemble the nodes grid-wide (K3), then one row-split CTA per node on the assembled working copies
    const int gx = std::max(1, std::min(ctx->prop.multiProcessorCount * 4, (int)(((size_t)R.H * W + 255) / 256)));
    const int gy = (int)std::min<int64_t>(n, 65535);
    k_assemble_nodes<<<dim3(gx, gy), 256, 0, st>>>(n, R.H, W, Hcap, (const double *)R.m.p, (const int *)R.pos.p,
                                                   (const int *)d_off, (const double *)d_sign, (const int *)d_var,
                                                   (const double *)d_val, (double *)d_work);
    CU(ctx, cudaGetLastError());
    ctx->launches++;
    a.assembled = 1;
    if ((rc = launch_simplex(ctx, plan, a, "nd", st))) return rc;
  } else if (use_grid_path(ctx, n, plan)) {
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
    // the nodes of a wave are independent: give each of up to kGridLanes concurrent cooperative launches its share of
    // the SMs (a K4 pivot on an L2-resident tableau is bound by its two grid barriers, which get cheaper with fewer
    // CTAs, while the wave as a whole uses the whole GPU)
    constexpr int kGridLanes = 6;
    const int lanes = (int)std::min<int64_t>(n, kGridLanes);
    while ((int)ctx->aux_streams.size() < lanes) {
      cudaStream_t s2;
      cudaEvent_t e2;
      CU(ctx, cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
      CU(ctx, cudaEventCreateWithFlags(&e2, cudaEventDisableTiming));
      ctx->aux_streams.push_back(s2);
      ctx->aux_events.push_back(e2);
    }
    if (!ctx->fork_event) CU(ctx, cudaEventCreateWithFlags(&ctx->fork_event, cudaEventDisableTiming));
    const int share = lanes > 1 ? std::max(8, ctx->prop.multiProcessorCount / lanes) : 0;
    if (lanes > 1) {
      CU(ctx, cudaEventRecord(ctx->fork_event, st));
      for (int l = 0; l < lanes; l++) CU(ctx, cudaStreamWaitEvent(ctx->aux_streams[l], ctx->fork_event, 0));
    }
    for (int64_t j = 0; j < n; j++) {
      const int hj = R.H + (cut_offsets[j + 1] - cut_offsets[j]);
      const int l = (int)(j % lanes);
      cudaStream_t sj = lanes > 1 ? ctx->aux_streams[l] : st;
      if ((rc = launch_grid(ctx, hj, W, (double *)d_work + (size_t)j * Hcap * W, opt, (int *)d_status + j,
                            (double *)d_value + j, (long long *)d_piv + 2 * j, (double *)d_rhs + (size_t)j * Hcap,
                            (int *)d_pos + (size_t)j * (W + Hcap), (int *)d_vr + (size_t)j * (W + Hcap), sj,
                            (const int *)R.var.p, W + R.H, share, lanes > 1 ? "_l" + std::to_string(l) : "", a.rows_out)))
        return rc;
    }
    if (lanes > 1) {
      for (int l = 0; l < lanes; l++) {
        CU(ctx, cudaEventRecord(ctx->aux_events[l], ctx->aux_streams[l]));
        CU(ctx, cudaStreamWaitEvent(st, ctx->aux_events[l], 0));
      }
    }
  } else if ((rc = launch_simplex(ctx, plan, a, "nd", st))) {
    return rc;
  }
  if (!zero_copy) CU(ctx, cudaMemcpyAsync(h_out, d_outb, out_bytes, cudaMemcpyDeviceToHost, st));
  if (matrices_out) CU(ctx, cudaMemcpyAsync(matrices_out, d_out, mat_bytes, cudaMemcpyDeviceToHost, st));
  const auto tw2 = std::chrono::steady_clock::now();
  CU(ctx, cudaStreamSynchronize(st));
  const auto tw3 = std::chrono::steady_clock::now();
  const char *ho = (const char *)h_out;
  if (status) std::memcpy(status, ho + o_status, (size_t)n * 4);
  if (value) std::memcpy(value, ho + o_value, (size_t)n * 8);
  if (pivots) std::memcpy(pivots, ho + o_piv, (size_t)n * 16);
  if (rhs_out) std::memcpy(rhs_out, ho + o_rhs, (size_t)n * Hcap * 8);
  if (pos_out) std::memcpy(pos_out, ho + o_pos, (size_t)n * (W + Hcap) * 4);
  if (var_out) std::memcpy(var_out, ho + o_var, (size_t)n * (W + Hcap) * 4);
  {
    const auto tw4 = std::chrono::steady_clock::now();
    auto ns = [](auto a, auto b) { return (int64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(b - a).count(); };
    ctx->wave_ns[0] += ns(tw0, tw1);
    ctx->wave_ns[1] += ns(tw1, tw2);
    ctx->wave_ns[2] += ns(tw2, tw3);
    ctx->wave_ns[3] += ns(tw3, tw4);
  }
  return check_device_status(ctx, status, n);
}

// Evaluates the n node LPs of one wave (cut lists in CSR form, offsets starting at 0) into arrays strided by
// height + maxcuts.  The default evaluator runs them on the ctx's GPU; the multi-GPU driver deals them over its ranks.
using WaveEval = std::function<int(int64_t n, const int32_t *off, const double *sign, const int32_t *var,
                                   const double *val, int maxcuts, int32_t *status, double *value, int64_t *piv,
                                   double *rhs, int32_t *pos, int32_t *var_out)>;
// Called after every wave with the incumbent (inf while none): the multi-GPU driver runs its min-allreduce here.
using WaveHook = std::function<int(int64_t wave_index, double best_eval)>;

int branch_and_cut_impl(yalps_ctx *ctx, const WaveEval *wave_eval, const WaveHook *wave_hook, const int32_t *ints,
                        int32_t nints, double sign, double init_result, const yalps_options *opt, int32_t *status,
                        double *result, int32_t *out_height, double *rhs_out, int32_t *pos_out, int32_t *var_out,
                        int64_t *stats);

}  // namespace

extern "C" {

int yalps_branch_and_cut(yalps_ctx *ctx, const int32_t *ints, int32_t nints, double sign, double init_result,
                         const yalps_options *opt, int32_t *status, double *result, int32_t *out_height,
                         double *rhs_out, int32_t *pos_out, int32_t *var_out, int64_t *stats) {
  return branch_and_cut_impl(ctx, nullptr, nullptr, ints, nints, sign, init_result, opt, status, result, out_height,
                             rhs_out, pos_out, var_out, stats);
}

}  // extern "C"

namespace {

// The whole search in one persistent launch (bnb_kernel.cuh).  Returns 0 when the search ran to its end on the device
// (outputs filled), 1 when this search is not for the device kernel (node tableaus beyond one CTA's shared memory,
// checkCycles) or one of its pools overflowed -- the caller then runs the wave driver --, < 0 on errors.
int device_search(yalps_ctx *ctx, const int32_t *ints, int32_t nints, double sign, double init_result, int32_t init_var,
                  double init_val, const yalps_options *opt, int32_t *status, double *result, int32_t *out_height,
                  double *rhs_out, int32_t *pos_out, int32_t *var_out, int64_t *stats) {
  Root &R = ctx->root;
  const int W = R.W, H = R.H;
  if (opt->check_cycles) return 1;
  const BnbConfig *cfg = bnb_config_for(W);
  if (!cfg) return 1;
  const int nw = cfg->nwc * cfg->nwr;
  // cut rows a node may carry on the device: as many as shared memory allows, at most 2*|integers| (:108) and 96
  const size_t smem_room = (size_t)ctx->smem_optin - 4096;  // the kernel's static shared memory (staged cut list)
  int extra = std::min(2 * nints, kBnbMaxCuts);
  while (extra >= 8 && SmemLayout(H + extra, W, true, nw, true).total > smem_room) extra -= 8;
  if (extra < std::min(2 * nints, 8)) return 1;
  const int Hcap = H + extra;
  const size_t worker_smem = SmemLayout(Hcap, W, true, nw, true).total;
  const size_t smem = std::max(worker_smem, (size_t)96 << 10);
  const int heap_cap = (int)(smem / 12);
  CU(ctx, raise_smem_limit(ctx->device, cfg->fn, (int)smem));
  const std::string okey = "bnb:" + std::to_string((size_t)cfg->fn) + ":" + std::to_string(smem);
  int occ = 0;
  auto it = ctx->occ_cache.find(okey);
  if (it == ctx->occ_cache.end()) {
    CU(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, cfg->fn, nw * 32, smem));
    ctx->occ_cache[okey] = occ;
  } else {
    occ = it->second;
  }
  if (occ < 1) return 1;
  // one scheduler + one worker per remaining SM (speculative expansion keeps a few thousand node LPs queued); the
  // concurrent searches of yalps_multi_solve_many share the SMs (ctx->bnb_workers) so that they are co-resident
  int max_workers = ctx->bnb_workers > 0 ? ctx->bnb_workers : ctx->prop.multiProcessorCount - 1;
  if (const char *env = getenv("YALPS_BNB_WORKERS")) max_workers = std::max(1, atoi(env));
  const int grid = std::min(occ * ctx->prop.multiProcessorCount, 1 + max_workers);
  if (grid < 2) return 1;

  // the replay creates at most 2 * maxIterations + 2 nodes; speculation (bnb_kernel.cuh) gets as many again
  double max_nodes = 2.0 * opt->max_iterations + 4.0;
  const int base_nodes = (int)std::min(max_nodes, 65536.0);
  int spec_depth = 2;
  if (const char *env = getenv("YALPS_BNB_SPEC")) spec_depth = std::max(0, atoi(env));
  const int spec_cap = spec_depth > 0 ? base_nodes : 0;
  const int node_cap = base_nodes + spec_cap;
  const unsigned long long cut_cap = 1ULL << 21;
  const int cand_cap = 4096;
  void *d_ctl, *d_nodes, *d_queue, *d_cuts, *d_crhs, *d_cpos, *d_cvar, *d_ints, *d_rank;
  int rc;
  if ((rc = dev_ensure(ctx, "kb_ctl", sizeof(BnbControl), &d_ctl))) return rc;
  if ((rc = dev_ensure(ctx, "kb_nodes", (size_t)node_cap * sizeof(BnbNode), &d_nodes))) return rc;
  if ((rc = dev_ensure(ctx, "kb_queue", (size_t)node_cap * sizeof(int), &d_queue))) return rc;
  if ((rc = dev_ensure(ctx, "kb_cuts", (size_t)cut_cap * sizeof(BnbCut), &d_cuts))) return rc;
  if ((rc = dev_ensure(ctx, "kb_crhs", (size_t)cand_cap * Hcap * 8, &d_crhs))) return rc;
  if ((rc = dev_ensure(ctx, "kb_cpos", (size_t)cand_cap * (W + Hcap) * 4, &d_cpos))) return rc;
  if ((rc = dev_ensure(ctx, "kb_cvar", (size_t)cand_cap * (W + Hcap) * 4, &d_cvar))) return rc;
  if ((rc = dev_ensure(ctx, "kb_ints", (size_t)std::max(nints, 1) * 4, &d_ints))) return rc;
  if ((rc = dev_ensure(ctx, "kb_rank", (size_t)(W + H) * 4, &d_rank))) return rc;
  // one pinned staging block: [ints | int_rank] up, control block + best candidate down
  const size_t up_bytes = (size_t)(nints + W + H) * 4;
  const size_t down_bytes = sizeof(BnbControl) + (size_t)Hcap * 8 + (size_t)2 * (W + Hcap) * 4;
  void *h_up, *h_down;
  if ((rc = pin_ensure(ctx, "kb_h_up", up_bytes + 16, &h_up))) return rc;
  if ((rc = pin_ensure(ctx, "kb_h_down", down_bytes + 16, &h_down))) return rc;
  int32_t *h_ints = (int32_t *)h_up, *h_rank = h_ints + nints;
  std::memcpy(h_ints, ints, (size_t)nints * 4);
  for (int k = 0; k < W + H; k++) h_rank[k] = -1;
  for (int i = nints - 1; i >= 0; i--)
    if (ints[i] >= 0 && ints[i] < W + H) h_rank[ints[i]] = i;  // first occurrence wins, like the reference's scan
  cudaStream_t st = ctx->streams[0];
  CU(ctx, cudaMemcpyAsync(d_ints, h_ints, (size_t)nints * 4, cudaMemcpyHostToDevice, st));
  CU(ctx, cudaMemcpyAsync(d_rank, h_rank, (size_t)(W + H) * 4, cudaMemcpyHostToDevice, st));
  CU(ctx, cudaMemsetAsync(d_ctl, 0, sizeof(BnbControl), st));
  CU(ctx, cudaMemsetAsync(d_queue, 0, (size_t)node_cap * sizeof(int), st));
  CU(ctx, cudaMemsetAsync(d_nodes, 0xff, (size_t)node_cap * sizeof(BnbNode), st));  // the cleared state of the result blocks

  BnbArgs a{};
  a.root = (const double *)R.m.p;
  a.root_pos = (const int *)R.pos.p;
  a.root_var = (const int *)R.var.p;
  a.H = H;
  a.W = W;
  a.Hcap = Hcap;
  a.ints = (const int *)d_ints;
  a.int_rank = (const int *)d_rank;
  a.nints = nints;
  a.sign = sign;
  a.init_result = init_result;
  a.init_var = init_var;
  a.init_value = init_val;
  a.precision = opt->precision;
  a.max_pivots = opt->max_pivots;
  a.tolerance = opt->tolerance;
  a.timeout_ms = opt->timeout_ms;
  a.max_iterations = opt->max_iterations;
  a.ctl = (BnbControl *)d_ctl;
  a.nodes = (BnbNode *)d_nodes;
  a.node_cap = node_cap;
  a.queue = (int *)d_queue;
  a.sched_cap = base_nodes;
  a.spec_cap = spec_cap;
  a.spec_depth = spec_depth;
  a.cuts = (BnbCut *)d_cuts;
  a.cut_cap = cut_cap;
  a.cand_rhs = (double *)d_crhs;
  a.cand_pos = (int *)d_cpos;
  a.cand_var = (int *)d_cvar;
  a.cand_cap = cand_cap;
  a.heap_cap = heap_cap;
  a.rows_out = ctx->rows_per_lp ? nullptr : ctx->d_rows;
  const auto t0 = std::chrono::steady_clock::now();
  CU(ctx, launch_bnb(cfg, a, grid, smem, st));
  ctx->launches++;
  BnbControl *hc = (BnbControl *)h_down;
  CU(ctx, cudaMemcpyAsync(hc, d_ctl, sizeof(BnbControl), cudaMemcpyDeviceToHost, st));
  CU(ctx, cudaStreamSynchronize(st));
  if (getenv("YALPS_BNB_DEBUG"))
    fprintf(stderr, "k_bnb scheduler: %lld iterations, %lld cycles, %lld waiting for results (%lld pops waited), %lld in pop(); overflow flags %d, %d nodes created (%d speculative), %d candidates, %llu cuts\n",
            hc->iters, hc->t_total, hc->t_wait, hc->n_wait, hc->t_heap, hc->overflow, hc->created, hc->spec_count, hc->cand_top, hc->cut_top);
  if (getenv("YALPS_BNB_DEBUG") && hc->w_nodes)
    fprintf(stderr, "k_bnb workers: %llu nodes (finished before the stop), cycles per node: cut list %llu, assembly %llu, simplex %llu (%.1f pivots), mostFractionalVar + publish %llu\n",
            hc->w_nodes, hc->w_cuts / hc->w_nodes, hc->w_asm / hc->w_nodes, hc->w_simplex / hc->w_nodes,
            (double)hc->node_pivots / (double)std::max<long long>(hc->iters, 1), hc->w_post / hc->w_nodes);
  if (hc->overflow) return 1;  // a pool ran out: the wave driver has no such limits
  if (hc->found) {
    if (hc->best_cand < 0) return fail(ctx, YALPS_ERR_CUDA, "device search: incumbent without candidate arrays");
    const int h = hc->best_height;
    double *h_rhs = (double *)((char *)h_down + sizeof(BnbControl));
    int32_t *h_pos = (int32_t *)(h_rhs + Hcap), *h_var = h_pos + (W + Hcap);
    CU(ctx, cudaMemcpyAsync(h_rhs, (double *)d_crhs + (size_t)hc->best_cand * Hcap, (size_t)h * 8, cudaMemcpyDeviceToHost, st));
    CU(ctx, cudaMemcpyAsync(h_pos, (int *)d_cpos + (size_t)hc->best_cand * (W + Hcap), (size_t)(W + h) * 4, cudaMemcpyDeviceToHost, st));
    CU(ctx, cudaMemcpyAsync(h_var, (int *)d_cvar + (size_t)hc->best_cand * (W + Hcap), (size_t)(W + h) * 4, cudaMemcpyDeviceToHost, st));
    CU(ctx, cudaStreamSynchronize(st));
    *out_height = h;
    std::memcpy(rhs_out, h_rhs, (size_t)h * 8);
    std::memcpy(pos_out, h_pos, (size_t)(W + h) * 4);
    std::memcpy(var_out, h_var, (size_t)(W + h) * 4);
  } else {  // bestTableau = root (:119)
    *out_height = H;
    std::memcpy(rhs_out, R.h_rhs.data(), (size_t)H * 8);
    std::memcpy(pos_out, R.h_pos.data(), (size_t)(W + H) * 4);
    std::memcpy(var_out, R.h_var.data(), (size_t)(W + H) * 4);
  }
  *status = hc->status;
  *result = hc->result;
  if (stats) {
    const int64_t us = std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count();
    stats[0] = hc->iters;
    stats[1] = hc->node_pivots;
    stats[2] = hc->max_cuts;
    stats[3] = hc->max_heap;
    stats[4] = 1;  // one launch
    stats[5] = hc->created;
    stats[6] = us;
    stats[7] = us;
  }
  return 0;
}

int branch_and_cut_impl(yalps_ctx *ctx, const WaveEval *wave_eval, const WaveHook *wave_hook, const int32_t *ints,
                        int32_t nints, double sign, double init_result, const yalps_options *opt, int32_t *status,
                        double *result, int32_t *out_height, double *rhs_out, int32_t *pos_out, int32_t *var_out,
                        int64_t *stats) {
  if (!ctx) return YALPS_ERR_ARGUMENT;
  Root &R = ctx->root;
  if (!R.valid) return fail(ctx, YALPS_ERR_ARGUMENT, "no root tableau: call yalps_bnb_set_root first");
  if (!opt || !status || !result || !out_height || !rhs_out || !pos_out || !var_out || nints < 0 || (nints && !ints))
    return fail(ctx, YALPS_ERR_ARGUMENT, "bad arguments");
  const int W = R.W, H = R.H;
  const double precision = opt->precision;
  int64_t st_nodes = 0, st_pivots = 0, st_maxcuts = 0, st_maxheap = 0, st_waves = 0, st_devnodes = 0, st_wave_us = 0;
  const auto t_begin = std::chrono::steady_clock::now();

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
    stats[6] = st_wave_us;
    stats[7] = std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t_begin).count();
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
  // searches whose node LPs fit one CTA run entirely on the device (bnb_kernel.cuh); everything else, and any search
  // that overflows the device pools, takes the wave driver below
  if (!wave_eval && ctx->bnb_mode != 1) {
    const int rc = device_search(ctx, ints, nints, sign, init_result, init_var, init_val, opt, status, result, out_height,
                                 rhs_out, pos_out, var_out, stats);
    if (rc <= 0) return rc;
    if (ctx->bnb_mode == 2) return fail(ctx, YALPS_ERR_TOO_LARGE, "this search does not fit the device-resident branch-and-cut kernel");
  }

  std::vector<Branch> arena;
  arena.reserve(4096);
  BranchHeap heap;
  heap.arena = &arena;
  arena.push_back(Branch{init_result, {Cut{-1.0, init_var, std::ceil(init_val)}}});
  arena.push_back(Branch{init_result, {Cut{1.0, init_var, std::floor(init_val)}}});
  heap.push(0);
  heap.push(1);

  // branch id -> (wave, slot) of its cached node result, -1 = not evaluated
  std::vector<int32_t> res_wave, res_slot;
  std::vector<std::unique_ptr<WaveBuf>> waves;
  size_t cached_bytes = 0, oldest_wave = 0;
  const size_t kCacheBudget = (size_t)768 << 20;
  auto cached = [&](int32_t id) { return (size_t)id < res_wave.size() && res_wave[id] >= 0; };

  const double threshold = init_result * (1.0 - sign * opt->tolerance);
  const double stop_time = opt->timeout_ms + now_ms();
  bool timedout = now_ms() >= stop_time;
  bool found = false;
  double best_eval = std::numeric_limits<double>::infinity();
  int best_height = 0;
  std::vector<double> best_rhs;
  std::vector<int32_t> best_pos, best_var;
  double iter = 0;

  std::vector<int32_t> w_off;
  std::vector<double> w_sign, w_val;
  std::vector<int32_t> w_var;
  BranchHeap peek;
  peek.arena = &arena;
  int cur_wave = std::min(16, std::max(ctx->wave, 1)), last_wave_n = 0, consumed_since_wave = 0;
  double avg_consumed = 8.0;

  auto run_wave = [&](int32_t needed) -> int {
    // the needed branch first, then the branches the heap would pop next (on a copy of the index heap)
    // adaptive look-ahead: the number of waves is set by the depth of the search's dives, not by the wave size
    // (Large Farm MIP: 135 waves for every size from 16 to 256), so speculation beyond what the replay consumes
    // only costs host and device work.  Width = 1.5 x the running mean of the nodes consumed per wave.
    if (last_wave_n > 0) {
      avg_consumed = 0.75 * avg_consumed + 0.25 * (double)consumed_since_wave;
      cur_wave = std::max(8, std::min(std::max(ctx->wave, 8), (int)(1.5 * avg_consumed) + 2));
    }
    consumed_since_wave = 0;
    auto wb = std::make_unique<WaveBuf>();
    wb->ids.push_back(needed);
    peek.a = heap.a;
    while ((int)wb->ids.size() < cur_wave && !peek.empty()) {
      const int32_t b = peek.pop();
      if (arena[b].eval > best_eval) break;  // would be pruned (:124)
      if (!cached(b)) wb->ids.push_back(b);
    }
    const int64_t n = (int64_t)wb->ids.size();
    last_wave_n = (int)n;
    w_off.assign(1, 0);
    w_sign.clear();
    w_var.clear();
    w_val.clear();
    int maxcuts = 0;
    for (int32_t id : wb->ids) {
      for (const Cut &c : arena[id].cuts) {
        w_sign.push_back(c.sign);
        w_var.push_back(c.variable);
        w_val.push_back(c.value);
      }
      w_off.push_back((int32_t)w_sign.size());
      maxcuts = std::max(maxcuts, (int)arena[id].cuts.size());
    }
    wb->n = (int)n;
    wb->Hcap = H + maxcuts;
    wb->status.resize(n);
    wb->value.resize(n);
    wb->piv.resize(2 * n);
    wb->rhs.resize((size_t)n * wb->Hcap);
    wb->pos.resize((size_t)n * (W + wb->Hcap));
    wb->var.resize((size_t)n * (W + wb->Hcap));
    const auto t0 = std::chrono::steady_clock::now();
    if (wave_eval) {
      if (int rc = (*wave_eval)(n, w_off.data(), w_sign.data(), w_var.data(), w_val.data(), maxcuts, wb->status.data(),
                                wb->value.data(), wb->piv.data(), wb->rhs.data(), wb->pos.data(), wb->var.data()))
        return rc;
    } else if (int rc = bnb_solve_nodes_impl(ctx, n, w_off.data(), w_sign.data(), w_var.data(), w_val.data(), opt,
                                             wb->status.data(), wb->value.data(), wb->piv.data(), wb->rhs.data(),
                                             wb->pos.data(), wb->var.data(), nullptr, 0)) {
      return rc;
    }
    if (wave_hook)
      if (int rc = (*wave_hook)(st_waves, best_eval)) return rc;
    st_wave_us += std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count();
    st_waves++;
    st_devnodes += n;
    if (res_wave.size() < arena.size()) {
      res_wave.resize(arena.capacity() > arena.size() ? arena.capacity() : arena.size() * 2, -1);
      res_slot.resize(res_wave.size(), -1);
    }
    for (int64_t j = 0; j < n; j++) {
      res_wave[wb->ids[j]] = (int32_t)waves.size();
      res_slot[wb->ids[j]] = (int32_t)j;
    }
    cached_bytes += wb->bytes();
    waves.push_back(std::move(wb));
    // bounded cache: forget the oldest waves (their unused speculation is recomputed if it is ever needed)
    while (cached_bytes > kCacheBudget && oldest_wave + 1 < waves.size()) {
      WaveBuf &old = *waves[oldest_wave];
      for (int32_t id : old.ids)
        if (res_wave[id] == (int32_t)oldest_wave) res_wave[id] = -1;
      cached_bytes -= old.bytes();
      waves[oldest_wave].reset();
      oldest_wave++;
    }
    return 0;
  };

  while (iter < opt->max_iterations && !heap.empty() && best_eval >= threshold && !timedout) {  // :122
    st_maxheap = std::max<int64_t>(st_maxheap, (int64_t)heap.a.size());
    const int32_t br = heap.pop();
    if (arena[br].eval > best_eval) break;  // :124

    if (!cached(br))
      if (int rc = run_wave(br)) return rc;
    const WaveBuf &wb = *waves[res_wave[br]];
    const int slot = res_slot[br];
    res_wave[br] = -1;  // consumed
    consumed_since_wave++;
    const int32_t n_status = wb.status[slot];
    const double n_value = wb.value[slot];
    const int n_height = H + (int)arena[br].cuts.size();
    const double *n_rhs = wb.rhs.data() + (size_t)slot * wb.Hcap;
    const int32_t *n_pos = wb.pos.data() + (size_t)slot * (W + wb.Hcap);
    const int32_t *n_var = wb.var.data() + (size_t)slot * (W + wb.Hcap);
    st_nodes++;
    st_pivots += wb.piv[2 * slot] + wb.piv[2 * slot + 1];
    st_maxcuts = std::max<int64_t>(st_maxcuts, (int64_t)arena[br].cuts.size());

    if (n_status == YALPS_OPTIMAL && n_value < best_eval) {  // :130
      int32_t variable;
      double value, frac;
      most_fractional(n_rhs, n_pos, W, ints, nints, &variable, &value, &frac);
      if (frac <= precision) {  // integer solution, new incumbent (:132-139)
        found = true;
        best_eval = n_value;
        best_height = n_height;
        best_rhs.assign(n_rhs, n_rhs + n_height);
        best_pos.assign(n_pos, n_pos + W + n_height);
        best_var.assign(n_var, n_var + W + n_height);
      } else {  // branch (:141-156)
        Branch upper, lower;
        upper.cuts.reserve(arena[br].cuts.size() + 1);
        lower.cuts.reserve(arena[br].cuts.size() + 1);
        for (const Cut &cut : arena[br].cuts) {
          if (cut.variable == variable) {
            if (cut.sign < 0)
              lower.cuts.push_back(cut);
            else
              upper.cuts.push_back(cut);
          } else {
            upper.cuts.push_back(cut);
            lower.cuts.push_back(cut);
          }
        }
        lower.cuts.push_back(Cut{1.0, variable, std::floor(value)});
        upper.cuts.push_back(Cut{-1.0, variable, std::ceil(value)});
        upper.eval = lower.eval = n_value;
        arena.push_back(std::move(upper));
        heap.push((int32_t)arena.size() - 1);
        arena.push_back(std::move(lower));
        heap.push((int32_t)arena.size() - 1);
        if (res_wave.size() < arena.size()) {
          res_wave.resize(arena.size() * 2, -1);
          res_slot.resize(res_wave.size(), -1);
        }
      }
    }
    std::vector<Cut>().swap(arena[br].cuts);  // the cut list of a consumed branch is no longer needed
    timedout = now_ms() >= stop_time;  // :162
    iter++;
  }

  const bool unfinished = (timedout || iter >= opt->max_iterations) && !heap.empty() && best_eval >= threshold;  // :167
  *status = unfinished ? YALPS_TIMEDOUT : (!found ? YALPS_INFEASIBLE : YALPS_OPTIMAL);
  *result = found ? best_eval : std::numeric_limits<double>::quiet_NaN();
  if (found)
    write_best(best_rhs.data(), best_pos.data(), best_var.data(), best_height);
  else
    write_best(R.h_rhs.data(), R.h_pos.data(), R.h_var.data(), H);  // bestTableau = root (:119)
  write_stats();
  if (getenv("YALPS_BNB_DEBUG") && st_waves) {
    const auto total = std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t_begin).count();
    fprintf(stderr, "wave driver: %lld waves, %lld device nodes, %lld us in all; per wave (us): host before the first CUDA call %.1f, "
                    "enqueue %.1f, wait %.1f, copy out %.1f, replay and wave set-up %.1f\n",
            (long long)st_waves, (long long)st_devnodes, (long long)total, ctx->wave_ns[0] / 1e3 / st_waves,
            ctx->wave_ns[1] / 1e3 / st_waves, ctx->wave_ns[2] / 1e3 / st_waves, ctx->wave_ns[3] / 1e3 / st_waves,
            ((double)total - (ctx->wave_ns[0] + ctx->wave_ns[1] + ctx->wave_ns[2] + ctx->wave_ns[3]) / 1e3) / st_waves);
  }
  for (auto &x : ctx->wave_ns) x = 0;
  return 0;
}

}  // namespace

extern "C" {

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
  // A MILP root that is too big for the zero-copy small-call path stays on the device: branch and cut needs the whole
  // final root tableau (applyCuts, src/branchAndCut.ts:28,38-42), and bringing it to the host only to upload it again
  // costs two PCIe trips of the matrix (SURVEY 8b "bnb_adopt_root").
  const bool keep_on_device = milp && cells * 8 > ((size_t)256 << 10);
  std::vector<double> final_m;
  if (milp && !keep_on_device) final_m.resize(cells);
  // root LP (src/YALPS.ts:79)
  ctx->keep_final = keep_on_device;
  ctx->kept_final = nullptr;
  const int root_rc = yalps_solve_batch(ctx, 1, height, width, matrix, opt, &st, &val, piv, rhs_out, pos_out, var_out,
                                        (milp && !keep_on_device) ? final_m.data() : nullptr);
  ctx->keep_final = false;
  if (root_rc) return root_rc;
  if (ctx->coo_nnz >= 0 && ctx->coo_dup) return 1;  // yalps_solve_sparse resolves the duplicates and calls again
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
  if (keep_on_device) {
    if (!ctx->kept_final) return fail(ctx, YALPS_ERR_CUDA, "root tableau was not kept on the device");
    if (((long long)height + 2 * nints) * width >= (1LL << 31))
      return fail(ctx, YALPS_ERR_TOO_LARGE, "(height+extra)*width must be < 2^31");
    if (int rc = adopt_root_device(ctx, height, width, ctx->kept_final, rhs_out, pos_out, var_out, 2 * nints)) return rc;
  } else if (int rc = yalps_bnb_set_root(ctx, height, width, final_m.data(), pos_out, var_out, 2 * nints)) {
    return rc;
  }
  return yalps_branch_and_cut(ctx, ints, nints, sign, val, opt, status, result, out_height, rhs_out, pos_out, var_out,
                              stats);
}

int yalps_solve_sparse(yalps_ctx *ctx, int32_t height, int32_t width, int64_t nnz, const int32_t *cell, const double *val,
                       const int32_t *ints, int32_t nints, double sign, const yalps_options *opt, int32_t *status,
                       double *result, int32_t *out_height, double *rhs_out, int32_t *pos_out, int32_t *var_out,
                       int32_t *root_status, double *root_value, int64_t *root_pivots, int64_t *stats) {
  if (!ctx) return YALPS_ERR_ARGUMENT;
  if (height < 1 || width < 1 || nnz < 0 || (nnz && (!cell || !val)))
    return fail(ctx, YALPS_ERR_ARGUMENT, "bad arguments");
  if ((long long)height * width >= (1LL << 31)) return fail(ctx, YALPS_ERR_TOO_LARGE, "height*width must be < 2^31");
  const int32_t cells = (int32_t)((long long)height * width);
  for (int64_t i = 0; i < nnz; i++)
    if (cell[i] < 0 || cell[i] >= cells)
      return fail(ctx, YALPS_ERR_ARGUMENT, "cell %lld = %d is outside the %dx%d tableau", (long long)i, cell[i], height, width);
  // small tableaus take the zero-copy latency path of yalps_solve_batch, which reads a dense host image
  if ((size_t)cells * 8 <= ((size_t)768 << 10)) {
    std::vector<double> dense((size_t)cells, 0.0);
    for (int64_t i = 0; i < nnz; i++) dense[cell[i]] = val[i];
    return yalps_solve(ctx, height, width, dense.data(), ints, nints, sign, opt, status, result, out_height, rhs_out,
                       pos_out, var_out, root_status, root_value, root_pivots, stats);
  }
  // the matrix argument below is a placeholder that is never dereferenced while coo_nnz >= 0 (solve_host scatters
  // the pairs on the device instead of copying a host matrix)
  ctx->coo_cell = cell;
  ctx->coo_val = val;
  ctx->coo_nnz = nnz;
  ctx->coo_dup = 0;
  int rc = yalps_solve(ctx, height, width, reinterpret_cast<const double *>(val ? (const void *)val : (const void *)ctx), ints,
                       nints, sign, opt, status, result, out_height, rhs_out, pos_out, var_out, root_status, root_value,
                       root_pivots, stats);
  if (rc == 1) {  // duplicates with different values: keep the last store of every cell (src/tableau.ts:100-117), again
    std::unordered_map<int32_t, int64_t> last;
    last.reserve((size_t)nnz);
    for (int64_t i = 0; i < nnz; i++) last[cell[i]] = i;
    std::vector<int32_t> ucell;
    std::vector<double> uval;
    for (int64_t i = 0; i < nnz; i++)
      if (last[cell[i]] == i) {
        ucell.push_back(cell[i]);
        uval.push_back(val[i]);
      }
    ctx->coo_cell = ucell.data();
    ctx->coo_val = uval.data();
    ctx->coo_nnz = (int64_t)ucell.size();
    ctx->coo_dup = 0;
    rc = yalps_solve(ctx, height, width, reinterpret_cast<const double *>((const void *)val), ints, nints, sign, opt, status,
                     result, out_height, rhs_out, pos_out, var_out, root_status, root_value, root_pivots, stats);
    if (rc == 1) rc = fail(ctx, YALPS_ERR_CUDA, "cell scatter did not verify after the duplicates were resolved");
  }
  ctx->coo_cell = nullptr;
  ctx->coo_val = nullptr;
  ctx->coo_nnz = -1;
  return rc;
}

}  // extern "C"
