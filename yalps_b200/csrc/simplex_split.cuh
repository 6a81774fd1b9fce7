// simplex_split.cuh -- latency-oriented work split of the same algorithm as simplex_device.cuh
// (src/simplex.ts:5-142 of the reference), for LPs that get a whole multi-warp CTA: single solve() calls,
// branch-and-cut node waves, and batches whose tableaus leave room for only one or two CTAs per SM.
//
// A pivot on one CTA is a dependent chain (select column -> select row -> normalise -> update), and on this
// machine a dependent FP64 op costs ~10 cycles, a division ~130, a warp reduction ~25 per REDUX, a CTA barrier
// 30-90 (scripts/micro/latency.cu).  The kernel is organised to keep that chain short:
//   * a CTA is NWR row groups x NWC column warps.  Thread (rg, ctid) owns the vector-columns ctid, ctid+NTC, ...
//     (KC of them; every row group holds the same normalised pivot-row cells in registers) and updates only the
//     rows dealt to its row group;
//   * the rows the reference rewrites (|pivot-column coefficient| > 1e-16, src/simplex.ts:31) are compacted
//     into a list while the pivot column is read: the update costs ceil(R / (NWR*RU)) straight-line blocks per
//     thread instead of ceil(H / RU), which is what sparse Netlib-style tableaus need (R << H);
//   * the list carries (coef, -coef/q) pairs and the row index, so the owner of the pivot column writes the final
//     pivot-column cell right after its own rank-1 store, and the RHS cells are updated by the last row group
//     next to the update: no fix-up pass, two CTA barriers per pivot plus one for the cross-warp selection;
//   * all quotients by the pivot element (pivot row, pivot column, RHS) share one reciprocal refinement
//     (fastdiv.cuh, bit-identical to IEEE division), and zero numerators skip nvcc's division slow path;
//   * cheap scans (objective row, RHS column) are done redundantly by every warp when one warp covers them, so
//     they need no barrier; only the ratio tests (one division per candidate) are split over the whole CTA.
// Arithmetic, skip rules and tie-breaking are those of simplex_device.cuh: results are bit-identical.
#pragma once

#include "fastdiv.cuh"
#include "simplex_device.cuh"

namespace yalps {

#ifdef YALPS_TIMING
#define YT_MARK(k) do { const long long now_ = clock64(); s.yt[k] += now_ - s.yt[15]; s.yt[15] = now_; } while (0)
#else
#define YT_MARK(k)
#endif

struct SplitScratch {
  double *cc;     // [2*Hpad] (coef, -coef/q) of the active rows of the current pivot, compacted
  int *list;      // [Hpad]   their row indices
  int *cnt;       // number of active rows (reset by thread 0 after every pivot)
  double *misc;   // [2] normalised RHS of the pivot row, its non-zero flag
  unsigned *red;  // [192] cross-warp reduction scratch
  int *hist;
  int hist_cap;
  // resume (nullable): [phase, pivots already done in phase 1, in phase 2] -- the LP continues a trajectory that was
  // followed elsewhere up to this tableau (replica path sharing, aux_kernels.cuh)
  const long long *resume;
  // trace (nullable, one LP per launch): per pivot k the choice and the raw pivot column, and a snapshot of the
  // tableau and of variableAtPosition BEFORE pivot k; entry K (= number of pivots) describes how the LP ended
  int *tr_steps;       // [cap + 1][4]: phase, row, col, 0  (kNone where nothing was chosen)
  double *tr_q;        // [cap]
  double *tr_col;      // [cap][tr_hp]
  double *tr_snap;     // [cap + 1][H * W], reference layout
  int *tr_var;         // [cap + 1][W + H]
  int tr_hp, tr_cap;
#ifdef YALPS_TIMING
  long long *yt;  // [16] per-thread phase cycle counters (debug builds only)
#endif
};

// RU active rows for this thread's KC vector-columns: loads first, then arithmetic and stores.
// kTail: some of the RU rows may be past the end of the list (n_in valid rows).
template <int NTC, int KC, int VW, int RU, bool kPartial, bool kTail>
__device__ __forceinline__ void update_block_split(double *__restrict__ Ac, int ldA, const int *list, const double *cc,
                                                   int i, int n_in, const double (&p)[KC][VW], unsigned st,
                                                   unsigned full, int jc_off) {
  int ridx[RU];
  double coef[RU], cnew[RU];
#pragma unroll
  for (int j = 0; j < RU; j++) {
    const double2 c2 = *reinterpret_cast<const double2 *>(cc + 2 * (i + j));  // padded: always readable
    ridx[j] = (!kTail || j < n_in) ? list[i + j] : 0;
    coef[j] = c2.x;
    cnew[j] = c2.y;
  }
  Cells<VW> x[RU][KC];
#pragma unroll
  for (int j = 0; j < RU; j++)
#pragma unroll
    for (int k = 0; k < KC; k++) {
      const bool need = kPartial ? (((st >> (k * VW)) & ((1u << VW) - 1u)) != 0u) : (((full >> k) & 1u) != 0u);
      if ((!kTail || j < n_in) && need) x[j][k].load(Ac + (size_t)ridx[j] * ldA + (size_t)VW * NTC * k);
    }
#pragma unroll
  for (int j = 0; j < RU; j++) {
    const bool on = !kTail || j < n_in;
    double *rowp = Ac + (size_t)ridx[j] * ldA;
#pragma unroll
    for (int k = 0; k < KC; k++) {
      double *dst = rowp + (size_t)VW * NTC * k;
      if (VW == 2) {
        const double t0 = __dsub_rn(x[j][k].get(0), __dmul_rn(coef[j], p[k][0]));
        const double t1 = __dsub_rn(x[j][k].get(1), __dmul_rn(coef[j], p[k][VW - 1]));
        if (on && ((full >> k) & 1u)) {
          *reinterpret_cast<double2 *>(dst) = make_double2(t0, t1);
        } else if (kPartial && on) {
          if ((st >> (k * VW)) & 1u) dst[0] = t0;
          if ((st >> (k * VW + 1)) & 1u) dst[1] = t1;
        }
      } else {
        const double t0 = __dsub_rn(x[j][k].get(0), __dmul_rn(coef[j], p[k][0]));
        if (on && ((full >> k) & 1u)) dst[0] = t0;
      }
    }
    // owner of the pivot column: the cell becomes -coef/q (src/simplex.ts:36), after this thread's own store
    if (jc_off >= 0 && on) rowp[jc_off] = cnew[j];
  }
}

template <int NTC, int KC, int VW, int RU, int NWR, bool kPartial>
__device__ __forceinline__ void update_split(double *__restrict__ Ac, int ldA, int R, int rg, const int *list,
                                             const double *cc, const double (&p)[KC][VW], unsigned st, unsigned full,
                                             int jc_off) {
  for (int i = rg * RU; i < R; i += NWR * RU) {
    if (i + RU <= R)
      update_block_split<NTC, KC, VW, RU, kPartial, false>(Ac, ldA, list, cc, i, RU, p, st, full, jc_off);
    else
      update_block_split<NTC, KC, VW, RU, kPartial, true>(Ac, ldA, list, cc, i, R - i, p, st, full, jc_off);
  }
}

// src/simplex.ts:5-39.
template <int NWC, int KC, int NWR, int VW>
__device__ __forceinline__ int pivot_split(const LpView &t, const SplitScratch &s, int row, int col) {
  constexpr int NTC = NWC * 32, NT = NTC * NWR;
  constexpr int RU = (VW == 1) ? (KC >= 8 ? 1 : 8 / KC) : (KC >= 4 ? 1 : 4 / KC);
  const int tid = threadIdx.x, lane = tid & 31;
  const int ctid = tid % NTC, rg = tid / NTC;
  double *__restrict__ A = t.A;
  double *__restrict__ b = t.b;
  const int ldA = t.ldA, ldb = t.ldb, H = t.H, Wm1 = t.W - 1;
  const int jc = col - 1;

  // ---- issue every load of this phase before the first division
  const double q = A[(size_t)row * ldA + jc];
  Cells<VW> v[KC];
  {
    const double *Arow = A + (size_t)row * ldA + VW * ctid;
#pragma unroll
    for (int k = 0; k < KC; k++)
      if (VW * (ctid + NTC * k) < Wm1) v[k].load(Arow + (size_t)VW * NTC * k);
  }
  double cell0 = 0.0;  // pivot-column cell of row `tid` (first pass of the column loop); RHS for the pivot row
  if (tid < H) cell0 = (tid == row) ? b[(size_t)row * ldb] : A[(size_t)tid * ldA + jc];
  // HBM/L2-resident tableaus (VW == 1): the cells of the next passes of the column loop leave in the same flight -- a
  // 64-thread CTA walks the 151 rows of SC105 in three passes, and each pass used to be a dependent trip to L2
  constexpr int CX = (VW == 1) ? 2 : 0;
  double cellx[CX > 0 ? CX : 1];
  if constexpr (CX > 0) {
#pragma unroll
    for (int u = 0; u < CX; u++) {
      const int r = tid + (u + 1) * NT;
      cellx[u] = 0.0;
      if (r < H) cellx[u] = (r == row) ? b[(size_t)row * ldb] : A[(size_t)r * ldA + jc];
    }
  }
  const Recip rq(q);

  // ---- normalise the pivot row into registers (:16-25), every row group for itself
  double p[KC][VW];
  unsigned st = 0, full = 0, valid = 0, partial = 0;
  int jc_off = -1;  // owner thread of the pivot column: offset of that cell from this thread's column base
#pragma unroll
  for (int k = 0; k < KC; k++) {
    const int j0 = VW * (ctid + NTC * k);
#pragma unroll
    for (int e = 0; e < VW; e++) {
      p[k][e] = 0.0;
      const int j = j0 + e;
      if (j < Wm1) {
        valid |= 1u << (k * VW + e);
        if (j == jc) jc_off = VW * NTC * k + e;
        const double x = (j == jc) ? 1.0 : v[k].get(e);
        if (fabs(x) > kTiny) {
          p[k][e] = rq.quot(x);
          st |= 1u << (k * VW + e);
        }
      } else if (VW == 2 && j0 < Wm1) {
        st |= 1u << (k * VW + e);  // padding cell next to the last column: rewriting it is harmless
      }
    }
    const unsigned m = (st >> (k * VW)) & ((1u << VW) - 1u);
    if (m == (1u << VW) - 1u)
      full |= 1u << k;
    else if (m)
      partial = 1u;
  }
  YT_MARK(2);
  // ---- pivot column: -coef/q per row (:36), normalised RHS of the pivot row (:19 for c = 0), and the
  // compacted list of rows the update rewrites (:31)
  auto column_pass = [&](const int r0, const double cell_in) {
    const int r = r0 + tid;
    bool act = false;
    double cell = 0.0, quo = 0.0;
    if (r < H) {
      cell = cell_in;
      const double num = (r == row) ? cell : -cell;
      const bool nz = fabs(num) > kTiny;  // also false for NaN, as in the reference
      quo = nz ? rq.quot(num) : 0.0;
      if (r == row) {
        s.misc[0] = quo;
        s.misc[1] = nz ? 1.0 : 0.0;
      } else {
        act = nz;
      }
    }
    const unsigned m = __ballot_sync(0xffffffffu, act);
    if (m) {
      int base = 0;
      if (lane == 0) base = atomicAdd(s.cnt, __popc(m));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (act) {
        const int k = base + __popc(m & ((1u << lane) - 1u));
        s.list[k] = r;
        *reinterpret_cast<double2 *>(s.cc + 2 * k) = make_double2(cell, quo);
      }
    }
  };
  column_pass(0, cell0);
  if constexpr (CX > 0) {
#pragma unroll
    for (int u = 0; u < CX; u++)
      if ((u + 1) * NT < H) column_pass((u + 1) * NT, cellx[u]);
  }
  for (int r0 = (CX + 1) * NT; r0 < H; r0 += NT) {
    const int r = r0 + tid;
    double cell = 0.0;
    if (r < H) cell = (r == row) ? b[(size_t)row * ldb] : A[(size_t)r * ldA + jc];
    column_pass(r0, cell);
  }
  if (tid == 0) {  // basis bookkeeping (:7-12)
    const int leaving = t.var[t.W + row];
    t.var[t.W + row] = t.var[col];
    t.var[col] = leaving;
  }
  YT_MARK(3);
  __syncthreads();
  const int R = *s.cnt;
  const bool any_partial = (VW == 2) && __any_sync(0xffffffffu, partial);  // per warp: warps need not agree
  YT_MARK(4);

  // ---- RHS cells of the active rows (:34 for c = 0) and of the pivot row, by the last row group: they are
  // independent of the coefficient block, so no barrier separates them from the update
  if (rg == NWR - 1) {
    const double p0 = s.misc[0];
    if (s.misc[1] != 0.0) {
      for (int i = ctid; i < R; i += NTC) {
        const int r = s.list[i];
        const double x = b[(size_t)r * ldb];
        b[(size_t)r * ldb] = __dsub_rn(x, __dmul_rn(s.cc[2 * i], p0));
      }
    }
    if (ctid == NTC - 1) b[(size_t)row * ldb] = p0;
  }
  YT_MARK(5);
  // ---- rank-1 update of the active rows, pivot-column cell included
  if (any_partial)
    update_split<NTC, KC, VW, RU, NWR, true>(A + VW * ctid, ldA, R, rg, s.list, s.cc, p, st, full, jc_off);
  else
    update_split<NTC, KC, VW, RU, NWR, false>(A + VW * ctid, ldA, R, rg, s.list, s.cc, p, st, full, jc_off);
  // pivot row (:19,22,25), by the last row group (it has the fewest update blocks)
  if (rg == NWR - 1) {
    double *Arow = A + (size_t)row * ldA + VW * ctid;
#pragma unroll
    for (int k = 0; k < KC; k++) {
      double *dst = Arow + (size_t)VW * NTC * k;
      if (VW == 2) {
        if ((valid >> (k * VW)) & 1u) *reinterpret_cast<double2 *>(dst) = make_double2(p[k][0], p[k][VW - 1]);
      } else {
        if ((valid >> k) & 1u) dst[0] = p[k][0];
      }
    }
  }
  YT_MARK(6);
  __syncthreads();
  YT_MARK(7);
  if (tid == 0) *s.cnt = 0;  // next written after the barrier of the next cross-warp selection
  return R;
}

// src/simplex.ts:106-142 (phase1) falling through to 66-103 (phase2); the whole CTA executes this uniformly.
template <int NWC, int KC, int NWR, int VW>
__device__ __forceinline__ LpResult simplex_cta_split(const LpView &t, const SplitScratch &s, double precision,
                                                      double max_pivots, int check_cycles) {
  constexpr int NTC = NWC * 32, NW = NWC * NWR, NT = NW * 32;
  static_assert(NW > 1, "the split kernel needs a cross-warp barrier between pivots");
  const int tid = threadIdx.x, lane = tid & 31, ctid = tid % NTC;
  const double *__restrict__ A = t.A;
  const double *__restrict__ b = t.b;
  const int ldA = t.ldA, ldb = t.ldb, H = t.H, Wm1 = t.W - 1;
  const double INF = d_inf();
  // `(double)iter < max_pivots` for integer iter == `iter < ceil(max_pivots)` (NaN / non-positive budgets: none)
  const long long budget =
      !(max_pivots > 0.0) ? 0LL : (max_pivots >= 9.0e18 ? 0x7fffffffffffffffLL : (long long)ceil(max_pivots));
  const bool warp_rows = H <= 32 * 8;  // one warp scans the RHS column by itself

  LpResult res;
  res.status = ST_CYCLED;
  res.value = d_nan();
  res.p1 = res.p2 = 0;
  res.rows = 0;
  int phase = 1, parity = 0, hist_len = 0;
  long long iter = 0;
  if (s.resume) {  // continue a trajectory: same phase, same per-phase counter
    phase = (int)s.resume[0];
    res.p1 = s.resume[1];
    res.p2 = s.resume[2];
    iter = phase == 1 ? res.p1 : res.p2;
  }
  int end_row = kNone, end_col = kNone;  // what the terminal step had chosen (trace)
  // snapshot of the tableau (reference layout) and of variableAtPosition into trace slot k
  auto trace_snapshot = [&](long long k) {
    const int W = t.W;
    double *dst = s.tr_snap + (size_t)k * H * W;
    for (int e = tid; e < H * W; e += NT) {
      const int r = e / W, c = e - r * W;
      dst[e] = c == 0 ? b[(size_t)r * ldb] : A[(size_t)r * ldA + (c - 1)];
    }
    int *dv = s.tr_var + (size_t)k * (W + H);
    for (int i = tid; i < W + H; i += NT) dv[i] = t.var[i];
  };

  for (;;) {
    if (iter >= budget) break;  // per-phase budget exhausted -> "cycled" (:102,:141)
    int row = kNone, col = kNone;
    if (phase == 1) {
      // leaving row: first index of the most negative RHS below -precision (:111-119)
      double bv = INF;
      int bi = kNone;
      if (warp_rows) {  // every warp scans all rows: no barrier
#pragma unroll 4
        for (int r = 1 + lane; r < H; r += 32) {
          const double v = b[(size_t)r * ldb];
          if (v < -precision && v < bv) {
            bv = v;
            bi = r;
          }
        }
        const unsigned long long key = bi == kNone ? no_key<false>() : order_key(bv);
        row = warp_best<false>((unsigned)(key >> 32), (unsigned)key, bi).idx;
      } else {
        for (int r = 1 + tid; r < H; r += NT) {
          const double v = b[(size_t)r * ldb];
          if (v < -precision && v < bv) {
            bv = v;
            bi = r;
          }
        }
        row = block_best<false, NW>(bi == kNone ? no_key<false>() : order_key(bv), bi, s.red, parity);
      }
      YT_MARK(0);
      if (row == kNone) {  // feasible: phase 2 with a fresh counter and history (:120, :67-69)
        phase = 2;
        iter = 0;
        hist_len = 0;
        continue;
      }
      // entering column: first index of max -M[0,c]/M[row,c] over M[row,c] < -precision (:123-134);
      // one division per candidate, split over the whole CTA
      bv = -INF;
      bi = kNone;
      {
        // (one cell per thread and step: a node LP of 101 columns on 256 threads has one division per thread instead of
        // two in a row on 50 of them)
        const double *Arow = A + (size_t)row * ldA;
        auto cost_ratio = [&](const int j, const double coef, const double obj) {
          if (coef < -precision) {
            const double ratio = div_rn(-obj, coef);
            if (ratio > bv) {  // bv starts at -inf: -inf and NaN ratios never win, as in the reference
              bv = ratio;
              bi = j + 1;
            }
          }
        };
        // HBM/L2-resident tableaus: pivot-row and objective-row cells of the first passes leave in one flight
        constexpr int PX = (VW == 1) ? 2 : 0;
        if constexpr (PX > 0) {
          double cf[PX], ob[PX];
#pragma unroll
          for (int u = 0; u < PX; u++) {
            const int j = tid + u * NT;
            cf[u] = j < Wm1 ? Arow[j] : 0.0;
            ob[u] = j < Wm1 ? A[j] : 0.0;
          }
#pragma unroll
          for (int u = 0; u < PX; u++) {
            const int j = tid + u * NT;
            if (j < Wm1) cost_ratio(j, cf[u], ob[u]);
          }
        }
        for (int j = tid + PX * NT; j < Wm1; j += NT) {
          const double coef = Arow[j];
          if (coef < -precision) cost_ratio(j, coef, A[j]);
        }
      }
      col = block_best<true, NW>(bi == kNone ? no_key<true>() : order_key(bv), bi, s.red, parity);
      YT_MARK(1);
      if (col == kNone) {
        res.status = ST_INFEASIBLE;
        end_row = row;
        break;
      }
    } else {
      // entering column: first index of the largest reduced cost above precision (:71-79); every row group scans
      // its own copy of the columns (with one column warp that is every warp by itself: no barrier)
      double bv = -INF;
      int bi = kNone;
#pragma unroll
      for (int k = 0; k < KC; k++) {
        const int j0 = VW * (ctid + NTC * k);
        if (j0 < Wm1) {
          Cells<VW> ob;
          ob.load(A + j0);
#pragma unroll
          for (int e = 0; e < VW; e++) {
            const double v = ob.get(e);
            if (j0 + e < Wm1 && v > precision && v > bv) {
              bv = v;
              bi = j0 + e + 1;
            }
          }
        }
      }
      {
        const unsigned long long key = bi == kNone ? no_key<true>() : order_key(bv);
        if (NWC == 1)
          col = warp_best<true>((unsigned)(key >> 32), (unsigned)key, bi).idx;
        else
          col = block_best<true, NW>(key, bi, s.red, parity);
      }
      YT_MARK(0);
      if (col == kNone) {
        res.status = ST_OPTIMAL;
        res.value = round_to_precision(b[0], precision);
        break;
      }
      // leaving row: ratio test with the reference's early break (:83-95) == lowest r whose ratio is
      // <= precision if any, else first index of the minimum ratio.  Ratios <= precision get key -inf.
      bv = INF;
      bi = kNone;
      auto ratio_test = [&](const int r, const double v) {
        if (v > precision) {
          const double ratio = div_rn(b[(size_t)r * ldb], v);
          if (ratio < INF) {  // +inf and NaN never win (`ratio < minRatio` with minRatio = Infinity)
            const double key = (ratio <= precision) ? -INF : ratio;
            if (bi == kNone || key < bv) {
              bv = key;
              bi = r;
            }
          }
        }
      };
      // HBM/L2-resident tableaus: the column cells of the first passes leave in one flight (see pivot_split)
      constexpr int SX = (VW == 1) ? 3 : 0;
      if constexpr (SX > 0) {
        double vv[SX];
#pragma unroll
        for (int u = 0; u < SX; u++) {
          const int r = 1 + tid + u * NT;
          vv[u] = r < H ? A[(size_t)r * ldA + (col - 1)] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < SX; u++) {
          const int r = 1 + tid + u * NT;
          if (r < H) ratio_test(r, vv[u]);
        }
      }
      for (int r = 1 + tid + SX * NT; r < H; r += NT) ratio_test(r, A[(size_t)r * ldA + (col - 1)]);
      row = block_best<false, NW>(bi == kNone ? no_key<false>() : order_key(bv), bi, s.red, parity);
      YT_MARK(1);
      if (row == kNone) {
        res.status = ST_UNBOUNDED;
        res.value = (double)col;
        end_col = col;
        break;
      }
    }

    if (s.tr_steps) {  // trace: the choice, the pivot element, the raw pivot column, the tableau before the pivot
      const long long k = res.p1 + res.p2;
      if (k < s.tr_cap) {
        if (tid == 0) {
          s.tr_steps[4 * k] = phase;
          s.tr_steps[4 * k + 1] = row;
          s.tr_steps[4 * k + 2] = col;
          s.tr_steps[4 * k + 3] = 0;
          s.tr_q[k] = A[(size_t)row * ldA + (col - 1)];
        }
        for (int r = tid; r < H; r += NT) s.tr_col[(size_t)k * s.tr_hp + r] = A[(size_t)r * ldA + (col - 1)];
        trace_snapshot(k);
      }
    }

    if (check_cycles) {  // (:98, :137)
      if (hist_len >= s.hist_cap) {
        res.status = ST_ERR_HISTORY;
        break;
      }
      if (tid == 0) {
        s.hist[2 * hist_len] = t.var[t.W + row];
        s.hist[2 * hist_len + 1] = t.var[col];
      }
      hist_len++;
      __syncthreads();
      if (history_has_cycle<NT>(s.hist, hist_len)) break;  // "cycled", NaN
    }

    res.rows += (unsigned)pivot_split<NWC, KC, NWR, VW>(t, s, row, col);
    if (phase == 1)
      res.p1++;
    else
      res.p2++;
    iter++;
  }
  if (s.tr_steps) {  // how the LP ended, and the final tableau
    const long long k = res.p1 + res.p2;
    if (k <= s.tr_cap) {
      if (tid == 0) {
        s.tr_steps[4 * k] = phase;
        s.tr_steps[4 * k + 1] = end_row;
        s.tr_steps[4 * k + 2] = end_col;
        s.tr_steps[4 * k + 3] = 1;
      }
      trace_snapshot(k);
    }
  }
  return res;
}

}  // namespace yalps
