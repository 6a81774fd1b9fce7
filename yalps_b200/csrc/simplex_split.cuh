// simplex_split.cuh -- latency-oriented work split of the same algorithm as simplex_device.cuh
// (src/simplex.ts:5-142 of the reference), for LPs that get a whole multi-warp CTA: single solve() calls,
// branch-and-cut node waves, and batches whose tableaus leave room for only one or two CTAs per SM.
//
// A CTA is NWR row groups x NWC column warps.  Thread (rg, ctid) owns the vector-columns ctid, ctid+NTC, ...
// (KC of them; every row group holds the same normalised pivot-row cells in registers) and updates only the
// rows dealt to its row group.  What makes the pivot short:
//   * the rows the reference rewrites (|pivot-column coefficient| > 1e-16, src/simplex.ts:31) are compacted
//     into a list while the pivot column is read: the update costs ceil(R / (NWR*RU)) straight-line blocks per
//     thread instead of ceil(H / RU), which is what sparse Netlib-style tableaus need (R << H);
//   * the list carries (coef, -coef/q) pairs, so the owner of the pivot column writes the final pivot-column
//     cell during the update itself and the RHS cells are updated next to it: no fix-up pass, two CTA barriers
//     per pivot plus one per selection;
//   * every row loop covers H rows with all NT threads (one division per thread, not ceil(H/32)).
// Arithmetic, skip rules and tie-breaking are those of simplex_device.cuh: results are bit-identical.
#pragma once

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
#ifdef YALPS_TIMING
  long long *yt;  // [16] per-thread phase cycle counters (debug builds only)
#endif
};

// RU active rows (indices ridx) for this thread's KC vector-columns: loads first, then arithmetic and stores.
// jcbits marks the cell of the pivot column (owner thread only): it receives -coef/q (cnew) instead of the
// rank-1 value (src/simplex.ts:36).
template <int NTC, int KC, int VW, int RU, bool kPartial>
__device__ __forceinline__ void update_rows_split(double *__restrict__ Ac, int ldA, const int (&ridx)[RU],
                                                  const double (&coef)[RU], const double (&cnew)[RU],
                                                  const double (&p)[KC][VW], unsigned st, unsigned full,
                                                  unsigned jcbits) {
  Cells<VW> x[RU][KC];
#pragma unroll
  for (int i = 0; i < RU; i++)
#pragma unroll
    for (int k = 0; k < KC; k++) {
      const bool need = kPartial ? (((st >> (k * VW)) & ((1u << VW) - 1u)) != 0u) : (((full >> k) & 1u) != 0u);
      if (coef[i] != 0.0 && need) x[i][k].load(Ac + (size_t)ridx[i] * ldA + (size_t)VW * NTC * k);
    }
#pragma unroll
  for (int i = 0; i < RU; i++)
#pragma unroll
    for (int k = 0; k < KC; k++) {
      double *dst = Ac + (size_t)ridx[i] * ldA + (size_t)VW * NTC * k;
      const bool on = coef[i] != 0.0;
      if (VW == 2) {
        double t0 = __dsub_rn(x[i][k].get(0), __dmul_rn(coef[i], p[k][0]));
        double t1 = __dsub_rn(x[i][k].get(1), __dmul_rn(coef[i], p[k][VW - 1]));
        if ((jcbits >> (k * VW)) & 1u) t0 = cnew[i];
        if ((jcbits >> (k * VW + 1)) & 1u) t1 = cnew[i];
        if (on && ((full >> k) & 1u)) {
          *reinterpret_cast<double2 *>(dst) = make_double2(t0, t1);
        } else if (kPartial && on) {
          if ((st >> (k * VW)) & 1u) dst[0] = t0;
          if ((st >> (k * VW + 1)) & 1u) dst[1] = t1;
        }
      } else {
        double t0 = __dsub_rn(x[i][k].get(0), __dmul_rn(coef[i], p[k][0]));
        if ((jcbits >> k) & 1u) t0 = cnew[i];
        if (on && ((full >> k) & 1u)) dst[0] = t0;
      }
    }
}

template <int NTC, int KC, int VW, int RU, int NWR, bool kPartial>
__device__ __forceinline__ void update_split(double *__restrict__ Ac, int ldA, int R, int rg, const int *list,
                                             const double *cc, const double (&p)[KC][VW], unsigned st, unsigned full,
                                             unsigned jcbits) {
  for (int i = rg * RU; i < R; i += NWR * RU) {
    int ridx[RU];
    double coef[RU], cnew[RU];
#pragma unroll
    for (int j = 0; j < RU; j++) {
      const bool in = i + j < R;
      const double2 c2 = *reinterpret_cast<const double2 *>(cc + 2 * (i + j));  // padded: always readable
      ridx[j] = in ? list[i + j] : 0;
      coef[j] = in ? c2.x : 0.0;
      cnew[j] = c2.y;
    }
    update_rows_split<NTC, KC, VW, RU, kPartial>(Ac, ldA, ridx, coef, cnew, p, st, full, jcbits);
  }
}

// src/simplex.ts:5-39.
template <int NWC, int KC, int NWR, int VW>
__device__ __forceinline__ void pivot_split(const LpView &t, const SplitScratch &s, int row, int col) {
  constexpr int NTC = NWC * 32, NT = NTC * NWR;
  constexpr int RU = (VW == 1) ? (KC >= 8 ? 1 : 8 / KC) : (KC >= 4 ? 1 : 4 / KC);
  const int tid = threadIdx.x, lane = tid & 31;
  const int ctid = tid % NTC, rg = tid / NTC;
  double *__restrict__ A = t.A;
  double *__restrict__ b = t.b;
  const int ldA = t.ldA, ldb = t.ldb, H = t.H, Wm1 = t.W - 1;
  const int jc = col - 1;
  const double q = A[(size_t)row * ldA + jc];

  // ---- normalise the pivot row into registers (:16-25), every row group for itself
  double p[KC][VW];
  unsigned st = 0, full = 0, valid = 0, partial = 0, jcbits = 0;
  {
    const double *Arow = A + (size_t)row * ldA + VW * ctid;
#pragma unroll
    for (int k = 0; k < KC; k++) {
      const int j0 = VW * (ctid + NTC * k);
      Cells<VW> v;
      if (j0 < Wm1) v.load(Arow + (size_t)VW * NTC * k);
#pragma unroll
      for (int e = 0; e < VW; e++) {
        p[k][e] = 0.0;
        const int j = j0 + e;
        if (j < Wm1) {
          valid |= 1u << (k * VW + e);
          if (j == jc) jcbits |= 1u << (k * VW + e);
          const double x = (j == jc) ? 1.0 : v.get(e);
          if (fabs(x) > kTiny) {
            p[k][e] = __ddiv_rn(x, q);
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
  }
  YT_MARK(2);
  // ---- pivot column: -coef/q per row (:36), normalised RHS of the pivot row (:19 for c = 0), and the
  // compacted list of rows the update rewrites (:31)
  for (int r0 = 0; r0 < H; r0 += NT) {
    const int r = r0 + tid;
    bool act = false;
    double coef = 0.0, quo = 0.0;
    if (r < H) {
      coef = A[(size_t)r * ldA + jc];
      const double num = (r == row) ? b[(size_t)row * ldb] : -coef;
      const bool nz = fabs(num) > kTiny;  // also false for NaN, as in the reference
      quo = nz ? __ddiv_rn(num, q) : 0.0;
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
        *reinterpret_cast<double2 *>(s.cc + 2 * k) = make_double2(coef, quo);
      }
    }
  }
  if (tid == 0) {  // basis bookkeeping (:7-12)
    const int leaving = t.var[t.W + row];
    t.var[t.W + row] = t.var[col];
    t.var[col] = leaving;
  }
  YT_MARK(3);
  const bool any_partial = __syncthreads_or((int)partial) != 0;
  const int R = *s.cnt;
  YT_MARK(4);

  // ---- RHS cells of the active rows (:34 for c = 0): independent of the coefficient block, no barrier needed
  {
    const double p0 = s.misc[0];
    if (s.misc[1] != 0.0) {
      for (int i = tid; i < R; i += NT) {
        const int r = s.list[i];
        const double x = b[(size_t)r * ldb];
        b[(size_t)r * ldb] = __dsub_rn(x, __dmul_rn(s.cc[2 * i], p0));
      }
    }
    if (tid == NT - 1) b[(size_t)row * ldb] = p0;
  }
  YT_MARK(5);
  // ---- rank-1 update of the active rows, pivot-column cell included
  if (any_partial)
    update_split<NTC, KC, VW, RU, NWR, true>(A + VW * ctid, ldA, R, rg, s.list, s.cc, p, st, full, jcbits);
  else
    update_split<NTC, KC, VW, RU, NWR, false>(A + VW * ctid, ldA, R, rg, s.list, s.cc, p, st, full, jcbits);
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
  if (tid == 0) *s.cnt = 0;  // next written after the barrier of the next selection
}

// src/simplex.ts:106-142 (phase1) falling through to 66-103 (phase2); the whole CTA executes this uniformly.
template <int NWC, int KC, int NWR, int VW>
__device__ __forceinline__ LpResult simplex_cta_split(const LpView &t, const SplitScratch &s, double precision,
                                                      double max_pivots, int check_cycles) {
  constexpr int NW = NWC * NWR, NT = NW * 32;
  static_assert(NW > 1, "the split kernel needs a cross-warp barrier between pivots");
  const int tid = threadIdx.x;
  const double *__restrict__ A = t.A;
  const double *__restrict__ b = t.b;
  const int ldA = t.ldA, ldb = t.ldb, H = t.H, Wm1 = t.W - 1;
  const double INF = d_inf();

  LpResult res;
  res.status = ST_CYCLED;
  res.value = d_nan();
  res.p1 = res.p2 = 0;
  int phase = 1, parity = 0, hist_len = 0;
  long long iter = 0;

  for (;;) {
    if (!((double)iter < max_pivots)) break;  // per-phase budget exhausted -> "cycled" (:102,:141)
    int row, col;
    if (phase == 1) {
      // leaving row: first index of the most negative RHS below -precision (:111-119)
      double bv = INF;
      int bi = kNone;
      for (int r = 1 + tid; r < H; r += NT) {
        const double v = b[(size_t)r * ldb];
        if (v < -precision && v < bv) {
          bv = v;
          bi = r;
        }
      }
      row = block_best<false, NW>(bi == kNone ? no_key<false>() : order_key(bv), bi, s.red, parity);
      YT_MARK(0);
      if (row == kNone) {  // feasible: phase 2 with a fresh counter and history (:120, :67-69)
        phase = 2;
        iter = 0;
        hist_len = 0;
        continue;
      }
      // entering column: first index of max -M[0,c]/M[row,c] over M[row,c] < -precision (:123-134)
      bv = -INF;
      bi = kNone;
      {
        const double *Arow = A + (size_t)row * ldA;
        for (int j0 = VW * tid; j0 < Wm1; j0 += VW * NT) {
          Cells<VW> cf, ob;
          cf.load(Arow + j0);
          ob.load(A + j0);
#pragma unroll
          for (int e = 0; e < VW; e++) {
            const double coef = cf.get(e);
            if (j0 + e < Wm1 && coef < -precision) {
              const double ratio = __ddiv_rn(-ob.get(e), coef);
              if (ratio > bv) {  // bv starts at -inf: -inf and NaN ratios never win, as in the reference
                bv = ratio;
                bi = j0 + e + 1;
              }
            }
          }
        }
      }
      col = block_best<true, NW>(bi == kNone ? no_key<true>() : order_key(bv), bi, s.red, parity);
      YT_MARK(1);
      if (col == kNone) {
        res.status = ST_INFEASIBLE;
        break;
      }
    } else {
      // entering column: first index of the largest reduced cost above precision (:71-79)
      double bv = -INF;
      int bi = kNone;
      for (int j0 = VW * tid; j0 < Wm1; j0 += VW * NT) {
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
      col = block_best<true, NW>(bi == kNone ? no_key<true>() : order_key(bv), bi, s.red, parity);
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
      for (int r = 1 + tid; r < H; r += NT) {
        const double v = A[(size_t)r * ldA + (col - 1)];
        if (v > precision) {
          const double ratio = __ddiv_rn(b[(size_t)r * ldb], v);
          if (ratio < INF) {  // +inf and NaN never win (`ratio < minRatio` with minRatio = Infinity)
            const double key = (ratio <= precision) ? -INF : ratio;
            if (bi == kNone || key < bv) {
              bv = key;
              bi = r;
            }
          }
        }
      }
      row = block_best<false, NW>(bi == kNone ? no_key<false>() : order_key(bv), bi, s.red, parity);
      YT_MARK(1);
      if (row == kNone) {
        res.status = ST_UNBOUNDED;
        res.value = (double)col;
        break;
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

    pivot_split<NWC, KC, NWR, VW>(t, s, row, col);
    if (phase == 1)
      res.p1++;
    else
      res.p2++;
    iter++;
  }
  return res;
}

}  // namespace yalps
