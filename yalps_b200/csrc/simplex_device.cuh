// simplex_device.cuh -- device-side two-phase tableau simplex for one LP per CTA.
//
// Parallel formulation of src/simplex.ts (reference paths relative to the YALPS
// repository): `pivot` 5-39, `hasCycle` 44-63, `phase2` 66-103, `phase1` 106-142.
// The reference is a sequential scalar loop; here every selection is a
// lexicographic (value, index) arg-reduction that reproduces the sequential
// "strict comparison, first index wins" outcome, and the Gauss-Jordan update is a
// column-parallel rank-1 update.  Arithmetic is IEEE binary64 with the reference's
// rounding sequence: products and differences are separate roundings
// (__dmul_rn / __dsub_rn are never contracted to FMA) and every quotient is a
// true division (__ddiv_rn), so trajectories are bit-identical.
//
// Data layout of one LP (LpView): the RHS column (column 0 of the reference
// tableau) is split from the coefficient block,
//     b[r*ldb]          = M[r, 0]
//     A[r*ldA + (c-1)]  = M[r, c],  1 <= c < W
// so that for the shared-memory resident kernel A rows are 16-byte aligned and a
// lane updates two adjacent cells per ld/st.shared.v2.f64 (VW = 2).  The HBM
// resident kernel views the reference layout in place (A = M+1, b = M, ldA = ldb
// = W) with 8-byte accesses (VW = 1).
//
// Work split: thread t owns the vector-columns t, t+NT, ... (KC of them, pivot-row
// values kept in registers from the normalisation to the update) and walks down
// ALL rows, so the per-row pivot-column coefficient is one broadcast load per
// warp and the inner loop is  ld - mul - sub - st  with no index arithmetic.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "fastdiv.cuh"

namespace yalps {

constexpr double kTiny = 1e-16;  // sparsity threshold of src/simplex.ts:18,31
constexpr int kNone = 0x7fffffff;

enum : int {
  ST_OPTIMAL = 0,
  ST_INFEASIBLE = 1,
  ST_UNBOUNDED = 2,
  ST_TIMEDOUT = 3,
  ST_CYCLED = 4,
  ST_ERR_HISTORY = -4,
  ST_ERR_PEER = -5,  // grid-wide / multi-GPU kernels: a participant of an exchange never showed up (spin budget exhausted)
  ST_ERR_POOL = -6,  // device-resident branch and cut: a speculative node that a pool (cuts, cut rows, candidates) could not serve
};

__device__ __forceinline__ double d_inf() { return __longlong_as_double(0x7ff0000000000000LL); }
__device__ __forceinline__ double d_nan() { return __longlong_as_double(0x7ff8000000000000LL); }

// JS Math.round: halves toward +inf, keeps -0 (src/util.ts:2-3).
__device__ __forceinline__ double js_round(double x) {
  if (!(fabs(x) < 4503599627370496.0)) return x;
  double r = floor(x);
  if (__dsub_rn(x, r) >= 0.5) r = __dadd_rn(r, 1.0);
  if (r == 0.0 && (__double2hiint(x) < 0)) r = -0.0;
  return r;
}

// src/util.ts:1-4
__device__ __forceinline__ double round_to_precision(double num, double precision) {
  const double rounding = js_round(__ddiv_rn(1.0, precision));
  const double shifted = __dadd_rn(num, 2.220446049250313e-16);
  return __ddiv_rn(js_round(__dmul_rn(shifted, rounding)), rounding);
}

template <int NW>
__device__ __forceinline__ void cta_sync() {
  if (NW == 1)
    __syncwarp();
  else
    __syncthreads();
}

// ---- arg-reductions on order-preserving integer keys ---------------------------------------------------
// key(x) is monotone in x with key(-0) == key(+0); a lane without a candidate carries kNoKey and kNone.
// "Best" = largest (kMax) or smallest key, lowest index on ties: the parallel equivalent of the reference's
// ascending scans with strict `>` / `<`.  Three redux.sync per warp instead of a 5-step shuffle butterfly.
__device__ __forceinline__ unsigned long long order_key(double x) {
  const long long bits = __double_as_longlong(__dadd_rn(x, 0.0));  // -0 -> +0
  return bits < 0 ? ~(unsigned long long)bits : ((unsigned long long)bits | 0x8000000000000000ULL);
}

template <bool kMax>
__device__ __forceinline__ unsigned long long no_key() {
  return kMax ? 0ULL : ~0ULL;
}

struct Best {
  unsigned hi, lo;
  int idx;
};

template <bool kMax>
__device__ __forceinline__ Best warp_best(unsigned hi, unsigned lo, int idx) {
  Best w;
  if (kMax) {
    w.hi = __reduce_max_sync(0xffffffffu, hi);
    w.lo = __reduce_max_sync(0xffffffffu, hi == w.hi ? lo : 0u);
  } else {
    w.hi = __reduce_min_sync(0xffffffffu, hi);
    w.lo = __reduce_min_sync(0xffffffffu, hi == w.hi ? lo : 0xffffffffu);
  }
  w.idx = (int)__reduce_min_sync(0xffffffffu, (hi == w.hi && lo == w.lo) ? (unsigned)idx : (unsigned)kNone);
  return w;
}

// CTA-wide: every thread returns the winning index (kNone if no lane had a candidate).
// red holds 2 x 3 x 32 words, used alternately (parity) so that one barrier per reduction suffices.
template <bool kMax, int NW>
__device__ __forceinline__ int block_best(unsigned long long key, int idx, unsigned *red, int &parity) {
  Best w = warp_best<kMax>((unsigned)(key >> 32), (unsigned)key, idx);
  if (NW == 1) return w.idx;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned *r = red + parity * 96;
  parity ^= 1;
  if (lane == 0) {
    r[warp] = w.hi;
    r[32 + warp] = w.lo;
    r[64 + warp] = (unsigned)w.idx;
  }
  __syncthreads();
  const unsigned long long nk = no_key<kMax>();
  const unsigned hi = lane < NW ? r[lane] : (unsigned)(nk >> 32);
  const unsigned lo = lane < NW ? r[32 + lane] : (unsigned)nk;
  const int id = lane < NW ? (int)r[64 + lane] : kNone;
  return warp_best<kMax>(hi, lo, id).idx;
}

struct LpView {
  double *A;  // coefficient block, A[r*ldA + j] = M[r, j+1]
  double *b;  // RHS column, b[r*ldb] = M[r, 0]
  int ldA, ldb;
  int H, W;   // reference tableau shape (W counts the RHS column)
  int *var;   // variableAtPosition[W+H]; positionOfVariable is its inverse and is rebuilt on output
};

struct Scratch {
  double *colbuf;     // [H]  pivot column before the update; 0 for rows the update must not touch
  double *colnew;     // [H]  -coef/q, element r at colnew[r*ldc]
  int ldc;
  double *misc;       // [2]  normalised RHS of the pivot row, its non-zero flag
  unsigned *red;      // [192] cross-warp reduction scratch (NW > 1 only)
  int *hist;       // [2*hist_cap] (leaving var, entering var) pairs, checkCycles only
  int hist_cap;
  const long long *resume;  // nullable: [phase, pivots done in phase 1, in phase 2] of a trajectory to continue
};

struct LpResult {
  int status;
  double value;
  long long p1, p2;
  // rows rewritten by the rank-1 updates (R of SURVEY 8d, summed over the pivots): THIS THREAD's share in K1/K2 (every
  // thread counts the rows whose pivot-column cell it handled), the whole LP's total in the row-split kernels
  unsigned long long rows;
};

// src/simplex.ts:44-63 after the push: does the history end in two identical runs of length 6..len/2 ?
template <int NT>
__device__ __forceinline__ bool history_has_cycle(const int *hist, int len) {
  bool found = false;
  for (int L = 6 + (int)threadIdx.x; L <= len / 2 && !found; L += NT) {
    bool cyc = true;
    for (int i = 0; i < L; i++) {
      const int item = len - 1 - i;
      if (hist[2 * item] != hist[2 * (item - L)] || hist[2 * item + 1] != hist[2 * (item - L) + 1]) {
        cyc = false;
        break;
      }
    }
    found = cyc;
  }
  return __syncthreads_or(found) != 0;
}

template <int VW>
struct Cells;
template <>
struct Cells<1> {
  double x;
  __device__ __forceinline__ void load(const double *p) { x = *p; }
  __device__ __forceinline__ double get(int) const { return x; }
};
template <>
struct Cells<2> {
  double2 v;
  __device__ __forceinline__ void load(const double *p) { v = *reinterpret_cast<const double2 *>(p); }
  __device__ __forceinline__ double get(int e) const { return e ? v.y : v.x; }
};

// RU rows of the rank-1 update for this thread's KC vector-columns, as one straight-line block: all loads are
// issued before the arithmetic so that RU*KC independent ld-mul-sub-st chains are in flight per thread.
//   p     normalised pivot-row cells (0.0 where the old cell was flushed)
//   st    bit k*VW+e: cell e of vector-column k is rewritten (old pivot-row cell > 1e-16, or padding)
//   full  bit k: all VW cells of vector-column k are rewritten
//   coef  pivot-column coefficient of each row of the block; 0.0 marks a row the update must not touch
// kPartial = some thread of the CTA has a vector-column with only one of its two cells rewritten.
template <int NT, int KC, int VW, int RU, bool kPartial>
__device__ __forceinline__ void update_rows(double *__restrict__ Ar, int ldA, const double (&p)[KC][VW], unsigned st,
                                            unsigned full, const double (&coef)[RU]) {
  Cells<VW> x[RU][KC];
#pragma unroll
  for (int i = 0; i < RU; i++)
#pragma unroll
    for (int k = 0; k < KC; k++) {
      const bool need = kPartial ? (((st >> (k * VW)) & ((1u << VW) - 1u)) != 0u) : (((full >> k) & 1u) != 0u);
      if (coef[i] != 0.0 && need) x[i][k].load(Ar + (size_t)i * ldA + (size_t)VW * NT * k);
    }
#pragma unroll
  for (int i = 0; i < RU; i++)
#pragma unroll
    for (int k = 0; k < KC; k++) {
      double *dst = Ar + (size_t)i * ldA + (size_t)VW * NT * k;
      const bool on = coef[i] != 0.0;
      if (VW == 2) {
        const double t0 = __dsub_rn(x[i][k].get(0), __dmul_rn(coef[i], p[k][0]));
        const double t1 = __dsub_rn(x[i][k].get(1), __dmul_rn(coef[i], p[k][VW - 1]));
        if (on && ((full >> k) & 1u)) {
          *reinterpret_cast<double2 *>(dst) = make_double2(t0, t1);
        } else if (kPartial && on) {
          if ((st >> (k * VW)) & 1u) dst[0] = t0;
          if ((st >> (k * VW + 1)) & 1u) dst[1] = t1;
        }
      } else {
        const double t0 = __dsub_rn(x[i][k].get(0), __dmul_rn(coef[i], p[k][0]));
        if (on && ((full >> k) & 1u)) dst[0] = t0;
      }
    }
}

template <int NT, int KC, int VW, int RU, bool kPartial>
__device__ __forceinline__ void update_all(double *__restrict__ Abase, int ldA, int H, const double *colbuf,
                                           const double (&p)[KC][VW], unsigned st, unsigned full) {
  int r = 0;
  for (; r + RU <= H; r += RU) {
    double coef[RU];
    if (RU % 2 == 0) {  // colbuf is 16-byte aligned and r is a multiple of RU: two coefficients per load
#pragma unroll
      for (int i = 0; i < RU; i += 2) {
        const double2 c2 = *reinterpret_cast<const double2 *>(colbuf + r + i);
        coef[i] = c2.x;
        coef[i + (RU > 1 ? 1 : 0)] = c2.y;
      }
    } else {
#pragma unroll
      for (int i = 0; i < RU; i++) coef[i] = colbuf[r + i];
    }
    bool any = false;
#pragma unroll
    for (int i = 0; i < RU; i++) any |= coef[i] != 0.0;
    if (any) update_rows<NT, KC, VW, RU, kPartial>(Abase + (size_t)r * ldA, ldA, p, st, full, coef);
  }
  for (; r < H; r++) {
    double coef[1] = {colbuf[r]};
    if (coef[0] != 0.0) update_rows<NT, KC, VW, 1, kPartial>(Abase + (size_t)r * ldA, ldA, p, st, full, coef);
  }
}

// src/simplex.ts:5-39.
template <int NW, int KC, int VW>
__device__ __forceinline__ int pivot_cta(const LpView &t, const Scratch &s, int row, int col) {
  constexpr int NT = NW * 32;
  constexpr int RU = KC == 1 ? 8 : (KC == 2 ? 4 : (KC <= 4 ? 2 : 1));
  const int tid = threadIdx.x;
  double *__restrict__ A = t.A;
  double *__restrict__ b = t.b;
  const int ldA = t.ldA, ldb = t.ldb, H = t.H, Wm1 = t.W - 1;
  const int jc = col - 1;
  const double q = A[(size_t)row * ldA + jc];
  // One divisor for every quotient of the pivot: branch-free quotients, one acceptance test per batch and the exact
  // division out of line for the operands the fast sequence rejects (fastdiv.cuh RecipBatch).  The per-quotient
  // test-and-branch forms (__ddiv_rn, Recip::quot) cost 10-14 % in the throughput kernels.
  RecipBatch rq(q);
  // ---- normalise the pivot row into registers (:16-25); A and b are only read in this phase.
  // The pivot cell itself becomes 1/q (:25): same division code path with numerator 1.
  double p[KC][VW];
  unsigned st = 0, full = 0, valid = 0, partial = 0;
  {
    const double *Arow = A + (size_t)row * ldA + VW * tid;
#pragma unroll
    for (int k = 0; k < KC; k++) {
      const int j0 = VW * (tid + NT * k);
      Cells<VW> v;
      if (j0 < Wm1) v.load(Arow + (size_t)VW * NT * k);
#pragma unroll
      for (int e = 0; e < VW; e++) {
        p[k][e] = 0.0;
        const int j = j0 + e;
        if (j < Wm1) {
          valid |= 1u << (k * VW + e);
          const double x = (j == jc) ? 1.0 : v.get(e);
          const bool nzx = fabs(x) > kTiny;
          const double quo = rq.quot(x, nzx);
          if (nzx) {
            p[k][e] = quo;
            st |= 1u << (k * VW + e);  // the pivot column is rewritten too and fixed up after the update
          }
        } else if (VW == 2 && j0 < Wm1) {
          st |= 1u << (k * VW + e);    // padding cell next to the last column: rewriting it is harmless
        }
      }
      const unsigned m = (st >> (k * VW)) & ((1u << VW) - 1u);
      if (m == (1u << VW) - 1u)
        full |= 1u << k;
      else if (m)
        partial = 1u;
    }
  }
  if (!rq.ok) {  // rare: some quotient of this thread needs the exact division (the pivot row is still unchanged)
    const double *Arow = A + (size_t)row * ldA + VW * tid;
#pragma unroll
    for (int k = 0; k < KC; k++)
#pragma unroll
      for (int e = 0; e < VW; e++) {
        const int j = VW * (tid + NT * k) + e;
        if (j < Wm1) {
          const double x = (j == jc) ? 1.0 : Arow[(size_t)VW * NT * k + e];
          if (fabs(x) > kTiny) p[k][e] = div_rn_slow(x, q);
        }
      }
    rq.reset();
  }
  const bool any_partial = (VW == 2) && (NW == 1 ? __any_sync(0xffffffffu, partial) : __syncthreads_or((int)partial));
  // old pivot column, -coef/q, and the RHS cell of the pivot row (:19 for c = 0, :28-36): one division each
  int rewritten = 0;  // rows of mine that the update touches
  for (int r = tid; r < H; r += NT) {
    const double coef = A[(size_t)r * ldA + jc];
    const double num = (r == row) ? b[(size_t)row * ldb] : -coef;
    const bool nz = fabs(num) > kTiny;  // also false for NaN, as in the reference
    double quo = rq.quot(num, nz);
    if (!rq.ok) {
      quo = div_rn_slow(num, q);
      rq.reset();
    }
    quo = nz ? quo : 0.0;
    if (r == row) {
      s.misc[0] = quo;
      s.misc[1] = nz ? 1.0 : 0.0;
      s.colbuf[r] = 0.0;  // the pivot row is not updated by the rank-1 pass
    } else {
      s.colbuf[r] = nz ? coef : 0.0;  // row skip (:31)
      s.colnew[(size_t)r * s.ldc] = quo;
      rewritten += nz;
    }
  }
  if (tid == 0) {  // basis bookkeeping (:7-12)
    const int leaving = t.var[t.W + row];
    t.var[t.W + row] = t.var[col];
    t.var[col] = leaving;
  }
  cta_sync<NW>();

  // ---- rank-1 update of every active row (the pivot-column cells get a throw-away value here)
  if (any_partial)
    update_all<NT, KC, VW, RU, true>(A + VW * tid, ldA, H, s.colbuf, p, st, full);
  else
    update_all<NT, KC, VW, RU, false>(A + VW * tid, ldA, H, s.colbuf, p, st, full);
  // pivot row (:19,22,25)
  {
    double *Arow = A + (size_t)row * ldA + VW * tid;
#pragma unroll
    for (int k = 0; k < KC; k++) {
      double *dst = Arow + (size_t)VW * NT * k;
      if (VW == 2) {
        if ((valid >> (k * VW)) & 1u) *reinterpret_cast<double2 *>(dst) = make_double2(p[k][0], p[k][VW - 1]);
      } else {
        if ((valid >> k) & 1u) dst[0] = p[k][0];
      }
    }
  }
  cta_sync<NW>();
  // RHS column (:34 for c = 0) and pivot column (:36)
  {
    const double p0 = s.misc[0];
    const bool nz0 = s.misc[1] != 0.0;
    for (int r = tid; r < H; r += NT) {
      const double coef = s.colbuf[r];
      if (r == row) {
        b[(size_t)r * ldb] = p0;
      } else if (coef != 0.0) {
        if (nz0) {
          const double x = b[(size_t)r * ldb];
          b[(size_t)r * ldb] = __dsub_rn(x, __dmul_rn(coef, p0));
        }
        A[(size_t)r * ldA + jc] = s.colnew[(size_t)r * s.ldc];
      }
    }
  }
  cta_sync<NW>();
  return rewritten;
}

// src/simplex.ts:106-142 (phase1) falling through to 66-103 (phase2); the whole CTA executes this uniformly.
template <int NW, int KC, int VW>
__device__ __forceinline__ LpResult simplex_cta(const LpView &t, const Scratch &s, double precision, double max_pivots,
                                                int check_cycles) {
  constexpr int NT = NW * 32;
  const int tid = threadIdx.x;
  const double *__restrict__ A = t.A;
  const double *__restrict__ b = t.b;
  const int ldA = t.ldA, ldb = t.ldb, H = t.H, Wm1 = t.W - 1;
  const double INF = d_inf();

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
        const double *Arow = A + (size_t)row * ldA + VW * tid;
        const double *A0 = A + VW * tid;
#pragma unroll
        for (int k = 0; k < KC; k++) {
          const int j0 = VW * (tid + NT * k);
          if (j0 < Wm1) {
            Cells<VW> cf, ob;
            cf.load(Arow + (size_t)VW * NT * k);
            ob.load(A0 + (size_t)VW * NT * k);
#pragma unroll
            for (int e = 0; e < VW; e++) {
              const double coef = cf.get(e);
              if (j0 + e < Wm1 && coef < -precision) {
                const double ratio = div_rn(-ob.get(e), coef);
                if (ratio > bv) {  // bv starts at -inf: -inf and NaN ratios never win, as in the reference
                  bv = ratio;
                  bi = j0 + e + 1;
                }
              }
            }
          }
        }
      }
      col = block_best<true, NW>(bi == kNone ? no_key<true>() : order_key(bv), bi, s.red, parity);
      if (col == kNone) {
        res.status = ST_INFEASIBLE;
        break;
      }
    } else {
      // entering column: first index of the largest reduced cost above precision (:71-79)
      double bv = -INF;
      int bi = kNone;
      {
        const double *A0 = A + VW * tid;
#pragma unroll
        for (int k = 0; k < KC; k++) {
          const int j0 = VW * (tid + NT * k);
          if (j0 < Wm1) {
            Cells<VW> ob;
            ob.load(A0 + (size_t)VW * NT * k);
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
      }
      col = block_best<true, NW>(bi == kNone ? no_key<true>() : order_key(bv), bi, s.red, parity);
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
          const double ratio = div_rn(b[(size_t)r * ldb], v);
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

    res.rows += (unsigned)pivot_cta<NW, KC, VW>(t, s, row, col);
    if (phase == 1)
      res.p1++;
    else
      res.p2++;
    iter++;
  }
  return res;
}

}  // namespace yalps
