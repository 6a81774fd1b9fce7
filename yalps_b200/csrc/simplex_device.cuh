// simplex_device.cuh -- device-side two-phase tableau simplex for one LP per CTA.
//
// Parallel formulation of src/simplex.ts (reference paths relative to the YALPS
// repository): `pivot` 5-39, `hasCycle` 44-63, `phase2` 66-103, `phase1` 106-142.
// The reference is a sequential scalar loop; here every selection is a
// lexicographic (value, index) arg-reduction that reproduces the sequential
// "strict comparison, first index wins" outcome, and the Gauss-Jordan update is a
// row-parallel rank-1 update.  Arithmetic is IEEE binary64 with the reference's
// rounding sequence: products and differences are separate roundings
// (__dmul_rn / __dsub_rn are never contracted to FMA) and every quotient is a
// true division (__ddiv_rn), so trajectories are bit-identical.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace yalps {

constexpr double kTiny = 1e-16;      // sparsity threshold of src/simplex.ts:18,31
constexpr int kNone = 0x7fffffff;

enum : int {
  ST_OPTIMAL = 0,
  ST_INFEASIBLE = 1,
  ST_UNBOUNDED = 2,
  ST_TIMEDOUT = 3,
  ST_CYCLED = 4,
  ST_ERR_HISTORY = -4,
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

struct Arg {
  double v;
  int i;
};

// "candidate (v,i) beats incumbent (bv,bi)": strict comparison on the value, lowest index on ties --
// the parallel equivalent of the reference's ascending scans with `>` / `<`.
template <bool kMax>
__device__ __forceinline__ bool beats(double v, int i, double bv, int bi) {
  return kMax ? (v > bv || (v == bv && i < bi)) : (v < bv || (v == bv && i < bi));
}

template <bool kMax>
__device__ __forceinline__ Arg warp_arg(Arg a) {
#pragma unroll
  for (int off = 16; off; off >>= 1) {
    const double v = __shfl_xor_sync(0xffffffffu, a.v, off);
    const int i = __shfl_xor_sync(0xffffffffu, a.i, off);
    if (beats<kMax>(v, i, a.v, a.i)) {
      a.v = v;
      a.i = i;
    }
  }
  return a;
}

template <int NW>
__device__ __forceinline__ void cta_sync() {
  if (NW == 1)
    __syncwarp();
  else
    __syncthreads();
}

// CTA-wide arg-reduction; every thread returns the same winner.  red_v/red_i hold 2x32 slots and are
// used alternately (parity) so that one barrier per reduction is enough.
template <bool kMax, int NW>
__device__ __forceinline__ Arg block_arg(Arg a, double *red_v, int *red_i, int &parity) {
  a = warp_arg<kMax>(a);
  if (NW == 1) return a;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double *rv = red_v + parity * 32;
  int *ri = red_i + parity * 32;
  parity ^= 1;
  if (lane == 0) {
    rv[warp] = a.v;
    ri[warp] = a.i;
  }
  __syncthreads();
  Arg b;
  b.v = lane < NW ? rv[lane] : (kMax ? -d_inf() : d_inf());
  b.i = lane < NW ? ri[lane] : kNone;
  return warp_arg<kMax>(b);
}

struct LpView {
  double *M;  // tableau, row stride ld (shared or global memory)
  int ld;
  int H, W;
  int *pos;  // positionOfVariable[W+H]
  int *var;  // variableAtPosition[W+H]
};

struct Scratch {
  double *prow;      // [W]  normalised pivot row
  double *colbuf;    // [H]  pivot column before the update
  double *colnew;    // [H]  -coef/q
  double *red_v;     // [64]
  int *red_i;        // [64]
  unsigned *nzmask;  // [ceil(W/32)] bit c%32 of word c/32: |old pivot-row cell| > 1e-16
  int *hist;         // [2*hist_cap] (leaving var, entering var) pairs, checkCycles only
  int hist_cap;
};

struct LpResult {
  int status;
  double value;
  long long p1, p2;
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

// src/simplex.ts:5-39.  NW warps, KC pivot-row cells per lane kept in registers per column chunk.
template <int NW, int KC>
__device__ __forceinline__ void pivot_cta(const LpView &t, const Scratch &s, int row, int col) {
  constexpr int NT = NW * 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double *__restrict__ M = t.M;
  const int ld = t.ld, H = t.H, W = t.W;
  const double q = M[(size_t)row * ld + col];

  // ---- snapshot: normalised pivot row, old pivot column, -coef/q (M itself is not modified here)
  for (int cbase = warp * 32; cbase < W; cbase += NT) {
    const int c = cbase + lane;
    bool nz = false;
    if (c < W) {
      const double v = M[(size_t)row * ld + c];
      nz = fabs(v) > kTiny;
      s.prow[c] = nz ? __ddiv_rn(v, q) : 0.0;
    }
    const unsigned m = __ballot_sync(0xffffffffu, nz);
    if (lane == 0) s.nzmask[cbase >> 5] = m;
  }
  for (int r = tid; r < H; r += NT) {
    const double coef = M[(size_t)r * ld + col];
    s.colbuf[r] = coef;
    if (r != row && fabs(coef) > kTiny) s.colnew[r] = __ddiv_rn(-coef, q);
  }
  if (tid == 0) {  // basis bookkeeping, src/simplex.ts:7-12
    const int leaving = t.var[W + row];
    const int entering = t.var[col];
    t.var[W + row] = entering;
    t.var[col] = leaving;
    t.pos[leaving] = col;
    t.pos[entering] = W + row;
  }
  cta_sync<NW>();

  // ---- update: warp w owns rows w, w+NW, ...; lanes span the columns of the chunk
  const double qinv = __ddiv_rn(1.0, q);
  for (int cb = 0; cb < W; cb += 32 * KC) {
    double p[KC];
    unsigned nzb = 0, vb = 0;
    int kcol = -1;
#pragma unroll
    for (int k = 0; k < KC; k++) {
      const int c = cb + lane + 32 * k;
      p[k] = 0.0;
      if (c < W) {
        p[k] = s.prow[c];
        vb |= 1u << k;
        if (c == col)
          kcol = k;
        else if ((s.nzmask[c >> 5] >> lane) & 1u)
          nzb |= 1u << k;
      }
    }
    const bool colchunk = (col >= cb) && (col < cb + 32 * KC);
    for (int r = warp; r < H; r += NW) {
      double *__restrict__ Mr = M + (size_t)r * ld + cb + lane;
      if (r == row) {
#pragma unroll
        for (int k = 0; k < KC; k++)
          if ((vb >> k) & 1u) Mr[32 * k] = (k == kcol) ? qinv : p[k];
        continue;
      }
      const double coef = s.colbuf[r];
      if (!(fabs(coef) > kTiny)) continue;  // row skip, src/simplex.ts:31
#pragma unroll
      for (int k = 0; k < KC; k++) {
        if ((nzb >> k) & 1u) {
          const double x = Mr[32 * k];
          Mr[32 * k] = __dsub_rn(x, __dmul_rn(coef, p[k]));
        }
      }
      if (colchunk && kcol >= 0) Mr[32 * kcol] = s.colnew[r];
    }
  }
  cta_sync<NW>();
}

// src/simplex.ts:106-142 (phase1) falling through to 66-103 (phase2); whole CTA executes this uniformly.
template <int NW, int KC>
__device__ __forceinline__ LpResult simplex_cta(const LpView &t, const Scratch &s, double precision, double max_pivots,
                                                int check_cycles) {
  constexpr int NT = NW * 32;
  const int tid = threadIdx.x;
  const double *__restrict__ M = t.M;
  const int ld = t.ld, H = t.H, W = t.W;
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
      Arg a = {INF, kNone};
      for (int r = 1 + tid; r < H; r += NT) {
        const double v = M[(size_t)r * ld];
        if (v < -precision && v < a.v) {
          a.v = v;
          a.i = r;
        }
      }
      a = block_arg<false, NW>(a, s.red_v, s.red_i, parity);
      if (a.i == kNone) {  // feasible: phase 2 with a fresh counter and history (:120, :67-69)
        phase = 2;
        iter = 0;
        hist_len = 0;
        continue;
      }
      row = a.i;
      // entering column: first index of max -M[0,c]/M[row,c] over M[row,c] < -precision (:123-134)
      Arg b = {-INF, kNone};
      for (int c = 1 + tid; c < W; c += NT) {
        const double coef = M[(size_t)row * ld + c];
        if (coef < -precision) {
          const double ratio = __ddiv_rn(-M[c], coef);
          if (ratio > b.v) {  // b.v starts at -inf: -inf and NaN ratios never win, as in the reference
            b.v = ratio;
            b.i = c;
          }
        }
      }
      b = block_arg<true, NW>(b, s.red_v, s.red_i, parity);
      if (b.i == kNone) {
        res.status = ST_INFEASIBLE;
        break;
      }
      col = b.i;
    } else {
      // entering column: first index of the largest reduced cost above precision (:71-79)
      Arg a = {-INF, kNone};
      for (int c = 1 + tid; c < W; c += NT) {
        const double v = M[c];
        if (v > precision && v > a.v) {
          a.v = v;
          a.i = c;
        }
      }
      a = block_arg<true, NW>(a, s.red_v, s.red_i, parity);
      if (a.i == kNone) {
        res.status = ST_OPTIMAL;
        res.value = round_to_precision(M[0], precision);
        break;
      }
      col = a.i;
      // leaving row: ratio test with the reference's early break (:83-95) == lowest r whose ratio is
      // <= precision if any, else first index of the minimum ratio.  Ratios <= precision get key -inf.
      Arg b = {INF, kNone};
      for (int r = 1 + tid; r < H; r += NT) {
        const double v = M[(size_t)r * ld + col];
        if (v > precision) {
          const double ratio = __ddiv_rn(M[(size_t)r * ld], v);
          if (ratio < INF) {  // +inf and NaN never win (`ratio < minRatio` with minRatio = Infinity)
            const double key = (ratio <= precision) ? -INF : ratio;
            if (b.i == kNone || key < b.v) {
              b.v = key;
              b.i = r;
            }
          }
        }
      }
      b = block_arg<false, NW>(b, s.red_v, s.red_i, parity);
      if (b.i == kNone) {
        res.status = ST_UNBOUNDED;
        res.value = (double)col;
        break;
      }
      row = b.i;
    }

    if (check_cycles) {  // (:98, :137)
      if (hist_len >= s.hist_cap) {
        res.status = ST_ERR_HISTORY;
        break;
      }
      if (tid == 0) {
        s.hist[2 * hist_len] = t.var[W + row];
        s.hist[2 * hist_len + 1] = t.var[col];
      }
      hist_len++;
      __syncthreads();
      if (history_has_cycle<NT>(s.hist, hist_len)) break;  // "cycled", NaN
    }

    pivot_cta<NW, KC>(t, s, row, col);
    if (phase == 1)
      res.p1++;
    else
      res.p2++;
    iter++;
  }
  return res;
}

}  // namespace yalps
