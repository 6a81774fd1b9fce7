// reg_kernel.cuh -- K1r: one small LP per warp with the whole tableau in REGISTERS.
//
// K1 is bound by the shared-memory data pipe (every cell is read and written through it once per pivot).
// For tableaus of at most HR rows and 65 columns the cells fit in the register file instead: lane l owns the two
// coefficient columns 2l+1, 2l+2 of EVERY row (2*HR doubles), lanes 0..HR-2 additionally own one RHS cell each,
// and M[0,0] is kept redundantly by all lanes.  Per pivot only the pivot column goes through shared memory (one
// lane writes it, everybody reads it back as broadcast loads); the rank-1 update is pure register arithmetic.
// Same algorithm, selection rules and rounding sequence as simplex_device.cuh (src/simplex.ts:5-142).
//
// Register arrays must be indexed statically, so every row loop is fully unrolled and the two places that need
// a dynamic row (reading / writing the pivot row) use a warp-uniform switch.
//
// STATUS: experimental, selected only with YALPS_PATH_REG.  Measured on B200 (config 2): 908 warp-instructions and
// 113 shared-memory wavefronts per pivot (K1: 1,051 and 401), but 454 M pivots/s against K1's 575 M: the unrolled
// code is ~64 KB, the per-pivot hot path ~15 KB, and with 12 warps per SM at different program counters the
// instruction caches miss constantly (ncu: 4.2 "no instruction" stall cycles per issued instruction, issue slots
// 32 % busy).  Kept because it is bit-exact (tests/test_gpu_simplex.py) and documents the experiment.
#pragma once

#include "kernels.cuh"

namespace yalps {

template <int HR>
struct RegTab {
  double t[HR][2];
};

// pivot row read / write: `row` is warp-uniform, the switch compiles to a jump table
template <int HR>
__device__ __forceinline__ void reg_get_row(const RegTab<HR> &T, int row, double &x0, double &x1) {
  x0 = x1 = 0.0;
  switch (row) {
#define YALPS_CASE(R)              \
  case R:                          \
    asm volatile("");              \
    if (R < HR) {                  \
      x0 = T.t[R < HR ? R : 0][0]; \
      x1 = T.t[R < HR ? R : 0][1]; \
    }                              \
    break;
    YALPS_CASE(0) YALPS_CASE(1) YALPS_CASE(2) YALPS_CASE(3) YALPS_CASE(4) YALPS_CASE(5) YALPS_CASE(6) YALPS_CASE(7)
    YALPS_CASE(8) YALPS_CASE(9) YALPS_CASE(10) YALPS_CASE(11) YALPS_CASE(12) YALPS_CASE(13) YALPS_CASE(14)
    YALPS_CASE(15) YALPS_CASE(16) YALPS_CASE(17) YALPS_CASE(18) YALPS_CASE(19) YALPS_CASE(20) YALPS_CASE(21)
    YALPS_CASE(22) YALPS_CASE(23) YALPS_CASE(24) YALPS_CASE(25) YALPS_CASE(26) YALPS_CASE(27) YALPS_CASE(28)
    YALPS_CASE(29) YALPS_CASE(30) YALPS_CASE(31) YALPS_CASE(32)
#undef YALPS_CASE
    default:
      break;
  }
}

template <int HR>
__device__ __forceinline__ void reg_set_row(RegTab<HR> &T, int row, double x0, double x1) {
  switch (row) {
#define YALPS_CASE(R)              \
  case R:                          \
    asm volatile("");              \
    if (R < HR) {                  \
      T.t[R < HR ? R : 0][0] = x0; \
      T.t[R < HR ? R : 0][1] = x1; \
    }                              \
    break;
    YALPS_CASE(0) YALPS_CASE(1) YALPS_CASE(2) YALPS_CASE(3) YALPS_CASE(4) YALPS_CASE(5) YALPS_CASE(6) YALPS_CASE(7)
    YALPS_CASE(8) YALPS_CASE(9) YALPS_CASE(10) YALPS_CASE(11) YALPS_CASE(12) YALPS_CASE(13) YALPS_CASE(14)
    YALPS_CASE(15) YALPS_CASE(16) YALPS_CASE(17) YALPS_CASE(18) YALPS_CASE(19) YALPS_CASE(20) YALPS_CASE(21)
    YALPS_CASE(22) YALPS_CASE(23) YALPS_CASE(24) YALPS_CASE(25) YALPS_CASE(26) YALPS_CASE(27) YALPS_CASE(28)
    YALPS_CASE(29) YALPS_CASE(30) YALPS_CASE(31) YALPS_CASE(32)
#undef YALPS_CASE
    default:
      break;
  }
}

// Rank-1 update of rows 1..H-1 in registers.  cb[r] = {pivot-column coefficient or 0 (row left alone), -coef/q};
// own0 / own1: my first / second column is the pivot column, its cell becomes -coef/q.
// kDense (H == HR and every lane rewrites both of its cells): the owner lane's pivot-column cell is computed like
// any other and then replaced, so a block needs no per-lane store predicates.  Rows go in blocks of 4 with the
// four broadcast loads issued first; a block whose rows are all active is one straight-line sequence.
// The code is unrolled over rows (register arrays need static indices), so there is exactly ONE copy of the hot
// variant: instruction-cache footprint, not instruction count, is what limits this kernel.
template <int HR, bool kDense>
__device__ __forceinline__ void reg_update(RegTab<HR> &T, const double2 *cb, int H, double pn0, double pn1, bool st0,
                                           bool st1, bool own0, bool own1) {
  auto one_row = [&](int r, const double2 cc) {
    if (kDense) {
      const double n0 = __dsub_rn(T.t[r][0], __dmul_rn(cc.x, pn0));
      const double n1 = __dsub_rn(T.t[r][1], __dmul_rn(cc.x, pn1));
      T.t[r][0] = own0 ? cc.y : n0;
      T.t[r][1] = own1 ? cc.y : n1;
    } else {
      if (st0) T.t[r][0] = __dsub_rn(T.t[r][0], __dmul_rn(cc.x, pn0));
      if (st1) T.t[r][1] = __dsub_rn(T.t[r][1], __dmul_rn(cc.x, pn1));
      if (own0) T.t[r][0] = cc.y;
      if (own1) T.t[r][1] = cc.y;
    }
  };
#pragma unroll
  for (int r0 = 1; r0 < HR; r0 += 4) {
    if (kDense || r0 < H) {
      double2 cc[4];
#pragma unroll
      for (int i = 0; i < 4; i++)
        cc[i] = (r0 + i < HR && (kDense || r0 + i < H)) ? cb[r0 + i < HR ? r0 + i : 0] : make_double2(0.0, 0.0);
      if (kDense && cc[0].x != 0.0 && cc[1].x != 0.0 && cc[2].x != 0.0 && cc[3].x != 0.0) {
#pragma unroll
        for (int i = 0; i < 4; i++)
          if (r0 + i < HR) one_row(r0 + i < HR ? r0 + i : 0, cc[i]);
      } else {
#pragma unroll
        for (int i = 0; i < 4; i++)
          if (r0 + i < HR && cc[i].x != 0.0) one_row(r0 + i < HR ? r0 + i : 0, cc[i]);
      }
    }
  }
}

constexpr int kRegMaxRows = 33;  // rows 1..32 map to lanes 0..31 for the RHS column
constexpr int kRegMaxCols = 65;  // 64 coefficient columns = 32 lanes x 2

// shared memory per CTA (= one warp): cb[HR] {coef or raw column value, -coef/q}, var[W+H]
template <int HR>
struct RegSmem {
  static constexpr size_t off_cb = 0;
  static constexpr size_t off_var = (size_t)HR * 16;
  static constexpr size_t total = off_var + (size_t)(kRegMaxCols + HR + 3) / 4 * 16;
};

template <int HR>
__global__ void __launch_bounds__(32, 10) k_simplex_reg(const BatchArgs a) {
  __shared__ __align__(16) unsigned char smem_raw[RegSmem<HR>::total];
  double2 *cb = reinterpret_cast<double2 *>(smem_raw + RegSmem<HR>::off_cb);
  int *var = reinterpret_cast<int *>(smem_raw + RegSmem<HR>::off_var);
  __shared__ long long s_lp;
  const int lane = threadIdx.x;
  const double precision = a.precision, INF = d_inf();

  for (long long static_lp = blockIdx.x;; static_lp += gridDim.x) {
    long long lp = static_lp;
    if (a.counter) {
      if (lane == 0) s_lp = (long long)atomicAdd(a.counter, 1ULL);
      __syncwarp();
      lp = s_lp;
      __syncwarp();
    }
    if (lp >= a.n) break;
    if (a.index) lp = a.index[lp];
    int H, W;
    size_t moff, roff, poff;
    if (a.heights) {
      H = a.heights[lp];
      W = a.widths[lp];
      moff = (size_t)a.mat_off[lp];
      roff = (size_t)a.rhs_off[lp];
      poff = (size_t)a.pos_off[lp];
    } else {
      H = a.H;
      W = a.W;
      moff = (size_t)lp * H * W;
      roff = (size_t)lp * H;
      poff = (size_t)lp * (W + H);
    }
    const int Wm1 = W - 1;
    const int j0 = 2 * lane;
    const bool v0 = j0 < Wm1, v1 = j0 + 1 < Wm1;  // my two columns exist
    const bool has_b = lane + 1 < H;               // I own the RHS cell of row lane+1

    // ---- load: global -> registers (coalesced 16 bytes per lane per row)
    RegTab<HR> T;
    const double *src = a.in + moff;
    {
      const double *rp = src + 1 + j0;
#pragma unroll
      for (int r = 0; r < HR; r++) {
        T.t[r][0] = (r < H && v0) ? rp[0] : 0.0;
        T.t[r][1] = (r < H && v1) ? rp[1] : 0.0;
        rp += W;
      }
    }
    double bv = has_b ? src[(size_t)(lane + 1) * W] : 0.0;
    double b0 = src[0];
    for (int k = lane; k < W + H; k += 32) var[k] = k;
    __syncwarp();

    int status = ST_CYCLED;
    double value = d_nan();
    long long p1 = 0, p2 = 0, iter = 0;
    int phase = 1;

    for (;;) {
      if (!((double)iter < a.max_pivots)) break;
      int row, col;
      double pr0 = 0.0, pr1 = 0.0;  // old pivot row cells of my two columns
      if (phase == 1) {
        // leaving row: first index of the most negative RHS below -precision (:111-119)
        const bool cand = has_b && bv < -precision;
        row = warp_best<false>(cand ? (unsigned)(order_key(bv) >> 32) : 0xffffffffu,
                               cand ? (unsigned)order_key(bv) : 0xffffffffu, cand ? lane + 1 : kNone)
                  .idx;
        if (row == kNone) {
          phase = 2;
          iter = 0;
          continue;
        }
        // entering column: first index of max -M[0,c]/M[row,c] over M[row,c] < -precision (:123-134)
        reg_get_row<HR>(T, row, pr0, pr1);
        double best = -INF;
        int bi = kNone;
        if (v0 && pr0 < -precision) {
          const double ratio = __ddiv_rn(-T.t[0][0], pr0);
          if (ratio > best) {
            best = ratio;
            bi = j0 + 1;
          }
        }
        if (v1 && pr1 < -precision) {
          const double ratio = __ddiv_rn(-T.t[0][1], pr1);
          if (ratio > best) {
            best = ratio;
            bi = j0 + 2;
          }
        }
        const unsigned long long key = bi == kNone ? no_key<true>() : order_key(best);
        col = warp_best<true>((unsigned)(key >> 32), (unsigned)key, bi).idx;
        if (col == kNone) {
          status = ST_INFEASIBLE;
          break;
        }
      } else {
        // entering column: first index of the largest reduced cost above precision (:71-79)
        double best = -INF;
        int bi = kNone;
        if (v0 && T.t[0][0] > precision) {
          best = T.t[0][0];
          bi = j0 + 1;
        }
        if (v1 && T.t[0][1] > precision && T.t[0][1] > best) {
          best = T.t[0][1];
          bi = j0 + 2;
        }
        const unsigned long long key = bi == kNone ? no_key<true>() : order_key(best);
        col = warp_best<true>((unsigned)(key >> 32), (unsigned)key, bi).idx;
        if (col == kNone) {
          status = ST_OPTIMAL;
          value = round_to_precision(b0, precision);
          break;
        }
      }
      const int jc = col - 1, lc = jc >> 1, ec = jc & 1;

      // ---- the pivot column goes through shared memory: its owner lane writes it, everybody reads it back
      if (lane == lc) {
#pragma unroll
        for (int r = 0; r < HR; r++)
          if (r < H) cb[r].x = ec ? T.t[r][1] : T.t[r][0];
      }
      __syncwarp();
      const double cmine = has_b ? cb[lane + 1].x : 0.0;  // pivot-column cell of my RHS row

      if (phase == 2) {
        // leaving row: ratio test with the reference's early break (:83-95)
        bool cand = false;
        double keyv = INF;
        if (has_b && cmine > precision) {
          const double ratio = __ddiv_rn(bv, cmine);
          if (ratio < INF) {
            cand = true;
            keyv = (ratio <= precision) ? -INF : ratio;
          }
        }
        const unsigned long long key = cand ? order_key(keyv) : no_key<false>();
        row = warp_best<false>((unsigned)(key >> 32), (unsigned)key, cand ? lane + 1 : kNone).idx;
        if (row == kNone) {
          status = ST_UNBOUNDED;
          value = (double)col;
          break;
        }
        reg_get_row<HR>(T, row, pr0, pr1);
      }

      // ---- pivot(row, col) (:5-39)
      const double q = cb[row].x;
      const double c0raw = cb[0].x;
      __syncwarp();  // every lane has read the raw column before its cells are replaced by {coef, -coef/q}
      // normalised pivot row cells of my columns; the pivot cell itself becomes 1/q
      const double x0 = (j0 == jc) ? 1.0 : pr0, x1 = (j0 + 1 == jc) ? 1.0 : pr1;
      const bool n0 = v0 && fabs(x0) > kTiny, n1 = v1 && fabs(x1) > kTiny;
      const double pn0 = n0 ? __ddiv_rn(x0, q) : 0.0, pn1 = n1 ? __ddiv_rn(x1, q) : 0.0;
      const bool st0 = n0 && j0 != jc, st1 = n1 && j0 + 1 != jc;  // cells the rank-1 pass rewrites
      // rows 1..H-1: one lane each (the lane of the pivot row normalises the RHS cell instead)
      const bool is_prow = has_b && lane + 1 == row;
      const double num = is_prow ? bv : -cmine;
      const bool nzq = has_b && fabs(num) > kTiny;
      const double quo = nzq ? __ddiv_rn(num, q) : 0.0;
      const double coef_mine = (nzq && !is_prow) ? cmine : 0.0;  // 0 = my row is skipped (:31) or is the pivot row
      if (has_b) cb[lane + 1] = make_double2(coef_mine, quo);
      // row 0 (objective row): every lane redundantly
      const bool act0 = fabs(c0raw) > kTiny;
      const double cn0 = act0 ? __ddiv_rn(-c0raw, q) : 0.0;
      const double p0 = __shfl_sync(0xffffffffu, quo, row - 1);        // normalised RHS of the pivot row
      const bool nz0 = __shfl_sync(0xffffffffu, nzq ? 1 : 0, row - 1);  // ... was above 1e-16
      if (lane == 0) {  // basis bookkeeping (:7-12)
        const int leaving = var[W + row];
        var[W + row] = var[col];
        var[col] = leaving;
      }
      __syncwarp();

      // rank-1 update in registers; rows whose coefficient is 0 (skipped rows, the pivot row) are left alone
      const bool fix0 = lane == lc && ec == 0, fix1 = lane == lc && ec == 1;
      if (act0) {  // row 0
        if (st0) T.t[0][0] = __dsub_rn(T.t[0][0], __dmul_rn(c0raw, pn0));
        if (st1) T.t[0][1] = __dsub_rn(T.t[0][1], __dmul_rn(c0raw, pn1));
        if (fix0) T.t[0][0] = cn0;
        if (fix1) T.t[0][1] = cn0;
        if (nz0) b0 = __dsub_rn(b0, __dmul_rn(c0raw, p0));
      }
      {
        // dense pivot row: every lane rewrites both cells (the owner lane's pivot-column cell included, it is
        // replaced by -coef/q right away), so no per-lane store predicates are needed
        const bool w0 = st0 || (lane == lc && ec == 0), w1 = st1 || (lane == lc && ec == 1);
        const bool dense = __all_sync(0xffffffffu, w0 && w1);
        const bool own0 = lane == lc && ec == 0, own1 = lane == lc && ec == 1;
        if (H == HR && dense)
          reg_update<HR, true>(T, cb, H, pn0, pn1, st0, st1, own0, own1);
        else
          reg_update<HR, false>(T, cb, H, pn0, pn1, st0, st1, own0, own1);
      }
      reg_set_row<HR>(T, row, pn0, pn1);
      if (is_prow)
        bv = quo;
      else if (coef_mine != 0.0 && nz0)
        bv = __dsub_rn(bv, __dmul_rn(coef_mine, p0));
      __syncwarp();

      if (phase == 1)
        p1++;
      else
        p2++;
      iter++;
    }

    // ---- outputs
    if (lane == 0) {
      if (a.status) a.status[lp] = status;
      if (a.value) a.value[lp] = value;
      if (a.pivots) {
        a.pivots[2 * lp] = p1;
        a.pivots[2 * lp + 1] = p2;
      }
      if (a.rhs_out) a.rhs_out[roff] = b0;
    }
    if (a.rhs_out && has_b) a.rhs_out[roff + lane + 1] = bv;
    __syncwarp();
    if (a.pos_out)
      for (int k = lane; k < W + H; k += 32) a.pos_out[poff + var[k]] = k;
    if (a.var_out)
      for (int k = lane; k < W + H; k += 32) a.var_out[poff + k] = var[k];
    if (a.mat_out) {
      double *dst = a.mat_out + moff;
      if (lane == 0) dst[0] = b0;
      if (has_b) dst[(size_t)(lane + 1) * W] = bv;
#pragma unroll
      for (int r = 0; r < HR; r++) {
        if (r < H) {
          if (v0) dst[(size_t)r * W + 1 + j0] = T.t[r][0];
          if (v1) dst[(size_t)r * W + 2 + j0] = T.t[r][1];
        }
      }
    }
    __syncwarp();
  }
}

}  // namespace yalps
