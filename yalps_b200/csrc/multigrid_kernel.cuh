// multigrid_kernel.cuh -- K4m: ONE large LP spread over SEVERAL GPUs, one cooperative persistent
// kernel per GPU, the GPUs talking to one another from INSIDE those kernels through peer memory (NVLink / NVSwitch
// loads and stores, no host in the loop, no NCCL call per pivot).
//
// Same algorithm and rounding sequence as grid_kernel.cuh (src/simplex.ts:5-142); what changes is where the rows live:
//   * global row r belongs to rank r % G and is its local row r / G (round robin, so that the row skip of sparse
//     pivots costs every rank the same); a rank holds only its rows: G GPUs hold a tableau G times the size of one HBM;
//   * every CTA of every rank keeps the private shared-memory copies of the objective row and of the RHS column that
//     K4 keeps, and makes the pivot choice redundantly from them -- the ranks agree on (row, col) without a
//     candidate exchange, because they all look at the same bits;
//   * per pivot two things cross the fabric, both pushed by their owners with stores into the receivers' exchange
//     buffers:
//       - the pivot column: every rank sends its H/G cells to every rank (an all-gather of H*8 bytes);
//       - the raw pivot row: its owner sends W*8 bytes to every other rank (a broadcast), in `row_parts` slices
//         pushed by different CTAs;
//     every cell travels as ONE 16-byte store {low word, sequence, high word, sequence} (the flag-in-data protocol of
//     NCCL's LL: each 8-byte half is valid the moment its sequence number matches), so there is no fence and no
//     separate flag on the way: a cell costs one one-way trip over NVLink; receivers poll the slots in their OWN memory
//     with volatile 16-byte loads (first version: data, fence.sys, release flag, acquire spin = 8 us per exchange on
//     two B200s; this one: see profiles/);
//   * exchange buffers are double-buffered by pivot parity: a rank can only be one exchange ahead of the slowest one
//     (it needs that rank's share of the next pivot column), so a buffer is never rewritten while someone reads it;
//   * the update of the local rows and the two local grid barriers per pivot are K4's.
// Every spin has a cycle budget: a rank whose peer never shows up (failed launch, kernels that could not be
// co-resident) ends with ST_ERR_PEER instead of hanging the GPU.
#pragma once

#include "grid_kernel.cuh"

namespace yalps {

constexpr int kMaxGridRanks = 16;
constexpr int kMaxRowParts = 8;

struct MultiGridArgs {
  double *M;  // local rows, Hl x W (global row g + G * lr at local row lr), updated in place
  int H, W;   // global shape
  int G, g;   // ranks, this rank
  int Hl, Hlmax, Wpad;
  int row_parts;
  uint4 *colx[kMaxGridRanks];  // per rank: [2][G][Hlmax] slots, pivot-column shares indexed by sender
  uint4 *rowx[kMaxGridRanks];  // per rank: [2][Wpad] slots, raw pivot row
  int *var;
  int *pos_out;
  double *rhs_out;
  int *status;
  double *value;
  long long *pivots;
  double precision, max_pivots;
  int check_cycles;
  int *hist;
  int hist_cap;
  int *flags;
  unsigned long long *barrier;
  unsigned long long *rows_out;
  long long spin_limit;  // cycles a spin may last
  long long *giveup;     // [8] where the first CTA of this rank that gave up was: wait id, sequence, phase, CTA, row, col
};

// One cell of an exchange buffer: {low word, sequence, high word, sequence}; 8-byte halves are written atomically.
__device__ __forceinline__ void ll_store(uint4 *slot, double v, unsigned seq) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(slot), "r"((unsigned)__double2loint(v)), "r"(seq),
               "r"((unsigned)__double2hiint(v)), "r"(seq)
               : "memory");
}
// Polls a slot of this rank's own buffer until both halves carry `seq`; false when the cycle budget ran out.
__device__ __forceinline__ bool ll_load(const uint4 *slot, unsigned seq, double &v, long long limit) {
  unsigned lo, f1, hi, f2;
  long long t0 = 0;
  for (;;) {
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(lo), "=r"(f1), "=r"(hi), "=r"(f2) : "l"(slot) : "memory");
    if (f1 == seq && f2 == seq) break;
    if (t0 == 0)
      t0 = clock64();
    else if (clock64() - t0 > limit)
      return false;
  }
  v = __hiloint2double((int)hi, (int)lo);
  return true;
}

// K4's grid barrier with a budget; *dead (shared memory) is set when it runs out.  Returns with the CTA synchronised.
__device__ __forceinline__ void grid_barrier_budget(unsigned long long *counter, unsigned long long &epoch, int *dead,
                                                    long long limit) {
  __syncthreads();
  epoch++;
  if (threadIdx.x == 0) {
    const unsigned long long goal = epoch * gridDim.x;
    __threadfence();
    asm volatile("red.release.gpu.global.add.u64 [%0], 1;" ::"l"(counter) : "memory");
    unsigned long long seen;
    const long long t0 = clock64();
    for (;;) {
      asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(counter) : "memory");
      if (seen >= goal) break;
      if (clock64() - t0 > limit) {
        *dead = 400;
        break;
      }
    }
    __threadfence();
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kGridThreads, 1) k_simplex_grid_multi(const MultiGridArgs a) {
  constexpr int NT = kGridThreads, NW = kGridWarps;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int dead;
  unsigned long long epoch = 0;
  const GridSmem L(a.H, a.W);
  double *prow = reinterpret_cast<double *>(smem_raw + L.off_prow);
  double *row0 = reinterpret_cast<double *>(smem_raw + L.off_row0);
  double *colbuf = reinterpret_cast<double *>(smem_raw + L.off_col);
  double *bcol = reinterpret_cast<double *>(smem_raw + L.off_b);
  unsigned *nzmask = reinterpret_cast<unsigned *>(smem_raw + L.off_nz);
  unsigned *red = reinterpret_cast<unsigned *>(smem_raw + L.off_red);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int H = a.H, W = a.W, G = a.G, g = a.g, Hl = a.Hl, Hlmax = a.Hlmax;
  double *__restrict__ M = a.M;
  const double precision = a.precision, INF = d_inf();
  const int gwarp = blockIdx.x * NW + warp, gwarps = gridDim.x * NW;
  const long long limit = a.spin_limit;
  const int RP = a.row_parts;
  if (tid == 0) dead = 0;

  for (int k = blockIdx.x * NT + tid; k < W + H; k += gridDim.x * NT) a.var[k] = k;
  __syncthreads();

  // all-gather of global column `c` (sequence number seq, parity par) into colbuf[0..H): every rank's CTAs d < G push
  // the rank's share to rank d; every CTA then waits for the G shares addressed to its rank.  False: a peer is missing.
  auto gather_column = [&](int c, unsigned long long seq, int par) -> bool {
    if ((int)blockIdx.x < G) {
      const int dst = blockIdx.x;
      uint4 *out = a.colx[dst] + ((size_t)par * G + g) * Hlmax;
      for (int lr = tid; lr < Hl; lr += NT) ll_store(out + lr, M[(size_t)lr * W + c], (unsigned)seq);
    }
    const uint4 *in = a.colx[g] + (size_t)par * G * Hlmax;
    bool ok = true;
    for (int idx = tid; idx < G * Hlmax; idx += NT) {
      const int src = idx / Hlmax, lr = idx - src * Hlmax;
      const int r = lr * G + src;
      double v;
      if (r < H) {
        if (ok && ll_load(in + idx, (unsigned)seq, v, limit))
          colbuf[r] = v;
        else
          ok = false;
      }
    }
    if (!ok) dead = 100;
    __syncthreads();
    return !dead;
  };
  // the owner's sender CTAs push the raw pivot row to the other ranks
  auto push_row = [&](int lrow, unsigned long long seq, int par) {
    if ((int)blockIdx.x < G * RP) {
      const int dst = blockIdx.x / RP, part = blockIdx.x - dst * RP;
      if (dst != g) {
        const int per = (((W + RP - 1) / RP) + 1) & ~1;
        const int c0 = part * per, c1 = min(W, c0 + per);
        uint4 *out = a.rowx[dst] + (size_t)par * a.Wpad;
        const double *src = M + (size_t)lrow * W;
        for (int c = c0 + tid; c < c1; c += NT) ll_store(out + c, src[c], (unsigned)seq);
      }
    }
  };
  // private copies of the objective row (rank 0 owns row 0: a row broadcast) and of the RHS column (a column gather)
  {
    // (sequence number 1 -> parity 1, like every later exchange: buffer = parity of the sequence number)
    if (g == 0) push_row(0, 1, 1);
    if (g == 0) {
      for (int c = tid; c < W; c += NT) row0[c] = M[c];
    } else {
      const uint4 *in = a.rowx[g] + a.Wpad;
      bool ok = true;
      for (int c = tid; c < W; c += NT) {
        double v;
        if (ok && ll_load(in + c, 1u, v, limit))
          row0[c] = v;
        else
          ok = false;
      }
      if (!ok) dead = 200;
    }
    __syncthreads();
    if (!dead && gather_column(0, 1, 1))
      for (int r = tid; r < H; r += NT) bcol[r] = colbuf[r];
  }
  grid_barrier_budget(a.barrier, epoch, &dead, limit);

  int status = ST_CYCLED;
  double value = d_nan();
  long long p1 = 0, p2 = 0, iter = 0;
  unsigned long long seq = 1;  // exchanges done so far (the initial ones used sequence number 1)
  unsigned long long rows_mine = 0;
  int phase = 1, parity = 0, hist_len = 0;
  bool have_carry = false;
  double carry_cv = -INF, carry_rv = INF;
  int carry_ci = kNone, carry_ri = kNone;

  for (;;) {
    if (dead) {
      status = ST_ERR_PEER;
      if (tid == 0 && atomicCAS((unsigned long long *)a.giveup, 0ULL, (unsigned long long)dead) == 0ULL) {
        a.giveup[1] = (long long)seq;
        a.giveup[2] = phase;
        a.giveup[3] = blockIdx.x;
        a.giveup[4] = p1;
        a.giveup[5] = p2;
      }
      break;
    }
    if (!((double)iter < a.max_pivots)) break;
    int row, col;
    const unsigned long long s = seq + 1;  // this pivot's sequence number; exchange buffers by its parity
    const int par = (int)(s & 1);
    // the pivot row, raw, into prow (from local memory on its owner, from the exchange buffer elsewhere); `scan`
    // sees every (c, coef) in ascending c per thread
    auto fetch_row = [&](int r, auto &&scan) -> bool {
      const int owner = r % G, lrow = r / G;
      if (owner == g) {
        push_row(lrow, s, par);
        for (int c = tid; c < W; c += NT) {
          const double coef = M[(size_t)lrow * W + c];
          prow[c] = coef;
          scan(c, coef);
        }
      } else {
        const uint4 *in = a.rowx[g] + (size_t)par * a.Wpad;
        bool ok = true;
        for (int c = tid; c < W; c += NT) {
          double coef;
          if (ok && ll_load(in + c, (unsigned)s, coef, limit)) {
            prow[c] = coef;
            scan(c, coef);
          } else {
            ok = false;
          }
        }
        if (!ok) dead = 300;
        __syncthreads();
        if (dead) return false;
      }
      return true;
    };
    if (phase == 1) {
      // leaving row from the private RHS copy (:111-119)
      double bv = INF;
      int bi = kNone;
      if (have_carry) {
        bv = carry_rv;
        bi = carry_ri;
      } else {
        for (int r = 1 + tid; r < H; r += NT) {
          const double v = bcol[r];
          if (v < -precision && v < bv) {
            bv = v;
            bi = r;
          }
        }
      }
      row = block_best<false, NW>(bi == kNone ? no_key<false>() : order_key(bv), bi, red, parity);
      if (row == kNone) {
        phase = 2;
        iter = 0;
        hist_len = 0;
        continue;
      }
      // exchange 1: the pivot row; entering column from it and the private objective row (:123-134)
      bv = -INF;
      bi = kNone;
      if (!fetch_row(row, [&](int c, double coef) {
            if (c >= 1 && coef < -precision) {
              const double ratio = __ddiv_rn(-row0[c], coef);
              if (ratio > bv) {
                bv = ratio;
                bi = c;
              }
            }
          }))
        continue;
      col = block_best<true, NW>(bi == kNone ? no_key<true>() : order_key(bv), bi, red, parity);
      if (col == kNone) {
        status = ST_INFEASIBLE;
        break;
      }
      // exchange 2: the pivot column
      if (!gather_column(col, s, par)) continue;
    } else {
      // entering column from the private objective row (:71-79)
      double bv = -INF;
      int bi = kNone;
      if (have_carry) {
        bv = carry_cv;
        bi = carry_ci;
      } else {
        for (int c = 1 + tid; c < W; c += NT) {
          const double v = row0[c];
          if (v > precision && v > bv) {
            bv = v;
            bi = c;
          }
        }
      }
      col = block_best<true, NW>(bi == kNone ? no_key<true>() : order_key(bv), bi, red, parity);
      if (col == kNone) {
        status = ST_OPTIMAL;
        value = round_to_precision(row0[0], precision);
        break;
      }
      // exchange 1: the pivot column; ratio test against the private RHS copy (:83-95)
      if (!gather_column(col, s, par)) continue;
      bv = INF;
      bi = kNone;
      for (int r = 1 + tid; r < H; r += NT) {
        const double v = colbuf[r];
        if (v > precision) {
          const double ratio = __ddiv_rn(bcol[r], v);
          if (ratio < INF) {
            const double key = (ratio <= precision) ? -INF : ratio;
            if (bi == kNone || key < bv) {
              bv = key;
              bi = r;
            }
          }
        }
      }
      row = block_best<false, NW>(bi == kNone ? no_key<false>() : order_key(bv), bi, red, parity);
      if (row == kNone) {
        status = ST_UNBOUNDED;
        value = (double)col;
        break;
      }
      // exchange 2: the pivot row
      if (!fetch_row(row, [](int, double) {})) continue;
    }
    seq = s;
    __syncthreads();  // raw pivot row and column are complete in shared memory

    if (a.check_cycles) {  // CTA 0 of every rank keeps its own history (same inputs, same verdict)
      if (blockIdx.x == 0) {
        int verdict = 0;
        if (hist_len >= a.hist_cap) {
          verdict = 2;
        } else {
          if (tid == 0) {
            a.hist[2 * hist_len] = a.var[W + row];
            a.hist[2 * hist_len + 1] = a.var[col];
          }
          __syncthreads();
          if (history_has_cycle<NT>(a.hist, hist_len + 1)) verdict = 1;
        }
        if (tid == 0) a.flags[iter & 1] = verdict;
        __threadfence();
      }
      hist_len++;
      grid_barrier_budget(a.barrier, epoch, &dead, limit);
      if (dead) continue;
      const int verdict = *reinterpret_cast<volatile int *>(a.flags + (iter & 1));
      if (verdict == 2) status = ST_ERR_HISTORY;
      if (verdict) break;
    }

    // ---- the fused stage of K4 (grid_kernel.cuh): normalised pivot row + non-zero flags, private objective row and
    // RHS column, "0 = row left alone" pivot column, candidates of the next pivot's first selection
    const double q = prow[col];
    const double coef0 = colbuf[0];
    const double braw = prow[0];
    __syncthreads();
    {
      const bool act0 = fabs(coef0) > kTiny;
      const bool nz0 = fabs(braw) > kTiny;
      const double p0 = nz0 ? __ddiv_rn(braw, q) : 0.0;
      carry_cv = -INF;
      carry_ci = kNone;
      for (int cbase = warp * 32; cbase < W; cbase += NT) {
        const int c = cbase + lane;
        bool nz = false;
        if (c < W) {
          const double v = (c == col) ? 1.0 : prow[c];
          nz = fabs(v) > kTiny;
          const double pn = nz ? __ddiv_rn(v, q) : 0.0;
          prow[c] = pn;
          if (c == col) nz = false;
          double o = row0[c];
          if (act0) {
            if (nz)
              o = __dsub_rn(o, __dmul_rn(coef0, pn));
            else if (c == col)
              o = __ddiv_rn(-coef0, q);
            row0[c] = o;
          }
          if (c == 0) bcol[0] = o;
          if (c >= 1 && o > precision && o > carry_cv) {
            carry_cv = o;
            carry_ci = c;
          }
        }
        const unsigned m = __ballot_sync(0xffffffffu, nz);
        if (lane == 0) nzmask[cbase >> 5] = m;
      }
      carry_rv = INF;
      carry_ri = kNone;
      for (int r = tid; r < H; r += NT) {
        const double coef = colbuf[r];
        const bool on = r != row && fabs(coef) > kTiny;
        colbuf[r] = on ? coef : 0.0;
        rows_mine += on;
        if (r >= 1) {
          double bnew = bcol[r];
          if (r == row) {
            bnew = p0;
            bcol[r] = bnew;
          } else if (on && nz0) {
            bnew = __dsub_rn(bnew, __dmul_rn(coef, p0));
            bcol[r] = bnew;
          }
          if (bnew < -precision && bnew < carry_rv) {
            carry_rv = bnew;
            carry_ri = r;
          }
        }
      }
      have_carry = true;
    }
    __syncthreads();
    if (blockIdx.x == 0 && tid == 0) {  // basis bookkeeping (:7-12), every rank its own copy
      const int leaving = a.var[W + row];
      a.var[W + row] = a.var[col];
      a.var[col] = leaving;
    }
    // the rank's sender CTAs have read the local rows of this pivot: the local rows may change now
    grid_barrier_budget(a.barrier, epoch, &dead, limit);
    if (dead) continue;

    // ---- rank-1 update (:28-38) of the LOCAL rows: (local row, column segment) items, round-robin over the warps
    {
      const int nseg = (W + kSegCols - 1) / kSegCols;
      const int step_r = gwarps / nseg, step_s = gwarps - step_r * nseg;
      int lr = gwarp / nseg, seg = gwarp - lr * nseg;
      while (lr < Hl) {
        const int r = lr * G + g;
        const int c0 = seg * kSegCols + lane;
        double *__restrict__ Mr = M + (size_t)lr * W + c0;
        const double coef = colbuf[r];
        const bool full_seg = (seg + 1) * kSegCols <= W;
        if (r == row) {
#pragma unroll
          for (int u = 0; u < 8; u++)
            if (full_seg || c0 + 32 * u < W) Mr[32 * u] = prow[c0 + 32 * u];
        } else if (coef != 0.0) {
          const uint4 w0 = *reinterpret_cast<const uint4 *>(nzmask + seg * 8);
          const uint4 w1 = *reinterpret_cast<const uint4 *>(nzmask + seg * 8 + 4);
          const unsigned bits = ((w0.x >> lane) & 1u) | (((w0.y >> lane) & 1u) << 1) | (((w0.z >> lane) & 1u) << 2) |
                                (((w0.w >> lane) & 1u) << 3) | (((w1.x >> lane) & 1u) << 4) |
                                (((w1.y >> lane) & 1u) << 5) | (((w1.z >> lane) & 1u) << 6) |
                                (((w1.w >> lane) & 1u) << 7);
          double x[8];
          if (full_seg) {
#pragma unroll
            for (int u = 0; u < 8; u++) x[u] = Mr[32 * u];
#pragma unroll
            for (int u = 0; u < 8; u++)
              if ((bits >> u) & 1u) Mr[32 * u] = __dsub_rn(x[u], __dmul_rn(coef, prow[c0 + 32 * u]));
          } else {
#pragma unroll
            for (int u = 0; u < 8; u++)
              if (c0 + 32 * u < W) x[u] = Mr[32 * u];
#pragma unroll
            for (int u = 0; u < 8; u++)
              if (c0 + 32 * u < W && ((bits >> u) & 1u)) Mr[32 * u] = __dsub_rn(x[u], __dmul_rn(coef, prow[c0 + 32 * u]));
          }
          if (col >= seg * kSegCols && col < (seg + 1) * kSegCols && ((col - lane) & 31) == 0)
            M[(size_t)lr * W + col] = __ddiv_rn(-coef, q);
        }
        lr += step_r;
        seg += step_s;
        if (seg >= nseg) {
          seg -= nseg;
          lr++;
        }
      }
    }
    grid_barrier_budget(a.barrier, epoch, &dead, limit);

    if (phase == 1)
      p1++;
    else
      p2++;
    iter++;
  }

  // ---- outputs: every rank reports its verdict (the host checks that they agree); rank 0's are the caller's
  if (blockIdx.x == 0 && tid == 0) {
    if (a.status) a.status[0] = status;
    if (a.value) a.value[0] = value;
    if (a.pivots) {
      a.pivots[0] = p1;
      a.pivots[1] = p2;
    }
  }
  if (a.rows_out && blockIdx.x == 0 && rows_mine != 0) atomicAdd(a.rows_out, rows_mine);
  if (a.rhs_out && blockIdx.x == 0)  // the private RHS copy is the RHS column, bit for bit
    for (int r = tid; r < H; r += NT) a.rhs_out[r] = bcol[r];
  if (a.pos_out)
    for (int k = blockIdx.x * NT + tid; k < W + H; k += gridDim.x * NT) a.pos_out[a.var[k]] = k;
}

}  // namespace yalps
