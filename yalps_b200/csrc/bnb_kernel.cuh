// bnb_kernel.cuh -- KB: the whole branch-and-cut search of src/branchAndCut.ts:89-176 inside ONE persistent launch.
//
// The host wave driver (bnb.inl) pays a launch, a synchronisation and a host replay per wave, and a search is a chain
// of dependent waves (Large Farm MIP: 135 of them, ~95 us each for ~35 us of kernel time).  Here the replay itself
// runs on the device:
//   * CTA 0 is the SCHEDULER: one thread executes the reference's loop statement by statement -- the binary heap of
//     npm heap@0.2.7 (= CPython heapq: push = append + sift towards the root, pop = move the last leaf to the root +
//     sift to a leaf + sift back) in shared memory, pruning (:124), incumbent update (:130-139), branching (:141-156),
//     tolerance exit / maxIterations / timeout (:122,162) and the final status rule (:167-173);
//   * every other CTA is a WORKER: it takes the next created node (nodes are evaluated in creation order, i.e. as soon
//     as the scheduler pushes them on the heap -- every node is root + its cut list, so it is independent of the
//     replay), materialises the node's cut list from its parent's (:141-154), assembles root + cut rows in shared memory
//     (applyCuts, :22-61), runs the row-split simplex of simplex_split.cuh on it, evaluates mostFractionalVar (:64-85)
//     and publishes (status, rounded result, branching variable, its value and fraction, pivots); integer-feasible
//     optimal nodes also publish RHS column and both permutations (the incumbent candidates solution() reads);
//   * scheduler and workers talk through global memory with release/acquire accesses; all CTAs are co-resident
//     (cooperative launch), so the spin-waits are inside one kernel.
// What the reference pops, prunes, accepts and branches on is reproduced exactly (the tests compare node and
// node-pivot counts, final basis and RHS bits with the oracle); only the evaluation order of the node LPs differs.
// Capacity limits (node pool, cut pool, candidate pool, heap, cut rows per node) are reported through `overflow`; the
// host then repeats the search with the wave driver, which has none.
#pragma once

#include "kernels.cuh"

namespace yalps {

constexpr int kBnbMaxCuts = 96;  // cut rows a node may carry in the device-resident search

struct BnbCut {
  double sign;
  double value;
  int var;
  int pad;
};

struct __align__(64) BnbNode {
  // ---- result block, written by the worker, read by the scheduler with four 16-byte loads (one round trip to L2)
  double result;    // rounded objective (optimal) / NaN
  double bval;      // mostFractionalVar: value
  double bfrac;     //                    fraction
  long long pivots;
  int status;
  int bvar;         //                    variable
  int cut_len;
  int cand;         // candidate slot (integer-feasible optimal node) or -1
  int cut_begin;
  int done;         // release-stored last
  int pad0, pad1;
  // ---- written by the scheduler when the branch is created
  double eval;      // parent's rounded result (heap key, :155-156)
  double new_sign, new_value;
  int new_var;
  int parent;       // -1: child of the root
  double pad2[3];
};
static_assert(sizeof(BnbNode) == 128, "BnbNode layout");

struct BnbControl {
  unsigned long long next_ticket;  // workers: atomicAdd
  unsigned long long cut_top;      // cut pool bump pointer
  int created;                     // nodes created so far (release-stored by the scheduler)
  int stop;                        // the scheduler is done
  int cand_top;
  int overflow;                    // 1 nodes, 2 cuts, 4 candidates, 8 heap, 16 cut rows
  // results (scheduler)
  int status, found, best_node, best_height, best_cand, pad;
  double result;
  long long iters, node_pivots, max_cuts, max_heap;
};

struct BnbArgs {
  // root (src/branchAndCut.ts:89: the root-optimal tableau and its permutations)
  const double *root;
  const int *root_pos, *root_var;
  int H, W, Hcap;          // root shape; Hcap = H + the most cut rows a node may carry here
  const int *ints;         // TableauModel.integers
  const int *int_rank;     // [W + H] index into `ints` of an integer variable, -1 otherwise
  int nints;
  double sign, init_result;
  int init_var;
  double init_value;
  // options
  double precision, max_pivots, tolerance, timeout_ms, max_iterations;
  // pools
  BnbControl *ctl;
  BnbNode *nodes;
  int node_cap;
  BnbCut *cuts;
  unsigned long long cut_cap;
  double *cand_rhs;        // [cand_cap][Hcap]
  int *cand_pos, *cand_var;  // [cand_cap][W + Hcap]
  int cand_cap;
  int heap_cap;
  unsigned long long *rows_out;
};

__device__ __forceinline__ int ld_acquire(const int *p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int *p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

template <int NWC, int KC, int NWR>
__global__ void __launch_bounds__(NWC *NWR * 32, 1) k_bnb(const BnbArgs a) {
  constexpr int NW = NWC * NWR, NT = NW * 32, VW = 2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_node;
  const int tid = threadIdx.x;
  BnbControl *ctl = a.ctl;
  const int W = a.W, rootH = a.H;

  if (blockIdx.x == 0) {
    // =========================== scheduler (src/branchAndCut.ts:89-176) ===========================
    if (tid != 0) return;
    // heap of (key, node) pairs in shared memory: the comparator is x[0] - y[0] < 0 on the branch evals (:100)
    double *hkey = reinterpret_cast<double *>(smem_raw);
    int *hid = reinterpret_cast<int *>(smem_raw + (size_t)a.heap_cap * 8);
    int hn = 0;
    auto toward_root = [&](int start, int pos) {
      const double k = hkey[pos];
      const int id = hid[pos];
      while (pos > start) {
        const int parent = (pos - 1) >> 1;
        if (!(k - hkey[parent] < 0)) break;
        hkey[pos] = hkey[parent];
        hid[pos] = hid[parent];
        pos = parent;
      }
      hkey[pos] = k;
      hid[pos] = id;
    };
    auto push = [&](double k, int id) {
      hkey[hn] = k;
      hid[hn] = id;
      hn++;
      toward_root(0, hn - 1);
    };
    auto pop = [&](double *k_out) -> int {
      hn--;
      const double lk = hkey[hn];
      const int lid = hid[hn];
      if (hn == 0) {
        *k_out = lk;
        return lid;
      }
      const double tk = hkey[0];
      const int tidx = hid[0];
      int pos = 0, child = 1;
      while (child < hn) {  // sift the hole to a leaf, smaller child first (heapq._siftup)
        const int right = child + 1;
        if (right < hn && !(hkey[child] - hkey[right] < 0)) child = right;
        hkey[pos] = hkey[child];
        hid[pos] = hid[child];
        pos = child;
        child = 2 * pos + 1;
      }
      hkey[pos] = lk;
      hid[pos] = lid;
      toward_root(0, pos);
      *k_out = tk;
      return tidx;
    };
    int created = 0;
    auto create = [&](double eval, int parent, double sign, int var, double value) -> bool {
      if (created >= a.node_cap) {
        atomicOr(&ctl->overflow, 1);
        return false;
      }
      if (hn >= a.heap_cap) {
        atomicOr(&ctl->overflow, 8);
        return false;
      }
      BnbNode *nd = a.nodes + created;
      nd->eval = eval;
      nd->new_sign = sign;
      nd->new_value = value;
      nd->new_var = var;
      nd->parent = parent;
      nd->done = 0;
      push(eval, created);
      created++;
      return true;
    };
    const unsigned long long t_start = global_ns();
    auto timed_out = [&]() -> bool {
      if (!(a.timeout_ms < 1.0e300)) return false;  // +inf: never
      return (double)(global_ns() - t_start) * 1e-6 >= a.timeout_ms;
    };

    // the root's two children (:101-102)
    bool ok = create(a.init_result, -1, -1.0, a.init_var, ceil(a.init_value));
    ok = ok && create(a.init_result, -1, 1.0, a.init_var, floor(a.init_value));
    st_release(&ctl->created, created);

    const double threshold = a.init_result * (1.0 - a.sign * a.tolerance);  // :114
    bool timedout = timed_out();
    bool found = false;
    double best_eval = d_inf();
    int best_node = -1;
    double iter = 0;
    long long node_pivots = 0, max_cuts = 0, max_heap = 0;
    while (ok && iter < a.max_iterations && hn > 0 && best_eval >= threshold && !timedout) {  // :122
      if (hn > max_heap) max_heap = hn;
      double ev;
      const int br = pop(&ev);
      if (ev > best_eval) break;  // :124
      BnbNode *nd = a.nodes + br;
      while (!ld_acquire(&nd->done)) {
      }
      if (*reinterpret_cast<volatile int *>(&ctl->overflow)) {  // a pool ran out: this node's result may be void
        ok = false;
        break;
      }
      // the worker's result block in one round trip
      double r_result, r_bval, r_bfrac;
      long long r_pivots;
      int r_status, r_bvar, r_cut_len, r_cand;
      asm volatile("ld.volatile.global.v2.f64 {%0, %1}, [%2];" : "=d"(r_result), "=d"(r_bval) : "l"(&nd->result) : "memory");
      asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=d"(r_bfrac), "=l"(r_pivots) : "l"(&nd->bfrac) : "memory");
      asm volatile("ld.volatile.global.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(r_status), "=r"(r_bvar), "=r"(r_cut_len), "=r"(r_cand) : "l"(&nd->status) : "memory");
      const int n_status = r_status;
      const double n_value = r_result;
      node_pivots += r_pivots;
      const int ncuts = r_cut_len;
      if (ncuts > max_cuts) max_cuts = ncuts;
      if (n_status == ST_OPTIMAL && n_value < best_eval) {  // :130
        const double frac = r_bfrac;
        if (frac <= a.precision) {  // integer solution: new incumbent (:132-139)
          found = true;
          best_eval = n_value;
          best_node = br;
        } else {  // branch (:141-156); the workers build the children's cut lists
          const int variable = r_bvar;
          const double value = r_bval;
          ok = create(n_value, br, -1.0, variable, ceil(value));       // upper first (:155)
          ok = ok && create(n_value, br, 1.0, variable, floor(value));  // then lower (:156)
          st_release(&ctl->created, created);  // release: the two node records above are visible before the count
        }
      }
      timedout = timed_out();  // :162
      iter++;
    }
    const bool unfinished = (timedout || iter >= a.max_iterations) && hn > 0 && best_eval >= threshold;  // :167
    ctl->status = unfinished ? ST_TIMEDOUT : (!found ? ST_INFEASIBLE : ST_OPTIMAL);
    ctl->found = found ? 1 : 0;
    ctl->result = found ? best_eval : d_nan();
    ctl->best_node = best_node;
    ctl->best_cand = best_node >= 0 ? *reinterpret_cast<volatile int *>(&a.nodes[best_node].cand) : -1;
    ctl->best_height = best_node >= 0 ? rootH + *reinterpret_cast<volatile int *>(&a.nodes[best_node].cut_len) : rootH;
    ctl->iters = (long long)iter;
    ctl->node_pivots = node_pivots;
    ctl->max_cuts = max_cuts;
    ctl->max_heap = max_heap;
    if (!ok) atomicOr(&ctl->overflow, 32);  // the search was abandoned: the host repeats it with the wave driver
    __threadfence();
    st_release(&ctl->stop, 1);
    return;
  }

  // =========================== workers ===========================
  const SmemLayout L(a.Hcap, W, true, NW, true);
  unsigned long long rows_total = 0;
  for (;;) {
    if (tid == 0) {
      const unsigned long long ticket = atomicAdd(&ctl->next_ticket, 1ULL);
      int node = -1;
      if (ticket < (unsigned long long)a.node_cap) {
        for (;;) {
          if ((unsigned long long)ld_acquire(&ctl->created) > ticket) {
            node = (int)ticket;
            break;
          }
          if (ld_acquire(&ctl->stop)) {  // re-check: the last nodes may have been created just before the stop
            if ((unsigned long long)ld_acquire(&ctl->created) > ticket) node = (int)ticket;
            break;
          }
          __nanosleep(64);
        }
        // nodes created before the stop but never needed are skipped once the scheduler is done
        if (node >= 0 && ld_acquire(&ctl->stop)) node = -1;
      }
      s_node = node;
    }
    __syncthreads();
    const int node = s_node;
    __syncthreads();
    if (node < 0) break;
    BnbNode *nd = a.nodes + node;

    // ---- the node's cut list from its parent's (:141-154): same-direction cuts on the branching variable are dropped.
    // Everything another CTA wrote during this launch is read through L2 (__ldcg): L1 lines may predate those writes.
    // Warp 0 filters the parent's list 32 cuts at a time (order kept), appends the new cut, and leaves the list both in
    // the pool (for this node's children) and in shared memory together with each cut's root position (for applyCuts).
    __shared__ unsigned long long s_cut_begin;
    __shared__ int s_ncuts;
    __shared__ BnbCut s_cuts[kBnbMaxCuts];
    __shared__ int s_cutpos[kBnbMaxCuts];
    if (tid < 32) {
      const int lane = tid;
      const int parent = __ldcg(&nd->parent);
      const int new_var = __ldcg(&nd->new_var);
      const double new_sign = __ldcg(&nd->new_sign), new_value = __ldcg(&nd->new_value);
      int pbeg = 0, plen = 0;
      if (parent >= 0) {
        pbeg = __ldcg(&a.nodes[parent].cut_begin);
        plen = __ldcg(&a.nodes[parent].cut_len);
      }
      unsigned long long begin = 0;
      if (lane == 0) begin = atomicAdd(&ctl->cut_top, (unsigned long long)(plen + 1));
      begin = __shfl_sync(0xffffffffu, begin, 0);
      int n = 0;
      if (begin + plen + 1 > a.cut_cap) {
        if (lane == 0) atomicOr(&ctl->overflow, 2);
        n = -1;
      } else if (plen + 1 > kBnbMaxCuts) {
        if (lane == 0) atomicOr(&ctl->overflow, 16);
        n = -1;
      } else {
        for (int i0 = 0; i0 < plen; i0 += 32) {
          const int i = i0 + lane;
          BnbCut c;
          c.sign = c.value = 0.0;
          c.var = -1;
          c.pad = 0;
          bool keep = false;
          if (i < plen) {
            c.sign = __ldcg(&a.cuts[pbeg + i].sign);
            c.value = __ldcg(&a.cuts[pbeg + i].value);
            c.var = __ldcg(&a.cuts[pbeg + i].var);
            keep = c.var != new_var || (new_sign < 0 ? !(c.sign < 0) : (c.sign < 0));
          }
          const unsigned m = __ballot_sync(0xffffffffu, keep);
          if (keep) {
            const int k = n + __popc(m & ((1u << lane) - 1u));
            s_cuts[k] = c;
            s_cutpos[k] = a.root_pos[c.var];
            a.cuts[begin + k] = c;
          }
          n += __popc(m);
        }
        if (lane == 0) {
          BnbCut c;
          c.sign = new_sign;
          c.value = new_value;
          c.var = new_var;
          c.pad = 0;
          s_cuts[n] = c;
          s_cutpos[n] = a.root_pos[new_var];
          a.cuts[begin + n] = c;
        }
        n++;
      }
      if (lane == 0) {
        s_cut_begin = begin;
        s_ncuts = n;
      }
    }
    __syncthreads();
    const int ncuts = s_ncuts;
    if (ncuts < 0 || rootH + ncuts > a.Hcap) {
      if (tid == 0) {
        if (ncuts >= 0) atomicOr(&ctl->overflow, 16);
        nd->cut_begin = 0;
        nd->cut_len = 0;
        nd->status = ST_CYCLED;
        nd->result = d_nan();
        nd->pivots = 0;
        nd->bfrac = 0.0;
        nd->cand = -1;
        __threadfence();
        st_release(&nd->done, 1);
      }
      __syncthreads();
      continue;
    }
    const int H = rootH + ncuts;

    // ---- applyCuts (:22-61) into the shared-memory (A, b) layout of kernels.cuh
    LpView t;
    t.H = H;
    t.W = W;
    t.A = reinterpret_cast<double *>(smem_raw + L.off_A);
    t.ldA = t.ldb = SmemLayout::ld_for(W);
    t.b = t.A + (t.ldA - 2);
    t.var = reinterpret_cast<int *>(smem_raw + L.off_var);
    SplitScratch ss{};
    ss.cc = reinterpret_cast<double *>(smem_raw + L.off_cc);
    ss.list = reinterpret_cast<int *>(smem_raw + L.off_list);
    ss.cnt = reinterpret_cast<int *>(smem_raw + L.off_cnt);
    ss.misc = reinterpret_cast<double *>(smem_raw + L.off_misc);
    ss.red = reinterpret_cast<unsigned *>(smem_raw + L.off_red);
    ss.hist = nullptr;
    ss.hist_cap = 0;
    if (tid == 0) *ss.cnt = 0;
    const int ldA = t.ldA, ldb = t.ldb;
    {
      const int lane = tid & 31;
      const unsigned sA0 = (unsigned)__cvta_generic_to_shared(t.A) + 8u * (unsigned)lane;
      const unsigned sB0 = (unsigned)__cvta_generic_to_shared(t.b);
      const unsigned row_bytes = 8u * (unsigned)ldA;
      const int full = (W - 1) / 32, tail = (W - 1) - 32 * full;
      unsigned sA = sA0 + row_bytes * (unsigned)(tid >> 5), sB = sB0 + 8u * (unsigned)ldb * (unsigned)(tid >> 5);
      const double *g = a.root + (size_t)(tid >> 5) * W;
      for (int r = tid >> 5; r < rootH; r += NW) {
        if (lane == 0) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sB), "l"(g) : "memory");
        const double *gc = g + 1 + lane;
        unsigned d = sA;
        for (int k = 0; k < full; k++) {
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gc) : "memory");
          gc += 32;
          d += 256u;
        }
        if (lane < tail) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gc) : "memory");
        g += (size_t)NW * W;
        sA += row_bytes * NW;
        sB += 8u * (unsigned)ldb * NW;
      }
    }
    for (int e = tid; e < ncuts * W; e += NT) {  // one cell of one cut row per step (:33-42)
      const int i = e / W, c = e - i * W;
      const double sign = s_cuts[i].sign, value = s_cuts[i].value;
      const int p = s_cutpos[i];
      const int r = rootH + i;
      double v;
      if (p < W) {
        v = c == 0 ? __dmul_rn(sign, value) : (c == p ? sign : 0.0);
      } else {
        const double x = a.root[(size_t)(p - W) * W + c];
        v = c == 0 ? __dmul_rn(sign, __dsub_rn(value, x)) : __dmul_rn(-sign, x);
      }
      if (c == 0)
        t.b[(size_t)r * ldb] = v;
      else
        t.A[(size_t)r * ldA + (c - 1)] = v;
    }
    {
      const int nroot = W + rootH;
      for (int k = tid; k < W + H; k += NT) t.var[k] = k < nroot ? a.root_var[k] : k;
      const int pad = ldA - 2 - (W - 1);
      for (int k = tid; k < H * pad; k += NT) t.A[(size_t)(k / pad) * ldA + (W - 1) + (k % pad)] = 0.0;
    }
    cp_async_wait_all();
    __syncthreads();

    // ---- simplex(tableau, options) (:127)
    const LpResult res = simplex_cta_split<NWC, KC, NWR, VW>(t, ss, a.precision, a.max_pivots, 0);
    __syncthreads();
    rows_total += res.rows;

    // ---- mostFractionalVar (:64-85): first integer variable (in `integers` order) with the largest fraction
    int bvar = 0;
    double bval = 0.0, bfrac = 0.0;
    if (res.status == ST_OPTIMAL) {
      double fv = 0.0;
      int fi = kNone;
      for (int p = W + tid; p < W + H; p += NT) {
        const int v = t.var[p];
        const int rank = v < W + rootH ? a.int_rank[v] : -1;
        if (rank >= 0) {
          const double val = t.b[(size_t)(p - W) * ldb];
          const double fr = fabs(__dsub_rn(val, js_round(val)));
          if (fr > 0.0 && (fi == kNone || fr > fv || (fr == fv && rank < fi))) {
            fv = fr;
            fi = rank;
          }
        }
      }
      int parity = 0;
      const int win = block_best<true, NW>(fi == kNone ? no_key<true>() : order_key(fv), fi, ss.red, parity);
      if (win != kNone) {
        bvar = a.ints[win];
        // its row: the position of variable bvar
        __shared__ int s_row;
        for (int p = W + tid; p < W + H; p += NT)
          if (t.var[p] == bvar) s_row = p - W;
        __syncthreads();
        bval = t.b[(size_t)s_row * ldb];
        bfrac = fabs(__dsub_rn(bval, js_round(bval)));
      }
      __syncthreads();
    }

    // ---- publish
    int cand = -1;
    if (res.status == ST_OPTIMAL && bfrac <= a.precision) {  // incumbent candidate: what solution() reads (:175)
      __shared__ int s_cand;
      if (tid == 0) {
        s_cand = atomicAdd(&ctl->cand_top, 1);
        if (s_cand >= a.cand_cap) {
          atomicOr(&ctl->overflow, 4);
          s_cand = -1;
        }
      }
      __syncthreads();
      cand = s_cand;
      if (cand >= 0) {
        double *o_rhs = a.cand_rhs + (size_t)cand * a.Hcap;
        int *o_pos = a.cand_pos + (size_t)cand * (W + a.Hcap), *o_var = a.cand_var + (size_t)cand * (W + a.Hcap);
        for (int r = tid; r < H; r += NT) o_rhs[r] = t.b[(size_t)r * ldb];
        for (int k = tid; k < W + H; k += NT) {
          o_var[k] = t.var[k];
          o_pos[t.var[k]] = k;
        }
      }
    }
    __syncthreads();
    if (tid == 0) {
      nd->cut_begin = (int)s_cut_begin;
      nd->cut_len = ncuts;
      nd->status = res.status;
      nd->result = res.value;
      nd->pivots = res.p1 + res.p2;
      nd->bvar = bvar;
      nd->bval = bval;
      nd->bfrac = bfrac;
      nd->cand = cand;
      __threadfence();
      st_release(&nd->done, 1);
    }
    __syncthreads();
  }
  if (a.rows_out && tid == 0 && rows_total) atomicAdd(a.rows_out, rows_total);
}

}  // namespace yalps
