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
// SPECULATIVE EXPANSION: a search is a chain of dives -- the scheduler pops a node, branches, and the best node of the
// heap is one of the two children it has just created, whose LP (~30 us) has not even started.  So a worker that finds
// its node optimal and fractional creates the node's two children ITSELF (same cuts the scheduler would derive from the
// same result, :141-156) and queues them for evaluation, `spec_depth` generations ahead of the replay at most and only
// while the node's result is below the incumbent the scheduler has published (the reference's own test at :130, taken
// at an earlier time: a superset of what the replay will branch on).  When the replay branches on such a node it adopts
// the two children instead of creating them.  Node ids are pool indices (allocated by whoever creates the node): they
// only name nodes, the heap order depends on keys and push order, which stay the reference's.
// Capacity limits (node pool, cut pool, candidate pool, heap, cut rows per node) are reported through `overflow`; the
// host then repeats the search with the wave driver, which has none.  Speculation has its own share of the node pool
// and simply stops when that is used up.
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
  // ---- result block: four 16-byte chunks, each written by the worker with ONE vector store and each self-validating
  // against the cleared state of the pool (every byte 0xff before the launch: result / pivots / status / done can never
  // hold that), so the scheduler reads a finished node in a single round trip to L2 -- four loads in flight, no
  // acquire in front of them -- and simply reads again while a chunk is still in its cleared state
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
  int child0, child1;  // the children a worker created speculatively (upper branch first, :155-156), -1: none
  // ---- written by whoever creates the branch (the scheduler, or the worker that evaluated the parent)
  double eval;      // parent's rounded result (heap key, :155-156)
  double new_sign, new_value;
  int new_var;
  int parent;       // -1: child of the root
  int depth;        // generations of speculation behind this node (0: created by the scheduler)
  int parent_cut_len;    // the parent's cut list (its creator knows it: one dependent L2 trip less for the worker)
  int parent_cut_begin;
  int pad1;
  double pad2;
};
static_assert(sizeof(BnbNode) == 128, "BnbNode layout");

struct BnbControl {
  unsigned long long next_ticket;  // (unused since the two-queue dispatch; kept for the layout)
  unsigned long long cut_top;      // cut pool bump pointer
  double best_eval;                // the scheduler's incumbent (+inf while none): bound of the workers' speculation
  // two evaluation queues: the scheduler's nodes (ids [0, sched_cap), single producer, no atomics on its path) are
  // served before the speculative ones (ids [sched_cap, node_cap), produced by the workers)
  int s_head;                      // scheduler queue: entries claimed by workers (CAS)
  int p_head, p_tail;              // speculative queue: claimed (CAS) / appended (atomicAdd)
  int spec_count;                  // nodes created speculatively so far
  int created;                     // nodes created in total (written by the scheduler at the end)
  int stop;                        // the scheduler is done
  int cand_top;
  int overflow;                    // 1 nodes, 2 cuts, 4 candidates, 8 heap, 16 cut rows
  // results (scheduler)
  int status, found, best_node, best_height, best_cand, pad;
  double result;
  long long iters, node_pivots, max_cuts, max_heap;
  long long t_total, t_wait, n_wait, t_heap;  // scheduler cycles: whole loop / waiting for results / pops that waited / heap work
  unsigned long long w_nodes, w_cuts, w_asm, w_simplex, w_post;  // workers (thread 0, summed): nodes and cycles per stage
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
  int *queue;              // [node_cap] node id + 1, 0 = not yet written (zeroed before the launch): the scheduler's
                           // queue in [0, sched_cap), the speculative one in [sched_cap, node_cap)
  int sched_cap, spec_cap, spec_depth;  // node_cap = sched_cap + spec_cap; generations of speculation ahead of the replay
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

// A worker's result block: everything it wrote for this node (cut list, candidate arrays, speculative children) is
// fenced in front of four 16-byte stores, one per self-validating chunk (see BnbNode).
__device__ __forceinline__ void publish_result(BnbNode *nd, double result, double bval, double bfrac, long long pivots, int status,
                                               int bvar, int cut_len, int cand, int cut_begin, int child0, int child1) {
  __threadfence();
  asm volatile("st.volatile.global.v2.f64 [%0], {%1, %2};" ::"l"(&nd->result), "d"(result), "d"(bval) : "memory");
  asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(&nd->bfrac), "l"(__double_as_longlong(bfrac)), "l"(pivots) : "memory");
  asm volatile("st.volatile.global.v4.s32 [%0], {%1, %2, %3, %4};" ::"l"(&nd->status), "r"(status), "r"(bvar), "r"(cut_len), "r"(cand)
               : "memory");
  asm volatile("st.volatile.global.v4.s32 [%0], {%1, %2, %3, %4};" ::"l"(&nd->cut_begin), "r"(cut_begin), "r"(1), "r"(child0), "r"(child1)
               : "memory");
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
    int created = 0;  // the scheduler's own nodes: ids and queue slots [0, sched_cap), no atomics on this path
    auto create = [&](double eval, int parent, double sign, int var, double value, int pbeg, int plen) -> bool {
      if (hn >= a.heap_cap) {
        atomicOr(&ctl->overflow, 8);
        return false;
      }
      const int id = created;
      if (id >= a.sched_cap) {
        atomicOr(&ctl->overflow, 1);
        return false;
      }
      created++;
      BnbNode *nd = a.nodes + id;
      nd->eval = eval;
      nd->new_sign = sign;
      nd->new_value = value;
      nd->new_var = var;
      nd->parent = parent;
      nd->depth = 0;
      nd->parent_cut_begin = pbeg;
      nd->parent_cut_len = plen;
      push(eval, id);
      return true;
    };
    // the records of the nodes created since the last call become visible before their queue entries (one fence for the
    // two children of a branch; the fence also orders the result block the scheduler has just read -- and with it the
    // parent's cut list -- before the entries, for the workers that claim them)
    int published = 0;
    auto publish_created = [&]() {
      if (published == created) return;
      __threadfence();
      for (; published < created; published++)
        *reinterpret_cast<volatile int *>(a.queue + published) = published + 1;
    };
    auto adopt = [&](double eval, int id) -> bool {  // a child a worker has created (and queued) already
      if (hn >= a.heap_cap) {
        atomicOr(&ctl->overflow, 8);
        return false;
      }
      push(eval, id);
      return true;
    };
    const unsigned long long t_start = global_ns();
    auto timed_out = [&]() -> bool {
      if (!(a.timeout_ms < 1.0e300)) return false;  // +inf: never
      return (double)(global_ns() - t_start) * 1e-6 >= a.timeout_ms;
    };

    *reinterpret_cast<volatile double *>(&ctl->best_eval) = d_inf();
    // the root's two children (:101-102)
    bool ok = create(a.init_result, -1, -1.0, a.init_var, ceil(a.init_value), 0, 0);
    ok = ok && create(a.init_result, -1, 1.0, a.init_var, floor(a.init_value), 0, 0);
    publish_created();

    const double threshold = a.init_result * (1.0 - a.sign * a.tolerance);  // :114
    bool timedout = timed_out();
    bool found = false;
    double best_eval = d_inf();
    int best_node = -1, best_cand = -1, best_len = 0;
    long long t_wait = 0, n_wait = 0, t_heap = 0;
    const long long t_loop0 = clock64();
    double iter = 0;
    long long node_pivots = 0, max_cuts = 0, max_heap = 0;
    while (ok && iter < a.max_iterations && hn > 0 && best_eval >= threshold && !timedout) {  // :122
      if (hn > max_heap) max_heap = hn;
      // the worker's result block in one round trip (see BnbNode); a pool overflow voids results: checked with it.  The
      // node that pop() is about to return is the heap's top: its loads are issued first and travel while the heap is
      // re-ordered (~1,300 cycles of shared-memory work that does not depend on them)
      BnbNode *nd = a.nodes + hid[0];
      double r_result, r_bval, r_bfrac;
      long long r_pivots;
      int r_status, r_bvar, r_cut_len, r_cand, r_cut_begin, r_done, r_child0, r_child1, r_overflow;
      auto read_block = [&]() {
        asm volatile("ld.volatile.global.v2.f64 {%0, %1}, [%2];" : "=d"(r_result), "=d"(r_bval) : "l"(&nd->result) : "memory");
        asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=d"(r_bfrac), "=l"(r_pivots) : "l"(&nd->bfrac) : "memory");
        asm volatile("ld.volatile.global.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(r_status), "=r"(r_bvar), "=r"(r_cut_len), "=r"(r_cand) : "l"(&nd->status) : "memory");
        asm volatile("ld.volatile.global.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(r_cut_begin), "=r"(r_done), "=r"(r_child0), "=r"(r_child1) : "l"(&nd->cut_begin) : "memory");
        asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(r_overflow) : "l"(&ctl->overflow) : "memory");
      };
      read_block();
      double ev;
      const long long tp0 = clock64();
      const int br = pop(&ev);
      t_heap += clock64() - tp0;
      if (ev > best_eval) break;  // :124
      const long long tw0 = clock64();
      int tries = 1;
      while (!r_overflow && !(__double_as_longlong(r_result) != -1LL && r_pivots != -1LL && r_status != -1 && r_done == 1)) {
        tries++;
        read_block();
      }
      t_wait += clock64() - tw0;
      n_wait += tries > 1;
      (void)r_cut_begin;
      if (r_status == ST_ERR_POOL && !r_overflow) {
        // a speculative node that a pool could not serve has become part of the replay: it is evaluated again as one of the
        // replay's own nodes (result block back to its cleared state, generation 0, a slot of the scheduler's queue -- the
        // slot's own node id stays unused) and awaited in place, so the heap never sees the difference
        publish_created();
        if (created >= a.sched_cap) {
          atomicOr(&ctl->overflow, 1);
          ok = false;
          break;
        }
        asm volatile("st.volatile.global.v4.s32 [%0], {%1, %1, %1, %1};" ::"l"(&nd->result), "r"(-1) : "memory");
        asm volatile("st.volatile.global.v4.s32 [%0], {%1, %1, %1, %1};" ::"l"(&nd->bfrac), "r"(-1) : "memory");
        asm volatile("st.volatile.global.v4.s32 [%0], {%1, %1, %1, %1};" ::"l"(&nd->status), "r"(-1) : "memory");
        asm volatile("st.volatile.global.v4.s32 [%0], {%1, %1, %1, %1};" ::"l"(&nd->cut_begin), "r"(-1) : "memory");
        *reinterpret_cast<volatile int *>(&nd->depth) = 0;
        __threadfence();
        *reinterpret_cast<volatile int *>(a.queue + created) = br + 1;
        created++;
        published++;
        do read_block();
        while (!r_overflow && !(__double_as_longlong(r_result) != -1LL && r_pivots != -1LL && r_status != -1 && r_done == 1));
      }
      if (r_overflow) {  // a pool ran out: this node's result may be void
        ok = false;
        break;
      }
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
          best_cand = r_cand;
          best_len = r_cut_len;
          *reinterpret_cast<volatile double *>(&ctl->best_eval) = best_eval;  // the workers stop speculating above it
        } else if (r_child0 >= 0) {  // branch (:141-156) on children the node's worker has created (and queued) already
          ok = adopt(n_value, r_child0);        // upper first (:155)
          ok = ok && adopt(n_value, r_child1);  // then lower (:156)
        } else {  // branch (:141-156); the workers build the children's cut lists
          const int variable = r_bvar;
          const double value = r_bval;
          ok = create(n_value, br, -1.0, variable, ceil(value), r_cut_begin, r_cut_len);       // upper first (:155)
          ok = ok && create(n_value, br, 1.0, variable, floor(value), r_cut_begin, r_cut_len);  // then lower (:156)
          publish_created();
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
    ctl->best_cand = best_node >= 0 ? best_cand : -1;
    ctl->best_height = best_node >= 0 ? rootH + best_len : rootH;
    ctl->iters = (long long)iter;
    ctl->node_pivots = node_pivots;
    ctl->max_cuts = max_cuts;
    ctl->max_heap = max_heap;
    ctl->t_total = clock64() - t_loop0;
    ctl->t_wait = t_wait;
    ctl->n_wait = n_wait;
    ctl->t_heap = t_heap;
    ctl->created = created + min(*reinterpret_cast<volatile int *>(&ctl->spec_count), a.spec_cap);
    if (!ok) atomicOr(&ctl->overflow, 32);  // the search was abandoned: the host repeats it with the wave driver
    __threadfence();
    st_release(&ctl->stop, 1);
    return;
  }

  // =========================== workers ===========================
  const SmemLayout L(a.Hcap, W, true, NW, true);
  unsigned long long rows_total = 0;
  unsigned long long w_nodes = 0, w_cuts = 0, w_asm = 0, w_simplex = 0, w_post = 0;  // (thread 0; YALPS_BNB_DEBUG prints the sums)
  int own_next = -1;  // (thread 0) the first child this worker has just created: it goes on with it without the queue
  for (;;) {
    if (tid == 0) {
      // claim the next entry: the scheduler's queue first (the replay is waiting for those), then the speculative one
      int node = own_next;
      own_next = -1;
      while (node < 0) {
        const int h = *reinterpret_cast<volatile int *>(&ctl->s_head);
        if (h < a.sched_cap) {
          const int q = ld_acquire(a.queue + h);
          if (q) {
            if (atomicCAS(&ctl->s_head, h, h + 1) == h) {
              node = q - 1;
              break;
            }
            continue;
          }
        }
        const int hp = *reinterpret_cast<volatile int *>(&ctl->p_head);
        if (hp < a.spec_cap) {
          const int q = ld_acquire(a.queue + a.sched_cap + hp);
          if (q) {
            if (atomicCAS(&ctl->p_head, hp, hp + 1) == hp) {
              node = q - 1;
              break;
            }
            continue;
          }
        }
        if (ld_acquire(&ctl->stop)) break;
        __nanosleep(32);
      }
      // nodes created before the stop but never needed are skipped once the scheduler is done
      if (node >= 0 && ld_acquire(&ctl->stop)) node = -1;
      s_node = node;
    }
    __syncthreads();
    const int node = s_node;
    __syncthreads();
    if (node < 0) break;
    BnbNode *nd = a.nodes + node;
    long long wt0 = clock64();

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
      const int pbeg = __ldcg(&nd->parent_cut_begin), plen = __ldcg(&nd->parent_cut_len);  // (0, 0 for the root's children)
      (void)parent;
      unsigned long long begin = 0;
      if (lane == 0) begin = atomicAdd(&ctl->cut_top, (unsigned long long)(plen + 1));
      begin = __shfl_sync(0xffffffffu, begin, 0);
      int n = 0;
      // (a pool that runs out under a SPECULATIVE node only voids that node: the search is abandoned when -- if ever --
      // the replay pops it)
      if (begin + plen + 1 > a.cut_cap) {
        if (lane == 0 && __ldcg(&nd->depth) == 0) atomicOr(&ctl->overflow, 2);
        n = -1;
      } else if (plen + 1 > kBnbMaxCuts) {
        if (lane == 0 && __ldcg(&nd->depth) == 0) atomicOr(&ctl->overflow, 16);
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
    if (tid == 0) {
      const long long now = clock64();
      w_cuts += now - wt0;
      wt0 = now;
    }
    const int ncuts = s_ncuts;
    if (ncuts < 0 || rootH + ncuts > a.Hcap) {
      if (tid == 0) {
        const bool spec = __ldcg(&nd->depth) > 0;
        if (ncuts >= 0 && !spec) atomicOr(&ctl->overflow, 16);
        publish_result(nd, d_nan(), 0.0, 0.0, 0, spec ? ST_ERR_POOL : ST_CYCLED, 0, 0, -1, 0, -1, -1);
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
    if (tid == 0) {
      const long long now = clock64();
      w_asm += now - wt0;
      wt0 = now;
    }

    // ---- simplex(tableau, options) (:127)
    const LpResult res = simplex_cta_split<NWC, KC, NWR, VW>(t, ss, a.precision, a.max_pivots, 0);
    __syncthreads();
    if (tid == 0) {
      const long long now = clock64();
      w_simplex += now - wt0;
      wt0 = now;
    }
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
        const bool spec = __ldcg(&nd->depth) > 0;
        // (speculative nodes leave half of the pool to the replay's own nodes; one that finds no slot is void, see the cut pool)
        if (spec && *reinterpret_cast<volatile int *>(&ctl->cand_top) >= a.cand_cap / 2) {
          s_cand = -2;
        } else {
          s_cand = atomicAdd(&ctl->cand_top, 1);
          if (s_cand >= a.cand_cap) {
            if (!spec) atomicOr(&ctl->overflow, 4);
            s_cand = -2;
          }
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
      // speculative expansion (see the header): the two children the replay will create if it branches on this node
      int c0 = -1, c1 = -1;
      const int depth = __ldcg(&nd->depth);
      if (res.status == ST_OPTIMAL && bfrac > a.precision && depth < a.spec_depth &&
          res.value < *reinterpret_cast<volatile double *>(&ctl->best_eval) && !ld_acquire(&ctl->stop)) {
        const int sp = atomicAdd(&ctl->spec_count, 2);
        if (sp + 2 <= a.spec_cap) {
          const int id = a.sched_cap + sp;
          {
            c0 = id;
            c1 = id + 1;
#pragma unroll
            for (int k = 0; k < 2; k++) {
              BnbNode *ch = a.nodes + id + k;
              ch->eval = res.value;
              ch->new_sign = k == 0 ? -1.0 : 1.0;                 // upper first (:155), then lower (:156)
              ch->new_value = k == 0 ? ceil(bval) : floor(bval);
              ch->new_var = bvar;
              ch->parent = node;
              ch->depth = depth + 1;
              ch->parent_cut_begin = (int)s_cut_begin;
              ch->parent_cut_len = ncuts;
            }
          }
        }
      }
      publish_result(nd, res.value, bval, bfrac, res.p1 + res.p2, cand == -2 ? ST_ERR_POOL : res.status, bvar, ncuts, cand, (int)s_cut_begin, c0,
                     c1);
      w_post += clock64() - wt0;
      w_nodes++;
      if (c0 >= 0) {  // (release: this node's cut list and the children's records are visible to whoever claims them)
        // a dive goes through one of the two children next: this worker continues with the first one itself (no trip
        // through the queue, no claim), the second one is queued for whoever is idle
        const int slot = atomicAdd(&ctl->p_tail, 1);
        st_release(a.queue + a.sched_cap + slot, c1 + 1);
        own_next = c0;
      }
    }
    __syncthreads();
  }
  if (a.rows_out && tid == 0 && rows_total) atomicAdd(a.rows_out, rows_total);
  if (tid == 0 && w_nodes) {
    atomicAdd(&ctl->w_nodes, w_nodes);
    atomicAdd(&ctl->w_cuts, w_cuts);
    atomicAdd(&ctl->w_asm, w_asm);
    atomicAdd(&ctl->w_simplex, w_simplex);
    atomicAdd(&ctl->w_post, w_post);
  }
}

}  // namespace yalps
