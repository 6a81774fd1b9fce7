/*
 * oracle/yalps_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Sequential CPU restatement of the YALPS v0.5.6 hot path, written to be the
 * checker for the CUDA kernels in yalps_b200/csrc/.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library; the product path never does.
 *
 * Each function cites the reference lines it follows (paths relative to
 * /root/reference).  The loops are deliberately literal and scalar: same
 * operation order, two-rounding mul-sub (build with -ffp-contract=off), true
 * divisions, strict comparisons, first-index ties, early break in the ratio
 * test, per-phase pivot budgets.
 *
 * Parity pin: checked by tests/test_oracle_*.py against the reference's own
 * fixtures -- tests/cases/<name>.json (status exact, objective 1e-5 relative as in
 * tests/helpers/validate.ts:4-16) and benchmarks/netlib/index.json.  Below
 * 1e-5 (pivot counts, final bases) the reference ships no golden data, so the
 * trajectory itself is "parity unpinned": the oracle is its own witness there.
 *
 * The npm `heap` 0.2.7 dependency (package.json:157) is not vendored in the
 * reference; it is a port of CPython's heapq (push = append + sift toward the
 * root, pop = move last to root + sift to a leaf then back up).  Restated
 * below from the published algorithm; call sites src/branchAndCut.ts:100-102,
 * 123, 155-156.
 */
#include <math.h>
#include <float.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <pthread.h>
#include <stdatomic.h>

enum { ST_OPTIMAL = 0, ST_INFEASIBLE = 1, ST_UNBOUNDED = 2, ST_TIMEDOUT = 3, ST_CYCLED = 4 };

typedef struct {
  double *m;
  int32_t width, height;
  int32_t *pos; /* positionOfVariable */
  int32_t *var; /* variableAtPosition */
} tableau_t;

typedef struct {
  double precision;
  double max_pivots; /* may be +inf (benchmarks/runners.ts:10) */
  int32_t check_cycles;
  double tolerance;
  double timeout_ms; /* may be +inf */
  double max_iterations;
} options_t;

typedef struct {
  int64_t phase1_pivots, phase2_pivots;
} counters_t;

/* JS Math.round: nearest integer, halves toward +inf, keeps -0 (src/util.ts:2-3). */
static double js_round(double x) {
  if (!(fabs(x) < 4503599627370496.0)) return x; /* NaN, inf, already integral */
  double r = floor(x);
  if (x - r >= 0.5) r += 1.0;
  if (r == 0.0 && signbit(x)) r = -0.0;
  return r;
}

/* src/util.ts:1-4 */
double oracle_round_to_precision(double num, double precision) {
  double rounding = js_round(1.0 / precision);
  return js_round((num + DBL_EPSILON) * rounding) / rounding;
}

#define IDX(t, r, c) ((t)->m[(size_t)(r) * (size_t)(t)->width + (size_t)(c)])

/* src/simplex.ts:5-39 */
static void pivot(tableau_t *t, int row, int col, int *nz) {
  const int W = t->width, H = t->height;
  const double quotient = IDX(t, row, col);
  const int32_t leaving = t->var[W + row];
  const int32_t entering = t->var[col];
  t->var[W + row] = entering;
  t->var[col] = leaving;
  t->pos[leaving] = col;
  t->pos[entering] = W + row;

  int nnz = 0;
  for (int c = 0; c < W; c++) {
    const double value = IDX(t, row, c);
    if (fabs(value) > 1e-16) {
      IDX(t, row, c) = value / quotient;
      nz[nnz++] = c;
    } else {
      IDX(t, row, c) = 0.0;
    }
  }
  IDX(t, row, col) = 1.0 / quotient;

  for (int r = 0; r < H; r++) {
    if (r == row) continue;
    const double coef = IDX(t, r, col);
    if (fabs(coef) > 1e-16) {
      for (int i = 0; i < nnz; i++) {
        const int c = nz[i];
        const double prod = coef * IDX(t, row, c); /* rounded product ... */
        IDX(t, r, c) = IDX(t, r, c) - prod;        /* ... then rounded difference */
      }
      IDX(t, r, col) = -coef / quotient;
    }
  }
}

typedef struct {
  int32_t *pairs;
  size_t len, cap;
} history_t;

/* src/simplex.ts:44-63 */
static int has_cycle(history_t *h, const tableau_t *t, int row, int col) {
  if (h->len == h->cap) {
    h->cap = h->cap ? h->cap * 2 : 64;
    h->pairs = (int32_t *)realloc(h->pairs, h->cap * 2 * sizeof(int32_t));
  }
  h->pairs[2 * h->len] = t->var[t->width + row];
  h->pairs[2 * h->len + 1] = t->var[col];
  h->len++;
  for (size_t length = 6; length <= h->len / 2; length++) {
    int cycle = 1;
    for (size_t i = 0; i < length; i++) {
      const size_t item = h->len - 1 - i;
      if (h->pairs[2 * item] != h->pairs[2 * (item - length)] ||
          h->pairs[2 * item + 1] != h->pairs[2 * (item - length) + 1]) {
        cycle = 0;
        break;
      }
    }
    if (cycle) return 1;
  }
  return 0;
}

/* src/simplex.ts:66-103 */
static int phase2(tableau_t *t, const options_t *o, double *result, int *nz, counters_t *cnt) {
  history_t hist = {0, 0, 0};
  const double precision = o->precision;
  int status = ST_CYCLED;
  *result = NAN;
  for (double iter = 0; iter < o->max_pivots; iter++) {
    int col = 0;
    double value = precision;
    for (int c = 1; c < t->width; c++) {
      const double reduced = IDX(t, 0, c);
      if (reduced > value) {
        value = reduced;
        col = c;
      }
    }
    if (col == 0) {
      status = ST_OPTIMAL;
      *result = oracle_round_to_precision(IDX(t, 0, 0), precision);
      break;
    }

    int row = 0;
    double min_ratio = INFINITY;
    for (int r = 1; r < t->height; r++) {
      const double v = IDX(t, r, col);
      if (v <= precision) continue;
      const double rhs = IDX(t, r, 0);
      const double ratio = rhs / v;
      if (ratio < min_ratio) {
        row = r;
        min_ratio = ratio;
        if (ratio <= precision) break;
      }
    }
    if (row == 0) {
      status = ST_UNBOUNDED;
      *result = (double)col;
      break;
    }

    if (o->check_cycles && has_cycle(&hist, t, row, col)) {
      status = ST_CYCLED;
      *result = NAN;
      break;
    }
    pivot(t, row, col, nz);
    if (cnt) cnt->phase2_pivots++;
  }
  free(hist.pairs);
  return status;
}

/* src/simplex.ts:106-142 (exported as `simplex`, :144) */
static int phase1(tableau_t *t, const options_t *o, double *result, int *nz, counters_t *cnt) {
  history_t hist = {0, 0, 0};
  const double precision = o->precision;
  int status = ST_CYCLED;
  *result = NAN;
  for (double iter = 0; iter < o->max_pivots; iter++) {
    int row = 0;
    double rhs = -precision;
    for (int r = 1; r < t->height; r++) {
      const double v = IDX(t, r, 0);
      if (v < rhs) {
        rhs = v;
        row = r;
      }
    }
    if (row == 0) {
      free(hist.pairs);
      return phase2(t, o, result, nz, cnt);
    }

    int col = 0;
    double max_ratio = -INFINITY;
    for (int c = 1; c < t->width; c++) {
      const double coefficient = IDX(t, row, c);
      if (coefficient < -precision) {
        const double ratio = -IDX(t, 0, c) / coefficient;
        if (ratio > max_ratio) {
          max_ratio = ratio;
          col = c;
        }
      }
    }
    if (col == 0) {
      status = ST_INFEASIBLE;
      break;
    }

    if (o->check_cycles && has_cycle(&hist, t, row, col)) {
      status = ST_CYCLED;
      break;
    }
    pivot(t, row, col, nz);
    if (cnt) cnt->phase1_pivots++;
  }
  free(hist.pairs);
  return status;
}

/* Public: simplex on one tableau, in place.  pos/var must hold width+height ints. */
int oracle_simplex(double *matrix, int32_t width, int32_t height, int32_t *pos, int32_t *var,
                   double precision, double max_pivots, int32_t check_cycles, double *result,
                   int64_t *pivots /* [2] phase1, phase2; may be NULL */) {
  tableau_t t = {matrix, width, height, pos, var};
  options_t o = {precision, max_pivots, check_cycles, 0.0, INFINITY, 32768.0};
  counters_t cnt = {0, 0};
  int *nz = (int *)malloc(sizeof(int) * (size_t)(width > 0 ? width : 1));
  int st = phase1(&t, &o, result, nz, &cnt);
  free(nz);
  if (pivots) {
    pivots[0] = cnt.phase1_pivots;
    pivots[1] = cnt.phase2_pivots;
  }
  return st;
}

/* Batch of same-shape tableaus with identity pos/var, optionally multi-threaded
 * (pthreads; each thread claims chunks of 16 LPs from a shared counter).  This
 * is the CPU baseline leg of bench.py.
 * rhs_out: n*height (column 0 after the solve); pos_out: n*(width+height). */
typedef struct {
  int64_t n;
  double *matrices;
  int32_t width, height;
  double precision, max_pivots;
  int32_t check_cycles;
  int32_t *status;
  double *value;
  int64_t *pivots;
  double *rhs_out;
  int32_t *pos_out, *var_out;
  atomic_llong next;
} batch_job_t;

static void *batch_worker(void *arg) {
  batch_job_t *j = (batch_job_t *)arg;
  const int32_t width = j->width, height = j->height;
  const size_t cells = (size_t)width * (size_t)height;
  const int nv = width + height;
  int32_t *pos = (int32_t *)malloc(sizeof(int32_t) * (size_t)nv);
  int32_t *var = (int32_t *)malloc(sizeof(int32_t) * (size_t)nv);
  for (;;) {
    const int64_t lo = atomic_fetch_add(&j->next, 16);
    if (lo >= j->n) break;
    const int64_t hi = lo + 16 < j->n ? lo + 16 : j->n;
    for (int64_t i = lo; i < hi; i++) {
      for (int k = 0; k < nv; k++) pos[k] = var[k] = k;
      double *m = j->matrices + (size_t)i * cells;
      j->status[i] = oracle_simplex(m, width, height, pos, var, j->precision, j->max_pivots, j->check_cycles,
                                    &j->value[i], j->pivots ? j->pivots + 2 * i : NULL);
      if (j->rhs_out)
        for (int r = 0; r < height; r++) j->rhs_out[(size_t)i * height + r] = m[(size_t)r * width];
      if (j->pos_out) memcpy(j->pos_out + (size_t)i * nv, pos, sizeof(int32_t) * (size_t)nv);
      if (j->var_out) memcpy(j->var_out + (size_t)i * nv, var, sizeof(int32_t) * (size_t)nv);
    }
  }
  free(pos);
  free(var);
  return NULL;
}

int oracle_simplex_batch(int64_t n, double *matrices, int32_t width, int32_t height, double precision,
                         double max_pivots, int32_t check_cycles, int32_t *status, double *value,
                         int64_t *pivots /* n*2 */, double *rhs_out, int32_t *pos_out, int32_t *var_out,
                         int32_t nthreads) {
  batch_job_t job = {n, matrices, width, height, precision, max_pivots, check_cycles, status, value,
                     pivots, rhs_out, pos_out, var_out, 0};
  if (nthreads <= 1) {
    batch_worker(&job);
    return 0;
  }
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
  for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, batch_worker, &job);
  for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
  free(th);
  return 0;
}

/* ------------------------------------------------------------------------- */
/* Branch and cut (src/branchAndCut.ts)                                       */

typedef struct {
  double sign;
  int32_t variable;
  double value;
} cut_t;

typedef struct {
  double eval;
  cut_t *cuts;
  int32_t ncuts;
} branch_t;

/* heap@0.2.7 == CPython heapq with cmp(x,y) = x[0]-y[0] (branchAndCut.ts:100) */
typedef struct {
  branch_t *a;
  size_t len, cap;
} heap_t;

static int heap_lt(const branch_t *x, const branch_t *y) { return x->eval - y->eval < 0; }

static void heap_siftdown(heap_t *h, size_t startpos, size_t pos) { /* toward the root */
  branch_t newitem = h->a[pos];
  while (pos > startpos) {
    size_t parentpos = (pos - 1) >> 1;
    if (heap_lt(&newitem, &h->a[parentpos])) {
      h->a[pos] = h->a[parentpos];
      pos = parentpos;
      continue;
    }
    break;
  }
  h->a[pos] = newitem;
}

static void heap_siftup(heap_t *h, size_t pos) { /* to a leaf, then back up */
  const size_t endpos = h->len, startpos = pos;
  branch_t newitem = h->a[pos];
  size_t childpos = 2 * pos + 1;
  while (childpos < endpos) {
    size_t rightpos = childpos + 1;
    if (rightpos < endpos && !heap_lt(&h->a[childpos], &h->a[rightpos])) childpos = rightpos;
    h->a[pos] = h->a[childpos];
    pos = childpos;
    childpos = 2 * pos + 1;
  }
  h->a[pos] = newitem;
  heap_siftdown(h, startpos, pos);
}

static void heap_push(heap_t *h, branch_t b) {
  if (h->len == h->cap) {
    h->cap = h->cap ? h->cap * 2 : 64;
    h->a = (branch_t *)realloc(h->a, h->cap * sizeof(branch_t));
  }
  h->a[h->len++] = b;
  heap_siftdown(h, 0, h->len - 1);
}

static branch_t heap_pop(heap_t *h) {
  branch_t last = h->a[--h->len];
  if (h->len) {
    branch_t ret = h->a[0];
    h->a[0] = last;
    heap_siftup(h, 0);
    return ret;
  }
  return last;
}

/* src/branchAndCut.ts:22-61.  buf holds (height+maxExtra)*width doubles etc. */
static tableau_t apply_cuts(const tableau_t *root, double *bm, int32_t *bpos, int32_t *bvar, const cut_t *cuts,
                            int32_t ncuts) {
  const int width = root->width, height = root->height;
  memcpy(bm, root->m, sizeof(double) * (size_t)width * (size_t)height);
  for (int i = 0; i < ncuts; i++) {
    const double sign = cuts[i].sign, value = cuts[i].value;
    const size_t r = (size_t)(height + i) * (size_t)width;
    const int32_t pos = root->pos[cuts[i].variable];
    if (pos < width) {
      bm[r] = sign * value;
      for (int c = 1; c < width; c++) bm[r + c] = 0.0;
      bm[r + pos] = sign;
    } else {
      const size_t row = (size_t)(pos - width) * (size_t)width;
      bm[r] = sign * (value - bm[row]);
      for (int c = 1; c < width; c++) bm[r + c] = -sign * bm[row + c];
    }
  }
  const int nv = width + height;
  memcpy(bpos, root->pos, sizeof(int32_t) * (size_t)nv);
  memcpy(bvar, root->var, sizeof(int32_t) * (size_t)nv);
  for (int i = nv; i < nv + ncuts; i++) bpos[i] = bvar[i] = i;
  tableau_t t = {bm, width, height + ncuts, bpos, bvar};
  return t;
}

/* src/branchAndCut.ts:64-85 */
static void most_fractional_var(const tableau_t *t, const int32_t *ints, int32_t nints, int32_t *variable,
                                double *value, double *frac) {
  double highest = 0.0, val_out = 0.0;
  int32_t v_out = 0;
  for (int i = 0; i < nints; i++) {
    const int32_t iv = ints[i];
    const int32_t row = t->pos[iv] - t->width;
    if (row < 0) continue;
    const double val = IDX(t, row, 0);
    const double fr = fabs(val - js_round(val));
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

static double now_ms(void) { /* Date.now(): integer milliseconds */
  struct timespec ts;
  clock_gettime(CLOCK_REALTIME, &ts);
  return floor((double)ts.tv_sec * 1000.0 + (double)ts.tv_nsec / 1e6);
}

typedef struct {
  int64_t nodes, node_pivots, max_cuts, max_heap;
} bnb_stats_t;

/*
 * src/branchAndCut.ts:89-176.  root_* describe the root tableau after the
 * root LP was solved to "optimal" with value init_result.  On return the
 * best tableau's RHS column / pos / var are written to the out arrays (sized
 * for height + 2*nints rows) and *out_height gives its height.
 * node_log (optional, 4 doubles per evaluated node: status, result, ncuts,
 * pivots) records the per-node trajectory for the GPU parity tests.
 */
int oracle_branch_and_cut(double *root_m, int32_t width, int32_t height, int32_t *root_pos, int32_t *root_var,
                          const int32_t *ints, int32_t nints, double sign, double init_result, double precision,
                          double max_pivots, int32_t check_cycles, double tolerance, double timeout_ms,
                          double max_iterations, double *result, int32_t *out_height, double *out_rhs,
                          int32_t *out_pos, int32_t *out_var, int64_t *stats /* [4] */, double *node_log,
                          int64_t node_log_cap) {
  tableau_t root = {root_m, width, height, root_pos, root_var};
  options_t o = {precision, max_pivots, check_cycles, tolerance, timeout_ms, max_iterations};
  bnb_stats_t st = {0, 0, 0, 0};
  int32_t init_var;
  double init_val, init_frac;
  const tableau_t *best = &root;
  tableau_t best_store;
  int status;
  int *nz = (int *)malloc(sizeof(int) * (size_t)width);
  double *bufm[2] = {0, 0};
  int32_t *bufp[2] = {0, 0}, *bufv[2] = {0, 0};
  heap_t heap = {0, 0, 0};

  most_fractional_var(&root, ints, nints, &init_var, &init_val, &init_frac);
  if (init_frac <= precision) { /* :98 */
    status = ST_OPTIMAL;
    *result = init_result;
    goto done;
  }

  {
    cut_t *c1 = (cut_t *)malloc(sizeof(cut_t));
    cut_t *c2 = (cut_t *)malloc(sizeof(cut_t));
    c1[0] = (cut_t){-1.0, init_var, ceil(init_val)};
    c2[0] = (cut_t){1.0, init_var, floor(init_val)};
    heap_push(&heap, (branch_t){init_result, c1, 1});
    heap_push(&heap, (branch_t){init_result, c2, 1});
  }

  const int max_extra = nints * 2;
  const size_t mlen = (size_t)width * (size_t)(height + max_extra);
  const size_t plen = (size_t)(width + height + max_extra);
  for (int b = 0; b < 2; b++) {
    bufm[b] = (double *)calloc(mlen, sizeof(double));
    bufp[b] = (int32_t *)calloc(plen, sizeof(int32_t));
    bufv[b] = (int32_t *)calloc(plen, sizeof(int32_t));
  }
  int cand = 0; /* candidateBuffer index; solutionBuffer = 1 - cand */

  const double threshold = init_result * (1.0 - sign * tolerance);
  const double stop_time = timeout_ms + now_ms();
  int timedout = now_ms() >= stop_time;
  int found = 0;
  double best_eval = INFINITY;
  double iter = 0;

  while (iter < max_iterations && heap.len && best_eval >= threshold && !timedout) {
    if ((int64_t)heap.len > st.max_heap) st.max_heap = (int64_t)heap.len;
    branch_t br = heap_pop(&heap);
    if (br.eval > best_eval) {
      free(br.cuts);
      break;
    }
    tableau_t cur = apply_cuts(&root, bufm[cand], bufp[cand], bufv[cand], br.cuts, br.ncuts);
    double res;
    counters_t cnt = {0, 0};
    const int s = phase1(&cur, &o, &res, nz, &cnt);
    if (node_log && st.nodes < node_log_cap) {
      node_log[4 * st.nodes + 0] = (double)s;
      node_log[4 * st.nodes + 1] = res;
      node_log[4 * st.nodes + 2] = (double)br.ncuts;
      node_log[4 * st.nodes + 3] = (double)(cnt.phase1_pivots + cnt.phase2_pivots);
    }
    st.nodes++;
    st.node_pivots += cnt.phase1_pivots + cnt.phase2_pivots;
    if (br.ncuts > st.max_cuts) st.max_cuts = br.ncuts;
    if (s == ST_OPTIMAL && res < best_eval) {
      int32_t variable;
      double value, frac;
      most_fractional_var(&cur, ints, nints, &variable, &value, &frac);
      if (frac <= precision) {
        found = 1;
        best_eval = res;
        best_store = cur;
        best = &best_store;
        cand = 1 - cand; /* swap buffers :137-139 */
      } else {
        cut_t *upper = (cut_t *)malloc(sizeof(cut_t) * (size_t)(br.ncuts + 1));
        cut_t *lower = (cut_t *)malloc(sizeof(cut_t) * (size_t)(br.ncuts + 1));
        int nu = 0, nl = 0;
        for (int i = 0; i < br.ncuts; i++) {
          const cut_t cut = br.cuts[i];
          if (cut.variable == variable) {
            if (cut.sign < 0) lower[nl++] = cut;
            else upper[nu++] = cut;
          } else {
            upper[nu++] = cut;
            lower[nl++] = cut;
          }
        }
        lower[nl++] = (cut_t){1.0, variable, floor(value)};
        upper[nu++] = (cut_t){-1.0, variable, ceil(value)};
        heap_push(&heap, (branch_t){res, upper, nu});
        heap_push(&heap, (branch_t){res, lower, nl});
      }
    }
    free(br.cuts);
    timedout = now_ms() >= stop_time;
    iter++;
  }

  {
    const int unfinished = (timedout || iter >= max_iterations) && heap.len && best_eval >= threshold;
    status = unfinished ? ST_TIMEDOUT : !found ? ST_INFEASIBLE : ST_OPTIMAL;
    *result = found ? best_eval : NAN;
  }

done:
  *out_height = best->height;
  for (int r = 0; r < best->height; r++) out_rhs[r] = IDX(best, r, 0);
  memcpy(out_pos, best->pos, sizeof(int32_t) * (size_t)(best->width + best->height));
  memcpy(out_var, best->var, sizeof(int32_t) * (size_t)(best->width + best->height));
  if (stats) {
    stats[0] = st.nodes;
    stats[1] = st.node_pivots;
    stats[2] = st.max_cuts;
    stats[3] = st.max_heap;
  }
  for (size_t i = 0; i < heap.len; i++) free(heap.a[i].cuts);
  free(heap.a);
  for (int b = 0; b < 2; b++) {
    free(bufm[b]);
    free(bufp[b]);
    free(bufv[b]);
  }
  free(nz);
  return status;
}

/* applyCuts alone, for the device node-assembly parity test. */
void oracle_apply_cuts(const double *root_m, int32_t width, int32_t height, const int32_t *root_pos,
                       const int32_t *root_var, const double *cut_sign, const int32_t *cut_var,
                       const double *cut_value, int32_t ncuts, double *out_m, int32_t *out_pos, int32_t *out_var) {
  tableau_t root = {(double *)root_m, width, height, (int32_t *)root_pos, (int32_t *)root_var};
  cut_t *cuts = (cut_t *)malloc(sizeof(cut_t) * (size_t)(ncuts > 0 ? ncuts : 1));
  for (int i = 0; i < ncuts; i++) cuts[i] = (cut_t){cut_sign[i], cut_var[i], cut_value[i]};
  apply_cuts(&root, out_m, out_pos, out_var, cuts, ncuts);
  free(cuts);
}

/* prospectorHash / newRand (tests/helpers/util.ts:20-41), used by the synthetic generators */
static uint32_t prospector_hash(uint32_t x) {
  x ^= x >> 16;
  x *= 0x21f0aaadu;
  x ^= x >> 15;
  x *= 0xd35a2d97u;
  x ^= x >> 15;
  return x;
}

uint32_t oracle_prospector_hash(uint32_t x) { return prospector_hash(x); }

/* SURVEY.md 8(d) config 2/5 generator: tableau (m+1) x (n+1), row 0 = c, col 0 = b.
 * neg_rows > 0 negates the first neg_rows constraint rows with RHS -(0.5+U)
 * (phase-1 exercising variant).  Draw order: c_1..c_n, then per row a_k1..a_kn, b_k. */
void oracle_generate_synthetic(int64_t first, int64_t n, int32_t m, int32_t nvars, int32_t neg_rows,
                               uint32_t salt, double *out) {
  const size_t W = (size_t)nvars + 1, H = (size_t)m + 1;
  for (int64_t i = 0; i < n; i++) {
    double *t = out + (size_t)i * W * H;
    uint32_t seed = prospector_hash((uint32_t)(first + i) ^ salt);
    t[0] = 0.0;
    for (size_t c = 1; c < W; c++) {
      seed += 0x9e3779b9u;
      t[c] = (double)prospector_hash(seed) / 4294967296.0;
    }
    for (size_t k = 1; k < H; k++) {
      for (size_t c = 1; c < W; c++) {
        seed += 0x9e3779b9u;
        t[k * W + c] = (double)prospector_hash(seed) / 4294967296.0;
      }
      seed += 0x9e3779b9u;
      const double u = (double)prospector_hash(seed) / 4294967296.0;
      if ((int)k <= neg_rows) {
        for (size_t c = 1; c < W; c++) t[k * W + c] = -t[k * W + c];
        t[k * W] = -(0.5 + u);
      } else {
        t[k * W] = (double)nvars * (0.25 + 0.5 * u);
      }
    }
  }
}
