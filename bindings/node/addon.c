/*
 * bindings/node/addon.c -- Node N-API binding of libyalps_b200.so (include/yalps_b200.h).
 *
 * NOT LOADED ANYWHERE IN THIS REPOSITORY: the build image has no Node toolchain and no node_api.h.  It is the stub a
 * YALPS maintainer would add.  tests/test_bindings.py compiles it (gcc -fsyntax-only) against a minimal declaration
 * of the N-API functions it uses (tests/stubs/node_api.h), which is as far as it can be checked here.
 *
 * Every typed array is validated (element type and minimum length) BEFORE its pointer reaches the C ABI; a mismatch
 * throws a JS TypeError/RangeError instead of letting the library write past a JS buffer.
 *
 * Exposed to JS (all synchronous, like the reference's solve()):
 *   create(device) -> ctx                       createMulti(devices: Int32Array) -> multi
 *   simplex(ctx, height, width, matrix: Float64Array(h*w), pos: Int32Array(w+h), vars: Int32Array(w+h),
 *           options: Float64Array(6), out: Float64Array(2) [status, value]) -> rc
 *       simplex(tableau, options) of src/simplex.ts:144 IN PLACE: matrix, pos and vars are inputs and outputs
 *       (yalps_solve_batch_basis with every output aliasing its input).
 *   solve(ctx, height, width, matrix, ints: Int32Array, sign, options, status: Int32Array(3), result: Float64Array(2),
 *         rhs: Float64Array(h+2k), pos: Int32Array(w+h+2k), vars: Int32Array(w+h+2k), stats: BigInt64Array(8)) -> rc
 *   solveSparse(ctx, height, width, cells: Int32Array(nnz), values: Float64Array(nnz), ints, sign, options, status, result,
 *               rhs, pos, vars, stats) -> rc
 *       solve() with the tableau as the ordered update() stores of tableauModel (yalps_solve_sparse): big model
 *       tableaus are almost all zeros, which are then neither allocated in JS nor sent over PCIe.
 *   solveMany(multi, heights: Int32Array(n), widths: Int32Array(n), offsets: BigInt64Array(n), matrices: Float64Array,
 *             intsOffsets: BigInt64Array(n+1), ints: Int32Array, signs: Float64Array(n), options,
 *             status: Int32Array(n), result: Float64Array(n), outHeight: Int32Array(n),
 *             rhs: Float64Array(sum h+2k), pos: Int32Array(sum w+h+2k), vars: Int32Array(same)) -> rc
 *   lastError(ctx | multi | undefined) -> string
 */
#include <node_api.h>
#include <stdbool.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/yalps_b200.h"

#define NAPI_OK(call)                     \
  do {                                    \
    if ((call) != napi_ok) {              \
      napi_throw_error(env, NULL, #call); \
      return NULL;                        \
    }                                     \
  } while (0)

/* Pointer of a typed array of exactly `want` element type holding at least `min_len` elements; throws and returns
 * false otherwise.  `optional`: undefined/null is accepted and yields NULL. */
static bool typed_arg(napi_env env, napi_value v, napi_typedarray_type want, size_t min_len, bool optional,
                      const char *name, void **out, size_t *out_len) {
  *out = NULL;
  if (out_len) *out_len = 0;
  napi_valuetype vt;
  if (napi_typeof(env, v, &vt) != napi_ok) return false;
  if (vt == napi_undefined || vt == napi_null) {
    if (optional) return true;
    napi_throw_type_error(env, "YALPS_B200", name);
    return false;
  }
  bool is_typed = false;
  if (napi_is_typedarray(env, v, &is_typed) != napi_ok || !is_typed) {
    napi_throw_type_error(env, "YALPS_B200", name);
    return false;
  }
  napi_typedarray_type type;
  size_t len = 0, off = 0;
  void *data = NULL;
  napi_value buf;
  if (napi_get_typedarray_info(env, v, &type, &len, &data, &buf, &off) != napi_ok) return false;
  if (type != want) {
    napi_throw_type_error(env, "YALPS_B200", name);
    return false;
  }
  if (len < min_len) {
    napi_throw_range_error(env, "YALPS_B200", name);
    return false;
  }
  *out = data;
  if (out_len) *out_len = len;
  return true;
}

static bool read_options(napi_env env, napi_value v, yalps_options *o) {
  /* Float64Array [precision, maxPivots, tolerance, timeout, maxIterations, checkCycles] (Infinity allowed) */
  yalps_default_options(o);
  void *p;
  if (!typed_arg(env, v, napi_float64_array, 6, false, "options: Float64Array(6)", &p, NULL)) return false;
  const double *d = (const double *)p;
  o->precision = d[0];
  o->max_pivots = d[1];
  o->tolerance = d[2];
  o->timeout_ms = d[3];
  o->max_iterations = d[4];
  o->check_cycles = d[5] != 0.0;
  return true;
}

enum { TAG_CTX = 1, TAG_MULTI = 2 };
typedef struct handle {
  int tag;
  void *p;
} handle;

static void handle_finalize(napi_env env, void *data, void *hint) {
  (void)env;
  (void)hint;
  handle *h = (handle *)data;
  if (h->tag == TAG_CTX) yalps_destroy((yalps_ctx *)h->p);
  if (h->tag == TAG_MULTI) yalps_destroy_multi((yalps_multi *)h->p);
  free(h);
}

static void *handle_arg(napi_env env, napi_value v, int tag) {
  handle *h = NULL;
  if (napi_get_value_external(env, v, (void **)&h) != napi_ok || !h || h->tag != tag) {
    napi_throw_type_error(env, "YALPS_B200", tag == TAG_CTX ? "expected a ctx from create()" : "expected a multi from createMulti()");
    return NULL;
  }
  return h->p;
}

static napi_value wrap_handle(napi_env env, int tag, void *p) {
  handle *h = (handle *)malloc(sizeof *h);
  h->tag = tag;
  h->p = p;
  napi_value ext;
  NAPI_OK(napi_create_external(env, h, handle_finalize, NULL, &ext));
  return ext;
}

static napi_value rc_value(napi_env env, int rc) {
  napi_value out;
  napi_create_int32(env, rc, &out);
  return out;
}

static napi_value js_create(napi_env env, napi_callback_info info) {
  size_t argc = 1;
  napi_value argv[1];
  NAPI_OK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
  int32_t device = 0;
  if (argc > 0) napi_get_value_int32(env, argv[0], &device);
  yalps_ctx *ctx = NULL;
  if (yalps_create(device, &ctx) != 0) {
    napi_throw_error(env, "YALPS_B200", yalps_last_error(NULL)); /* no CPU fallback */
    return NULL;
  }
  return wrap_handle(env, TAG_CTX, ctx);
}

static napi_value js_create_multi(napi_env env, napi_callback_info info) {
  size_t argc = 1;
  napi_value argv[1];
  NAPI_OK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
  void *dev;
  size_t ndev;
  if (argc < 1 || !typed_arg(env, argv[0], napi_int32_array, 1, false, "devices: Int32Array", &dev, &ndev)) return NULL;
  yalps_multi *m = NULL;
  if (yalps_create_multi((const int32_t *)dev, (int32_t)ndev, &m) != 0) {
    napi_throw_error(env, "YALPS_B200", yalps_multi_last_error(NULL));
    return NULL;
  }
  return wrap_handle(env, TAG_MULTI, m);
}

static napi_value js_simplex(napi_env env, napi_callback_info info) {
  size_t argc = 8;
  napi_value a[8];
  NAPI_OK(napi_get_cb_info(env, info, &argc, a, NULL, NULL));
  if (argc < 8) {
    napi_throw_type_error(env, "YALPS_B200", "simplex needs 8 arguments");
    return NULL;
  }
  yalps_ctx *ctx = (yalps_ctx *)handle_arg(env, a[0], TAG_CTX);
  if (!ctx) return NULL;
  int32_t h = 0, w = 0;
  napi_get_value_int32(env, a[1], &h);
  napi_get_value_int32(env, a[2], &w);
  if (h < 1 || w < 1) {
    napi_throw_range_error(env, "YALPS_B200", "height and width must be positive");
    return NULL;
  }
  const size_t cells = (size_t)h * (size_t)w, pv = (size_t)h + (size_t)w;
  void *m, *pos, *vars, *out;
  yalps_options o;
  if (!typed_arg(env, a[3], napi_float64_array, cells, false, "matrix: Float64Array(height*width)", &m, NULL) ||
      !typed_arg(env, a[4], napi_int32_array, pv, false, "positionOfVariable: Int32Array(width+height)", &pos, NULL) ||
      !typed_arg(env, a[5], napi_int32_array, pv, false, "variableAtPosition: Int32Array(width+height)", &vars, NULL) ||
      !read_options(env, a[6], &o) ||
      !typed_arg(env, a[7], napi_float64_array, 2, false, "out: Float64Array(2)", &out, NULL))
    return NULL;
  int32_t status = 0;
  double value = 0.0;
  /* every output aliases its input: the reference's simplex() mutates the tableau (src/simplex.ts:5-39) */
  const int rc = yalps_solve_batch_basis(ctx, 1, h, w, (const double *)m, (const int32_t *)pos, (const int32_t *)vars, &o,
                                         &status, &value, NULL, NULL, (int32_t *)pos, (int32_t *)vars, (double *)m);
  ((double *)out)[0] = (double)status;
  ((double *)out)[1] = value;
  return rc_value(env, rc);
}

/* simplexLarge(multi, height, width, matrix, positionOfVariable, variableAtPosition, options, out[2]): simplex() on ONE
 * fresh tableau whose rows are dealt over the GPUs of `multi` (yalps_multi_solve_large); the tableau is mutated in place
 * and both permutation arrays are overwritten, as src/simplex.ts:144 does. */
static napi_value js_simplex_large(napi_env env, napi_callback_info info) {
  size_t argc = 8;
  napi_value a[8];
  NAPI_OK(napi_get_cb_info(env, info, &argc, a, NULL, NULL));
  if (argc < 8) {
    napi_throw_type_error(env, "YALPS_B200", "simplexLarge needs 8 arguments");
    return NULL;
  }
  yalps_multi *mu = (yalps_multi *)handle_arg(env, a[0], TAG_MULTI);
  if (!mu) return NULL;
  int32_t h = 0, w = 0;
  napi_get_value_int32(env, a[1], &h);
  napi_get_value_int32(env, a[2], &w);
  if (h < 1 || w < 1) {
    napi_throw_range_error(env, "YALPS_B200", "height and width must be positive");
    return NULL;
  }
  const size_t cells = (size_t)h * (size_t)w, pv = (size_t)h + (size_t)w;
  void *m, *pos, *vars, *out;
  yalps_options o;
  if (!typed_arg(env, a[3], napi_float64_array, cells, false, "matrix: Float64Array(height*width)", &m, NULL) ||
      !typed_arg(env, a[4], napi_int32_array, pv, false, "positionOfVariable: Int32Array(width+height)", &pos, NULL) ||
      !typed_arg(env, a[5], napi_int32_array, pv, false, "variableAtPosition: Int32Array(width+height)", &vars, NULL) ||
      !read_options(env, a[6], &o) ||
      !typed_arg(env, a[7], napi_float64_array, 2, false, "out: Float64Array(2)", &out, NULL))
    return NULL;
  int32_t status = 0;
  double value = 0.0;
  const int rc = yalps_multi_solve_large(mu, h, w, (const double *)m, &o, &status, &value, NULL, NULL, (int32_t *)pos,
                                         (int32_t *)vars, (double *)m, NULL);
  ((double *)out)[0] = (double)status;
  ((double *)out)[1] = value;
  return rc_value(env, rc);
}

static napi_value js_solve(napi_env env, napi_callback_info info) {
  size_t argc = 13;
  napi_value a[13];
  NAPI_OK(napi_get_cb_info(env, info, &argc, a, NULL, NULL));
  if (argc < 13) {
    napi_throw_type_error(env, "YALPS_B200", "solve needs 13 arguments");
    return NULL;
  }
  yalps_ctx *ctx = (yalps_ctx *)handle_arg(env, a[0], TAG_CTX);
  if (!ctx) return NULL;
  int32_t h = 0, w = 0;
  double sign = 1.0;
  napi_get_value_int32(env, a[1], &h);
  napi_get_value_int32(env, a[2], &w);
  napi_get_value_double(env, a[5], &sign);
  if (h < 1 || w < 1) {
    napi_throw_range_error(env, "YALPS_B200", "height and width must be positive");
    return NULL;
  }
  void *m, *ints, *st, *res, *rhs, *pos, *vars, *stats;
  size_t nints = 0;
  yalps_options o;
  if (!typed_arg(env, a[3], napi_float64_array, (size_t)h * (size_t)w, false, "matrix: Float64Array(height*width)", &m, NULL) ||
      !typed_arg(env, a[4], napi_int32_array, 0, false, "ints: Int32Array", &ints, &nints) || !read_options(env, a[6], &o))
    return NULL;
  /* branch and cut appends up to 2*|integers| cut rows (src/branchAndCut.ts:108): the outputs must have room */
  const size_t rows = (size_t)h + 2 * nints, pv = (size_t)w + rows;
  if (!typed_arg(env, a[7], napi_int32_array, 3, false, "status: Int32Array(3)", &st, NULL) ||
      !typed_arg(env, a[8], napi_float64_array, 2, false, "result: Float64Array(2)", &res, NULL) ||
      !typed_arg(env, a[9], napi_float64_array, rows, false, "rhs: Float64Array(height + 2*ints.length)", &rhs, NULL) ||
      !typed_arg(env, a[10], napi_int32_array, pv, false, "pos: Int32Array(width + height + 2*ints.length)", &pos, NULL) ||
      !typed_arg(env, a[11], napi_int32_array, pv, false, "vars: Int32Array(width + height + 2*ints.length)", &vars, NULL) ||
      !typed_arg(env, a[12], napi_bigint64_array, 8, true, "stats: BigInt64Array(8)", &stats, NULL))
    return NULL;
  int32_t *s3 = (int32_t *)st; /* [status, out_height, root_status] */
  double *r2 = (double *)res;  /* [result, root_value] */
  const int rc = yalps_solve(ctx, h, w, (const double *)m, nints ? (const int32_t *)ints : NULL, (int32_t)nints, sign, &o,
                             &s3[0], &r2[0], &s3[1], (double *)rhs, (int32_t *)pos, (int32_t *)vars, &s3[2], &r2[1], NULL,
                             (int64_t *)stats);
  return rc_value(env, rc);
}

/* solveSparse(ctx, height, width, cells: Int32Array, values: Float64Array, ints, sign, options, status, result, rhs, pos,
 * vars, stats): yalps_solve_sparse -- js_solve with the initial tableau as the ordered update() stores of tableauModel
 * (src/tableau.ts:100-134) instead of the zero-filled Float64Array they land in. */
static napi_value js_solve_sparse(napi_env env, napi_callback_info info) {
  size_t argc = 14;
  napi_value a[14];
  NAPI_OK(napi_get_cb_info(env, info, &argc, a, NULL, NULL));
  if (argc < 14) {
    napi_throw_type_error(env, "YALPS_B200", "solveSparse needs 14 arguments");
    return NULL;
  }
  yalps_ctx *ctx = (yalps_ctx *)handle_arg(env, a[0], TAG_CTX);
  if (!ctx) return NULL;
  int32_t h = 0, w = 0;
  double sign = 1.0;
  napi_get_value_int32(env, a[1], &h);
  napi_get_value_int32(env, a[2], &w);
  napi_get_value_double(env, a[6], &sign);
  if (h < 1 || w < 1) {
    napi_throw_range_error(env, "YALPS_B200", "height and width must be positive");
    return NULL;
  }
  void *cells, *values, *ints, *st, *res, *rhs, *pos, *vars, *stats;
  size_t nnz = 0, nints = 0;
  yalps_options o;
  if (!typed_arg(env, a[3], napi_int32_array, 0, false, "cells: Int32Array", &cells, &nnz) ||
      !typed_arg(env, a[4], napi_float64_array, nnz, false, "values: Float64Array(cells.length)", &values, NULL) ||
      !typed_arg(env, a[5], napi_int32_array, 0, false, "ints: Int32Array", &ints, &nints) || !read_options(env, a[7], &o))
    return NULL;
  const size_t rows = (size_t)h + 2 * nints, pv = (size_t)w + rows;
  if (!typed_arg(env, a[8], napi_int32_array, 3, false, "status: Int32Array(3)", &st, NULL) ||
      !typed_arg(env, a[9], napi_float64_array, 2, false, "result: Float64Array(2)", &res, NULL) ||
      !typed_arg(env, a[10], napi_float64_array, rows, false, "rhs: Float64Array(height + 2*ints.length)", &rhs, NULL) ||
      !typed_arg(env, a[11], napi_int32_array, pv, false, "pos: Int32Array(width + height + 2*ints.length)", &pos, NULL) ||
      !typed_arg(env, a[12], napi_int32_array, pv, false, "vars: Int32Array(width + height + 2*ints.length)", &vars, NULL) ||
      !typed_arg(env, a[13], napi_bigint64_array, 8, true, "stats: BigInt64Array(8)", &stats, NULL))
    return NULL;
  int32_t *s3 = (int32_t *)st; /* [status, out_height, root_status] */
  double *r2 = (double *)res;  /* [result, root_value] */
  const int rc = yalps_solve_sparse(ctx, h, w, (int64_t)nnz, nnz ? (const int32_t *)cells : NULL,
                                    nnz ? (const double *)values : NULL, nints ? (const int32_t *)ints : NULL,
                                    (int32_t)nints, sign, &o, &s3[0], &r2[0], &s3[1], (double *)rhs, (int32_t *)pos,
                                    (int32_t *)vars, &s3[2], &r2[1], NULL, (int64_t *)stats);
  return rc_value(env, rc);
}

static napi_value js_solve_many(napi_env env, napi_callback_info info) {
  size_t argc = 15;
  napi_value a[15];
  NAPI_OK(napi_get_cb_info(env, info, &argc, a, NULL, NULL));
  if (argc < 15) {
    napi_throw_type_error(env, "YALPS_B200", "solveMany needs 15 arguments");
    return NULL;
  }
  yalps_multi *mu = (yalps_multi *)handle_arg(env, a[0], TAG_MULTI);
  if (!mu) return NULL;
  void *heights, *widths, *offs, *mats, *ioffs, *ints, *signs, *status, *result, *outh, *rhs, *pos, *vars;
  size_t n = 0, nmat = 0, nints = 0;
  yalps_options o;
  if (!typed_arg(env, a[1], napi_int32_array, 0, false, "heights: Int32Array(n)", &heights, &n) ||
      !typed_arg(env, a[2], napi_int32_array, n, false, "widths: Int32Array(n)", &widths, NULL) ||
      !typed_arg(env, a[3], napi_bigint64_array, n, false, "offsets: BigInt64Array(n)", &offs, NULL) ||
      !typed_arg(env, a[4], napi_float64_array, 0, false, "matrices: Float64Array", &mats, &nmat) ||
      !typed_arg(env, a[5], napi_bigint64_array, n + 1, false, "intsOffsets: BigInt64Array(n+1)", &ioffs, NULL) ||
      !typed_arg(env, a[6], napi_int32_array, 0, false, "ints: Int32Array", &ints, &nints) ||
      !typed_arg(env, a[7], napi_float64_array, n, false, "signs: Float64Array(n)", &signs, NULL) || !read_options(env, a[8], &o))
    return NULL;
  /* sizes implied by the shapes: every tableau inside `matrices`, every integer list inside `ints`, outputs with room
   * for 2*|integers_i| cut rows per model */
  size_t rows = 0, pvs = 0;
  for (size_t i = 0; i < n; i++) {
    const int32_t h = ((const int32_t *)heights)[i], w = ((const int32_t *)widths)[i];
    const int64_t off = ((const int64_t *)offs)[i], i0 = ((const int64_t *)ioffs)[i], i1 = ((const int64_t *)ioffs)[i + 1];
    if (h < 1 || w < 1 || off < 0 || (uint64_t)off + (uint64_t)h * (uint64_t)w > nmat || i0 < 0 || i1 < i0 || (uint64_t)i1 > nints) {
      napi_throw_range_error(env, "YALPS_B200", "model shapes / offsets do not fit the matrices or ints arrays");
      return NULL;
    }
    rows += (size_t)h + 2 * (size_t)(i1 - i0);
    pvs += (size_t)h + (size_t)w + 2 * (size_t)(i1 - i0);
  }
  if (!typed_arg(env, a[9], napi_int32_array, n, false, "status: Int32Array(n)", &status, NULL) ||
      !typed_arg(env, a[10], napi_float64_array, n, false, "result: Float64Array(n)", &result, NULL) ||
      !typed_arg(env, a[11], napi_int32_array, n, false, "outHeight: Int32Array(n)", &outh, NULL) ||
      !typed_arg(env, a[12], napi_float64_array, rows, false, "rhs: Float64Array(sum height_i + 2*ints_i)", &rhs, NULL) ||
      !typed_arg(env, a[13], napi_int32_array, pvs, false, "pos: Int32Array(sum width_i + height_i + 2*ints_i)", &pos, NULL) ||
      !typed_arg(env, a[14], napi_int32_array, pvs, false, "vars: Int32Array(sum width_i + height_i + 2*ints_i)", &vars, NULL))
    return NULL;
  const int rc = yalps_multi_solve_many(mu, (int64_t)n, (const int32_t *)heights, (const int32_t *)widths, (const int64_t *)offs,
                                        (const double *)mats, (const int64_t *)ioffs, (const int32_t *)ints, (const double *)signs,
                                        &o, 0, (int32_t *)status, (double *)result, (int32_t *)outh, (double *)rhs, (int32_t *)pos,
                                        (int32_t *)vars);
  return rc_value(env, rc);
}

static napi_value js_last_error(napi_env env, napi_callback_info info) {
  size_t argc = 1;
  napi_value a[1];
  NAPI_OK(napi_get_cb_info(env, info, &argc, a, NULL, NULL));
  const char *msg = yalps_last_error(NULL);
  handle *h = NULL;
  if (argc > 0 && napi_get_value_external(env, a[0], (void **)&h) == napi_ok && h) {
    if (h->tag == TAG_CTX) msg = yalps_last_error((const yalps_ctx *)h->p);
    if (h->tag == TAG_MULTI) msg = yalps_multi_last_error((const yalps_multi *)h->p);
  }
  napi_value s;
  NAPI_OK(napi_create_string_utf8(env, msg, NAPI_AUTO_LENGTH, &s));
  return s;
}

NAPI_MODULE_INIT() {
  napi_property_descriptor props[] = {
      {"create", NULL, js_create, NULL, NULL, NULL, napi_default, NULL},
      {"createMulti", NULL, js_create_multi, NULL, NULL, NULL, napi_default, NULL},
      {"simplex", NULL, js_simplex, NULL, NULL, NULL, napi_default, NULL},
      {"simplexLarge", NULL, js_simplex_large, NULL, NULL, NULL, napi_default, NULL},
      {"solve", NULL, js_solve, NULL, NULL, NULL, napi_default, NULL},
      {"solveSparse", NULL, js_solve_sparse, NULL, NULL, NULL, napi_default, NULL},
      {"solveMany", NULL, js_solve_many, NULL, NULL, NULL, napi_default, NULL},
      {"lastError", NULL, js_last_error, NULL, NULL, NULL, napi_default, NULL},
  };
  napi_define_properties(env, exports, sizeof(props) / sizeof(props[0]), props);
  return exports;
}
