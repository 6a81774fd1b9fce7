/*
 * bindings/node/addon.c -- Node N-API binding of libyalps_b200.so (include/yalps_b200.h).
 *
 * NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no Node toolchain and no node_api.h.
 * It is the stub a YALPS maintainer would add; it only forwards typed-array pointers to the C ABI.
 *
 * Exposed to JS:
 *   create(device) -> ctx (external)
 *   solveBatch(ctx, n, height, width, matrices: Float64Array, options: Float64Array(6),
 *              status: Int32Array, value: Float64Array, pivots: BigInt64Array,
 *              rhs: Float64Array, pos: Int32Array, vars: Int32Array) -> rc
 *   solve(ctx, height, width, matrix, ints: Int32Array, sign, options, out: {...typed arrays}) -> rc
 *   lastError(ctx) -> string
 */
#include <node_api.h>
#include <stdint.h>
#include <string.h>

#include "../../include/yalps_b200.h"

#define NAPI_OK(call)                         \
  do {                                        \
    if ((call) != napi_ok) {                  \
      napi_throw_error(env, NULL, #call);     \
      return NULL;                            \
    }                                         \
  } while (0)

static void *typed_data(napi_env env, napi_value v) {
  bool is_typed = false;
  if (napi_is_typedarray(env, v, &is_typed) != napi_ok || !is_typed) return NULL;
  void *data = NULL;
  size_t len;
  napi_typedarray_type type;
  napi_value buf;
  size_t off;
  napi_get_typedarray_info(env, v, &type, &len, &data, &buf, &off);
  return data;
}

static yalps_options read_options(napi_env env, napi_value v) {
  /* Float64Array [precision, maxPivots, tolerance, timeout, maxIterations, checkCycles] (Infinity allowed) */
  yalps_options o;
  yalps_default_options(&o);
  double *d = (double *)typed_data(env, v);
  if (d) {
    o.precision = d[0];
    o.max_pivots = d[1];
    o.tolerance = d[2];
    o.timeout_ms = d[3];
    o.max_iterations = d[4];
    o.check_cycles = d[5] != 0.0;
  }
  return o;
}

static void ctx_finalize(napi_env env, void *data, void *hint) {
  (void)env;
  (void)hint;
  yalps_destroy((yalps_ctx *)data);
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
  napi_value ext;
  NAPI_OK(napi_create_external(env, ctx, ctx_finalize, NULL, &ext));
  return ext;
}

static napi_value js_solve_batch(napi_env env, napi_callback_info info) {
  size_t argc = 12;
  napi_value a[12];
  NAPI_OK(napi_get_cb_info(env, info, &argc, a, NULL, NULL));
  yalps_ctx *ctx;
  NAPI_OK(napi_get_value_external(env, a[0], (void **)&ctx));
  int64_t n;
  int32_t h, w;
  napi_get_value_int64(env, a[1], &n);
  napi_get_value_int32(env, a[2], &h);
  napi_get_value_int32(env, a[3], &w);
  yalps_options o = read_options(env, a[5]);
  int rc = yalps_solve_batch(ctx, n, h, w, (const double *)typed_data(env, a[4]), &o, (int32_t *)typed_data(env, a[6]),
                             (double *)typed_data(env, a[7]), (int64_t *)typed_data(env, a[8]),
                             (double *)typed_data(env, a[9]), (int32_t *)typed_data(env, a[10]),
                             (int32_t *)typed_data(env, a[11]), NULL);
  napi_value out;
  napi_create_int32(env, rc, &out);
  return out;
}

static napi_value js_solve(napi_env env, napi_callback_info info) {
  /* (ctx, height, width, matrix, ints, sign, options, status:Int32Array(3), result:Float64Array(2),
   *  rhs, pos, vars, stats:BigInt64Array(8)) */
  size_t argc = 13;
  napi_value a[13];
  NAPI_OK(napi_get_cb_info(env, info, &argc, a, NULL, NULL));
  yalps_ctx *ctx;
  NAPI_OK(napi_get_value_external(env, a[0], (void **)&ctx));
  int32_t h, w;
  double sign;
  napi_get_value_int32(env, a[1], &h);
  napi_get_value_int32(env, a[2], &w);
  napi_get_value_double(env, a[5], &sign);
  size_t nints = 0;
  void *ints = NULL;
  napi_typedarray_type ty;
  napi_value buf;
  size_t off;
  napi_get_typedarray_info(env, a[4], &ty, &nints, &ints, &buf, &off);
  yalps_options o = read_options(env, a[6]);
  int32_t *st = (int32_t *)typed_data(env, a[7]); /* [status, out_height, root_status] */
  double *res = (double *)typed_data(env, a[8]);  /* [result, root_value] */
  int rc = yalps_solve(ctx, h, w, (const double *)typed_data(env, a[3]), (const int32_t *)ints, (int32_t)nints, sign,
                       &o, &st[0], &res[0], &st[1], (double *)typed_data(env, a[9]), (int32_t *)typed_data(env, a[10]),
                       (int32_t *)typed_data(env, a[11]), &st[2], &res[1], NULL, (int64_t *)typed_data(env, a[12]));
  napi_value out;
  napi_create_int32(env, rc, &out);
  return out;
}

static napi_value js_last_error(napi_env env, napi_callback_info info) {
  size_t argc = 1;
  napi_value a[1];
  NAPI_OK(napi_get_cb_info(env, info, &argc, a, NULL, NULL));
  yalps_ctx *ctx = NULL;
  if (argc > 0) napi_get_value_external(env, a[0], (void **)&ctx);
  napi_value s;
  napi_create_string_utf8(env, yalps_last_error(ctx), NAPI_AUTO_LENGTH, &s);
  return s;
}

NAPI_MODULE_INIT() {
  napi_property_descriptor props[] = {
      {"create", NULL, js_create, NULL, NULL, NULL, napi_default, NULL},
      {"solveBatch", NULL, js_solve_batch, NULL, NULL, NULL, napi_default, NULL},
      {"solve", NULL, js_solve, NULL, NULL, NULL, napi_default, NULL},
      {"lastError", NULL, js_last_error, NULL, NULL, NULL, napi_default, NULL},
  };
  napi_define_properties(env, exports, sizeof(props) / sizeof(props[0]), props);
  return exports;
}
