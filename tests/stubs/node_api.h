/*
 * tests/stubs/node_api.h -- TEST INFRASTRUCTURE, not Node's header.
 *
 * The build image has no Node toolchain, so bindings/node/addon.c cannot be built into a loadable addon here.
 * tests/test_bindings.py still compiles it with `gcc -fsyntax-only` against these declarations: the subset of the
 * N-API C interface the addon uses, with the signatures of Node's node_api.h / js_native_api.h (N-API version 6,
 * which introduced the BigInt typed arrays).  It proves that the addon is valid C and calls the C ABI with the
 * declared argument types; it does not prove that it runs.
 */
#ifndef YALPS_TEST_NODE_API_STUB_H
#define YALPS_TEST_NODE_API_STUB_H
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

typedef struct napi_env__ *napi_env;
typedef struct napi_value__ *napi_value;
typedef struct napi_callback_info__ *napi_callback_info;
typedef enum { napi_ok = 0, napi_invalid_arg, napi_object_expected, napi_generic_failure = 9 } napi_status;
typedef enum {
  napi_undefined, napi_null, napi_boolean, napi_number, napi_string, napi_symbol, napi_object, napi_function,
  napi_external, napi_bigint
} napi_valuetype;
typedef enum {
  napi_int8_array, napi_uint8_array, napi_uint8_clamped_array, napi_int16_array, napi_uint16_array, napi_int32_array,
  napi_uint32_array, napi_float32_array, napi_float64_array, napi_bigint64_array, napi_biguint64_array
} napi_typedarray_type;
typedef enum { napi_default = 0 } napi_property_attributes;
typedef napi_value (*napi_callback)(napi_env env, napi_callback_info info);
typedef void (*napi_finalize)(napi_env env, void *finalize_data, void *finalize_hint);
typedef struct {
  const char *utf8name;
  napi_value name;
  napi_callback method;
  napi_callback getter;
  napi_callback setter;
  napi_value value;
  napi_property_attributes attributes;
  void *data;
} napi_property_descriptor;

#define NAPI_AUTO_LENGTH SIZE_MAX
#define NAPI_MODULE_INIT() napi_value napi_register_module_v1(napi_env env, napi_value exports)

napi_status napi_get_cb_info(napi_env env, napi_callback_info cbinfo, size_t *argc, napi_value *argv, napi_value *this_arg, void **data);
napi_status napi_typeof(napi_env env, napi_value value, napi_valuetype *result);
napi_status napi_is_typedarray(napi_env env, napi_value value, bool *result);
napi_status napi_get_typedarray_info(napi_env env, napi_value typedarray, napi_typedarray_type *type, size_t *length, void **data, napi_value *arraybuffer, size_t *byte_offset);
napi_status napi_get_value_int32(napi_env env, napi_value value, int32_t *result);
napi_status napi_get_value_int64(napi_env env, napi_value value, int64_t *result);
napi_status napi_get_value_double(napi_env env, napi_value value, double *result);
napi_status napi_get_value_external(napi_env env, napi_value value, void **result);
napi_status napi_create_external(napi_env env, void *data, napi_finalize finalize_cb, void *finalize_hint, napi_value *result);
napi_status napi_create_int32(napi_env env, int32_t value, napi_value *result);
napi_status napi_create_string_utf8(napi_env env, const char *str, size_t length, napi_value *result);
napi_status napi_throw_error(napi_env env, const char *code, const char *msg);
napi_status napi_throw_type_error(napi_env env, const char *code, const char *msg);
napi_status napi_throw_range_error(napi_env env, const char *code, const char *msg);
napi_status napi_define_properties(napi_env env, napi_value object, size_t property_count, const napi_property_descriptor *properties);
#endif
