/* Test double of Node's <node_api.h>: the subset of the N-API C ABI that js/napi/jsrt_addon.cc uses, with the
 * signatures of the real header (Node 18 LTS, NAPI_VERSION 8).  No Node runtime exists in the build image, so the
 * addon is compiled against this header and driven by mock_runtime.cc (tests/test_napi_addon.py).  Test
 * infrastructure only; the product addon builds against the real header with node-gyp (js/napi/binding.gyp). */
#ifndef JSRT_TEST_NODE_API_H
#define JSRT_TEST_NODE_API_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct napi_env__* napi_env;
typedef struct napi_value__* napi_value;
typedef struct napi_callback_info__* napi_callback_info;
typedef enum {
    napi_ok, napi_invalid_arg, napi_object_expected, napi_string_expected, napi_name_expected, napi_function_expected,
    napi_number_expected, napi_boolean_expected, napi_array_expected, napi_generic_failure, napi_pending_exception
} napi_status;
typedef enum {
    napi_int8_array, napi_uint8_array, napi_uint8_clamped_array, napi_int16_array, napi_uint16_array, napi_int32_array,
    napi_uint32_array, napi_float32_array, napi_float64_array, napi_bigint64_array, napi_biguint64_array
} napi_typedarray_type;
typedef napi_value (*napi_callback)(napi_env env, napi_callback_info info);
typedef void (*napi_finalize)(napi_env env, void* finalize_data, void* finalize_hint);
#define NAPI_AUTO_LENGTH SIZE_MAX

napi_status napi_get_cb_info(napi_env env, napi_callback_info cbinfo, size_t* argc, napi_value* argv, napi_value* this_arg, void** data);
napi_status napi_get_buffer_info(napi_env env, napi_value value, void** data, size_t* length);
napi_status napi_get_typedarray_info(napi_env env, napi_value typedarray, napi_typedarray_type* type, size_t* length, void** data,
                                     napi_value* arraybuffer, size_t* byte_offset);
napi_status napi_get_value_int32(napi_env env, napi_value value, int32_t* result);
napi_status napi_get_value_int64(napi_env env, napi_value value, int64_t* result);
napi_status napi_get_value_double(napi_env env, napi_value value, double* result);
napi_status napi_get_value_external(napi_env env, napi_value value, void** result);
napi_status napi_create_external(napi_env env, void* data, napi_finalize finalize_cb, void* finalize_hint, napi_value* result);
napi_status napi_create_int32(napi_env env, int32_t value, napi_value* result);
napi_status napi_create_function(napi_env env, const char* utf8name, size_t length, napi_callback cb, void* data, napi_value* result);
napi_status napi_set_named_property(napi_env env, napi_value object, const char* utf8name, napi_value value);
napi_status napi_throw_error(napi_env env, const char* code, const char* msg);
napi_status napi_throw_type_error(napi_env env, const char* code, const char* msg);
napi_status napi_throw_range_error(napi_env env, const char* code, const char* msg);

#ifndef NODE_GYP_MODULE_NAME
#define NODE_GYP_MODULE_NAME jsrt_addon
#endif
/* the real macro registers `regfunc` with the runtime from a static constructor; the mock runtime calls this symbol */
#define NAPI_MODULE(modname, regfunc) napi_value napi_mock_module_init(napi_env env, napi_value exports) { return regfunc(env, exports); }
napi_value napi_mock_module_init(napi_env env, napi_value exports);

#ifdef __cplusplus
}
#endif
#endif
