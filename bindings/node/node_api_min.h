/* node_api_min.h — the subset of Node-API (N-API, ABI-stable since Node 8) that wsb200_napi.c
 * uses, declared locally because this build image ships no Node headers.  With a real Node
 * toolchain compile against <node_api.h> instead (-DWSB_HAVE_NODE_API_H); the declarations below
 * are the documented C signatures and are only used for the compile check in this repository. */
#ifndef WSB_NODE_API_MIN_H
#define WSB_NODE_API_MIN_H
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

typedef struct napi_env__* napi_env;
typedef struct napi_value__* napi_value;
typedef struct napi_callback_info__* napi_callback_info;
typedef enum { napi_ok = 0 } napi_status;
typedef enum {
  napi_int8_array, napi_uint8_array, napi_uint8_clamped_array, napi_int16_array, napi_uint16_array,
  napi_int32_array, napi_uint32_array, napi_float32_array, napi_float64_array
} napi_typedarray_type;
typedef enum { napi_undefined, napi_null, napi_boolean, napi_number, napi_string, napi_symbol, napi_object,
               napi_function, napi_external, napi_bigint } napi_valuetype;
typedef napi_value (*napi_callback)(napi_env env, napi_callback_info info);
typedef void (*napi_finalize)(napi_env env, void* data, void* hint);
typedef struct {
  const char* utf8name; napi_value name; napi_callback method; napi_callback getter; napi_callback setter;
  napi_value value; int attributes; void* data;
} napi_property_descriptor;
typedef struct napi_module {
  int nm_version; unsigned int nm_flags; const char* nm_filename;
  napi_value (*nm_register_func)(napi_env env, napi_value exports);
  const char* nm_modname; void* nm_priv; void* reserved[4];
} napi_module;

napi_status napi_get_cb_info(napi_env, napi_callback_info, size_t* argc, napi_value* argv, napi_value* this_arg, void** data);
napi_status napi_typeof(napi_env, napi_value, napi_valuetype*);
napi_status napi_get_value_int32(napi_env, napi_value, int32_t*);
napi_status napi_get_value_double(napi_env, napi_value, double*);
napi_status napi_get_value_bool(napi_env, napi_value, bool*);
napi_status napi_get_value_external(napi_env, napi_value, void**);
napi_status napi_get_typedarray_info(napi_env, napi_value, napi_typedarray_type*, size_t* length, void** data, napi_value* arraybuffer, size_t* byte_offset);
napi_status napi_create_external(napi_env, void* data, napi_finalize, void* hint, napi_value* result);
napi_status napi_create_double(napi_env, double, napi_value*);
napi_status napi_create_int32(napi_env, int32_t, napi_value*);
napi_status napi_get_undefined(napi_env, napi_value*);
napi_status napi_throw_error(napi_env, const char* code, const char* msg);
napi_status napi_define_properties(napi_env, napi_value object, size_t count, const napi_property_descriptor*);
void napi_module_register(napi_module*);
#endif
