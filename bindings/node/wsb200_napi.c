/* wsb200_napi.c — thin N-API addon over libwsb200's C ABI (include/wsb200.h).
 *
 * This is the shim the north star asks for: host code stays JavaScript (app.js under Node), the
 * simulation textures and the 8 draw calls per iteration are replaced by calls into libwsb200.so.
 * Every function is a 1:1 forward; typed arrays are passed by pointer (no copies on the JS side),
 * a non-zero status becomes a JS exception carrying wsb_last_error().
 *
 * Build (with a Node toolchain):  cc -shared -fPIC -DWSB_HAVE_NODE_API_H -I$NODE/include/node \
 *        -I../../include wsb200_napi.c -L../../2d-weather-sandbox_b200/csrc -lwsb200 -o wsb200.node
 * This image has no Node: `make -C bindings/node check` compiles it against node_api_min.h.
 */
#ifdef WSB_HAVE_NODE_API_H
#include <node_api.h>
#else
#include "node_api_min.h"
#endif
#include <string.h>

#include "wsb200.h"

#define MAX_ARGS 12
#define THROW(env) do { napi_throw_error((env), "WSB200", wsb_last_error()); return undefined(env); } while (0)

static napi_value undefined(napi_env env) { napi_value u; napi_get_undefined(env, &u); return u; }

static int args(napi_env env, napi_callback_info info, size_t want, napi_value* argv) {
  size_t argc = MAX_ARGS;
  if (napi_get_cb_info(env, info, &argc, argv, NULL, NULL) != napi_ok || argc < want) {
    napi_throw_error(env, "WSB200", "wrong number of arguments");
    return 0;
  }
  return 1;
}
static wsb_sim* sim_of(napi_env env, napi_value v) {
  void* p = NULL;
  if (napi_get_value_external(env, v, &p) != napi_ok || !p) napi_throw_error(env, "WSB200", "not a simulation handle");
  return (wsb_sim*)p;
}
static int32_t i32(napi_env env, napi_value v) { int32_t x = 0; napi_get_value_int32(env, v, &x); return x; }
static double f64(napi_env env, napi_value v) { double x = 0; napi_get_value_double(env, v, &x); return x; }
/* typed array -> pointer (NULL for null/undefined); *len = element count */
static void* ta(napi_env env, napi_value v, size_t* len) {
  napi_valuetype t; void* data = NULL; size_t n = 0; napi_typedarray_type tt;
  napi_typeof(env, v, &t);
  if (t == napi_null || t == napi_undefined) { if (len) *len = 0; return NULL; }
  if (napi_get_typedarray_info(env, v, &tt, &n, &data, NULL, NULL) != napi_ok) napi_throw_error(env, "WSB200", "expected a typed array");
  if (len) *len = n;
  return data;
}

/* create(width, height, nDroplets[, device]) -> handle          app.js:5149-5317, 4885-5002 */
static napi_value js_create(napi_env env, napi_callback_info info) {
  napi_value a[MAX_ARGS];
  if (!args(env, info, 3, a)) return undefined(env);
  wsb_config cfg; memset(&cfg, 0, sizeof cfg);
  cfg.abi_version = WSB_ABI_VERSION;
  cfg.width = i32(env, a[0]); cfg.height = i32(env, a[1]); cfg.n_droplets = i32(env, a[2]);
  cfg.n_ranks = 1; cfg.schedule = WSB_SCHEDULE_FUSED;
  wsb_sim* s = NULL;
  if (wsb_create(&cfg, &s)) THROW(env);
  napi_value h; napi_create_external(env, s, NULL, NULL, &h);
  return h;
}
/* destroy(handle) */
static napi_value js_destroy(napi_env env, napi_callback_info info) {
  napi_value a[MAX_ARGS];
  if (!args(env, info, 1, a)) return undefined(env);
  wsb_destroy(sim_of(env, a[0]));
  return undefined(env);
}
/* upload(handle, Float32Array base, Float32Array water, Int8Array wall, Float32Array droplets)   app.js:5189-5234 */
static napi_value js_upload(napi_env env, napi_callback_info info) {
  napi_value a[MAX_ARGS];
  if (!args(env, info, 5, a)) return undefined(env);
  if (wsb_upload(sim_of(env, a[0]), (const float*)ta(env, a[1], NULL), (const float*)ta(env, a[2], NULL),
                 (const int8_t*)ta(env, a[3], NULL), (const float*)ta(env, a[4], NULL))) THROW(env);
  return undefined(env);
}
/* setParams(handle, Float32Array params[30] in wsb_params order, enablePrecipitation)   app.js:3401-3443 */
static napi_value js_set_params(napi_env env, napi_callback_info info) {
  napi_value a[MAX_ARGS];
  if (!args(env, info, 3, a)) return undefined(env);
  size_t n = 0; const float* v = (const float*)ta(env, a[1], &n);
  if (!v || n < 30) { napi_throw_error(env, "WSB200", "setParams: need 30 floats"); return undefined(env); }
  wsb_params p; memset(&p, 0, sizeof p);
  memcpy(&p, v, 30 * sizeof(float));
  bool on = false; napi_get_value_bool(env, a[2], &on);
  p.enablePrecipitation = on ? 1 : 0;
  if (wsb_set_params(sim_of(env, a[0]), &p)) THROW(env);
  return undefined(env);
}
/* setProfiles(handle, initial_T, sounding_T|null, sounding_W|null, sounding_Vel|null)   app.js:5444-5502 */
static napi_value js_set_profiles(napi_env env, napi_callback_info info) {
  napi_value a[MAX_ARGS];
  if (!args(env, info, 5, a)) return undefined(env);
  if (wsb_set_profiles(sim_of(env, a[0]), (const float*)ta(env, a[1], NULL), (const float*)ta(env, a[2], NULL),
                       (const float*)ta(env, a[3], NULL), (const float*)ta(env, a[4], NULL))) THROW(env);
  return undefined(env);
}
/* setFrameInputs(handle, sunAngle, sunIntensity, userInputValues[4], userInputMove[2], userInputType,
 *                wrapHorizontally, airplaneValues[4])                 app.js:6557-6561, 5804-5808, 3335 */
static napi_value js_set_frame_inputs(napi_env env, napi_callback_info info) {
  napi_value a[MAX_ARGS];
  if (!args(env, info, 8, a)) return undefined(env);
  wsb_frame_inputs in; memset(&in, 0, sizeof in);
  in.sunAngle = (float)f64(env, a[1]); in.sunIntensity = (float)f64(env, a[2]);
  size_t n = 0; const float* v;
  if ((v = (const float*)ta(env, a[3], &n)) && n >= 4) memcpy(in.userInputValues, v, 16);
  if ((v = (const float*)ta(env, a[4], &n)) && n >= 2) memcpy(in.userInputMove, v, 8);
  in.userInputType = i32(env, a[5]);
  bool wrap = false; napi_get_value_bool(env, a[6], &wrap); in.wrapHorizontally = wrap ? 1 : 0;
  if ((v = (const float*)ta(env, a[7], &n)) && n >= 4) memcpy(in.airplaneValues, v, 16);
  if (wsb_set_frame_inputs(sim_of(env, a[0]), &in)) THROW(env);
  return undefined(env);
}
/* step(handle, nIterations) — asynchronous, like the reference's draw calls   app.js:5830-6005 */
static napi_value js_step(napi_env env, napi_callback_info info) {
  napi_value a[MAX_ARGS];
  if (!args(env, info, 2, a)) return undefined(env);
  if (wsb_step(sim_of(env, a[0]), i32(env, a[1]))) THROW(env);
  return undefined(env);
}
/* readRect(handle, field, view, x, y, w, h, TypedArray dst) — gl.readPixels */
static napi_value js_read_rect(napi_env env, napi_callback_info info) {
  napi_value a[MAX_ARGS];
  if (!args(env, info, 8, a)) return undefined(env);
  if (wsb_read_rect(sim_of(env, a[0]), i32(env, a[1]), i32(env, a[2]), i32(env, a[3]), i32(env, a[4]), i32(env, a[5]), i32(env, a[6]),
                    ta(env, a[7], NULL))) THROW(env);
  return undefined(env);
}
/* readDroplets(handle, buffer, first, count, Float32Array dst) — getBufferSubData */
static napi_value js_read_droplets(napi_env env, napi_callback_info info) {
  napi_value a[MAX_ARGS];
  if (!args(env, info, 5, a)) return undefined(env);
  if (wsb_read_droplets(sim_of(env, a[0]), i32(env, a[1]), i32(env, a[2]), i32(env, a[3]), (float*)ta(env, a[4], NULL))) THROW(env);
  return undefined(env);
}
/* getInactiveDroplets(handle) -> number   app.js:5957-5967 */
static napi_value js_get_inactive(napi_env env, napi_callback_info info) {
  napi_value a[MAX_ARGS];
  if (!args(env, info, 1, a)) return undefined(env);
  float v = 0;
  if (wsb_get_inactive_droplets(sim_of(env, a[0]), &v)) THROW(env);
  napi_value r; napi_create_double(env, v, &r);
  return r;
}
/* getLightning(handle, Float32Array out[4])   app.js:5985-5994 */
static napi_value js_get_lightning(napi_env env, napi_callback_info info) {
  napi_value a[MAX_ARGS];
  if (!args(env, info, 2, a)) return undefined(env);
  size_t n = 0; float* out = (float*)ta(env, a[1], &n);
  if (!out || n < 4) { napi_throw_error(env, "WSB200", "getLightning: need Float32Array(4)"); return undefined(env); }
  if (wsb_get_lightning(sim_of(env, a[0]), out)) THROW(env);
  return undefined(env);
}

static napi_value init(napi_env env, napi_value exports) {
  const napi_property_descriptor props[] = {
      {"create", 0, js_create, 0, 0, 0, 0, 0},
      {"destroy", 0, js_destroy, 0, 0, 0, 0, 0},
      {"upload", 0, js_upload, 0, 0, 0, 0, 0},
      {"setParams", 0, js_set_params, 0, 0, 0, 0, 0},
      {"setProfiles", 0, js_set_profiles, 0, 0, 0, 0, 0},
      {"setFrameInputs", 0, js_set_frame_inputs, 0, 0, 0, 0, 0},
      {"step", 0, js_step, 0, 0, 0, 0, 0},
      {"readRect", 0, js_read_rect, 0, 0, 0, 0, 0},
      {"readDroplets", 0, js_read_droplets, 0, 0, 0, 0, 0},
      {"getInactiveDroplets", 0, js_get_inactive, 0, 0, 0, 0, 0},
      {"getLightning", 0, js_get_lightning, 0, 0, 0, 0, 0},
  };
  napi_define_properties(env, exports, sizeof props / sizeof props[0], props);
  return exports;
}

static napi_module wsb_module = {1, 0, __FILE__, init, "wsb200", NULL, {0}};
__attribute__((constructor)) static void wsb_register(void) { napi_module_register(&wsb_module); }
