/* fake_napi_host.c — a stand-in for the Node runtime, just big enough to LOAD and DRIVE the N-API addon
 * (wsb200_napi.so) in an image without Node: it implements the Node-API subset of node_api_min.h over a
 * tagged-value struct, dlopen()s the addon (whose constructor calls napi_module_register, exactly as under
 * Node), runs its init function and then calls the registered JS-facing functions by name:
 *
 *     fake_napi_host <addon.so> --list                      print module name + exported functions (no GPU)
 *     fake_napi_host <addon.so> <input.bin> <output.bin>    create -> upload -> setParams -> setProfiles ->
 *                                                            setFrameInputs -> step -> readRect x2 -> readDroplets ->
 *                                                            getInactiveDroplets -> getLightning -> destroy
 *
 * These are the calls app.js:5149-5317 (allocation / upload), 3401-3443 (uniforms), 5830-6005 (loop) and the
 * gl.readPixels / getBufferSubData sites make once the shim is dropped in.  Test infrastructure:
 * tests/test_napi_addon.py compares <output.bin> with the same run through the ctypes host mirror. */
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "node_api_min.h"

struct napi_value__ {
  napi_valuetype type;
  double num;
  bool b;
  void* ptr;          /* external pointer or typed-array data */
  size_t len;         /* typed-array element count */
  napi_typedarray_type tt;
};
struct napi_callback_info__ { size_t argc; napi_value* argv; };
struct napi_env__ { int pending; char msg[512]; };

#define MAX_VALUES 256
static struct napi_value__ g_values[MAX_VALUES];
static int g_nvalues = 0;
static napi_value new_value(napi_valuetype t) {
  if (g_nvalues == MAX_VALUES) g_nvalues = 16;  /* the first few are long-lived (undefined, handle); recycle the rest */
  napi_value v = &g_values[g_nvalues++];
  memset(v, 0, sizeof *v);
  v->type = t;
  return v;
}
static napi_module* g_module = NULL;
static struct { const char* name; napi_callback fn; } g_props[64];
static int g_nprops = 0;

/* ---- the Node-API subset ---- */
void napi_module_register(napi_module* m) { g_module = m; }
napi_status napi_get_cb_info(napi_env env, napi_callback_info info, size_t* argc, napi_value* argv, napi_value* this_arg, void** data) {
  (void)env; (void)this_arg; (void)data;
  size_t n = info->argc < *argc ? info->argc : *argc;
  for (size_t i = 0; i < n; i++) argv[i] = info->argv[i];
  *argc = info->argc;
  return napi_ok;
}
napi_status napi_typeof(napi_env env, napi_value v, napi_valuetype* t) { (void)env; *t = v->type; return napi_ok; }
napi_status napi_get_value_int32(napi_env env, napi_value v, int32_t* out) { (void)env; if (v->type != napi_number) return (napi_status)1; *out = (int32_t)v->num; return napi_ok; }
napi_status napi_get_value_double(napi_env env, napi_value v, double* out) { (void)env; if (v->type != napi_number) return (napi_status)1; *out = v->num; return napi_ok; }
napi_status napi_get_value_bool(napi_env env, napi_value v, bool* out) { (void)env; if (v->type != napi_boolean) return (napi_status)1; *out = v->b; return napi_ok; }
napi_status napi_get_value_external(napi_env env, napi_value v, void** out) { (void)env; if (v->type != napi_external) return (napi_status)1; *out = v->ptr; return napi_ok; }
napi_status napi_get_typedarray_info(napi_env env, napi_value v, napi_typedarray_type* tt, size_t* length, void** data, napi_value* ab, size_t* off) {
  (void)env;
  if (v->type != napi_object || !v->ptr) return (napi_status)1;
  if (tt) *tt = v->tt;
  if (length) *length = v->len;
  if (data) *data = v->ptr;
  if (ab) *ab = NULL;
  if (off) *off = 0;
  return napi_ok;
}
napi_status napi_create_external(napi_env env, void* data, napi_finalize fin, void* hint, napi_value* result) {
  (void)env; (void)fin; (void)hint;
  *result = new_value(napi_external);
  (*result)->ptr = data;
  return napi_ok;
}
napi_status napi_create_double(napi_env env, double x, napi_value* r) { (void)env; *r = new_value(napi_number); (*r)->num = x; return napi_ok; }
napi_status napi_create_int32(napi_env env, int32_t x, napi_value* r) { (void)env; *r = new_value(napi_number); (*r)->num = x; return napi_ok; }
napi_status napi_get_undefined(napi_env env, napi_value* r) { (void)env; *r = &g_values[0]; return napi_ok; }
napi_status napi_throw_error(napi_env env, const char* code, const char* msg) {
  env->pending = 1;
  snprintf(env->msg, sizeof env->msg, "%s: %s", code ? code : "", msg ? msg : "");
  return napi_ok;
}
napi_status napi_define_properties(napi_env env, napi_value object, size_t count, const napi_property_descriptor* p) {
  (void)env; (void)object;
  for (size_t i = 0; i < count && g_nprops < 64; i++) { g_props[g_nprops].name = p[i].utf8name; g_props[g_nprops].fn = p[i].method; g_nprops++; }
  return napi_ok;
}

/* ---- "JavaScript" side ---- */
static struct napi_env__ g_env;
static napi_value num(double x) { napi_value v = new_value(napi_number); v->num = x; return v; }
static napi_value boolean(bool b) { napi_value v = new_value(napi_boolean); v->b = b; return v; }
static napi_value null_value(void) { return new_value(napi_null); }
static napi_value typed(void* data, size_t n, napi_typedarray_type tt) { napi_value v = new_value(napi_object); v->ptr = data; v->len = n; v->tt = tt; return v; }
static napi_value call(const char* name, size_t argc, napi_value* argv) {
  for (int i = 0; i < g_nprops; i++)
    if (!strcmp(g_props[i].name, name)) {
      struct napi_callback_info__ info = {argc, argv};
      napi_value r = g_props[i].fn(&g_env, &info);
      if (g_env.pending) { fprintf(stderr, "JS exception from %s: %s\n", name, g_env.msg); exit(3); }
      return r;
    }
  fprintf(stderr, "addon does not export %s\n", name);
  exit(2);
}
static void* read_block(FILE* f, size_t bytes) {
  void* p = malloc(bytes ? bytes : 1);
  if (fread(p, 1, bytes, f) != bytes) { fprintf(stderr, "short input file\n"); exit(2); }
  return p;
}

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: %s addon.so --list | addon.so in.bin out.bin\n", argv[0]); return 2; }
  g_values[0].type = napi_undefined;
  g_nvalues = 1;
  void* h = dlopen(argv[1], RTLD_NOW | RTLD_GLOBAL);
  if (!h) { fprintf(stderr, "dlopen(%s): %s\n", argv[1], dlerror()); return 2; }
  if (!g_module) { fprintf(stderr, "the addon did not call napi_module_register\n"); return 2; }
  napi_value exports = new_value(napi_object);
  g_module->nm_register_func(&g_env, exports);
  if (!strcmp(argv[2], "--list")) {
    printf("module %s:", g_module->nm_modname);
    for (int i = 0; i < g_nprops; i++) printf(" %s", g_props[i].name);
    printf("\n");
    return 0;
  }
  if (argc < 4) return 2;
  FILE* f = fopen(argv[2], "rb");
  if (!f) { perror(argv[2]); return 2; }
  int32_t hdr[4];  /* W, H, ND, iterations */
  if (fread(hdr, 4, 4, f) != 4) return 2;
  const size_t W = hdr[0], H = hdr[1], ND = hdr[2], n = W * H;
  float* base = read_block(f, n * 16);
  float* water = read_block(f, n * 16);
  signed char* wall = read_block(f, n * 4);
  float* drops = read_block(f, ND * 20);
  float* params = read_block(f, 30 * 4);
  int32_t* enable = read_block(f, 4);
  float* initial_T = read_block(f, (H + 1) * 4);
  float* fi = read_block(f, 14 * 4);  /* sunAngle, sunIntensity, userInputValues[4], userInputMove[2], type, wrap, airplane[4] */
  fclose(f);

  napi_value a[12];
  a[0] = num((double)W); a[1] = num((double)H); a[2] = num((double)ND);
  napi_value sim = call("create", 3, a);
  g_values[1] = *sim; sim = &g_values[1]; if (g_nvalues < 2) g_nvalues = 2;  /* keep the handle out of the recycled range */
  a[0] = sim; a[1] = typed(base, n * 4, napi_float32_array); a[2] = typed(water, n * 4, napi_float32_array);
  a[3] = typed(wall, n * 4, napi_int8_array); a[4] = ND ? typed(drops, ND * 5, napi_float32_array) : null_value();
  call("upload", 5, a);
  a[0] = sim; a[1] = typed(params, 30, napi_float32_array); a[2] = boolean(*enable != 0);
  call("setParams", 3, a);
  a[0] = sim; a[1] = typed(initial_T, H + 1, napi_float32_array); a[2] = null_value(); a[3] = null_value(); a[4] = null_value();
  call("setProfiles", 5, a);
  a[0] = sim; a[1] = num(fi[0]); a[2] = num(fi[1]); a[3] = typed(fi + 2, 4, napi_float32_array); a[4] = typed(fi + 6, 2, napi_float32_array);
  a[5] = num((double)((int32_t*)fi)[8]); a[6] = boolean(((int32_t*)fi)[9] != 0); a[7] = typed(fi + 10, 4, napi_float32_array);
  call("setFrameInputs", 8, a);
  a[0] = sim; a[1] = num((double)hdr[3]);
  call("step", 2, a);

  float* obase = calloc(n, 16);
  signed char* owall = calloc(n, 4);
  float* odrops = calloc(ND ? ND : 1, 20);
  float olight[4] = {0, 0, 0, 0};
  a[0] = sim; a[1] = num(0 /* WSB_FIELD_BASE */); a[2] = num(0); a[3] = num(0); a[4] = num(0); a[5] = num((double)W); a[6] = num((double)H);
  a[7] = typed(obase, n * 4, napi_float32_array);
  call("readRect", 8, a);
  a[0] = sim; a[1] = num(2 /* WSB_FIELD_WALL */); a[2] = num(0); a[3] = num(0); a[4] = num(0); a[5] = num((double)W); a[6] = num((double)H);
  a[7] = typed(owall, n * 4, napi_int8_array);
  call("readRect", 8, a);
  if (ND) {
    a[0] = sim; a[1] = num(2); a[2] = num(0); a[3] = num((double)ND); a[4] = typed(odrops, ND * 5, napi_float32_array);
    call("readDroplets", 5, a);
  }
  a[0] = sim;
  napi_value inactive = call("getInactiveDroplets", 1, a);
  float oinactive = (float)inactive->num;
  a[0] = sim; a[1] = typed(olight, 4, napi_float32_array);
  call("getLightning", 2, a);
  a[0] = sim;
  call("destroy", 1, a);

  f = fopen(argv[3], "wb");
  if (!f) { perror(argv[3]); return 2; }
  fwrite(obase, 16, n, f); fwrite(owall, 4, n, f); fwrite(odrops, 20, ND, f); fwrite(&oinactive, 4, 1, f); fwrite(olight, 4, 4, f);
  fclose(f);
  printf("ok %zux%zu, %d iterations through the N-API addon\n", W, H, hdr[3]);
  return 0;
}
