// host_cells.cpp — TEST INFRASTRUCTURE ONLY.
//
// The product's per-cell bodies (2d-weather-sandbox_b200/csrc/wsb_cells.cuh, wsb_math.cuh and the
// GlobalCtx / GlobalAt contexts of wsb_ref_kernels.cuh) are plain fp32 C++ once the CUDA qualifiers
// are compiled away, so this file builds them for the HOST with g++ (-ffp-contract=off: one
// rounding per operation, like nvcc -fmad=false) and drives them through the REFERENCE schedule
// (one pass = one loop over the grid, exactly the calls the k_ref_* kernels make).  The CPU test
// tests/test_host_cells.py compares the result with the oracle bit for bit: a disagreement that
// only shows on the GPU can then only come from the tile plumbing of the fused kernels, never from
// the cell arithmetic.  Nothing in the product links or loads this file; it is not a CPU path of
// the library (no particles, no fused schedule, no strips).
#include <cuda_runtime.h>  // vector types and make_*; __device__ / __forceinline__ vanish under g++

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#ifndef __CUDACC__
using std::max;
using std::min;
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
#endif

#include "../../2d-weather-sandbox_b200/csrc/wsb_ref_kernels.cuh"

using namespace wsb;

namespace {
struct Field {
  std::vector<float> data[4];
  Planes4 p;
  void alloc(size_t n) {
    for (int k = 0; k < 4; k++) { data[k].assign(n, 0.0f); p.c[k] = data[k].data(); }
  }
};
struct Host {
  int W, H;
  Geom g;
  DevParams dp;
  Field base[2], water[2], light[2];
  std::vector<int> wall[2];
  std::vector<float> curl;
  std::vector<float2> vort, dep;
  std::vector<float4> fb;
  std::vector<float> initial_T, sndT, sndW, sndV;
  bool even = true;
  long long iter = 0;
};
GlobalCtx ctx(Host& s, int b, int w, int wl, int l) {
  GlobalCtx c;
  c.base = s.base[b].p; c.water = s.water[w].p; c.wall = s.wall[wl].data(); c.vortf = s.vort.data(); c.light = s.light[l].p;
  c.fb = s.fb.data(); c.dep = s.dep.data(); c.g = s.g;
  return c;
}
void derived(Host& s) {
  s.dp.sinSun = (float)sin((double)s.dp.in.sunAngle);
  s.dp.cosSun = (float)cos((double)s.dp.in.sunAngle);
  s.dp.iterNum = (float)s.iter;
  s.dp.iterI = (int)s.dp.iterNum;
}
}  // namespace

extern "C" {
void* hc_create(int W, int H) {
  Host* s = new Host();
  s->W = W; s->H = H;
  const size_t n = (size_t)W * H;
  for (int k = 0; k < 2; k++) { s->base[k].alloc(n); s->water[k].alloc(n); s->light[k].alloc(n); s->wall[k].assign(n, 0); }
  s->curl.assign(n, 0.0f); s->vort.assign(n, make_float2(0.f, 0.f)); s->dep.assign(n, make_float2(0.f, 0.f));
  s->fb.assign(n, make_float4(0.f, 0.f, 0.f, 0.f));
  s->initial_T.assign(H + 2, 0.0f); s->sndT.assign(H + 2, 0.0f); s->sndW.assign(H + 2, 0.0f); s->sndV.assign(H + 2, 0.0f);
  Geom& g = s->g;  // as wsb_create (csrc/wsb200.cu), single domain
  g.Wg = W; g.H = H; g.pitch = W; g.gx0 = 0; g.wrap = 1; g.cx0 = 0; g.cx1 = W; g.cxGapAt = 0x7fffffff; g.cxGapLen = 0; g.ox0 = 0; g.ox1 = W;
  g.texelX = (float)(1.0 / (double)W); g.texelY = (float)(1.0 / (double)H);
  g.Wf = (float)W; g.Hf = (float)H;
  g.ltexelX = 1.0f / g.Wf; g.ltexelY = 1.0f / g.Hf;
  g.cellHeightComp = 300.0f / g.Hf;
  g.nearV = 0.9f;
  memset(&s->dp, 0, sizeof(s->dp));
  s->dp.in.userInputType = -1;
  derived(*s);
  return s;
}
void hc_destroy(void* h) { delete (Host*)h; }
// texels as the C ABI speaks them: base / water float4, wall char4; both ping-pong copies are filled (wsb_upload)
void hc_upload(void* h, const float* base, const float* water, const int8_t* wall) {
  Host& s = *(Host*)h;
  const size_t n = (size_t)s.W * s.H;
  for (int k = 0; k < 2; k++) {
    for (size_t i = 0; i < n; i++) {
      for (int ch = 0; ch < 4; ch++) { s.base[k].data[ch][i] = base[i * 4 + ch]; s.water[k].data[ch][i] = water[i * 4 + ch]; s.light[k].data[ch][i] = 0.0f; }
      int w; memcpy(&w, wall + i * 4, 4); s.wall[k][i] = w;
    }
  }
  std::fill(s.fb.begin(), s.fb.end(), make_float4(0.f, 0.f, 0.f, 0.f));
  std::fill(s.dep.begin(), s.dep.end(), make_float2(0.f, 0.f));
  s.even = true; s.iter = 0;
  derived(s);
}
void hc_set_params(void* h, const wsb_params* p) { ((Host*)h)->dp.p = *p; }
void hc_set_frame_inputs(void* h, const wsb_frame_inputs* in) { Host& s = *(Host*)h; s.dp.in = *in; derived(s); }
void hc_set_profiles(void* h, const float* t0, const float* st, const float* sw, const float* sv) {
  Host& s = *(Host*)h;
  const size_t n = (size_t)s.H + 1;
  if (t0) memcpy(s.initial_T.data(), t0, n * 4);
  if (st) memcpy(s.sndT.data(), st, n * 4);
  if (sw) memcpy(s.sndW.data(), sw, n * 4);
  if (sv) memcpy(s.sndV.data(), sv, n * 4);
}
void hc_set_iter(void* h, long long it) { Host& s = *(Host*)h; s.iter = it; derived(s); }
void hc_set_feedback(void* h, const float* fb, const float* dep) {
  Host& s = *(Host*)h;
  memcpy(s.fb.data(), fb, s.fb.size() * sizeof(float4));
  memcpy(s.dep.data(), dep, s.dep.size() * sizeof(float2));
}
// one pass of the REFERENCE schedule, WSB_PASS_* numbering (0 velocity .. 6 lighting; 8 = iter++)
void hc_run_pass(void* h, int pass) {
  Host& s = *(Host*)h;
  const Geom& g = s.g;
  derived(s);
  const int W = s.W, H = s.H;
  switch (pass) {
    case 0: {  // k_ref_velocity: frameBuff_0 -> frameBuff_1 (base, wall)
      GlobalCtx c = ctx(s, 0, 0, 0, 0);
      for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) {
        const size_t ci = (size_t)y * W + x;
        float4 b = c.base.ld(ci);
        const int w = c.wall[ci];
        velocity_cell(s.dp, b.x, b.y, b.z, c.bp(x + 1, y), c.bp(x, y + 1), as_char4(w).y);
        s.base[1].p.st(ci, b);
        s.wall[1][ci] = w;
      }
      break;
    }
    case 1: {  // k_ref_curl
      GlobalCtx c = ctx(s, 1, 1, 1, 0);
      for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) {
        const size_t ci = (size_t)y * W + x;
        s.curl[ci] = curl_cell(c.base.c[0][ci], c.base.c[1][ci], c.bx(x, y + 1), c.by(x + 1, y));
      }
      break;
    }
    case 2: {  // k_ref_vorticity
      const float* curl = s.curl.data();
      auto at = [&](int xx, int yy) { return curl[(size_t)wrap_y(yy, g.H) * g.pitch + wrap_x(g, xx)]; };
      for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) {
        const size_t ci = (size_t)y * W + x;
        s.vort[ci] = vorticity_cell(curl[ci], at(x - 1, y), at(x, y - 1), at(x + 1, y), at(x, y + 1));
      }
      break;
    }
    case 3: {  // k_ref_boundary: frameBuff_1 + light_0 + feedback -> frameBuff_0
      GlobalCtx c = ctx(s, 1, 1, 1, 0);
      for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) {
        const size_t ci = (size_t)y * W + x;
        float4 b, w; char4 wl;
        boundary_cell(GlobalAt{c, x, y}, g, s.dp, s.initial_T.data(), x, y, b, w, wl);
        s.base[0].p.st(ci, b); s.water[0].p.st(ci, w); s.wall[0][ci] = as_int(wl);
      }
      break;
    }
    case 4: {  // k_ref_advection<false>: frameBuff_0 -> frameBuff_1
      GlobalCtx c = ctx(s, 0, 0, 0, 0);
      for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) {
        const size_t ci = (size_t)y * W + x;
        float4 b, w; char4 wl; float vm = 0.0f;
        advection_cell<false>(c, g, s.dp, s.initial_T.data(), s.sndT.data(), s.sndW.data(), s.sndV.data(), x, y, b, w, wl, vm);
        s.base[1].p.st(ci, b); s.water[1].p.st(ci, w); s.wall[1][ci] = as_int(wl);
      }
      break;
    }
    case 5: {  // k_ref_pressure: frameBuff_1 -> frameBuff_0 (base, wall)
      GlobalCtx c = ctx(s, 1, 1, 1, 0);
      for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) {
        const size_t ci = (size_t)y * W + x;
        float4 b = c.base.ld(ci);
        char4 wYm = c.wall4(x, y - 1);
        pressure_cell(b.x, b.y, b.z, b.w, c.bx(x - 1, y), c.by(x, y - 1), c.bt(x, y - 1), wYm.x, wYm.y);
        s.base[0].p.st(ci, b); s.wall[0][ci] = c.wall[ci];
      }
      break;
    }
    case 6: {  // k_ref_lighting: frameBuff_1 + light_src -> light_dst (app.js:5912-5926)
      const int src = s.even ? 0 : 1, dst = s.even ? 1 : 0;
      GlobalCtx c = ctx(s, 1, 1, 1, src);
      for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) {
        const size_t ci = (size_t)y * W + x;
        s.light[dst].p.st(ci, lighting_cell(c, g, s.dp, x, y, (float)global_x(g, x) + 0.5f, c.base.c[3][ci], c.water.ld(ci),
                                            as_char4(c.wall[ci]), c.bt(x, y - 1)));
      }
      s.even = !s.even;
      break;
    }
    case 8: s.iter++; break;
    default: break;
  }
}
// field: 0 base, 1 water, 2 wall (as int32 texels), 3 light; dst receives packed texels
void hc_read(void* h, int field, int buf, void* dst) {
  Host& s = *(Host*)h;
  const size_t n = (size_t)s.W * s.H;
  if (field == 2) { memcpy(dst, s.wall[buf].data(), n * 4); return; }
  Field& f = field == 0 ? s.base[buf] : field == 1 ? s.water[buf] : s.light[buf];
  float* o = (float*)dst;
  for (size_t i = 0; i < n; i++) for (int ch = 0; ch < 4; ch++) o[i * 4 + ch] = f.data[ch][i];
}
}  // extern "C"
