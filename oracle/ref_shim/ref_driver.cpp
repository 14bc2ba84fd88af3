// ref_driver.cpp — runs the REFERENCE'S OWN simulation shaders on the CPU.  TEST INFRASTRUCTURE ONLY.
//
// The twelve gen_*.h headers included below are produced at build time by translate.py from the GLSL files
// where they lie in the reference checkout (shaders/common.glsl, vertex/simShader.vert, fragment/{velocity,
// curl,vorticity,boundary,advection,pressure,lighting,lightningLocation,precipitation,setup}Shader.frag,
// vertex/precipitationShader.vert) and compiled against glsl_shim.h; they are intermediates of the build (a temporary directory; only the .so is kept, under oracle/_ref/)
// and never enter the repository.  This file is the part of the reference that is JavaScript and therefore has
// to be restated: the GL objects of app.js:5118-5317 (textures and their sampling state), the sampler -> texture
// unit assignments of app.js:5486-5640, the draw loop of app.js:5830-6005 (which texture is bound to which unit
// and which framebuffer receives which outputs; attachments: frameBuff_N = base_N / water_N / wall_N app.js:5244-5252,
// lightFrameBuff_N 5282-5294, feedback + deposition 5308-5309, lightning 5317), transform feedback + additive point sprites of the particle
// pass, and the fixed-function steps GL performs around a shader (varying interpolation at pixel centres,
// RGBA8I saturation, point rasterisation, blending) in the canonical forms of DESIGN.md "Spec freeze".
//
// Used by tests/test_ref_shaders.py (here, where /root/reference exists) to hold oracle/wsb_oracle.cpp to the
// reference's own shader code bit for bit, and by tests/golden/make_ref_shader_golden.py to generate the golden
// vectors the GPU box checks the CUDA path against.
#include <omp.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "glsl_shim.h"
#include "gen_simVert.h"
#include "gen_velocity.h"
#include "gen_curl.h"
#include "gen_vorticity.h"
#include "gen_boundary.h"
#include "gen_advection.h"
#include "gen_pressure.h"
#include "gen_lighting.h"
#include "gen_precipVert.h"
#include "gen_precipFrag.h"
#include "gen_lightningLocation.h"
#include "gen_setup.h"

namespace {
using namespace glsl;

struct Params {  // same field order as wsb_params (include/wsb200.h); every field is a uniform of the same name
  float dragMultiplier, wind, vorticity, landEvaporation, waterEvaporation, dynamicWaterTemperature,
      evapHeat, waterWeight, meltingHeat, condensationRate, globalDrying, globalHeating,
      soundingForcing, globalEffectsStartAlt, globalEffectsEndAlt, waterTemperature,
      greenhouseGases, waterGreenHouseEffect, IR_rate, dryLapse, aboveZeroThreshold,
      subZeroThreshold, spawnChanceMult, snowDensity, fallSpeed, growthRate0C, growthRate_30C,
      freezingRate, meltingRate, evapRate;
  int32_t enablePrecipitation, reserved;
};
struct FrameInputs {  // same field order as wsb_frame_inputs
  float sunAngle, sunIntensity;
  float userInputValues[4];
  float userInputMove[2];
  int32_t userInputType, wrapHorizontally;
  float airplaneValues[4];
};

// every non-sampler uniform the simulation programs declare, under the name the shaders use
// (gl.getUniformLocation(program, name), app.js:5478-5640, 3401-3443, 5804-5808, 6557-6561)
struct UniformBag {
  vec2 texelSize, resolution;
  float dragMultiplier, wind, vorticity, landEvaporation, waterEvaporation, dynamicWaterTemperature, evapHeat,
      waterWeight, meltingHeat, condensationRate, globalDrying, globalHeating, soundingForcing,
      globalEffectsStartAlt, globalEffectsEndAlt, waterTemperature, greenhouseGases, waterGreenHouseEffect,
      IR_rate, dryLapse, aboveZeroThreshold, subZeroThreshold, spawnChanceMult, snowDensity, fallSpeed,
      growthRate0C, growthRate_30C, freezingRate, meltingRate, evapRate;
  float sunAngle, sunIntensity, iterNum, numDroplets, inactiveDroplets;
  vec4 userInputValues, airplaneValues;
  vec2 userInputMove;
  int userInputType;
  bool wrapHorizontally;
  vec4 initial_Tv[WSB_REF_PROFILE_VEC4S], realWorldSounding_Tv[WSB_REF_PROFILE_VEC4S], realWorldSounding_Wv[WSB_REF_PROFILE_VEC4S],
      realWorldSounding_Velv[WSB_REF_PROFILE_VEC4S];
  float simHeight, seed, heightMult;  // setupShader only
};

struct Sim {
  int w, h, nd;
  std::vector<float> base[2], water[2], light[2], fb, dep, curl, vort, drops[2], lightning;
  std::vector<int8_t> wall[2];
  // texture objects (app.js:5189-5317): storage + sampling state
  Texture tBase[2], tWater[2], tWall[2], tLight[2], tFb, tDep, tCurl, tVort, tLightning;
  const Texture* unit[16];  // gl.activeTexture / gl.bindTexture
  float initial_T[4 * WSB_REF_PROFILE_VEC4S], snd_T[4 * WSB_REF_PROFILE_VEC4S], snd_W[4 * WSB_REF_PROFILE_VEC4S], snd_Vel[4 * WSB_REF_PROFILE_VEC4S];
  float inactiveDroplets;  // the uniform, re-sent every 600 iterations (app.js:5957-5967)
  long iter;
  bool even;
  int last_drops;
  Params p;
  FrameInputs in;
};

void make_textures(Sim& s) {
  auto f32 = [&](Texture& t, std::vector<float>& v, int w, int h, int c, int filter, int wrapT) {
    t.f = v.data(); t.W = w; t.H = h; t.C = c; t.filter = filter; t.wrapS = GL_REPEAT; t.wrapT = wrapT;
  };
  for (int k = 0; k < 2; k++) {
    f32(s.tBase[k], s.base[k], s.w, s.h, 4, GL_NEAREST, GL_REPEAT);   // app.js:5191-5202 (WRAP_T line commented out)
    f32(s.tWater[k], s.water[k], s.w, s.h, 4, GL_NEAREST, GL_REPEAT); // app.js:5205-5216
    s.tWall[k].i8 = s.wall[k].data(); s.tWall[k].W = s.w; s.tWall[k].H = s.h;  // RGBA8I, NEAREST (app.js:5219-5230)
    f32(s.tLight[k], s.light[k], s.w, s.h, 4, GL_LINEAR, GL_CLAMP_TO_EDGE);  // app.js:5273-5290
  }
  f32(s.tCurl, s.curl, s.w, s.h, 1, GL_NEAREST, GL_REPEAT);      // R32F  app.js:5255-5258
  f32(s.tVort, s.vort, s.w, s.h, 2, GL_NEAREST, GL_REPEAT);      // RG32F app.js:5265-5268
  f32(s.tFb, s.fb, s.w, s.h, 4, GL_NEAREST, GL_REPEAT);          // app.js:5297-5300
  f32(s.tDep, s.dep, s.w, s.h, 2, GL_NEAREST, GL_REPEAT);        // app.js:5302-5305
  f32(s.tLightning, s.lightning, 1, 1, 4, GL_NEAREST, GL_REPEAT);  // app.js:5311-5314
}

UniformBag make_bag(const Sim& s) {
  UniformBag u;
  const Params& p = s.p;
  u.texelSize = vec2((float)(1.0 / (double)s.w), (float)(1.0 / (double)s.h));  // app.js:5436-5437 (JS doubles -> uniform2f)
  u.resolution = vec2((float)s.w, (float)s.h);
#define CP(n) u.n = p.n
  CP(dragMultiplier); CP(wind); CP(vorticity); CP(landEvaporation); CP(waterEvaporation); CP(dynamicWaterTemperature);
  CP(evapHeat); CP(waterWeight); CP(meltingHeat); CP(condensationRate); CP(globalDrying); CP(globalHeating);
  CP(soundingForcing); CP(globalEffectsStartAlt); CP(globalEffectsEndAlt); CP(waterTemperature); CP(greenhouseGases);
  CP(waterGreenHouseEffect); CP(IR_rate); CP(dryLapse); CP(aboveZeroThreshold); CP(subZeroThreshold);
  CP(spawnChanceMult); CP(snowDensity); CP(fallSpeed); CP(growthRate0C); CP(growthRate_30C); CP(freezingRate);
  CP(meltingRate); CP(evapRate);
#undef CP
  u.sunAngle = s.in.sunAngle; u.sunIntensity = s.in.sunIntensity;
  u.iterNum = (float)s.iter; u.numDroplets = (float)s.nd; u.inactiveDroplets = s.inactiveDroplets;
  u.userInputValues = vec4(s.in.userInputValues[0], s.in.userInputValues[1], s.in.userInputValues[2], s.in.userInputValues[3]);
  u.airplaneValues = vec4(s.in.airplaneValues[0], s.in.airplaneValues[1], s.in.airplaneValues[2], s.in.airplaneValues[3]);
  u.userInputMove = vec2(s.in.userInputMove[0], s.in.userInputMove[1]);
  u.userInputType = s.in.userInputType;
  u.wrapHorizontally = s.in.wrapHorizontally != 0;
  for (int i = 0; i < WSB_REF_PROFILE_VEC4S; i++) {  // gl.uniform4fv(..., Float32Array(504)) app.js:5476, 5492-5494
    u.initial_Tv[i] = vec4(s.initial_T[4 * i], s.initial_T[4 * i + 1], s.initial_T[4 * i + 2], s.initial_T[4 * i + 3]);
    u.realWorldSounding_Tv[i] = vec4(s.snd_T[4 * i], s.snd_T[4 * i + 1], s.snd_T[4 * i + 2], s.snd_T[4 * i + 3]);
    u.realWorldSounding_Wv[i] = vec4(s.snd_W[4 * i], s.snd_W[4 * i + 1], s.snd_W[4 * i + 2], s.snd_W[4 * i + 3]);
    u.realWorldSounding_Velv[i] = vec4(s.snd_Vel[4 * i], s.snd_Vel[4 * i + 1], s.snd_Vel[4 * i + 2], s.snd_Vel[4 * i + 3]);
  }
  u.simHeight = 0.0f; u.seed = 0.0f; u.heightMult = 0.0f;
  return u;
}

inline int8_t sat8(int v) { return (int8_t)(v < -128 ? -128 : (v > 127 ? 127 : v)); }  // RGBA8I store saturates
inline void st4(std::vector<float>& a, size_t i, const vec4& v) { memcpy(&a[i * 4], v.v, 16); }
inline void stw(std::vector<int8_t>& a, size_t i, const ivec4& v) { for (int c = 0; c < 4; c++) a[i * 4 + c] = sat8(v.v[c]); }

// gl.drawArrays(TRIANGLE_STRIP, 0, 4) of the screen-filling quad over a w x h framebuffer: the vertex shader's
// outputs are interpolated to pixel centres; canonical = simShader.vert evaluated AT the pixel centre, where the
// interpolated attribute vertTexCoord is exactly (x + 0.5, y + 0.5) (app.js:4766-4787).
template <class FS, class Write>
void draw_quad(int w, int h, Write&& write) {
#pragma omp parallel for schedule(static)
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      ref_simVert::Shader vs;
      vs.vertTexCoord = vec2((float)x + 0.5f, (float)y + 0.5f);
      vs.main();
      FS fs;
      fs.set_varyings(vs);
      fs.main();
      if (!fs.glsl_discarded) write((size_t)y * w + x, fs);
    }
}

#define BIND(u, tex) s.unit[u] = &(tex)  // gl.activeTexture(gl.TEXTURE0 + u); gl.bindTexture(gl.TEXTURE_2D, tex)

void pass_velocity(Sim& s, const UniformBag& u) {  // app.js:5832-5840
  typedef ref_velocity::Shader P;
  BIND(0, s.tBase[0]); BIND(1, s.tWall[0]);
  P::baseTex.t = s.unit[0]; P::wallTex.t = s.unit[1];  // uniform1i app.js:5506-5507
  P::bind_uniforms(u);
  draw_quad<P>(s.w, s.h, [&](size_t i, const P& f) { st4(s.base[1], i, f.base); stw(s.wall[1], i, f.wall); });  // frameBuff_1, attachments 0 and 2
}
void pass_curl(Sim& s, const UniformBag& u) {  // app.js:5843-5848
  typedef ref_curl::Shader P;
  BIND(0, s.tBase[1]);
  P::baseTex.t = s.unit[0];  // app.js:5535
  P::bind_uniforms(u);
  draw_quad<P>(s.w, s.h, [&](size_t i, const P& f) { s.curl[i] = f.curl; });
}
void pass_vorticity(Sim& s, const UniformBag& u) {  // app.js:5851-5856
  typedef ref_vorticity::Shader P;
  BIND(0, s.tCurl);
  P::curlTex.t = s.unit[0];  // app.js:5514
  P::bind_uniforms(u);
  draw_quad<P>(s.w, s.h, [&](size_t i, const P& f) { s.vort[2 * i] = f.vortForce.x; s.vort[2 * i + 1] = f.vortForce.y; });
}
void pass_boundary(Sim& s, const UniformBag& u) {  // app.js:5859-5880
  typedef ref_boundary::Shader P;
  BIND(0, s.tBase[1]); BIND(1, s.tWater[1]); BIND(2, s.tVort); BIND(3, s.tWall[1]); BIND(4, s.tLight[0]); BIND(5, s.tFb); BIND(6, s.tDep);
  P::baseTex.t = s.unit[0]; P::waterTex.t = s.unit[1]; P::vortForceTex.t = s.unit[2]; P::wallTex.t = s.unit[3];  // app.js:5517-5523
  P::lightTex.t = s.unit[4]; P::precipFeedbackTex.t = s.unit[5]; P::precipDepositionTex.t = s.unit[6];
  P::bind_uniforms(u);
  draw_quad<P>(s.w, s.h, [&](size_t i, const P& f) { st4(s.base[0], i, f.base); st4(s.water[0], i, f.water); stw(s.wall[0], i, f.wall); });
}
void pass_advection(Sim& s, const UniformBag& u) {  // app.js:5883-5892
  typedef ref_advection::Shader P;
  BIND(0, s.tBase[0]); BIND(1, s.tWater[0]); BIND(2, s.tWall[0]);
  P::baseTex.t = s.unit[0]; P::waterTex.t = s.unit[1]; P::wallTex.t = s.unit[2];  // app.js:5482-5484
  P::bind_uniforms(u);
  draw_quad<P>(s.w, s.h, [&](size_t i, const P& f) { st4(s.base[1], i, f.base); st4(s.water[1], i, f.water); stw(s.wall[1], i, f.wall); });
}
void pass_pressure(Sim& s, const UniformBag& u) {  // app.js:5895-5903
  typedef ref_pressure::Shader P;
  BIND(0, s.tBase[1]); BIND(1, s.tWall[1]);
  P::baseTex.t = s.unit[0]; P::wallTex.t = s.unit[1];  // app.js:5500-5501
  P::bind_uniforms(u);
  draw_quad<P>(s.w, s.h, [&](size_t i, const P& f) { st4(s.base[0], i, f.base); stw(s.wall[0], i, f.wall); });
}
void pass_lighting(Sim& s, const UniformBag& u) {  // app.js:5906-5930
  typedef ref_lighting::Shader P;
  BIND(0, s.tBase[1]); BIND(1, s.tWater[1]); BIND(2, s.tWall[1]);
  const int src = s.even ? 0 : 1, dst = s.even ? 1 : 0;
  BIND(3, s.tLight[src]);
  s.even = !s.even;
  P::baseTex.t = s.unit[0]; P::waterTex.t = s.unit[1]; P::wallTex.t = s.unit[2]; P::lightTex.t = s.unit[3];  // app.js:5541-5544
  P::bind_uniforms(u);
  draw_quad<P>(s.w, s.h, [&](size_t i, const P& f) { st4(s.light[dst], i, f.light); });  // attachment 1 (reflectedLight) is display-only
}

// Point rasterisation + blend ONE, ONE (app.js:5938-5954), canonical rule of DESIGN.md 2: window centre
// ((p + 1) / 2 * res); a pixel is covered when its centre lies in [c - size/2, c + size/2); clipped to the
// viewport, never wrapped; a point whose centre is outside the clip volume is discarded.
void draw_point(Sim& s, const ref_precipVert::Shader& vs) {
  const float posx = vs.gl_Position.x, posy = vs.gl_Position.y, size = vs.gl_PointSize;
  if (!(posx >= -1.0f && posx <= 1.0f && posy >= -1.0f && posy <= 1.0f)) return;
  const float xw = (posx + 1.0f) * 0.5f * (float)s.w, yw = (posy + 1.0f) * 0.5f * (float)s.h;
  const float half = size * 0.5f;
  const int xs = (int)ceilf(xw - half - 0.5f), ys = (int)ceilf(yw - half - 0.5f), n = (int)size;
  ref_precipFrag::Shader fs;  // flat varyings of a point: one evaluation serves every covered pixel
  fs.set_varyings(vs);
  fs.main();
  for (int j = ys; j < ys + n; j++) {
    if (j < 0 || j >= s.h) continue;
    for (int i = xs; i < xs + n; i++) {
      if (i < 0 || i >= s.w) continue;
      const size_t c = (size_t)j * s.w + i;
      for (int k = 0; k < 4; k++) s.fb[c * 4 + k] += fs.feedbackOut.v[k];
      s.dep[c * 2] += fs.depositionOut.x;
      s.dep[c * 2 + 1] += fs.depositionOut.y;
    }
  }
}

void pass_precipitation(Sim& s, const UniformBag& u) {  // app.js:5932-5984
  // srcVAO / destTF were chosen before `even` was toggled (app.js:5912-5927)
  const int src = s.even ? 1 : 0, dst = s.even ? 0 : 1;
  std::fill(s.fb.begin(), s.fb.end(), 0.0f);  // gl.clear of precipitationFeedbackFrameBuff (both attachments)
  std::fill(s.dep.begin(), s.dep.end(), 0.0f);
  if (!s.p.enablePrecipitation || s.nd == 0) return;
  typedef ref_precipVert::Shader V;
  BIND(0, s.tBase[1]); BIND(1, s.tWater[1]); BIND(2, s.tLightning);
  V::baseTex.t = s.unit[0]; V::waterTex.t = s.unit[1]; V::lightningDataTex.t = s.unit[2];  // app.js:5609-5611
  V::bind_uniforms(u);
  for (int n = 0; n < s.nd; n++) {  // gl.drawArrays(gl.POINTS, 0, NUM_DROPLETS) with transform feedback
    const float* d = &s.drops[src][(size_t)n * 5];  // VBO layout app.js:4901-4913: position, mass, density
    V vs;
    vs.dropPosition = vec2(d[0], d[1]);
    vs.mass = vec2(d[2], d[3]);
    vs.density = d[4];
    vs.main();
    float* o = &s.drops[dst][(size_t)n * 5];
    o[0] = vs.position_out.x; o[1] = vs.position_out.y; o[2] = vs.mass_out.x; o[3] = vs.mass_out.y; o[4] = vs.density_out;
    draw_point(s, vs);
  }
  s.last_drops = dst;
  if (s.iter % 600 == 0) s.inactiveDroplets = s.fb[0];  // readPixels(0,0,1,1) -> uniform inactiveDroplets (app.js:5957-5967)
  // lightningLocationProgram into the 1x1 lightningDataFrameBuff (app.js:5974-5983): one fragment
  typedef ref_lightningLocation::Shader L;
  BIND(0, s.tFb);
  L::precipFeedbackTex.t = s.unit[0];  // app.js:5630
  L::bind_uniforms(u);
  draw_quad<L>(1, 1, [&](size_t, const L& f) { memcpy(s.lightning.data(), f.lightningLocation.v, 16); });
}

void iteration(Sim& s) {  // app.js:5830-6005
  const UniformBag u = make_bag(s);
  ref_simVert::Shader::bind_uniforms(u);  // texelSize of simShader.vert: every program links it (app.js:5480-5640)
  pass_velocity(s, u);
  pass_curl(s, u);
  pass_vorticity(s, u);
  pass_boundary(s, u);
  pass_advection(s, u);
  pass_pressure(s, u);
  pass_lighting(s, u);
  pass_precipitation(s, u);
  s.iter++;
}

}  // namespace

extern "C" {

void* refsim_create(int w, int h, int n_droplets) {
  if (w < 1 || h < 1 || h > WSB_REF_MAX_ROWS) return nullptr;  // uniform vec4 initial_Tv[]: one entry per row + 1
  Sim* s = new Sim();
  s->w = w; s->h = h; s->nd = n_droplets;
  const size_t n = (size_t)w * h;
  for (int k = 0; k < 2; k++) {
    s->base[k].assign(n * 4, 0.0f); s->water[k].assign(n * 4, 0.0f); s->light[k].assign(n * 4, 0.0f);
    s->wall[k].assign(n * 4, 0); s->drops[k].assign((size_t)n_droplets * 5 + 1, 0.0f);
  }
  s->fb.assign(n * 4, 0.0f); s->dep.assign(n * 2, 0.0f); s->curl.assign(n, 0.0f); s->vort.assign(n * 2, 0.0f);
  s->lightning.assign(4, 0.0f);
  memset(s->initial_T, 0, sizeof(s->initial_T)); memset(s->snd_T, 0, sizeof(s->snd_T));
  memset(s->snd_W, 0, sizeof(s->snd_W)); memset(s->snd_Vel, 0, sizeof(s->snd_Vel));
  s->inactiveDroplets = 0.0f; s->iter = 0; s->even = true; s->last_drops = 0;
  memset(&s->p, 0, sizeof(s->p)); memset(&s->in, 0, sizeof(s->in));
  s->in.userInputType = -1;
  make_textures(*s);
  return s;
}
void refsim_destroy(void* h) { delete (Sim*)h; }

// app.js:5189-5234, 4917-4962: both ping-pong copies get the same data; light / feedback start at zero
void refsim_upload(void* h, const float* base, const float* water, const int8_t* wall, const float* drops) {
  Sim& s = *(Sim*)h;
  const size_t n = (size_t)s.w * s.h;
  for (int k = 0; k < 2; k++) {
    memcpy(s.base[k].data(), base, n * 16); memcpy(s.water[k].data(), water, n * 16); memcpy(s.wall[k].data(), wall, n * 4);
    if (drops && s.nd) memcpy(s.drops[k].data(), drops, (size_t)s.nd * 20);
    std::fill(s.light[k].begin(), s.light[k].end(), 0.0f);
  }
  std::fill(s.fb.begin(), s.fb.end(), 0.0f); std::fill(s.dep.begin(), s.dep.end(), 0.0f);
  std::fill(s.curl.begin(), s.curl.end(), 0.0f); std::fill(s.vort.begin(), s.vort.end(), 0.0f);
  std::fill(s.lightning.begin(), s.lightning.end(), 0.0f);
  s.inactiveDroplets = 0.0f; s.iter = 0; s.even = true; s.last_drops = 0;
}
void refsim_set_params(void* h, const void* p) { memcpy(&((Sim*)h)->p, p, sizeof(Params)); }
void refsim_set_frame_inputs(void* h, const void* in) { memcpy(&((Sim*)h)->in, in, sizeof(FrameInputs)); }
void refsim_set_profiles(void* h, const float* T0, const float* sT, const float* sW, const float* sV) {
  Sim& s = *(Sim*)h;
  const size_t n = (size_t)s.h + 1;
  auto put = [&](float* dst, const float* src) { memset(dst, 0, sizeof(s.initial_T)); if (src) memcpy(dst, src, n * 4); };
  if (T0) put(s.initial_T, T0);
  put(s.snd_T, sT); put(s.snd_W, sW); put(s.snd_Vel, sV);
}
void refsim_step(void* h, int n) { Sim& s = *(Sim*)h; for (int i = 0; i < n; i++) iteration(s); }
void refsim_run_pass(void* h, int pass) {  // pass ids as WSB_PASS_* (include/wsb200.h)
  Sim& s = *(Sim*)h;
  const UniformBag u = make_bag(s);
  ref_simVert::Shader::bind_uniforms(u);
  switch (pass) {
    case 0: pass_velocity(s, u); break;
    case 1: pass_curl(s, u); break;
    case 2: pass_vorticity(s, u); break;
    case 3: pass_boundary(s, u); break;
    case 4: pass_advection(s, u); break;
    case 5: pass_pressure(s, u); break;
    case 6: pass_lighting(s, u); break;
    case 7: pass_precipitation(s, u); break;
    case 8: s.iter++; break;
  }
}
float* refsim_field_f32(void* h, int field, int buf) {  // field ids as WSB_FIELD_*
  Sim& s = *(Sim*)h;
  switch (field) {
    case 0: return s.base[buf].data();
    case 1: return s.water[buf].data();
    case 3: return s.light[buf].data();
    case 4: return s.fb.data();
    case 5: return s.dep.data();
    case 6: return s.curl.data();
    case 7: return s.vort.data();
    case 8: return s.drops[buf].data();
    case 9: return s.lightning.data();
  }
  return nullptr;
}
int8_t* refsim_field_i8(void* h, int buf) { return ((Sim*)h)->wall[buf].data(); }
long refsim_get_iter(void* h) { return ((Sim*)h)->iter; }
void refsim_set_iter(void* h, long it) { ((Sim*)h)->iter = it; }
int refsim_get_even(void* h) { return ((Sim*)h)->even ? 1 : 0; }
void refsim_set_even(void* h, int e) { ((Sim*)h)->even = e != 0; }
int refsim_last_drops(void* h) { return ((Sim*)h)->last_drops; }
void refsim_set_last_drops(void* h, int b) { ((Sim*)h)->last_drops = b; }
float refsim_get_inactive(void* h) { return ((Sim*)h)->inactiveDroplets; }
void refsim_set_inactive(void* h, float v) { ((Sim*)h)->inactiveDroplets = v; }
void refsim_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }

// setupProgram (app.js:5729-5742): the terrain / initial state generator, one draw into frameBuff_0
void refsim_setup(int w, int h, float seed, float heightMult, float simHeight, float dryLapse, const float* initial_T,
                  float* base, float* water, int8_t* wall) {
  typedef ref_setup::Shader P;
  UniformBag u;
  memset((void*)&u, 0, sizeof(u));
  u.texelSize = vec2((float)(1.0 / (double)w), (float)(1.0 / (double)h));
  u.resolution = vec2((float)w, (float)h);
  u.dryLapse = dryLapse; u.simHeight = simHeight; u.seed = seed; u.heightMult = heightMult;
  for (int i = 0; i < WSB_REF_PROFILE_VEC4S; i++) {
    float t[4];
    for (int k = 0; k < 4; k++) t[k] = (4 * i + k <= h) ? initial_T[4 * i + k] : 0.0f;
    u.initial_Tv[i] = vec4(t[0], t[1], t[2], t[3]);
  }
  P::bind_uniforms(u);
  ref_simVert::Shader::bind_uniforms(u);
  draw_quad<P>(w, h, [&](size_t i, const P& f) {
    memcpy(base + i * 4, f.base.v, 16);
    memcpy(water + i * 4, f.water.v, 16);
    for (int c = 0; c < 4; c++) wall[i * 4 + c] = sat8(f.wall.v[c]);
  });
}

}  // extern "C"
