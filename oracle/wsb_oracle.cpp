// wsb_oracle.cpp — CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
//
// A plain C++ restatement of the reference's per-iteration simulation loop (2D-Weather-Sandbox,
// app.js:5830-6005) and of the GLSL passes it draws.  It exists to check the CUDA path in
// libwsb200.so; nothing in the product may link, import or call it (only tests/, bench.py's
// cpu_baseline / --impl reference leg and __graft_entry__.smoke()).
//
// PARITY: PINNED — to the reference's own shader code, executed, and to reference output.
// The reference ships no tests, golden vectors or expected outputs for this path and its own
// implementation (GLSL ES 3.00 under a browser) cannot run in this environment as it stands (no node / browser / GL).
// (1) oracle/_ref: oracle/ref_shim/ translates the reference's GLSL files mechanically (literal suffixes, globals ->
// struct members, nothing restated) from the checkout where they lie and compiles them for the host against a small
// GLSL-in-C++ header; the draw loop of app.js:5830-6005 is restated beside them.  tests/test_ref_shaders.py: every
// pass of this oracle equals the reference's shader of the same name BIT FOR BIT on all 14 shipped saves; whole runs
// (up to 300 iterations, particles on) on power-of-two crops of shipped saves, every brush / wall tool and the
// airplane, the particle life cycle, lightning, the slow processes at iteration multiples and the setup generator are
// bit-identical; BASELINE config 1 (the 100 x 100 save, 1000 iterations) ends with base / water / wall / droplets
// bit-identical.  The one quantity that is not: SUNLIGHT on grids whose 1 / size is not a power of two, where the
// hardware LINEAR fetch position is rounded in normalised (shader) vs pixel (frozen) coordinates — an
// ulp of the row number, below the 8-bit weights of real hardware.  tests/golden/ref_shader_golden.npz carries
// vectors produced by those shaders to the GPU box (tests/test_ref_shader_golden.py).
// (2) the 14 shipped saves are the reference's WebGL output (frameBuff_0 read back, app.js:6575-6628);
// tests/test_reference_saves.py: their integer wall planes TYPE / DISTANCE / VERT_DISTANCE are a fixed point of one
// more oracle iteration (0 changed bytes; only LAND <-> FIRE flips in burning saves), land-surface vegetation is
// unchanged outside a growth tick, reference-written invariants survive.
// (3) every pass is also reproduced bit for bit by a second, independent Python restatement of its shader
// (tests/test_oracle_numpy_*.py, tests/test_oracle_python_*.py).
// What remains outside: the behaviour of a particular WebGL implementation where GLSL ES leaves it open (precision of
// pow / sin, texture filtering weights, point rasterisation) — frozen canonically, see below.
// Where GLSL/WebGL leaves behaviour implementation-defined the canonical choice is written next to the code (and
// listed in DESIGN.md "Spec freeze").
//
// Arithmetic is fp32 with no contraction (build with -ffp-contract=off) so that the CUDA kernels
// (built with -fmad=false) can reproduce it operation for operation.
//
// Citations: paths are relative to the reference checkout; "common" = shaders/common.glsl.

#include <omp.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {

// ---------------------------------------------------------------------------------------------
// GLSL built-ins, restated with exactly one fp32 rounding per operation.
// ---------------------------------------------------------------------------------------------
inline float gmax(float a, float b) { return a < b ? b : a; }       // GLSL max: (x < y) ? y : x
inline float gmin(float a, float b) { return b < a ? b : a; }       // GLSL min: (y < x) ? y : x
inline float gclamp(float x, float lo, float hi) { return gmin(gmax(x, lo), hi); }
inline float gmix(float a, float b, float t) { return a * (1.0f - t) + b * t; }
inline float gfract(float x) { return x - floorf(x); }
inline float gmod(float x, float y) { return x - y * floorf(x / y); }
inline float glength2(float x, float y) { return sqrtf(x * x + y * y); }
inline int imax(int a, int b) { return a < b ? b : a; }
inline int imin(int a, int b) { return b < a ? b : a; }
inline float gsmoothstep(float e0, float e1, float x) {
  float t = gclamp((x - e0) / (e1 - e0), 0.0f, 1.0f);
  return t * t * (3.0f - 2.0f * t);
}

struct f4 {
  float v[4];
  float& operator[](int i) { return v[i]; }
  const float& operator[](int i) const { return v[i]; }
};
struct f2 { float x, y; };
struct i4 {
  int v[4];
  int& operator[](int i) { return v[i]; }
  const int& operator[](int i) const { return v[i]; }
};

// common:42-95 channel indices
enum { VX = 0, VY = 1, PRESSURE = 2, TEMPERATURE = 3 };
enum { TOTAL = 0, CLOUD = 1, PRECIPITATION = 2, SOIL_MOISTURE = 2, SMOKE = 3, SNOW = 3 };
enum { TYPE = 0, DISTANCE = 1, VERT_DISTANCE = 2, VEGETATION = 3 };
enum { WALLTYPE_INERT = 0, WALLTYPE_LAND = 1, WALLTYPE_WATER = 2, WALLTYPE_FIRE = 3,
       WALLTYPE_URBAN = 4, WALLTYPE_RUNWAY = 5, WALLTYPE_INDUSTRIAL = 6 };
enum { SUNLIGHT = 0, NET_HEATING = 1, IR_DOWN = 2, IR_UP = 3 };
enum { WATER = 0, ICE = 1 };
enum { MASS = 0, HEAT = 1, VAPOR = 2 };
enum { START_ITERNUM = 2, INTENSITY = 3 };
enum { RAIN_DEPOSITION = 0, SNOW_DEPOSITION = 1 };

// common:9-35 constants
const float lightHeatingConst = 0.000002f;
const float maxWaterTemp = 40.0f;
const float waterHeatExchangeRate = 0.0002f;
const float waterHeatCapacity = 50.0f;
const float fullWhiteSnowHeight = 10.0f;
const float snowMassToHeight = 0.05f;
const float snowMeltRate = 0.000015f;
const float ALBEDO_SNOW = 0.85f, ALBEDO_SNOW_FOREST = 0.30f, ALBEDO_FOREST = 0.10f,
            ALBEDO_DRYSOIL = 0.30f, ALBEDO_WETSOIL = 0.15f, ALBEDO_URBAN = 0.08f,
            ALBEDO_INDUSTRIAL = 0.08f, ALBEDO_RUNWAY = 0.04f, ALBEDO_WATER = 0.05f;
const float deg2rad = 0.0174533f;

// common:99-101
inline float map_range(float value, float min1, float max1, float min2, float max2) {
  return min2 + (value - min1) * (max2 - min2) / (max1 - min1);
}
inline float map_rangeC(float value, float min1, float max1, float min2, float max2) {
  return gclamp(map_range(value, min1, max1, min2, max2), gmin(min2, max2), gmax(min2, max2));
}

// common:103-137
inline uint32_t hash_u(uint32_t x) {
  x += (x << 10u);
  x ^= (x >> 6u);
  x += (x << 3u);
  x ^= (x >> 11u);
  x += (x << 15u);
  return x;
}
inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
inline float random2d(float sx, float sy) {
  uint32_t h = hash_u(f2u(sx) + hash_u(f2u(sy)));
  h &= 0x007FFFFFu;
  h |= 0x3F800000u;
  float r2 = u2f(h);
  return gmod(r2, 1.0f);
}

inline float CtoK(float c) { return c + 273.15f; }   // common:157
inline float KtoC(float k) { return k - 273.15f; }   // common:159

// common:177-180  pow(T/250, 17).  Canonical form (DESIGN.md spec freeze): the exponent is an
// integer, so the power is the multiply chain x^2, x^4, x^8, x^16, x^16*x.
inline float maxWater(float T) {
  float x = T / 250.0f;
  float x2 = x * x, x4 = x2 * x2, x8 = x4 * x4, x16 = x8 * x8;
  return x16 * x;
}
// common:258-261  pow(T*0.01, 4) * IR_constant; canonical: (x*x)*(x*x).
inline float IR_emitted(float T) {
  float x = T * 0.01f;
  float x2 = x * x;
  return (x2 * x2) * 5.670374419f;
}
// pow(m, 1/3) of precipitationShader.vert:195.  Canonical cube root: bit-level initial guess and
// four Newton steps, +,-,*,/ only, so CPU and GPU agree bit for bit.
inline float cbrt_canon(float m) {
  if (!(m > 0.0f)) return 0.0f;
  float y = u2f(f2u(m) / 3u + 709921077u);
  for (int k = 0; k < 4; k++) y = y - (y - m / (y * y)) * (1.0f / 3.0f);
  return y;
}

// ---------------------------------------------------------------------------------------------
struct Params {  // same field order as wsb_params (include/wsb200.h)
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

struct Sim {
  // W = width of the arrays held here.  For a whole-domain run W == Wg and x0 == 0.  For the
  // "fake cluster" strip tests the arrays hold a strip plus ghost columns: array column lx is
  // global column (x0 + lx) mod Wg, Wg the global width; neighbour fetches still wrap inside the
  // local array (the ghost columns absorb the garbage and are refreshed by the halo exchange).
  int W, H, Wg, x0, ND;
  std::vector<float> base[2], water[2], light[2], fb, dep, curl, vort, drops[2];
  std::vector<int8_t> wall[2];
  std::vector<float> initial_T, snd_T, snd_W, snd_Vel;  // H+1 entries
  float lightning[4];
  float inactiveDroplets;  // the uniform, latched every 600 iterations (app.js:5957-5967)
  long iter;
  bool even;
  int last_drops;  // droplet buffer written last
  Params p;
  FrameInputs in;
  float texelX, texelY;  // uniform texelSize = (1/sim_res_x, 1/sim_res_y) as f32 (app.js:5436)
};

inline int wrapi(int i, int n) { i %= n; return i < 0 ? i + n : i; }

struct View {  // nearest + REPEAT sampling of one texture
  const Sim* s;
  inline size_t idx(int x, int y) const { return (size_t)wrapi(y, s->H) * s->W + wrapi(x, s->W); }
};
inline f4 ld4(const std::vector<float>& a, size_t i) {
  f4 r; memcpy(r.v, &a[i * 4], 16); return r;
}
inline void st4(std::vector<float>& a, size_t i, const f4& v) { memcpy(&a[i * 4], v.v, 16); }
inline i4 ldw(const std::vector<int8_t>& a, size_t i) {
  i4 r; for (int c = 0; c < 4; c++) r.v[c] = a[i * 4 + c]; return r;
}
// RGBA8I store: canonical = saturate to [-128,127] (SURVEY 7 hard part 3a).
inline void stw(std::vector<int8_t>& a, size_t i, const i4& v) {
  for (int c = 0; c < 4; c++) a[i * 4 + c] = (int8_t)imin(imax(v.v[c], -128), 127);
}

inline int gx_of(const Sim& s, int lx) { return wrapi(s.x0 + lx, s.Wg); }

// simShader.vert:23-24.  Canonical fragment coordinate is exactly (x+0.5, y+0.5) (the *1.0000001
// at app.js:4766-4787 is a rasteriser work-around whose intent is "exactly x.5").
inline float fragX(const Sim& s, int lx) { return (float)gx_of(s, lx) + 0.5f; }
inline float fragY(int y) { return (float)y + 0.5f; }
inline float texY(const Sim& s, int y) { return fragY(y) * s.texelY; }
inline float texX(const Sim& s, int lx) { return fragX(s, lx) * s.texelX; }

inline float potentialToRealT(const Sim& s, float potential, float texCoordY) {  // common:151-153
  return potential - texCoordY * s.p.dryLapse;
}

// ---------------------------------------------------------------------------------------------
// pass 1 — velocityShader.frag:32-62   base_0, wall_0 -> base_1, wall_1
// ---------------------------------------------------------------------------------------------
void pass_velocity(Sim& s) {
  View v{&s};
  const auto& B = s.base[0]; const auto& Wl = s.wall[0];
#pragma omp parallel for schedule(static)
  for (int y = 0; y < s.H; y++)
    for (int x = 0; x < s.W; x++) {
      size_t c = v.idx(x, y);
      f4 base = ld4(B, c);
      f4 baseXpY0 = ld4(B, v.idx(x + 1, y));
      f4 baseX0Yp = ld4(B, v.idx(x, y + 1));
      i4 wall = ldw(Wl, c);
      if (wall[DISTANCE] == 0) {  // :40-44
        base[VX] = 0.0f;
        base[VY] = 0.0f;
      } else {
        base[VX] += base[PRESSURE] - baseXpY0[PRESSURE];  // :49
        base[VY] += base[PRESSURE] - baseX0Yp[PRESSURE];  // :50
        base[VX] *= 1.0f - s.p.dragMultiplier * 0.0002f;  // :52
        base[VY] *= 1.0f - s.p.dragMultiplier * 0.0002f;  // :53
        base[VX] += s.p.wind * 0.000001f;                 // :60
      }
      st4(s.base[1], c, base);
      stw(s.wall[1], c, wall);
    }
}

// pass 2 — curlShader.frag:12-19   base_1 -> curl
void pass_curl(Sim& s) {
  View v{&s};
  const auto& B = s.base[1];
#pragma omp parallel for schedule(static)
  for (int y = 0; y < s.H; y++)
    for (int x = 0; x < s.W; x++) {
      f4 cell = ld4(B, v.idx(x, y));
      f4 cellXpY0 = ld4(B, v.idx(x + 1, y));
      f4 cellX0Yp = ld4(B, v.idx(x, y + 1));
      s.curl[v.idx(x, y)] = cellX0Yp[0] - cell[0] - cellXpY0[1] + cell[1];
    }
}

// pass 3 — vorticityShader.frag:19-38   curl -> vortForce
void pass_vorticity(Sim& s) {
  View v{&s};
  const auto& C = s.curl;
#pragma omp parallel for schedule(static)
  for (int y = 0; y < s.H; y++)
    for (int x = 0; x < s.W; x++) {
      float curl = C[v.idx(x, y)];
      float curlXmY0 = C[v.idx(x - 1, y)];
      float curlX0Ym = C[v.idx(x, y - 1)];
      float curlXpY0 = C[v.idx(x + 1, y)];
      float curlX0Yp = C[v.idx(x, y + 1)];
      float fx = fabsf(curlX0Ym) - fabsf(curlX0Yp);
      float fy = fabsf(curlXpY0) - fabsf(curlXmY0);
      float magnitude = glength2(fx, fy) + 0.0001f;
      fx /= magnitude; fy /= magnitude;
      fx *= curl; fy *= curl;
      size_t c = v.idx(x, y);
      s.vort[c * 2] = fx; s.vort[c * 2 + 1] = fy;
    }
}

// ---------------------------------------------------------------------------------------------
// pass 4 — boundaryShader.frag:72-531
//   base_1, water_1, vortForce, wall_1, lightTexture_0 (always _0, app.js:5869), feedback,
//   deposition -> base_0, water_0, wall_0
// ---------------------------------------------------------------------------------------------
inline float calcEvaporation(const Sim& s, float T, float W, float V, float M) {  // :65-68
  return gmax((maxWater(T) - W) * s.p.landEvaporation * (V / 127.0f + 0.1f) * gmin(M + 1.0f, 50.0f) * 0.05f, 0.0f);
}
inline float calcFireIntensity(int veg, float moist, float precip) {  // :70
  return gmax((float)veg * 0.00025f - moist * 0.00020f - precip * 0.02f, 0.0f);
}

void pass_boundary(Sim& s) {
  View v{&s};
  const auto& B = s.base[1]; const auto& WT = s.water[1]; const auto& WL = s.wall[1];
  const auto& L = s.light[0];
  const float iterNum = (float)s.iter;
  const int iterI = (int)iterNum;
  // sin/cos of a uniform: evaluated once on the host in double, rounded to f32 (spec freeze).
  const float cosSun = (float)cos((double)s.in.sunAngle);
  const float sinSun = (float)sin((double)s.in.sunAngle);
  const float sinNegSun = -sinSun;
  const float exchangeRate = 0.015f;  // :54
#pragma omp parallel for schedule(static)
  for (int y = 0; y < s.H; y++)
    for (int x = 0; x < s.W; x++) {
      size_t c = v.idx(x, y);
      const size_t cXm = v.idx(x - 1, y), cXp = v.idx(x + 1, y), cYm = v.idx(x, y - 1), cYp = v.idx(x, y + 1);
      const float texCoordY = texY(s, y);
      const float texCoordYp = texCoordY + s.texelY;  // simShader.vert:28 texCoordX0Yp.y
      f4 base = ld4(B, c);
      f4 water = ld4(WT, c);
      f4 precipFeedback = ld4(s.fb, c);
      float realTemp = potentialToRealT(s, base[TEMPERATURE], texCoordY);  // :80
      i4 wall = ldw(WL, c);
      i4 wallXmY0 = ldw(WL, cXm), wallX0Ym = ldw(WL, cYm), wallXpY0 = ldw(WL, cXp), wallX0Yp = ldw(WL, cYp);
      f4 light = ld4(L, c);
      bool nextToWall = false;
      wall[VERT_DISTANCE] = wallX0Ym[VERT_DISTANCE] + 1;  // :92

      if (wall[DISTANCE] != 0) {  // fluid :94
        wall[TYPE] = wallX0Ym[TYPE];  // :96
        if (wall[TYPE] != WALLTYPE_WATER) base[TEMPERATURE] += light[NET_HEATING];  // :98-99
        base[TEMPERATURE] += precipFeedback[HEAT];                                  // :101
        float precipCoalescence = gmax(-precipFeedback[VAPOR], 0.0f);               // :104
        water[CLOUD] -= precipCoalescence;
        water[TOTAL] -= precipCoalescence;
        float precipEvaporation = gmax(precipFeedback[VAPOR], 0.0f);  // :109
        water[TOTAL] += precipEvaporation;
        water[PRECIPITATION] = gmax(water[PRECIPITATION] * 0.997f - 0.00001f + precipFeedback[MASS] * 0.005f, 0.0f);  // :115
        water[SMOKE] /= 1.0f + gmax(-precipFeedback[VAPOR] * 0.1f, 0.0f) + precipFeedback[MASS] * 0.000f;  // :119
        water[SMOKE] -= precipFeedback[MASS] * 0.0001f;                    // :121
        water[SMOKE] -= gmax((water[SMOKE] - 4.0f) * 0.01f, 0.0f);         // :124
        water[SMOKE] = gmax(water[SMOKE], 0.0f);                           // :126
        if (water[SMOKE] > 4.0f) water[SMOKE] -= water[PRECIPITATION] * 0.02f;  // :128-130

        // GRAVITY :132-148
        f4 baseX0Yp = ld4(B, cYp);
        const float gravMult = 0.0001f;
        int fy = (int)fragY(y);
        float gravityForce = ((base[TEMPERATURE] + baseX0Yp[TEMPERATURE]) * 0.5f - (s.initial_T[fy] + s.initial_T[fy + 1]) * 0.5f) * gravMult;
        gravityForce -= water[CLOUD] * gravMult * s.p.waterWeight;
        gravityForce -= precipFeedback[MASS] * gravMult * s.p.waterWeight;
        base[VY] += gravityForce;

        float snowCover = 0.0f, soilMoisture = 0.0f;
        if (wallX0Ym[DISTANCE] == 0) {  // :155-164
          nextToWall = true;
          wall[DISTANCE] = 1;
          f4 waterX0Ym = ld4(WT, cYm);
          snowCover = waterX0Ym[SNOW];
          soilMoisture = waterX0Ym[SOIL_MOISTURE];
          wall[VERT_DISTANCE] = 1;
        }
        if (wallXmY0[DISTANCE] == 0) {  // :166-177
          nextToWall = true;
          wall[DISTANCE] = 1;
          if (wallXmY0[TYPE] == WALLTYPE_WATER) { wall[TYPE] = WALLTYPE_LAND; wall[DISTANCE] = 0; }
          if (wallXpY0[DISTANCE] == 0) wall[DISTANCE] = 0;
        } else if (wallXpY0[DISTANCE] == 0) {  // :178-187
          nextToWall = true;
          wall[DISTANCE] = 1;
          if (wallXpY0[TYPE] == WALLTYPE_WATER) { wall[TYPE] = WALLTYPE_LAND; wall[DISTANCE] = 0; }
        }
        if (wallX0Yp[DISTANCE] == 0) {  // :188-196
          nextToWall = true;
          wall[DISTANCE] = 1;
          if (texCoordY < 0.99f) wall[DISTANCE] = 0;
        }

        // vorticity confinement :201-208
        float vf00x = s.vort[c * 2], vf00y = s.vort[c * 2 + 1];
        float vfXm_y = s.vort[cXm * 2 + 1];
        float vfYm_x = s.vort[cYm * 2];
        float velocityFactor = glength2(base[VX], base[VY]) * 0.1f;
        base[VX] += (vf00x + vfYm_x) * (s.p.vorticity + velocityFactor);
        base[VY] += (vf00y + vfXm_y) * (s.p.vorticity + velocityFactor);

        if (nextToWall) {  // :211-243
          if (wall[TYPE] != WALLTYPE_WATER) {
            float lightPower = 0.0f;
            if (wallX0Ym[DISTANCE] == 0) lightPower += gmax(light[SUNLIGHT] * cosSun, 0.0f);
            if (wallXmY0[DISTANCE] == 0) lightPower += gmax(light[SUNLIGHT] * sinSun, 0.0f);
            if (wallXpY0[DISTANCE] == 0) lightPower += gmax(light[SUNLIGHT] * sinNegSun, 0.0f);
            float albedoTotal = 1.0f;
            if (wall[TYPE] == WALLTYPE_LAND || wall[TYPE] == WALLTYPE_FIRE) {
              float albedoSoil = map_rangeC(soilMoisture, 0.0f, 20.0f, ALBEDO_DRYSOIL, ALBEDO_WETSOIL);
              albedoSoil = map_rangeC(snowCover, 0.0f, fullWhiteSnowHeight, albedoSoil, ALBEDO_SNOW);
              float fullVegetationAlbedo = map_range(snowCover, 0.0f, fullWhiteSnowHeight, ALBEDO_FOREST, ALBEDO_SNOW_FOREST);
              albedoTotal = map_range((float)wallX0Ym[VEGETATION], 0.0f, 127.0f, albedoSoil, fullVegetationAlbedo);
            } else if (wall[TYPE] == WALLTYPE_URBAN) {
              albedoTotal = ALBEDO_URBAN;
            } else if (wall[TYPE] == WALLTYPE_INDUSTRIAL) {
              albedoTotal = ALBEDO_INDUSTRIAL;
            } else if (wall[TYPE] == WALLTYPE_RUNWAY) {
              albedoTotal = ALBEDO_RUNWAY;
            }
            lightPower *= (1.0f - albedoTotal);
            lightPower *= lightHeatingConst;
            base[TEMPERATURE] += lightPower;
          }
        }

        if (!nextToWall) {  // :245-269
          int nearest = 255;
          if (wallX0Ym[DISTANCE] < nearest) nearest = wallX0Ym[DISTANCE];
          if (wallX0Yp[DISTANCE] < nearest) nearest = wallX0Yp[DISTANCE];
          if (wallXmY0[DISTANCE] < nearest) nearest = wallXmY0[DISTANCE];
          if (wallXpY0[DISTANCE] < nearest) nearest = wallXpY0[DISTANCE];
          wall[DISTANCE] = nearest + 1;
        }

        if (wall[VERT_DISTANCE] <= 5) {  // :273-303 surfaceWindSmootingDist
          if (wall[VERT_DISTANCE] == 1) {
            float surfaceDrag = 0.0015f;
            if (wall[TYPE] == WALLTYPE_URBAN)
              surfaceDrag = 0.040f;
            else if (wall[TYPE] == WALLTYPE_LAND || wall[TYPE] == WALLTYPE_FIRE)
              surfaceDrag = map_rangeC((float)wall[VEGETATION], 50.0f, 127.0f, 0.0015f, 0.020f);
            base[VX] -= fabsf(base[VX]) * base[VX] * surfaceDrag * 50.0f;  // :283
          }
          // exchangeWith (:56-62) smooths only VX, sequentially
          if (wallX0Yp[VERT_DISTANCE] <= 5) base[VX] -= (base[VX] - ld4(B, cYp)[VX]) * exchangeRate;  // :288-290
          if (wallX0Ym[VERT_DISTANCE] > 0) base[VX] -= (base[VX] - ld4(B, cYm)[VX]) * exchangeRate;   // :292-294
        }

        if (wall[VERT_DISTANCE] <= 8) {  // :305-372
          const float influenceDevider = 1.0f;  // wallVerticalInfluence = 1
          wall[VEGETATION] = wallX0Ym[VEGETATION];  // :310
          f4 waterInSurface = ld4(WT, cYm);         // :316
          int t = wall[TYPE];
          // C-style switch with deliberate fall-through: FIRE -> INDUSTRIAL -> URBAN -> LAND
          bool in = false;
          if (t == WALLTYPE_FIRE) {
            in = true;
            if (wall[VERT_DISTANCE] == 1) {  // :320-327
              float fireIntensity = calcFireIntensity(wall[VEGETATION], waterInSurface[SOIL_MOISTURE], water[PRECIPITATION]);
              fireIntensity = gmax(fireIntensity, 0.0f);
              base[TEMPERATURE] += fireIntensity;
              water[SMOKE] += fireIntensity * 2.0f;
              water[TOTAL] += fireIntensity * 0.50f;
            }
          }
          if (in || t == WALLTYPE_INDUSTRIAL) {
            in = true;
            if (wall[TYPE] == WALLTYPE_INDUSTRIAL) {  // :330-345
              int texFragX = (int)fragX(s, x) % 80;
              if (wall[VERT_DISTANCE] == 5 && (texFragX == 18 || texFragX == 22)) {
                water[TOTAL] += 0.25f;
                base[VX] *= 0.5f; base[VY] *= 0.5f;
                base[VY] += 0.05f;
              } else if (wall[VERT_DISTANCE] == 6 && texFragX == 29) {
                water[SMOKE] += 0.01f;
                base[TEMPERATURE] += 0.02f;
                base[VX] *= 0.5f; base[VY] *= 0.5f;
              }
            }
          }
          if (in || t == WALLTYPE_URBAN) {
            in = true;
            water[SMOKE] += 0.000002f;  // :348
          }
          if (in || t == WALLTYPE_LAND) {  // :350-362
            if (wall[VERT_DISTANCE] <= 1) {
              float evaporation = calcEvaporation(s, realTemp, water[TOTAL], (float)wall[VEGETATION], waterInSurface[SOIL_MOISTURE]) / influenceDevider;
              water[TOTAL] += evaporation;
              base[TEMPERATURE] -= evaporation * s.p.evapHeat * 0.5f;
              if (wall[VEGETATION] < 10 && water[SOIL_MOISTURE] < 5.0f) {
                water[SMOKE] = gmin(water[SMOKE] + (gmax(fabsf(base[VX]) - 0.12f, 0.0f) * 0.15f), 2.4f);
              }
            }
          } else if (t == WALLTYPE_WATER) {  // :363-370
            if (wall[VERT_DISTANCE] <= 1) {
              float LocalWaterTemperature = ld4(B, cYm)[TEMPERATURE];
              base[TEMPERATURE] += (LocalWaterTemperature - realTemp - 1.0f) / influenceDevider * waterHeatExchangeRate;
              water[TOTAL] += gmax((maxWater(LocalWaterTemperature) - water[TOTAL]) * s.p.waterEvaporation / influenceDevider, 0.0f);
            }
          }
        }
      } else {  // this is wall :373
        wall[VERT_DISTANCE] = wallX0Yp[VERT_DISTANCE] - 1;  // :375
        if (wall[VERT_DISTANCE] < 0) {                       // :377-388
          f4 wYp = ld4(WT, cYp);
          water[2] = wYp[2]; water[3] = wYp[3];
          wall[VEGETATION] = wallX0Yp[VEGETATION];
          if (wallX0Yp[DISTANCE] == 0) {
            if (wallX0Yp[TYPE] != WALLTYPE_WATER) {
              wall[TYPE] = wallX0Yp[TYPE];
            } else if (wall[TYPE] == WALLTYPE_WATER) {
              base[TEMPERATURE] = ld4(B, cYp)[TEMPERATURE];
            }
          }
        } else if (wall[VERT_DISTANCE] == 0) {  // surface layer :390
          f4 waterX0Yp = ld4(WT, cYp);
          float depR = s.dep[c * 2 + RAIN_DEPOSITION], depS = s.dep[c * 2 + SNOW_DEPOSITION];
          // the light texture is the only one with wrap T = CLAMP_TO_EDGE (app.js:5278)
          f4 lightAboveSurface = ld4(L, (size_t)imin(y + 1, s.H - 1) * s.W + x);
          int t = wall[TYPE];
          bool in = false;
          if (t == WALLTYPE_INDUSTRIAL) { in = true; wall[VEGETATION] = imin(wall[VEGETATION], 15); }  // :399-400
          if (in || t == WALLTYPE_URBAN) { in = true; wall[VEGETATION] = imin(wall[VEGETATION], 75); } // :401-402
          if (in || t == WALLTYPE_FIRE) {  // :403-414
            in = true;
            if (wall[TYPE] == WALLTYPE_FIRE) {
              float fireIntensity = calcFireIntensity(wall[VEGETATION], water[SOIL_MOISTURE], waterX0Yp[PRECIPITATION]);
              if (fireIntensity < 0.002f) {  // minimalFireIntensity
                wall[TYPE] = WALLTYPE_LAND;
              } else if (iterI % ((int)(10.0f / fireIntensity) + 1) == 0) {
                wall[VEGETATION] -= 1;
                if (wall[VEGETATION] < 10) wall[TYPE] = WALLTYPE_LAND;
              }
            }
          }
          if (in || t == WALLTYPE_LAND) {  // :415-475
            water[SOIL_MOISTURE] = gclamp(water[SOIL_MOISTURE] + depR * 0.1f, 0.0f, 1000.0f);
            water[SNOW] = gclamp(water[SNOW] + depS * snowMassToHeight, 0.0f, 4000.0f);
            f4 baseAboveSurface = ld4(B, cYp);
            f4 waterAboveSurface = ld4(WT, cYp);
            float realTempAboveSurface = potentialToRealT(s, baseAboveSurface[TEMPERATURE], texCoordYp);
            float evaporation = calcEvaporation(s, realTempAboveSurface, waterAboveSurface[TOTAL], (float)wall[VEGETATION], water[SOIL_MOISTURE]) * 0.10f;
            water[SOIL_MOISTURE] -= evaporation;
            if (iterI % 100 == 0) {  // :430-474
              const float snowSmoothingRate = 0.02f, moistureSmoothingRate = 0.02f;
              float numNeighbors = 0.0f, totalNeighborSnow = 0.0f, totalNeighborSoilMoisture = 0.0f;
              if (wallXmY0[VERT_DISTANCE] == 0 && (wallXmY0[TYPE] == WALLTYPE_LAND || wallXmY0[TYPE] == WALLTYPE_URBAN)) {
                f4 wn = ld4(WT, cXm);
                totalNeighborSnow += wn[SNOW];
                totalNeighborSoilMoisture += wn[SOIL_MOISTURE];
                numNeighbors += 1.0f;
              }
              if (wallXpY0[VERT_DISTANCE] == 0 && (wallXpY0[TYPE] == WALLTYPE_LAND || wallXpY0[TYPE] == WALLTYPE_URBAN)) {
                f4 wn = ld4(WT, cXp);
                totalNeighborSnow += wn[SNOW];
                totalNeighborSoilMoisture += wn[SOIL_MOISTURE];
                numNeighbors += 1.0f;
              }
              if (numNeighbors > 0.0f) {
                float avgNeighborSnow = totalNeighborSnow / numNeighbors;
                water[SNOW] += (avgNeighborSnow - water[SNOW]) * snowSmoothingRate;
                float avgNeighborSoilMoisture = totalNeighborSoilMoisture / numNeighbors;
                water[SOIL_MOISTURE] += (avgNeighborSoilMoisture - water[SOIL_MOISTURE]) * moistureSmoothingRate;
              }
              int vegetationGrowthRate = (int)(water[SOIL_MOISTURE] * sqrtf(lightAboveSurface[SUNLIGHT]) * 0.01f);  // :460
              // rate > 100: the interval is 0 and `% 0` is undefined in GLSL ES 3.00 (5.9) — frozen as "no growth tick" 
              const int growthInterval = vegetationGrowthRate > 0 ? (100 / vegetationGrowthRate) * 100 : 0;
              if (growthInterval > 0 && iterI % growthInterval == 0) {
                if ((int)map_rangeC(realTempAboveSurface, CtoK(0.0f), CtoK(25.0f), 0.0f, 127.0f) > wall[VEGETATION]) wall[VEGETATION] += 1;
              }
              int subInterval = iterI / 100;  // :467
              if (subInterval % ((int)(water[SOIL_MOISTURE] * 0.1f + water[SNOW] * 0.5f) + 10) == 0 && wall[VEGETATION] >= 20 &&
                  (wallXmY0[TYPE] == WALLTYPE_FIRE || wallXpY0[TYPE] == WALLTYPE_FIRE || waterX0Yp[SMOKE] > 4.5f)) {
                wall[TYPE] = WALLTYPE_FIRE;
              }
            }
          } else if (t == WALLTYPE_WATER) {  // :476-527
            const float waterTempUpdateInterval = 20.0f;
            if (s.p.dynamicWaterTemperature >= 1.0f && gmod(iterNum, waterTempUpdateInterval) < 0.5f) {
              float numNeighbors = 0.0f, totalNeighborTemp = 0.0f;
              if (wallXmY0[TYPE] == WALLTYPE_WATER) { totalNeighborTemp += ld4(B, cXm)[TEMPERATURE]; numNeighbors += 1.0f; }
              if (wallXpY0[TYPE] == WALLTYPE_WATER) { totalNeighborTemp += ld4(B, cXp)[TEMPERATURE]; numNeighbors += 1.0f; }
              if (numNeighbors > 0.0f) {
                float avgNeighborTemp = totalNeighborTemp / numNeighbors;
                base[TEMPERATURE] += (avgNeighborTemp - base[TEMPERATURE]) * 0.10f;
              }
              if (base[TEMPERATURE] > 500.0f) base[TEMPERATURE] = CtoK(25.0f);
              float airTemperature = potentialToRealT(s, ld4(B, cYp)[TEMPERATURE], texCoordYp);
              float netWaterHeating = 0.0f;
              netWaterHeating += (airTemperature - base[TEMPERATURE]) * waterHeatExchangeRate;
              netWaterHeating -= gmax((maxWater(base[TEMPERATURE]) - waterX0Yp[TOTAL]) * s.p.waterEvaporation, 0.0f) * s.p.evapHeat * 0.5f;
              float lightPower = gmax(lightAboveSurface[SUNLIGHT] * cosSun, 0.0f);
              lightPower *= (1.0f - ALBEDO_WATER);
              lightPower *= lightHeatingConst;
              netWaterHeating += lightPower;
              netWaterHeating += lightAboveSurface[NET_HEATING];
              base[TEMPERATURE] += netWaterHeating / waterHeatCapacity * waterTempUpdateInterval;
            }
            base[TEMPERATURE] = gclamp(base[TEMPERATURE], CtoK(0.0f), CtoK(maxWaterTemp));  // :522
            wall[VEGETATION] = 20;
            water[SOIL_MOISTURE] = 100.0f;
            water[SNOW] = 0.0f;
          }
        }
      }
      st4(s.base[0], c, base);
      st4(s.water[0], c, water);
      stw(s.wall[0], c, wall);
    }
}

// ---------------------------------------------------------------------------------------------
// pass 5 — advectionShader.frag:65-458   base_0, water_0, wall_0 -> base_1, water_1, wall_1
// ---------------------------------------------------------------------------------------------
struct Bilerp {  // common:194-254.  pos is in GLOBAL pixel coordinates; (lx,y) the local cell whose
                 // fragCoord the back-trace started from, used to map global texels to local ones.
  int ix, iy; float fx, fy;
};
inline Bilerp bilerp_setup(float posx, float posy) {
  float stx = posx - 0.5f, sty = posy - 0.5f;  // common:196
  float flx = floorf(stx), fly = floorf(sty);
  Bilerp b; b.ix = (int)flx; b.iy = (int)fly; b.fx = stx - flx; b.fy = sty - fly;  // fract
  return b;
}
inline f4 mix4(const f4& a, const f4& b, float t) {
  f4 r; for (int c = 0; c < 4; c++) r[c] = gmix(a[c], b[c], t); return r;
}

void pass_advection(Sim& s, bool dry) {
  View v{&s};
  const auto& B = s.base[0]; const auto& WT = s.water[0]; const auto& WL = s.wall[0];
  const Params& p = s.p;
  const float H = (float)s.H, Wf = (float)s.Wg;
  const float ltexelX = 1.0f / Wf, ltexelY = 1.0f / H;  // :69 texelSize = vec2(1.)/resolution
#pragma omp parallel for schedule(static)
  for (int y = 0; y < s.H; y++)
    for (int x = 0; x < s.W; x++) {
      size_t c = v.idx(x, y);
      const int gx = gx_of(s, x);
      const float fragCoordX = fragX(s, x), fragCoordY = fragY(y);
      const float texCoordX = texX(s, x), texCoordY = texY(s, y);
      i4 wall = ldw(WL, c);
      f4 base, water;
      float realTemp = 0.0f;
      // texel fetch in global pixel space, mapped into the local array
      auto LX = [&](int gix) { return x + (gix - gx); };

      if (wall[DISTANCE] != 0) {  // not wall :73
        f4 cellX0Y0 = ld4(B, c);
        f4 cellXmY0 = ld4(B, v.idx(x - 1, y));
        f4 cellX0Ym = ld4(B, v.idx(x, y - 1));
        f4 cellXpY0 = ld4(B, v.idx(x + 1, y));
        f4 cellX0Yp = ld4(B, v.idx(x, y + 1));
        f4 cellXmYp = ld4(B, v.idx(x - 1, y + 1));
        f4 cellXpYm = ld4(B, v.idx(x + 1, y - 1));
        // :85-89
        float velAtP_x = (cellXmY0[0] + cellX0Y0[0]) / 2.0f;
        float velAtP_y = (cellX0Ym[1] + cellX0Y0[1]) / 2.0f;
        float velAtVx_x = cellX0Y0[0];
        float velAtVx_y = (cellX0Ym[1] + cellXpY0[1] + cellX0Y0[1] + cellXpYm[1]) / 4.0f;
        float velAtVy_x = (cellXmY0[0] + cellX0Yp[0] + cellXmYp[0] + cellX0Y0[0]) / 4.0f;
        float velAtVy_y = cellX0Y0[1];

        auto bilerp = [&](const std::vector<float>& T, float px, float py) {  // common:194-214
          Bilerp b = bilerp_setup(px, py);
          f4 a = ld4(T, v.idx(LX(b.ix), b.iy));
          f4 bb = ld4(T, v.idx(LX(b.ix + 1), b.iy));
          f4 cc = ld4(T, v.idx(LX(b.ix), b.iy + 1));
          f4 d = ld4(T, v.idx(LX(b.ix + 1), b.iy + 1));
          return mix4(mix4(a, bb, b.fx), mix4(cc, d, b.fx), b.fy);
        };
        auto bilerpWall = [&](const std::vector<float>& T, float px, float py) {  // common:216-254
          Bilerp b = bilerp_setup(px, py);
          size_t ia = v.idx(LX(b.ix), b.iy), ib = v.idx(LX(b.ix + 1), b.iy);
          size_t ic = v.idx(LX(b.ix), b.iy + 1), id = v.idx(LX(b.ix + 1), b.iy + 1);
          f4 a = ld4(T, ia), bb = ld4(T, ib), cc = ld4(T, ic), d = ld4(T, id);
          int wa = WL[ia * 4 + 1], wb = WL[ib * 4 + 1], wc = WL[ic * 4 + 1], wd = WL[id * 4 + 1];
          float mixAB = b.fx, mixCD = b.fx, mixAB_CD = b.fy;
          if (wa == 0) mixAB = 1.0f; else if (wb == 0) mixAB = 0.0f;
          if (wc == 0) mixCD = 1.0f; else if (wd == 0) mixCD = 0.0f;
          if (wa == 0 && wb == 0) mixAB_CD = 1.0f; else if (wc == 0 && wd == 0) mixAB_CD = 0.0f;
          return mix4(mix4(a, bb, mixAB), mix4(cc, d, mixCD), mixAB_CD);
        };

        base[VX] = bilerp(B, fragCoordX - velAtVx_x, fragCoordY - velAtVx_y)[0];  // :93
        base[VY] = bilerp(B, fragCoordX - velAtVy_x, fragCoordY - velAtVy_y)[1];  // :94
        f4 bP = bilerpWall(B, fragCoordX - velAtP_x, fragCoordY - velAtP_y);      // :96-97
        base[PRESSURE] = bP[PRESSURE];
        base[TEMPERATURE] = bP[TEMPERATURE];
        if (dry) {
          water = ld4(WT, c);  // dry sweep (BASELINE config 2): water is not part of the state
        } else {
          f4 wP = bilerpWall(WT, fragCoordX - velAtP_x, fragCoordY - velAtP_y);   // :99
          water[0] = wP[0]; water[1] = wP[1]; water[3] = wP[3];
          water[PRECIPITATION] = bilerpWall(WT, fragCoordX - velAtP_x + 0.0f, fragCoordY - velAtP_y + 0.05f)[PRECIPITATION];  // :103

          realTemp = potentialToRealT(s, base[TEMPERATURE], texCoordY);  // :111
          float excessWater = water[TOTAL] - maxWater(realTemp);         // :115
          float overSaturation = excessWater - water[CLOUD];             // :117
          float condensation;
          if (overSaturation < 0.0f) condensation = overSaturation * 0.20f;
          else condensation = overSaturation * p.condensationRate;
          condensation = gmax(condensation, -water[CLOUD]);  // :126
          float dT = condensation * p.evapHeat * 1.0f;       // :128
          base[TEMPERATURE] += dT;
          realTemp += dT;
          water[CLOUD] += condensation;

          if (texCoordY > p.globalEffectsStartAlt && texCoordY < p.globalEffectsEndAlt) {  // :154-181
            water[TOTAL] -= gclamp(p.globalDrying, 0.0f, gmax(water[TOTAL] - maxWater(gmax(realTemp - 20.0f, CtoK(-80.0f))), 0.0f));
            base[TEMPERATURE] += p.globalHeating;
            int si = (int)(texCoordY * (1.0f / ltexelY));  // :162
            int sm = imax(si - 1, 0);                      // canonical: index y-1 clamped at 0
            float sndT = (s.snd_T[si] + s.snd_T[sm]) / 2.0f;
            float Tdiff = base[TEMPERATURE] - sndT;
            base[TEMPERATURE] -= Tdiff * 0.001f * p.soundingForcing;
            float sndW = (s.snd_W[si] + s.snd_W[sm]) / 2.0f;
            float Wdiff = water[TOTAL] - sndW;
            water[TOTAL] -= Wdiff * 0.001f * p.soundingForcing;
            float dragK = 1.0f - map_rangeC(p.soundingForcing, 0.1f, 1.0f, 0.0f, 0.001f);
            base[VX] *= dragK; base[VY] *= dragK;
            float sndV = (s.snd_Vel[si] + s.snd_Vel[sm]) / 2.0f;
            float velDiff = base[VX] - sndV;
            base[VX] -= velDiff * map_rangeC(p.soundingForcing, 0.9f, 1.0f, 0.0f, 0.001f);
          }
          water[TOTAL] = gmax(water[TOTAL], 0.0f);  // :187
        }
      } else {  // this is wall :189
        base = ld4(B, c);
        water = ld4(WT, c);
        if (wall[TYPE] == WALLTYPE_LAND) base[TEMPERATURE] = 1000.0f;  // :195-197
        if (!dry) {
          i4 wallX0Yp = ldw(WL, v.idx(x, y + 1));
          wall[VEGETATION] = imax(wall[VEGETATION], 0);            // :204
          water[SOIL_MOISTURE] = gmax(water[SOIL_MOISTURE], 0.0f); // :205
          if (wallX0Yp[DISTANCE] != 0) {                           // :207-226
            f4 baseX0Yp = ld4(B, v.idx(x, y + 1));
            // potentialToRealT(T) uses THIS cell's texCoord.y (common:151), as the shader does
            float tempC = KtoC(potentialToRealT(s, baseX0Yp[TEMPERATURE], texCoordY));
            if (water[SNOW] > 0.0f && tempC > 0.0f) {
              float melting = gmin(tempC * snowMeltRate, water[SNOW]);
              water[SNOW] -= melting;
              base[TEMPERATURE] += melting / snowMassToHeight * p.meltingHeat;
              water[SOIL_MOISTURE] += melting;
            }
            if (water[SOIL_MOISTURE] > 0.0f && tempC > 0.0f) {
              float evaporation = gmax((maxWater(CtoK(tempC)) - water[TOTAL]) * 0.00001f, 0.0f);
              water[SOIL_MOISTURE] -= evaporation;
            }
          }
        }
      }

      if (!dry) {
        // USER INPUT :229-401
        const float* uiv = s.in.userInputValues;
        bool inBrush = false;
        float weight = 1.0f;
        if (uiv[0] < -0.5f) {  // whole width brush
          if (fabsf(uiv[1] - texCoordY) < uiv[3] * ltexelY) inBrush = true;
        } else {
          float vx, vy = uiv[1] - texCoordY;
          if (s.in.wrapHorizontally) {
            float a = uiv[0], b = texCoordX;  // common:268-271 absHorizontalDist
            vx = gmin(gmin(fabsf(a - b), fabsf(1.0f + a - b)), 1.0f - a + b);
          } else {
            vx = fabsf(uiv[0] - texCoordX);
          }
          vx *= ltexelY / ltexelX;  // :247
          float distFromMouse = glength2(vx, vy);
          weight = gsmoothstep(uiv[3] * ltexelY, 0.0f, distFromMouse);
          if (distFromMouse < uiv[3] * ltexelY) inBrush = true;
        }
        const int uit = s.in.userInputType;
        const float intensity = uiv[2];
        if (inBrush) {
          if (uit == 1) {  // :259-262
            base[3] += intensity;
            if (wall[TYPE] == 2 && wall[DISTANCE] == 0) base[3] = gclamp(base[TEMPERATURE], CtoK(0.0f), CtoK(maxWaterTemp));
          } else if (uit == 2) {  // :263-278
            float cloudWaterChange = intensity;
            if (water[CLOUD] > 0.0f) { water[CLOUD] += cloudWaterChange; water[CLOUD] = gmax(water[CLOUD], 0.0f); }
            water[TOTAL] += cloudWaterChange;
            water[TOTAL] = gmax(water[TOTAL], 0.0f);
          } else if (uit == 3 && wall[DISTANCE] != 0) {  // :280-282
            water[SMOKE] += intensity;
            water[SMOKE] = gmin(gmax(water[SMOKE], 0.0f), 2.0f);
          } else if (uit == 4) {  // :284-290
            if (uiv[0] < -0.5f) {
              base[VX] += s.in.userInputMove[0] * 5.0f * weight * intensity;
            } else {
              base[VX] += s.in.userInputMove[0] * 5.0f * weight * intensity;
              base[VY] += s.in.userInputMove[1] * 5.0f * weight * intensity;
            }
          } else if (uit >= 10) {  // wall :291
            const int aboveDist = WL[v.idx(x, y + 1) * 4 + DISTANCE];
            if (intensity > 0.0f) {
              bool setWall = false;
              switch (uit) {
                case 10: wall[TYPE] = WALLTYPE_INERT; setWall = true; break;
                case 11: wall[TYPE] = WALLTYPE_LAND; setWall = true; break;
                case 12: wall[TYPE] = WALLTYPE_WATER; setWall = true; break;
                case 13:
                  if (wall[DISTANCE] == 0 && wall[TYPE] == WALLTYPE_LAND && aboveDist != 0) { wall[TYPE] = WALLTYPE_FIRE; setWall = true; }
                  break;
                case 14:
                  if (wall[DISTANCE] == 0 && (wall[TYPE] == WALLTYPE_LAND || wall[TYPE] == WALLTYPE_RUNWAY || wall[TYPE] == WALLTYPE_INDUSTRIAL) && aboveDist != 0) wall[TYPE] = WALLTYPE_URBAN;
                  break;
                case 15:
                  if (wall[DISTANCE] == 0 && (wall[TYPE] == WALLTYPE_LAND || wall[TYPE] == WALLTYPE_URBAN || wall[TYPE] == WALLTYPE_INDUSTRIAL) && aboveDist != 0) wall[TYPE] = WALLTYPE_RUNWAY;
                  break;
                case 16:
                  if (wall[DISTANCE] == 0 && (wall[TYPE] == WALLTYPE_LAND || wall[TYPE] == WALLTYPE_URBAN || wall[TYPE] == WALLTYPE_RUNWAY) && aboveDist != 0) wall[TYPE] = WALLTYPE_INDUSTRIAL;
                  break;
                case 20:
                  if (wall[DISTANCE] == 0 && wall[TYPE] != WALLTYPE_WATER && aboveDist != 0) water[SOIL_MOISTURE] += intensity * 10.0f;
                  break;
                case 21:
                  if (wall[DISTANCE] == 0 && (wall[TYPE] == WALLTYPE_LAND || wall[TYPE] == WALLTYPE_URBAN || wall[TYPE] == WALLTYPE_INDUSTRIAL) && aboveDist != 0) water[SNOW] += intensity * 0.5f;
                  break;
                case 22:
                  if (wall[DISTANCE] == 0 && (wall[TYPE] == WALLTYPE_LAND || wall[TYPE] == WALLTYPE_FIRE || wall[TYPE] == WALLTYPE_URBAN || wall[TYPE] == WALLTYPE_INDUSTRIAL) && aboveDist != 0) wall[VEGETATION] += 1;
                  break;
                default: break;
              }
              if (setWall) {  // :354-365
                wall[DISTANCE] = 0;
                base[TEMPERATURE] = 1000.0f;
                if (wall[TYPE] == WALLTYPE_LAND) water[SOIL_MOISTURE] = 25.0f;
                else if (wall[TYPE] == WALLTYPE_WATER) base[TEMPERATURE] = p.waterTemperature;
              }
            } else {  // :366-399
              if (wall[DISTANCE] == 0) {
                if (uit == 13) { if (wall[TYPE] == WALLTYPE_FIRE) wall[TYPE] = WALLTYPE_LAND; }
                else if (uit == 14) { if (wall[TYPE] == WALLTYPE_URBAN) wall[TYPE] = WALLTYPE_LAND; }
                else if (uit == 15) { if (wall[TYPE] == WALLTYPE_RUNWAY) wall[TYPE] = WALLTYPE_LAND; }
                else if (uit == 16) { if (wall[TYPE] == WALLTYPE_INDUSTRIAL) wall[TYPE] = WALLTYPE_LAND; }
                else if (uit == 20) { water[SOIL_MOISTURE] += intensity * 10.0f; }
                else if (uit == 21) { water[SNOW] += intensity * 0.5f; }
                else if (uit == 22) { wall[VEGETATION] = imax(wall[VEGETATION] - 1, 0); }
                else if (texCoordY > ltexelY) {
                  wall[DISTANCE] = 255;
                  base[VX] = 0.0f; base[VY] = 0.0f; base[PRESSURE] = 0.0f;
                  base[TEMPERATURE] = s.initial_T[(int)(texCoordY * (1.0f / ltexelY))];
                  water[TOTAL] = 0.0f; water[CLOUD] = 0.0f; water[PRECIPITATION] = 0.0f; water[SMOKE] = 0.0f;
                }
              }
            }
          }
        }

        if (wall[DISTANCE] == 0) {  // :403-409
          if (wall[TYPE] == WALLTYPE_WATER) water[TOTAL] = 1002.0f;
          else water[TOTAL] = 1001.0f;
        }

        // airplane :415-457
        const float* av = s.in.airplaneValues;
        float px, py = av[1] - texCoordY;
        if (s.in.wrapHorizontally) {
          float a = av[0], b = texCoordX;
          px = gmin(gmin(fabsf(a - b), fabsf(1.0f + a - b)), 1.0f - a + b);
        } else {
          px = fabsf(av[0] - texCoordX);
        }
        px *= ltexelY / ltexelX;
        px *= H; py *= H;  // :424 resolution.y
        if (av[3] < 0.0f) { px += 0.0f; py += -1.0f; }
        float distFromPlane = glength2(px, py);
        float planeInfluence = gmax(1.0f - distFromPlane, 0.0f) * 0.03f;
        if (av[3] < 0.0f) water[PRECIPITATION] += planeInfluence * 100.0f;
        if (av[3] > 0.9f) {
          if (distFromPlane < 1.5f) {
            if (wall[DISTANCE] == 0) {
              if (wall[TYPE] == WALLTYPE_LAND && wall[VERT_DISTANCE] == 0) wall[TYPE] = WALLTYPE_FIRE;
            } else {
              base[PRESSURE] += 0.05f;
              base[TEMPERATURE] = CtoK(50.0f);
              water[TOTAL] += 1.0f;
              water[SMOKE] += 10.0f;
            }
          }
        }
      }
      st4(s.base[1], c, base);
      st4(s.water[1], c, water);
      stw(s.wall[1], c, wall);
    }
}

// pass 6 — pressureShader.frag:16-43   base_1, wall_1 -> base_0, wall_0
void pass_pressure(Sim& s) {
  View v{&s};
  const auto& B = s.base[1]; const auto& WL = s.wall[1];
#pragma omp parallel for schedule(static)
  for (int y = 0; y < s.H; y++)
    for (int x = 0; x < s.W; x++) {
      size_t c = v.idx(x, y);
      f4 base = ld4(B, c);
      f4 baseXmY0 = ld4(B, v.idx(x - 1, y));
      f4 baseX0Ym = ld4(B, v.idx(x, y - 1));
      i4 wall = ldw(WL, c);
      i4 wallX0Ym = ldw(WL, v.idx(x, y - 1));
      if (wallX0Ym[1] == 0 && wallX0Ym[0] == 1) base[3] -= baseX0Ym[3] - 1000.0f;  // :24-26
      base[2] += (baseXmY0[0] - base[0] + baseX0Ym[1] - base[1]) * 0.45f;           // :42
      st4(s.base[0], c, base);
      stw(s.wall[0], c, wall);
    }
}

// ---------------------------------------------------------------------------------------------
// pass 7 — lightingShader.frag:38-171   base_1, water_1, wall_1, light[src] -> light[dst]
//   (second render target reflectedLight is display-only and not computed)
// ---------------------------------------------------------------------------------------------
void pass_lighting(Sim& s) {
  View v{&s};
  const int src = s.even ? 0 : 1, dst = s.even ? 1 : 0;  // app.js:5912-5926
  const auto& B = s.base[1]; const auto& WT = s.water[1]; const auto& WL = s.wall[1];
  const auto& L = s.light[src];
  const float Hf = (float)s.H;
  const float sinSun = (float)sin((double)s.in.sunAngle), cosSun = (float)cos((double)s.in.sunAngle);
  const float absSun = fabsf(s.in.sunAngle);
  (void)absSun;
  const Params& p = s.p;
#pragma omp parallel for schedule(static)
  for (int y = 0; y < s.H; y++)
    for (int x = 0; x < s.W; x++) {
      size_t c = v.idx(x, y);
      f4 light;
      if (fragY(y) >= Hf - 1.0f) {  // :40-41
        light[0] = s.in.sunIntensity; light[1] = 0.0f; light[2] = 0.0f; light[3] = 0.0f;
        st4(s.light[dst], c, light);
        continue;
      }
      const float texCoordY = texY(s, y);
      float cellHeightCompensation = 300.0f / Hf;  // :44
      // :48-49 texture(lightTex, texCoord + sunRay) with LINEAR filter, wrap S = REPEAT, wrap T =
      // CLAMP_TO_EDGE (app.js:5276-5290).  Canonical: full-fp32 bilinear in pixel space at
      // (x + 0.5 + sin a, y + 0.5 + cos a)  (hardware uses reduced-precision weights).
      float sunlight;
      {
        float px = fragX(s, x) + sinSun, py = fragY(y) + cosSun;  // global pixel space
        float stx = px - 0.5f, sty = py - 0.5f;
        float flx = floorf(stx), fly = floorf(sty);
        float fx = stx - flx, fy = sty - fly;
        int ix = x + ((int)flx - gx_of(s, x)), iy = (int)fly;  // global texel -> local column
        int y0 = imin(imax(iy, 0), s.H - 1), y1 = imin(imax(iy + 1, 0), s.H - 1);
        int x0 = wrapi(ix, s.W), x1 = wrapi(ix + 1, s.W);
        float a = L[((size_t)y0 * s.W + x0) * 4], b = L[((size_t)y0 * s.W + x1) * 4];
        float cc = L[((size_t)y1 * s.W + x0) * 4], d = L[((size_t)y1 * s.W + x1) * 4];
        sunlight = gmix(gmix(a, b, fx), gmix(cc, d, fx), fy);
      }
      float realTemp = potentialToRealT(s, ld4(B, c)[TEMPERATURE], texCoordY);  // :52
      f4 water = ld4(WT, c);
      i4 wall = ldw(WL, c);
      if (wall[DISTANCE] != 0) {  // :59
        float net_heating = 0.0f;
        if (fragY(y) < Hf - 2.0f) {  // :65-85
          float reflection = gmin(sqrtf(water[CLOUD] * 0.0010f + water[PRECIPITATION] * 0.00020f) * cellHeightCompensation, 1.0f);
          reflection += 0.0002f;
          float absorbtion = gmin(water[SMOKE] * 0.020f * cellHeightCompensation, 1.0f);
          float lightReflected = sunlight * reflection;
          float lightAbsorbed = sunlight * absorbtion;
          sunlight = gmax(0.0f, sunlight - lightReflected - lightAbsorbed);
          net_heating += lightAbsorbed * lightHeatingConst;
        }
        // light texture wrap T = CLAMP_TO_EDGE applies to the NEAREST-position fetches too (:88,:119)
        int yp = imin(y + 1, s.H - 1), ym = imax(y - 1, 0);
        float IR_down = L[((size_t)yp * s.W + x) * 4 + IR_DOWN];
        float IR_up = 0.0f;  // canonical 0 where the reference leaves it uninitialised (INERT surface)
        if (wall[VERT_DISTANCE] == 1) {  // :90-116
          switch (wall[TYPE]) {
            case WALLTYPE_RUNWAY: case WALLTYPE_URBAN: case WALLTYPE_INDUSTRIAL: case WALLTYPE_LAND:
              IR_up = IR_emitted(realTemp);
              net_heating += (IR_down - IR_up) * lightHeatingConst;
              break;
            case WALLTYPE_WATER: {
              float waterTemperature = ld4(B, v.idx(x, y - 1))[TEMPERATURE];
              IR_up = IR_emitted(waterTemperature);
              net_heating += (IR_down - IR_up) * lightHeatingConst;
              break;
            }
            case WALLTYPE_FIRE:
              IR_up = IR_emitted(realTemp + 100.0f);
              net_heating = 0.0f;
              break;
            default: break;
          }
        } else {  // :117-144
          IR_up = L[((size_t)ym * s.W + x) * 4 + IR_UP];
          float emissivity = p.greenhouseGases;
          emissivity += water[TOTAL] * p.waterGreenHouseEffect;
          emissivity += water[CLOUD] * 5.0f;
          emissivity *= cellHeightCompensation;
          emissivity = gmin(emissivity, 1.0f);
          float absorbedDown = IR_down * emissivity;
          float absorbedUp = IR_up * emissivity;
          float emitted = IR_emitted(realTemp) * emissivity;
          net_heating += (absorbedDown + absorbedUp - emitted * 2.0f) * lightHeatingConst;
          IR_down -= absorbedDown;
          IR_down += emitted;
          IR_up -= absorbedUp;
          IR_up += emitted;
        }
        net_heating *= p.IR_rate;  // :151
        light[0] = sunlight; light[1] = net_heating; light[2] = IR_down; light[3] = IR_up;
      } else {  // :155-168
        if (wall[TYPE] == WALLTYPE_WATER) { light[0] = sunlight * 0.90f; light[1] = 0.0f; light[2] = 0.0f; light[3] = 0.0f; }
        else { light[0] = 0.0f; light[1] = 0.0f; light[2] = 0.0f; light[3] = 0.0f; }
      }
      st4(s.light[dst], c, light);
    }
  s.even = !s.even;  // app.js:5927
}

// ---------------------------------------------------------------------------------------------
// pass 8 — feedback clear (app.js:5933-5934), precipitationShader.vert:66-293 with additive
//          point sprites (app.js:5938-5954), inactive latch (5957-5967),
//          lightningLocationShader.frag:24-38 (app.js:5975-5983)
// ---------------------------------------------------------------------------------------------
// Sprite rasterisation, canonical rule: window centre (xw,yw) = ((p+1)/2 * res);
// a pixel is covered when its centre lies in [c - size/2, c + size/2); sprites are clipped to the
// viewport, never wrapped; a point whose centre is outside the clip volume (|x|>1 or |y|>1) is
// discarded (OpenGL ES 3.0 point clipping).  Blend ONE,ONE in droplet order.
inline void splat(Sim& s, float posx, float posy, float size, const float fbv[4], const float depv[2]) {
  if (!(posx >= -1.0f && posx <= 1.0f && posy >= -1.0f && posy <= 1.0f)) return;
  float xw = (posx + 1.0f) * 0.5f * (float)s.W, yw = (posy + 1.0f) * 0.5f * (float)s.H;
  float half = size * 0.5f;
  int xs = (int)ceilf(xw - half - 0.5f), ys = (int)ceilf(yw - half - 0.5f);
  int n = (int)size;
  for (int j = ys; j < ys + n; j++) {
    if (j < 0 || j >= s.H) continue;
    for (int i = xs; i < xs + n; i++) {
      if (i < 0 || i >= s.W) continue;
      size_t c = (size_t)j * s.W + i;
      for (int k = 0; k < 4; k++) s.fb[c * 4 + k] += fbv[k];
      s.dep[c * 2] += depv[0];
      s.dep[c * 2 + 1] += depv[1];
    }
  }
}

void pass_precipitation(Sim& s) {
  // srcVAO/destTF are chosen before `even` is toggled by the lighting block (app.js:5912-5927);
  // pass_lighting has already toggled, so the source buffer is the one `even` now excludes.
  const int src = s.even ? 1 : 0, dst = s.even ? 0 : 1;
  std::fill(s.fb.begin(), s.fb.end(), 0.0f);
  std::fill(s.dep.begin(), s.dep.end(), 0.0f);
  if (!s.p.enablePrecipitation || s.ND == 0) return;
  const Params& p = s.p;
  const auto& B = s.base[1]; const auto& WT = s.water[1];
  const float iterNum = (float)s.iter;
  const float resX = (float)s.W, resY = (float)s.H;
  auto texel = [&](float tx, float ty) {  // NEAREST + REPEAT
    int ix = wrapi((int)floorf(tx * resX), s.W), iy = wrapi((int)floorf(ty * resY), s.H);
    return (size_t)iy * s.W + ix;
  };
  const float lightningStart = s.lightning[START_ITERNUM];  // 1x1 lightningDataTexture (previous latch)
  for (int n = 0; n < s.ND; n++) {
    const float* d = &s.drops[src][(size_t)n * 5];
    const float dropX = d[0], dropY = d[1], massW = d[2], massI = d[3], density = d[4];
    float newPosX = dropX, newPosY = dropY, newMassW = massW, newMassI = massI, newDensity = density;
    float feedback[4] = {0, 0, 0, 0}, deposition[2] = {0, 0};  // canonical zero-init of the varyings
    bool isActive = true, spawned = false, lightningSpawned = false;
    float pointSize = 1.0f, glPosX = 0.0f, glPosY = 0.0f;
    float texCoordX = 0, texCoordY = 0, realTemp = 0;
    f4 base{}, water{};
    if (massW < 0.0f) {  // inactive :72
      texCoordX = random2d(massW, dropX + iterNum * 0.3754f);      // :82
      texCoordY = random2d(massI, dropX + iterNum * 0.073162f);
      size_t c = texel(texCoordX, texCoordY);
      base = ld4(B, c); water = ld4(WT, c);
      realTemp = potentialToRealT(s, base[TEMPERATURE], texCoordY);  // :90 (droplet's texCoord.y)
      const float initalMass = 0.15f;
      float threshold = (realTemp > CtoK(0.0f)) ? p.aboveZeroThreshold : p.subZeroThreshold;  // :94-98
      if (water[CLOUD] > threshold && base[TEMPERATURE] < 500.0f) {  // :100
        float spawnChance = ((water[CLOUD] - threshold) / (s.inactiveDroplets + 10.0f)) * resX * resY * p.spawnChanceMult;  // :105
        float t10 = water[CLOUD] * 10.0f;
        float nrmRand = gfract(t10 * t10);  // :109 pow(x,2) canonical x*x
        if (spawnChance > nrmRand) {
          spawned = true;
          newPosX = (texCoordX - 0.5f) * 2.0f; newPosY = (texCoordY - 0.5f) * 2.0f;  // :113
          if (realTemp < CtoK(0.0f)) {  // :115
            newMassW = 0.0f;
            newMassI = initalMass;
            feedback[HEAT] += newMassI * p.meltingHeat;
            newDensity = p.snowDensity;
            const float lightningCloudDensityThreshold = 2.5f, lightningChanceMultiplier = 0.0033f;
            float cloudPlusPrecipDensity = water[CLOUD] + water[PRECIPITATION];
            float lightningSpawnChance = gmax((cloudPlusPrecipDensity - lightningCloudDensityThreshold) * lightningChanceMultiplier, 0.0f);
            const float minIterationsSinceLastLightningBolt = 30.0f;
            if (lightningStart < iterNum - minIterationsSinceLastLightningBolt &&
                random2d(base[TEMPERATURE] * 0.2324f, water[TOTAL] * 7.7f) < lightningSpawnChance) {  // :132
              lightningSpawned = true;
              isActive = false;
              pointSize = 1.0f;
              feedback[0] = texCoordX; feedback[1] = texCoordY;
              feedback[START_ITERNUM] = iterNum;
              feedback[INTENSITY] = gclamp(cloudPlusPrecipDensity / 10.0f + (random2d(texCoordX, texCoordY) - 0.5f), 0.01f, 4.0f);
              glPosX = -1.0f + s.texelX * 3.0f; glPosY = -1.0f + s.texelY;  // pixel (1,0) :139
            }
          } else {  // :141-145
            newMassW = initalMass;
            newMassI = 0.0f;
            newDensity = 1.0f;
          }
          feedback[VAPOR] -= initalMass;  // :146 (also lands on START_ITERNUM of a lightning record)
        }
      }
      if (spawned) {
        if (!lightningSpawned) { pointSize = 1.0f; glPosX = newPosX; glPosY = newPosY; }
      } else {  // :154-161
        isActive = false;
        pointSize = 1.0f;
        feedback[MASS] = 1.0f;
        glPosX = -1.0f + s.texelX; glPosY = -1.0f + s.texelY;  // pixel (0,0)
      }
    }
    if (isActive) {  // :164
      if (!spawned) {
        texCoordX = dropX / 2.0f + 0.5f; texCoordY = dropY / 2.0f + 0.5f;
        size_t c = texel(texCoordX, texCoordY);
        water = ld4(WT, c); base = ld4(B, c);
        realTemp = potentialToRealT(s, base[TEMPERATURE], texCoordY);
      }
      float totalMass = newMassW + newMassI;
      if (totalMass < 0.04f) {  // :175-181
        feedback[HEAT] = -(totalMass * p.evapHeat);
        feedback[VAPOR] = totalMass;
        newMassW = -2.0f - dropX; newMassI = dropY;  // disableDroplet :60-64
      } else if (newPosY < -1.0f || water[TOTAL] > 1000.0f) {  // :183
        size_t ca = texel(texCoordX, texCoordY + s.texelY);
        if (ld4(B, ca)[TEMPERATURE] > 500.0f) newPosY += s.texelY * 1.0f;  // :185-186
        deposition[RAIN_DEPOSITION] = newMassW;
        deposition[SNOW_DEPOSITION] = newMassI;
        newMassW = -2.0f - dropX; newMassI = dropY;
      } else {  // update :193
        float surfaceArea = cbrt_canon(totalMass);  // :195
        float growthRate = gmax(map_range(realTemp, CtoK(0.0f), CtoK(-30.0f), p.growthRate0C, p.growthRate_30C), p.growthRate0C);  // :198
        float growth = water[CLOUD] * growthRate * surfaceArea;
        if (realTemp < CtoK(0.0f) && water[CLOUD] > 0.0f && density == 1.0f) growth += surfaceArea * water[PRECIPITATION] * 0.0030f;  // :205-207
        feedback[VAPOR] -= growth * 1.0f;
        if (realTemp < CtoK(0.0f)) {  // :212-220
          newMassI += growth;
          feedback[HEAT] += growth * p.meltingHeat;
          float freezing = gmin((CtoK(0.0f) - realTemp) * p.freezingRate * surfaceArea, newMassW);
          newMassW -= freezing;
          newMassI += freezing;
          feedback[HEAT] += freezing * p.meltingHeat;
        } else {  // :222-232
          newMassW += growth;
          float melting = gmin((realTemp - CtoK(0.0f)) * p.meltingRate * surfaceArea, newMassI);
          newMassI -= melting;
          newMassW += melting;
          feedback[HEAT] -= melting * p.meltingHeat;
          newDensity = gmin(newDensity + (melting / totalMass) * 1.00f, 1.0f);
        }
        float dropletTemp = potentialToRealT(s, base[TEMPERATURE], texCoordY);  // :234
        if (newMassI > 0.0f) dropletTemp = gmin(dropletTemp, CtoK(0.0f));
        float evapAndSubli = gmax((maxWater(dropletTemp) - water[TOTAL]) * surfaceArea * p.evapRate, 0.0f);  // :239
        float evap = gmin(newMassW, evapAndSubli);
        float subli = gmin(newMassI, evapAndSubli - evap);
        newMassW -= evap;
        newMassI -= subli;
        feedback[VAPOR] += evap;
        feedback[VAPOR] += subli;
        feedback[HEAT] -= evap * p.evapHeat;
        feedback[HEAT] -= subli * p.evapHeat;
        feedback[HEAT] -= subli * p.meltingHeat;
        newPosX += base[VX] / resX * 2.0f;  // :258
        newPosY += base[VY] / resY * 2.0f;
        newPosY -= p.fallSpeed * newDensity * sqrtf(totalMass / surfaceArea);  // :259
        newPosX = gmod(newPosX + 1.0f, 2.0f) - 1.0f;                          // :270
        feedback[MASS] = totalMass;
      }
      const float pntSize = 12.0f, pntSurface = pntSize * pntSize;  // :275-287
      feedback[MASS] /= pntSurface;
      feedback[HEAT] /= pntSurface;
      feedback[VAPOR] /= pntSurface;
      deposition[RAIN_DEPOSITION] /= pntSize;
      deposition[SNOW_DEPOSITION] /= pntSize;
      pointSize = pntSize;
      glPosX = newPosX; glPosY = newPosY;
    }
    float* o = &s.drops[dst][(size_t)n * 5];
    o[0] = newPosX; o[1] = newPosY; o[2] = newMassW; o[3] = newMassI; o[4] = gmax(newDensity, 0.0f);  // :290-292
    splat(s, glPosX, glPosY, pointSize, feedback, deposition);
  }
  s.last_drops = dst;
  if (s.iter % 600 == 0) s.inactiveDroplets = s.fb[0];  // app.js:5957-5967 texel (0,0).MASS
  // lightningLocationShader.frag:24-38
  const float* nl = &s.fb[4];  // texel (1,0)
  if (!(nl[START_ITERNUM] < gmax(iterNum - 1.0f, 1.0f) || nl[START_ITERNUM] > iterNum)) memcpy(s.lightning, nl, 16);
}

void iteration(Sim& s) {  // app.js:5830-6005
  pass_velocity(s);
  pass_curl(s);
  pass_vorticity(s);
  pass_boundary(s);
  pass_advection(s, false);
  pass_pressure(s);
  pass_lighting(s);
  pass_precipitation(s);
  s.iter++;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// C interface for ctypes (tests, bench cpu_baseline)
// ---------------------------------------------------------------------------------------------
extern "C" {

void* oracle_create(int W, int H, int n_droplets, int Wg, int x0) {
  Sim* s = new Sim();
  s->W = W; s->H = H; s->Wg = Wg > 0 ? Wg : W; s->x0 = x0; s->ND = n_droplets;
  size_t n = (size_t)W * H;
  for (int k = 0; k < 2; k++) {
    s->base[k].assign(n * 4, 0.0f); s->water[k].assign(n * 4, 0.0f); s->light[k].assign(n * 4, 0.0f);
    s->wall[k].assign(n * 4, 0); s->drops[k].assign((size_t)n_droplets * 5, 0.0f);
  }
  s->fb.assign(n * 4, 0.0f); s->dep.assign(n * 2, 0.0f); s->curl.assign(n, 0.0f); s->vort.assign(n * 2, 0.0f);
  s->initial_T.assign(H + 2, 0.0f); s->snd_T.assign(H + 2, 0.0f); s->snd_W.assign(H + 2, 0.0f); s->snd_Vel.assign(H + 2, 0.0f);
  memset(s->lightning, 0, sizeof(s->lightning));
  s->inactiveDroplets = 0.0f; s->iter = 0; s->even = true; s->last_drops = 0;
  memset(&s->p, 0, sizeof(s->p)); memset(&s->in, 0, sizeof(s->in));
  s->in.userInputType = -1;
  s->texelX = (float)(1.0 / (double)s->Wg); s->texelY = (float)(1.0 / (double)H);
  return s;
}
void oracle_destroy(void* h) { delete (Sim*)h; }

// app.js:5189-5234, 4917-4962: both copies get the same data; the rest restarts at zero.
void oracle_upload(void* h, const float* base, const float* water, const int8_t* wall, const float* drops) {
  Sim& s = *(Sim*)h;
  size_t n = (size_t)s.W * s.H;
  for (int k = 0; k < 2; k++) {
    memcpy(s.base[k].data(), base, n * 16); memcpy(s.water[k].data(), water, n * 16); memcpy(s.wall[k].data(), wall, n * 4);
    if (drops && s.ND) memcpy(s.drops[k].data(), drops, (size_t)s.ND * 20);
    std::fill(s.light[k].begin(), s.light[k].end(), 0.0f);
  }
  std::fill(s.fb.begin(), s.fb.end(), 0.0f); std::fill(s.dep.begin(), s.dep.end(), 0.0f);
  std::fill(s.curl.begin(), s.curl.end(), 0.0f); std::fill(s.vort.begin(), s.vort.end(), 0.0f);
  memset(s.lightning, 0, sizeof(s.lightning));
  s.inactiveDroplets = 0.0f; s.iter = 0; s.even = true; s.last_drops = 0;
}
void oracle_set_params(void* h, const void* p) { memcpy(&((Sim*)h)->p, p, sizeof(Params)); }
void oracle_set_frame_inputs(void* h, const void* in) { memcpy(&((Sim*)h)->in, in, sizeof(FrameInputs)); }
void oracle_set_profiles(void* h, const float* T0, const float* sT, const float* sW, const float* sV) {
  Sim& s = *(Sim*)h;
  size_t n = (size_t)s.H + 1;
  if (T0) memcpy(s.initial_T.data(), T0, n * 4);
  if (sT) memcpy(s.snd_T.data(), sT, n * 4); else std::fill(s.snd_T.begin(), s.snd_T.end(), 0.0f);
  if (sW) memcpy(s.snd_W.data(), sW, n * 4); else std::fill(s.snd_W.begin(), s.snd_W.end(), 0.0f);
  if (sV) memcpy(s.snd_Vel.data(), sV, n * 4); else std::fill(s.snd_Vel.begin(), s.snd_Vel.end(), 0.0f);
}
void oracle_step(void* h, int n) { Sim& s = *(Sim*)h; for (int i = 0; i < n; i++) iteration(s); }
// velocity -> advection(base only) -> pressure: the dry sweep of BASELINE config 2.
void oracle_step_dry(void* h, int n) {
  Sim& s = *(Sim*)h;
  // the reference's velocity pass renders into frameBuff_1 and advection samples frameBuff_0
  // (the boundary pass sits between them, app.js:5876); with that pass left out the dry sweep
  // hands velocity's output over unchanged.
  for (int i = 0; i < n; i++) {
    pass_velocity(s);
    s.base[0] = s.base[1]; s.wall[0] = s.wall[1];
    pass_advection(s, true);
    pass_pressure(s);
    s.iter++;
  }
}
// pass ids as WSB_PASS_* in include/wsb200.h
void oracle_run_pass(void* h, int pass) {
  Sim& s = *(Sim*)h;
  switch (pass) {
    case 0: pass_velocity(s); break;
    case 1: pass_curl(s); break;
    case 2: pass_vorticity(s); break;
    case 3: pass_boundary(s); break;
    case 4: pass_advection(s, false); break;
    case 5: pass_pressure(s); break;
    case 6: pass_lighting(s); break;
    case 7: pass_precipitation(s); break;
    case 8: s.iter++; break;
    case 9: pass_advection(s, true); break;
  }
}
// raw access: field ids as WSB_FIELD_*, buf selects the ping-pong copy
float* oracle_field_f32(void* h, int field, int buf) {
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
    case 9: return s.lightning;
  }
  return nullptr;
}
int8_t* oracle_field_i8(void* h, int buf) { return ((Sim*)h)->wall[buf].data(); }
long oracle_get_iter(void* h) { return ((Sim*)h)->iter; }
void oracle_set_iter(void* h, long it) { ((Sim*)h)->iter = it; }
int oracle_get_even(void* h) { return ((Sim*)h)->even ? 1 : 0; }
int oracle_last_drops(void* h) { return ((Sim*)h)->last_drops; }
float oracle_get_inactive(void* h) { return ((Sim*)h)->inactiveDroplets; }
void oracle_set_inactive(void* h, float v) { ((Sim*)h)->inactiveDroplets = v; }

// OpenMP team size (torchrun exports OMP_NUM_THREADS=1; the CPU baseline wants every core)
void oracle_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }
int oracle_get_threads() { return omp_get_max_threads(); }

// unit-test hooks for the helpers
float oracle_maxWater(float T) { return maxWater(T); }
float oracle_IR_emitted(float T) { return IR_emitted(T); }
float oracle_cbrt(float m) { return cbrt_canon(m); }
float oracle_random2d(float a, float b) { return random2d(a, b); }
float oracle_map_rangeC(float v, float a, float b, float c, float d) { return map_rangeC(v, a, b, c, d); }
uint32_t oracle_hash(uint32_t x) { return hash_u(x); }

}  // extern "C"
