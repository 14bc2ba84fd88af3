// wsb_math.cuh — device math for the simulation kernels.
//
// Numerical contract ("spec freeze"): fp32, one rounding per written operation — the
// library is compiled with -fmad=false, IEEE division and square root — and every function the
// GLSL source leaves to the driver's libm is pinned to a closed form:
//   pow(x, 17) -> multiply chain, pow(x, 4) -> (x*x)*(x*x), pow(x, .5) -> sqrt, pow(x, 2) -> x*x,
//   pow(x, 1/3) -> wsb_cbrt (bit guess + 4 Newton steps), sin/cos(uniform) -> evaluated on the
//   host in double.
// References: shaders/common.glsl (reference checkout), cited per function.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace wsb {

// channel indices, common.glsl:42-95
enum { VX = 0, VY = 1, PRESSURE = 2, TEMPERATURE = 3 };
enum { WALLTYPE_INERT = 0, WALLTYPE_LAND = 1, WALLTYPE_WATER = 2, WALLTYPE_FIRE = 3,
       WALLTYPE_URBAN = 4, WALLTYPE_RUNWAY = 5, WALLTYPE_INDUSTRIAL = 6 };

// common.glsl:9-35
#define WSB_lightHeatingConst 0.000002f
#define WSB_maxWaterTemp 40.0f
#define WSB_waterHeatExchangeRate 0.0002f
#define WSB_waterHeatCapacity 50.0f
#define WSB_fullWhiteSnowHeight 10.0f
#define WSB_snowMassToHeight 0.05f
#define WSB_snowMeltRate 0.000015f
#define WSB_ALBEDO_SNOW 0.85f
#define WSB_ALBEDO_SNOW_FOREST 0.30f
#define WSB_ALBEDO_FOREST 0.10f
#define WSB_ALBEDO_DRYSOIL 0.30f
#define WSB_ALBEDO_WETSOIL 0.15f
#define WSB_ALBEDO_URBAN 0.08f
#define WSB_ALBEDO_INDUSTRIAL 0.08f
#define WSB_ALBEDO_RUNWAY 0.04f
#define WSB_ALBEDO_WATER 0.05f

// GLSL max/min semantics: max(x,y) = (x < y) ? y : x ; min(x,y) = (y < x) ? y : x
__device__ __forceinline__ float gmax(float a, float b) { return a < b ? b : a; }
__device__ __forceinline__ float gmin(float a, float b) { return b < a ? b : a; }
__device__ __forceinline__ float gclamp(float x, float lo, float hi) { return gmin(gmax(x, lo), hi); }
#ifdef WSB_EXP_FMAMIX  // timing experiment only (NOT the frozen semantics: results differ from the oracle)
__device__ __forceinline__ float gmix(float a, float b, float t) { return __fmaf_rn(t, b, __fmaf_rn(-t, a, a)); }
#else
__device__ __forceinline__ float gmix(float a, float b, float t) { return a * (1.0f - t) + b * t; }
#endif
__device__ __forceinline__ float gfract(float x) { return x - floorf(x); }
__device__ __forceinline__ float gmod(float x, float y) { return x - y * floorf(x / y); }
__device__ __forceinline__ float glength(float x, float y) { return sqrtf(x * x + y * y); }
__device__ __forceinline__ float gsmoothstep(float e0, float e1, float x) {
  float t = gclamp((x - e0) / (e1 - e0), 0.0f, 1.0f);
  return t * t * (3.0f - 2.0f * t);
}
__device__ __forceinline__ float4 mix4(float4 a, float4 b, float t) {
  return make_float4(gmix(a.x, b.x, t), gmix(a.y, b.y, t), gmix(a.z, b.z, t), gmix(a.w, b.w, t));
}

// common.glsl:99-101
__device__ __forceinline__ float map_range(float v, float min1, float max1, float min2, float max2) {
  return min2 + (v - min1) * (max2 - min2) / (max1 - min1);
}
__device__ __forceinline__ float map_rangeC(float v, float min1, float max1, float min2, float max2) {
  return gclamp(map_range(v, min1, max1, min2, max2), gmin(min2, max2), gmax(min2, max2));
}

// common.glsl:103-137
__device__ __forceinline__ uint32_t hash_u(uint32_t x) {
  x += (x << 10u);
  x ^= (x >> 6u);
  x += (x << 3u);
  x ^= (x >> 11u);
  x += (x << 15u);
  return x;
}
__device__ __forceinline__ float random2d(float sx, float sy) {
  uint32_t h = hash_u(__float_as_uint(sx) + hash_u(__float_as_uint(sy)));
  h &= 0x007FFFFFu;
  h |= 0x3F800000u;
  return gmod(__uint_as_float(h), 1.0f);
}

__device__ __forceinline__ float CtoK(float c) { return c + 273.15f; }  // common.glsl:157
__device__ __forceinline__ float KtoC(float k) { return k - 273.15f; }  // common.glsl:159

// common.glsl:177-180
__device__ __forceinline__ float maxWater(float T) {
  float x = T / 250.0f;
  float x2 = x * x, x4 = x2 * x2, x8 = x4 * x4, x16 = x8 * x8;
  return x16 * x;
}
// common.glsl:258-261
__device__ __forceinline__ float IR_emitted(float T) {
  float x = T * 0.01f;
  float x2 = x * x;
  return (x2 * x2) * 5.670374419f;
}
// pow(m, 1/3), precipitationShader.vert:195
__device__ __forceinline__ float wsb_cbrt(float m) {
  if (!(m > 0.0f)) return 0.0f;
  float y = __uint_as_float(__float_as_uint(m) / 3u + 709921077u);
#pragma unroll
  for (int k = 0; k < 4; k++) y = y - (y - m / (y * y)) * (1.0f / 3.0f);
  return y;
}

// the same texel as a packed word when every component is known to be in [-128, 127] already (no
// wall tool is active: type, distance and vegetation only take values a texel held before)
__device__ __forceinline__ int pack_wall_in_range(int t, int d, int v, int g) {
  return (t & 0xff) | ((d & 0xff) << 8) | ((v & 0xff) << 16) | (g << 24);
}
// RGBA8I store, canonical saturation to [-128,127]
__device__ __forceinline__ char4 pack_wall(int t, int d, int v, int g) {
  return make_char4((signed char)min(max(t, -128), 127), (signed char)min(max(d, -128), 127),
                    (signed char)min(max(v, -128), 127), (signed char)min(max(g, -128), 127));
}

}  // namespace wsb
