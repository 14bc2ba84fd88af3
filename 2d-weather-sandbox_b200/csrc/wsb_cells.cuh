// wsb_cells.cuh — per-cell bodies of the simulation passes, written once and instantiated with
// different "contexts": a context answers neighbour fetches either straight from HBM/L2 (the
// one-kernel-per-pass REFERENCE schedule) or from the shared-memory tiles of the fused kernels.
//
// Coordinates handed to a context are LOCAL array coordinates (x = column of this rank's padded
// strip, y = row); the context applies the periodic wrap of the reference's REPEAT textures
// (app.js:5191-5230).  Reference citations: shaders/fragment/*.frag, shaders/common.glsl.
#pragma once
#include "../../include/wsb200.h"
#include "wsb_math.cuh"

namespace wsb {

struct Geom {
  int Wg, H;      // global width, height
  int pitch;      // columns of the local arrays (owned + 2*ghost)
  int gx0;        // global x of local column 0 (negative in the left ghost zone of rank 0)
  int wrap;       // 1: single domain, local x indices wrap modulo pitch (== Wg)
  int cx0, cx1;   // local columns this launch computes ...
  int cxGapAt, cxGapLen;  // ... minus [cxGapAt, cxGapAt + cxGapLen): one launch covers the two edge tile columns of a strip
  int ox0, ox1;   // owned columns (without the ghost zones): max |v| is only taken over these
  float texelX, texelY;    // uniform texelSize: (float)(1.0/W), (float)(1.0/H)   app.js:5436
  float ltexelX, ltexelY;  // advectionShader.frag:69  vec2(1.)/resolution in fp32
  float Hf, Wf;
  float cellHeightComp;    // lightingShader.frag:44  300. / resolution.y — uniform-only, divided once on the host
  float nearV;             // fused kernels: |v| bound of the near back-trace (wsb_fused_kernels.cuh: near_tap); -1 = never
};

struct DevParams {
  wsb_params p;
  wsb_frame_inputs in;
  float sinSun, cosSun;  // of in.sunAngle, evaluated on the host in double
  float iterNum;         // uniform float iterNum
  int iterI;             // int(iterNum)
};

__device__ __forceinline__ int wrap_y(int y, int H) { return y < 0 ? y + H : (y >= H ? y - H : y); }
__device__ __forceinline__ int mod_i(int i, int n) { i %= n; return i < 0 ? i + n : i; }
// x index for small offsets (|offset| <= pitch)
__device__ __forceinline__ int wrap_x(const Geom& g, int x) {
  if (g.wrap) return x < 0 ? x + g.pitch : (x >= g.pitch ? x - g.pitch : x);
  return min(max(x, 0), g.pitch - 1);  // strips: stay in bounds; only ghost-edge garbage differs
}
// x index for data-dependent gathers (any offset)
__device__ __forceinline__ int gather_x(const Geom& g, int x) {
  if (g.wrap) return mod_i(x, g.pitch);
  return min(max(x, 0), g.pitch - 1);
}
// first local column of the tile column blockIdx.x of this launch
__device__ __forceinline__ int tile_col0(const Geom& g, int bx, int tileW) {
  int c0 = g.cx0 + bx * tileW;
  if (c0 >= g.cxGapAt) c0 += g.cxGapLen;
  return c0;
}
__device__ __forceinline__ int global_x(const Geom& g, int lx) {
  int gx = g.gx0 + lx;
  return gx < 0 ? gx + g.Wg : (gx >= g.Wg ? gx - g.Wg : gx);
}
__device__ __forceinline__ float potentialToRealT(const DevParams& d, float pot, float texCoordY) {
  return pot - texCoordY * d.p.dryLapse;  // common.glsl:151-153
}

#if defined(__CUDACC__) || defined(WSB_HOST_EMU)  // warp / CTA collectives: device, or the host emulation of tests/host_cells (the cell bodies below compile for the plain host too)
// Largest |v| component seen by advection since the upload — a RUNNING maximum, so the global
// atomic is only issued by whoever holds something larger than the current value (one L2 read
// otherwise).  Measured (profiles/r2_vmax_atomics.md): an unconditional same-address atomicMax per
// warp — 524 288 per launch at 16384 x 4096 — runs at ~0.43 atomics/ns and, depending on which L2
// slice the word lands in, set the whole kernel's duration (1.22 ms instead of 0.63 ms).
// Non-negative floats order like their bit patterns.  Safe with partially active warps.
__device__ __forceinline__ void report_vmax(float vm, unsigned* __restrict__ maxv) {
#ifdef WSB_EXP_NO_VMAX  // timing experiment only
  return;
#endif
  const unsigned m = __activemask();
  const unsigned r = __reduce_max_sync(m, __float_as_uint(vm));
  if ((threadIdx.x + threadIdx.y * blockDim.x) % 32 == (unsigned)(__ffs(m) - 1) && r > __ldcg(maxv)) atomicMax(maxv, r);
}
// CTA-wide variant for the fused kernels: warps combine in shared memory (`sMax`, zeroed by the
// caller before an earlier __syncthreads), thread 0 talks to global memory.  Every thread of the
// CTA must call it.
__device__ __forceinline__ void report_vmax_cta(float vm, unsigned* __restrict__ maxv, unsigned* sMax) {
#ifdef WSB_EXP_NO_VMAX
  return;
#endif
  const unsigned r = __reduce_max_sync(0xffffffffu, __float_as_uint(vm));
  if ((threadIdx.x & 31) == 0 && r != 0u) atomicMax(sMax, r);
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned v = *sMax;
    if (v > __ldcg(maxv)) atomicMax(maxv, v);
  }
}
#endif  // __CUDACC__ || WSB_HOST_EMU

// ---------------------------------------------------------------------------------------------
// velocityShader.frag:32-62
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void velocity_cell(const DevParams& d, float& vx, float& vy, float P, float PXp,
                                              float PYp, int wallDist) {
  if (wallDist == 0) {
    vx = 0.0f;
    vy = 0.0f;
  } else {
    vx += P - PXp;
    vy += P - PYp;
    vx *= 1.0f - d.p.dragMultiplier * 0.0002f;
    vy *= 1.0f - d.p.dragMultiplier * 0.0002f;
    vx += d.p.wind * 0.000001f;
  }
}

// curlShader.frag:12-19
__device__ __forceinline__ float curl_cell(float vx, float vy, float vxYp, float vyXp) { return vxYp - vx - vyXp + vy; }

// vorticityShader.frag:19-38
__device__ __forceinline__ float2 vorticity_cell(float c, float cXm, float cYm, float cXp, float cYp) {
  float fx = fabsf(cYm) - fabsf(cYp);
  float fy = fabsf(cXp) - fabsf(cXm);
  float magnitude = glength(fx, fy) + 0.0001f;
  fx /= magnitude;
  fy /= magnitude;
  fx *= c;
  fy *= c;
  return make_float2(fx, fy);
}

// pressureShader.frag:16-43 ; returns new (P, T)
__device__ __forceinline__ void pressure_cell(float vx, float vy, float& P, float& T, float vxXm, float vyYm,
                                              float TYm, int wallYmType, int wallYmDist) {
  if (wallYmDist == 0 && wallYmType == 1) T -= TYm - 1000.0f;
  P += (vxXm - vx + vyYm - vy) * 0.45f;
}

// ---------------------------------------------------------------------------------------------
// boundaryShader.frag:72-531
//   Ctx is positioned at the cell and answers fetches by OFFSET (dx, dy in {-1,0,1}), so a tiled
//   context turns them into compile-time-constant shared-memory offsets:
//        float bx/by/bt(dx,dy) of the post-velocity base; float4 base4(dx,dy); float4 water4(dx,dy);
//        char4 wall4(dx,dy); float2 vort(dx,dy); float4 light4(dx,dy) (y clamped); float4 fb4();
//        float2 dep2()
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float calcEvaporation(const DevParams& d, float T, float W, float V, float M) {
  return gmax((maxWater(T) - W) * d.p.landEvaporation * (V / 127.0f + 0.1f) * gmin(M + 1.0f, 50.0f) * 0.05f, 0.0f);
}
__device__ __forceinline__ float calcFireIntensity(int veg, float moist, float precip) {
  return gmax((float)veg * 0.00025f - moist * 0.00020f - precip * 0.02f, 0.0f);
}

template <class C>
__device__ void boundary_cell(const C& c, const Geom& g, const DevParams& d, const float* __restrict__ initial_T,
                              int x, int y, float4& base, float4& water, char4& wallOut) {
  const float texCoordY = ((float)y + 0.5f) * g.texelY;
  const float texCoordYp = texCoordY + g.texelY;
  base = c.base4(0, 0);
  water = c.water4(0, 0);
  const float4 fb = c.fb4();
  const float realTemp = potentialToRealT(d, base.w, texCoordY);
  const char4 w0 = c.wall4(0, 0);
  const char4 wXm = c.wall4(-1, 0), wYm = c.wall4(0, -1), wXp = c.wall4(1, 0), wYp = c.wall4(0, 1);
  const float4 light = c.light4(0, 0);
  int wType = w0.x, wDist = w0.y, wVert, wVeg = w0.w;
  bool nextToWall = false;
  wVert = (int)wYm.z + 1;  // :92

  if (wDist != 0) {  // fluid :94
    wType = wYm.x;   // :96
    if (wType != WALLTYPE_WATER) base.w += light.y;  // NET_HEATING :98
    base.w += fb.y;                                  // HEAT :101
    float precipCoalescence = gmax(-fb.z, 0.0f);     // :104
    water.y -= precipCoalescence;
    water.x -= precipCoalescence;
    float precipEvaporation = gmax(fb.z, 0.0f);
    water.x += precipEvaporation;
    water.z = gmax(water.z * 0.997f - 0.00001f + fb.x * 0.005f, 0.0f);               // :115
    {  // :119  smoke is washed out by precipitation; x / 1.0f == x exactly, so the IEEE division
       // sequence is skipped wherever no precipitation feedback arrived (almost every cell)
      const float washout = 1.0f + gmax(-fb.z * 0.1f, 0.0f) + fb.x * 0.000f;
      if (washout != 1.0f) water.w /= washout;
    }
    water.w -= fb.x * 0.0001f;                                                         // :121
    water.w -= gmax((water.w - 4.0f) * 0.01f, 0.0f);                                   // :124
    water.w = gmax(water.w, 0.0f);                                                     // :126
    if (water.w > 4.0f) water.w -= water.z * 0.02f;                                    // :128

    // gravity :132-148
    const float gravMult = 0.0001f;
    const float TYp = c.bt(0, 1);
    float gravityForce = ((base.w + TYp) * 0.5f - (initial_T[y] + initial_T[y + 1]) * 0.5f) * gravMult;
    gravityForce -= water.y * gravMult * d.p.waterWeight;
    gravityForce -= fb.x * gravMult * d.p.waterWeight;
    base.y += gravityForce;

    float snowCover = 0.0f, soilMoisture = 0.0f;
    if (wYm.y == 0) {  // :155
      nextToWall = true;
      wDist = 1;
      float4 waterYm = c.water4(0, -1);
      snowCover = waterYm.w;
      soilMoisture = waterYm.z;
      wVert = 1;
    }
    if (wXm.y == 0) {  // :166
      nextToWall = true;
      wDist = 1;
      if (wXm.x == WALLTYPE_WATER) { wType = WALLTYPE_LAND; wDist = 0; }
      if (wXp.y == 0) wDist = 0;
    } else if (wXp.y == 0) {  // :178
      nextToWall = true;
      wDist = 1;
      if (wXp.x == WALLTYPE_WATER) { wType = WALLTYPE_LAND; wDist = 0; }
    }
    if (wYp.y == 0) {  // :188
      nextToWall = true;
      wDist = 1;
      if (texCoordY < 0.99f) wDist = 0;
    }

    // vorticity confinement :201-208
    {
      const float2 vf = c.vort(0, 0);
      const float2 vfXm = c.vort(-1, 0);
      const float2 vfYm = c.vort(0, -1);
      float velocityFactor = glength(base.x, base.y) * 0.1f;
      float k = d.p.vorticity + velocityFactor;
      base.x += (vf.x + vfYm.x) * k;
      base.y += (vf.y + vfXm.y) * k;
    }

    if (nextToWall) {  // :211-243
      if (wType != WALLTYPE_WATER) {
        float lightPower = 0.0f;
        if (wYm.y == 0) lightPower += gmax(light.x * d.cosSun, 0.0f);
        if (wXm.y == 0) lightPower += gmax(light.x * d.sinSun, 0.0f);
        if (wXp.y == 0) lightPower += gmax(light.x * (-d.sinSun), 0.0f);
        float albedoTotal = 1.0f;
        if (wType == WALLTYPE_LAND || wType == WALLTYPE_FIRE) {
          float albedoSoil = map_rangeC(soilMoisture, 0.0f, 20.0f, WSB_ALBEDO_DRYSOIL, WSB_ALBEDO_WETSOIL);
          albedoSoil = map_rangeC(snowCover, 0.0f, WSB_fullWhiteSnowHeight, albedoSoil, WSB_ALBEDO_SNOW);
          float fullVegetationAlbedo = map_range(snowCover, 0.0f, WSB_fullWhiteSnowHeight, WSB_ALBEDO_FOREST, WSB_ALBEDO_SNOW_FOREST);
          albedoTotal = map_range((float)wYm.w, 0.0f, 127.0f, albedoSoil, fullVegetationAlbedo);
        } else if (wType == WALLTYPE_URBAN) {
          albedoTotal = WSB_ALBEDO_URBAN;
        } else if (wType == WALLTYPE_INDUSTRIAL) {
          albedoTotal = WSB_ALBEDO_INDUSTRIAL;
        } else if (wType == WALLTYPE_RUNWAY) {
          albedoTotal = WSB_ALBEDO_RUNWAY;
        }
        lightPower *= (1.0f - albedoTotal);
        lightPower *= WSB_lightHeatingConst;
        base.w += lightPower;
      }
    } else {  // :245-269
      int nearest = 255;
      if (wYm.y < nearest) nearest = wYm.y;
      if (wYp.y < nearest) nearest = wYp.y;
      if (wXm.y < nearest) nearest = wXm.y;
      if (wXp.y < nearest) nearest = wXp.y;
      wDist = nearest + 1;
    }

    if (wVert <= 5) {  // :273-303
      if (wVert == 1) {
        float surfaceDrag = 0.0015f;
        if (wType == WALLTYPE_URBAN)
          surfaceDrag = 0.040f;
        else if (wType == WALLTYPE_LAND || wType == WALLTYPE_FIRE)
          surfaceDrag = map_rangeC((float)wVeg, 50.0f, 127.0f, 0.0015f, 0.020f);
        base.x -= fabsf(base.x) * base.x * surfaceDrag * 50.0f;
      }
      const float exchangeRate = 0.015f;
      if (wYp.z <= 5) base.x -= (base.x - c.bx(0, 1)) * exchangeRate;
      if (wYm.z > 0) base.x -= (base.x - c.bx(0, -1)) * exchangeRate;
    }

    if (wVert <= 8) {  // :305-372
      wVeg = wYm.w;
      const float4 waterInSurface = c.water4(0, -1);
      const int t = wType;
      bool in = false;
      if (t == WALLTYPE_FIRE) {
        in = true;
        if (wVert == 1) {
          float fireIntensity = calcFireIntensity(wVeg, waterInSurface.z, water.z);
          fireIntensity = gmax(fireIntensity, 0.0f);
          base.w += fireIntensity;
          water.w += fireIntensity * 2.0f;
          water.x += fireIntensity * 0.50f;
        }
      }
      if (in || t == WALLTYPE_INDUSTRIAL) {
        in = true;
        if (wType == WALLTYPE_INDUSTRIAL) {
          int texFragX = (int)((float)global_x(g, x) + 0.5f) % 80;
          if (wVert == 5 && (texFragX == 18 || texFragX == 22)) {
            water.x += 0.25f;
            base.x *= 0.5f;
            base.y *= 0.5f;
            base.y += 0.05f;
          } else if (wVert == 6 && texFragX == 29) {
            water.w += 0.01f;
            base.w += 0.02f;
            base.x *= 0.5f;
            base.y *= 0.5f;
          }
        }
      }
      if (in || t == WALLTYPE_URBAN) {
        in = true;
        water.w += 0.000002f;
      }
      if (in || t == WALLTYPE_LAND) {
        if (wVert <= 1) {
          float evaporation = calcEvaporation(d, realTemp, water.x, (float)wVeg, waterInSurface.z) / 1.0f;
          water.x += evaporation;
          base.w -= evaporation * d.p.evapHeat * 0.5f;
          if (wVeg < 10 && water.z < 5.0f) water.w = gmin(water.w + (gmax(fabsf(base.x) - 0.12f, 0.0f) * 0.15f), 2.4f);
        }
      } else if (t == WALLTYPE_WATER) {
        if (wVert <= 1) {
          float LocalWaterTemperature = c.bt(0, -1);
          base.w += (LocalWaterTemperature - realTemp - 1.0f) / 1.0f * WSB_waterHeatExchangeRate;
          water.x += gmax((maxWater(LocalWaterTemperature) - water.x) * d.p.waterEvaporation / 1.0f, 0.0f);
        }
      }
    }
  } else {  // wall :373
    wVert = (int)wYp.z - 1;
    if (wVert < 0) {  // :377
      float4 wtYp = c.water4(0, 1);
      water.z = wtYp.z;
      water.w = wtYp.w;
      wVeg = wYp.w;
      if (wYp.y == 0) {
        if (wYp.x != WALLTYPE_WATER) {
          wType = wYp.x;
        } else if (wType == WALLTYPE_WATER) {
          base.w = c.bt(0, 1);
        }
      }
    } else if (wVert == 0) {  // surface layer :390
      const float4 waterYp = c.water4(0, 1);
      const float2 dep = c.dep2();
      const float4 lightAbove = c.light4(0, 1);
      const int t = wType;
      bool in = false;
      if (t == WALLTYPE_INDUSTRIAL) { in = true; wVeg = min(wVeg, 15); }
      if (in || t == WALLTYPE_URBAN) { in = true; wVeg = min(wVeg, 75); }
      if (in || t == WALLTYPE_FIRE) {
        in = true;
        if (wType == WALLTYPE_FIRE) {
          float fireIntensity = calcFireIntensity(wVeg, water.z, waterYp.z);
          if (fireIntensity < 0.002f) {
            wType = WALLTYPE_LAND;
          } else if (d.iterI % ((int)(10.0f / fireIntensity) + 1) == 0) {
            wVeg -= 1;
            if (wVeg < 10) wType = WALLTYPE_LAND;
          }
        }
      }
      if (in || t == WALLTYPE_LAND) {  // :415-475
        water.z = gclamp(water.z + dep.x * 0.1f, 0.0f, 1000.0f);
        water.w = gclamp(water.w + dep.y * WSB_snowMassToHeight, 0.0f, 4000.0f);
        const float TAbove = c.bt(0, 1);
        float realTempAboveSurface = potentialToRealT(d, TAbove, texCoordYp);
        float evaporation = calcEvaporation(d, realTempAboveSurface, waterYp.x, (float)wVeg, water.z) * 0.10f;
        water.z -= evaporation;
        if (d.iterI % 100 == 0) {
          float numNeighbors = 0.0f, totalNeighborSnow = 0.0f, totalNeighborSoilMoisture = 0.0f;
          if (wXm.z == 0 && (wXm.x == WALLTYPE_LAND || wXm.x == WALLTYPE_URBAN)) {
            float4 wn = c.water4(-1, 0);
            totalNeighborSnow += wn.w;
            totalNeighborSoilMoisture += wn.z;
            numNeighbors += 1.0f;
          }
          if (wXp.z == 0 && (wXp.x == WALLTYPE_LAND || wXp.x == WALLTYPE_URBAN)) {
            float4 wn = c.water4(1, 0);
            totalNeighborSnow += wn.w;
            totalNeighborSoilMoisture += wn.z;
            numNeighbors += 1.0f;
          }
          if (numNeighbors > 0.0f) {
            float avgNeighborSnow = totalNeighborSnow / numNeighbors;
            water.w += (avgNeighborSnow - water.w) * 0.02f;
            float avgNeighborSoilMoisture = totalNeighborSoilMoisture / numNeighbors;
            water.z += (avgNeighborSoilMoisture - water.z) * 0.02f;
          }
          int vegetationGrowthRate = (int)(water.z * sqrtf(lightAbove.x) * 0.01f);
          // growth rates above 100 make the reference's interval (100 / rate) * 100 zero and its `%` undefined
          // (GLSL ES 3.00 5.9); frozen as "no growth tick" (DESIGN 2), identically in the oracle
          const int growthInterval = vegetationGrowthRate > 0 ? (100 / vegetationGrowthRate) * 100 : 0;
          if (growthInterval > 0 && d.iterI % growthInterval == 0) {
            if ((int)map_rangeC(realTempAboveSurface, CtoK(0.0f), CtoK(25.0f), 0.0f, 127.0f) > wVeg) wVeg += 1;
          }
          int subInterval = d.iterI / 100;
          if (subInterval % ((int)(water.z * 0.1f + water.w * 0.5f) + 10) == 0 && wVeg >= 20 &&
              (wXm.x == WALLTYPE_FIRE || wXp.x == WALLTYPE_FIRE || waterYp.w > 4.5f)) {
            wType = WALLTYPE_FIRE;
          }
        }
      } else if (t == WALLTYPE_WATER) {  // :476-527
        const float waterTempUpdateInterval = 20.0f;
        if (d.p.dynamicWaterTemperature >= 1.0f && gmod(d.iterNum, waterTempUpdateInterval) < 0.5f) {
          float numNeighbors = 0.0f, totalNeighborTemp = 0.0f;
          if (wXm.x == WALLTYPE_WATER) { totalNeighborTemp += c.bt(-1, 0); numNeighbors += 1.0f; }
          if (wXp.x == WALLTYPE_WATER) { totalNeighborTemp += c.bt(1, 0); numNeighbors += 1.0f; }
          if (numNeighbors > 0.0f) {
            float avgNeighborTemp = totalNeighborTemp / numNeighbors;
            base.w += (avgNeighborTemp - base.w) * 0.10f;
          }
          if (base.w > 500.0f) base.w = CtoK(25.0f);
          float airTemperature = potentialToRealT(d, c.bt(0, 1), texCoordYp);
          float netWaterHeating = 0.0f;
          netWaterHeating += (airTemperature - base.w) * WSB_waterHeatExchangeRate;
          netWaterHeating -= gmax((maxWater(base.w) - waterYp.x) * d.p.waterEvaporation, 0.0f) * d.p.evapHeat * 0.5f;
          float lightPower = gmax(lightAbove.x * d.cosSun, 0.0f);
          lightPower *= (1.0f - WSB_ALBEDO_WATER);
          lightPower *= WSB_lightHeatingConst;
          netWaterHeating += lightPower;
          netWaterHeating += lightAbove.y;
          base.w += netWaterHeating / WSB_waterHeatCapacity * waterTempUpdateInterval;
        }
        base.w = gclamp(base.w, CtoK(0.0f), CtoK(WSB_maxWaterTemp));
        wVeg = 20;
        water.z = 100.0f;
        water.w = 0.0f;
      }
    }
  }
  wallOut = pack_wall(wType, wDist, wVert, wVeg);
}

// ---------------------------------------------------------------------------------------------
// advectionShader.frag:65-458
//   Ctx: float bx/by/bp/bt(x,y); float4 base4(x,y); float wt0/wt1/wt2/wt3(x,y); float4 water4(x,y);
//        char4 wall4(x,y); int wdist(x,y) — callable with ANY coordinates (data-dependent
//        back-trace), the context wraps; sbx/sby/sbt/swdist(x,y): the same fetches for the fixed
//        +-1 stencil, which a tiled context can serve without a bounds check.
//   DRY = the "dry sweep" of BASELINE config 2: only the base field is advected.
// ---------------------------------------------------------------------------------------------
struct BilerpSetup { int ix, iy; float fx, fy; };
__device__ __forceinline__ BilerpSetup bilerp_setup(float posx, float posy) {  // common.glsl:196-199
  float stx = posx - 0.5f, sty = posy - 0.5f;
  float flx = floorf(stx), fly = floorf(sty);
  BilerpSetup b;
  b.ix = (int)flx; b.iy = (int)fly; b.fx = stx - flx; b.fy = sty - fly;
  return b;
}
struct WallMix { float ab, cd, abcd; };
__device__ __forceinline__ WallMix wall_mix(int wa, int wb, int wc, int wd, float fx, float fy) {  // common.glsl:234-251
  WallMix m;
  m.ab = fx; m.cd = fx; m.abcd = fy;
  if (wa == 0) m.ab = 1.0f; else if (wb == 0) m.ab = 0.0f;
  if (wc == 0) m.cd = 1.0f; else if (wd == 0) m.cd = 0.0f;
  if (wa == 0 && wb == 0) m.abcd = 1.0f; else if (wc == 0 && wd == 0) m.abcd = 0.0f;
  return m;
}
__device__ __forceinline__ float mix2d(float a, float b, float cc, float dd, float mab, float mcd, float mabcd) {
  return gmix(gmix(a, b, mab), gmix(cc, dd, mcd), mabcd);
}
__device__ __forceinline__ float absHorizontalDist(float a, float b) {  // common.glsl:268-271
  return gmin(gmin(fabsf(a - b), fabsf(1.0f + a - b)), 1.0f - a + b);
}

// :85-89 velocities at the three staggered points of the cell
struct AdvVel { float Px, Py, Vxx, Vxy, Vyx, Vyy; };
__device__ __forceinline__ AdvVel adv_velocities(float vx00, float vy00, float vxXm, float vyYm, float vyXp, float vxYp,
                                                 float vxXmYp, float vyXpYm) {
  AdvVel a;
  a.Px = (vxXm + vx00) / 2.0f;
  a.Py = (vyYm + vy00) / 2.0f;
  a.Vxx = vx00;
  a.Vxy = (vyYm + vyXp + vy00 + vyXpYm) / 4.0f;
  a.Vyx = (vxXm + vxYp + vxXmYp + vx00) / 4.0f;
  a.Vyy = vy00;
  return a;
}

// :111-187 air cell after the gathers: condensation / evaporation with latent heat, global
// drying / heating / sounding forcing, clamp of total water
__device__ __forceinline__ void adv_air_thermo(const Geom& g, const DevParams& d, const float* __restrict__ sndT,
                                               const float* __restrict__ sndW, const float* __restrict__ sndV,
                                               float texCoordY, float4& base, float4& water) {
  const wsb_params& p = d.p;
  float realTemp = potentialToRealT(d, base.w, texCoordY);  // :111
  float excessWater = water.x - maxWater(realTemp);         // :115
  float overSaturation = excessWater - water.y;
  float condensation;
  if (overSaturation < 0.0f) condensation = overSaturation * 0.20f;
  else condensation = overSaturation * p.condensationRate;
  condensation = gmax(condensation, -water.y);
  float dT = condensation * p.evapHeat * 1.0f;
  base.w += dT;
  realTemp += dT;
  water.y += condensation;
  if (texCoordY > p.globalEffectsStartAlt && texCoordY < p.globalEffectsEndAlt) {  // :154-181
    // Each sub-block is skipped when its (uniform) rate is exactly +0: the skipped update then is
    // x -= +0 / x += +0 / x *= 1, which returns x itself for every finite x (the GUI defaults are
    // all zero, so the idle simulation pays for none of this).
    if (__float_as_uint(p.globalDrying) != 0u)
      water.x -= gclamp(p.globalDrying, 0.0f, gmax(water.x - maxWater(gmax(realTemp - 20.0f, CtoK(-80.0f))), 0.0f));
    if (__float_as_uint(p.globalHeating) != 0u) base.w += p.globalHeating;
    if (__float_as_uint(p.soundingForcing) != 0u) {
      int si = (int)(texCoordY * (1.0f / g.ltexelY));
      int sm = max(si - 1, 0);
      float Tdiff = base.w - (sndT[si] + sndT[sm]) / 2.0f;
      base.w -= Tdiff * 0.001f * p.soundingForcing;
      float Wdiff = water.x - (sndW[si] + sndW[sm]) / 2.0f;
      water.x -= Wdiff * 0.001f * p.soundingForcing;
      float dragK = 1.0f - map_rangeC(p.soundingForcing, 0.1f, 1.0f, 0.0f, 0.001f);
      base.x *= dragK;
      base.y *= dragK;
      float velDiff = base.x - (sndV[si] + sndV[sm]) / 2.0f;
      base.x -= velDiff * map_rangeC(p.soundingForcing, 0.9f, 1.0f, 0.0f, 0.001f);
    }
  }
  water.x = gmax(water.x, 0.0f);  // :187
}

// :189-227 wall cell: base / water enter as the pass-through copies of this cell.
// aboveDist / TAbove: wall distance and (pre-advection) temperature of the cell above.
template <bool DRY>
__device__ __forceinline__ void adv_wall_cell(const DevParams& d, float texCoordY, int wType, int aboveDist, float TAbove,
                                              int& wVeg, float4& base, float4& water) {
  if (wType == WALLTYPE_LAND) base.w = 1000.0f;
  if (!DRY) {
    wVeg = max(wVeg, 0);
    water.z = gmax(water.z, 0.0f);
    if (aboveDist != 0) {  // surface layer :207
      float tempC = KtoC(potentialToRealT(d, TAbove, texCoordY));
      if (water.w > 0.0f && tempC > 0.0f) {
        float melting = gmin(tempC * WSB_snowMeltRate, water.w);
        water.w -= melting;
        base.w += melting / WSB_snowMassToHeight * d.p.meltingHeat;
        water.z += melting;
      }
      if (water.z > 0.0f && tempC > 0.0f) {
        float evaporation = gmax((maxWater(CtoK(tempC)) - water.x) * 0.00001f, 0.0f);
        water.z -= evaporation;
      }
    }
  }
}

// :229-457 user input (brush), wall marker, airplane.  Inert for userInputType < 1 and
// airplaneValues[3] in [0, 0.9] — the idle-frame values — in which case only the wall marker runs.
__device__ __forceinline__ void adv_user_input(const Geom& g, const DevParams& d, const float* __restrict__ initial_T,
                                               float texCoordX, float texCoordY, int aboveDist, int& wType, int& wDist,
                                               int wVert, int& wVeg, float4& base, float4& water) {
  const wsb_params& p = d.p;
  const float* uiv = d.in.userInputValues;
  const int uit = d.in.userInputType;
  if (uit >= 1) {  // every branch below needs userInputType in {1,2,3,4} or >= 10
    bool inBrush = false;
    float weight = 1.0f;
    if (uiv[0] < -0.5f) {
      if (fabsf(uiv[1] - texCoordY) < uiv[3] * g.ltexelY) inBrush = true;
    } else {
      float vx, vy = uiv[1] - texCoordY;
      if (d.in.wrapHorizontally) vx = absHorizontalDist(uiv[0], texCoordX);
      else vx = fabsf(uiv[0] - texCoordX);
      vx *= g.ltexelY / g.ltexelX;
      float distFromMouse = glength(vx, vy);
      weight = gsmoothstep(uiv[3] * g.ltexelY, 0.0f, distFromMouse);
      if (distFromMouse < uiv[3] * g.ltexelY) inBrush = true;
    }
    if (inBrush) {
      const float intensity = uiv[2];
      if (uit == 1) {
        base.w += intensity;
        if (wType == 2 && wDist == 0) base.w = gclamp(base.w, CtoK(0.0f), CtoK(WSB_maxWaterTemp));
      } else if (uit == 2) {
        if (water.y > 0.0f) { water.y += intensity; water.y = gmax(water.y, 0.0f); }
        water.x += intensity;
        water.x = gmax(water.x, 0.0f);
      } else if (uit == 3 && wDist != 0) {
        water.w += intensity;
        water.w = gmin(gmax(water.w, 0.0f), 2.0f);
      } else if (uit == 4) {
        base.x += d.in.userInputMove[0] * 5.0f * weight * intensity;
        if (!(uiv[0] < -0.5f)) base.y += d.in.userInputMove[1] * 5.0f * weight * intensity;
      } else if (uit >= 10) {
        if (intensity > 0.0f) {
          bool setWall = false;
          switch (uit) {
            case 10: wType = WALLTYPE_INERT; setWall = true; break;
            case 11: wType = WALLTYPE_LAND; setWall = true; break;
            case 12: wType = WALLTYPE_WATER; setWall = true; break;
            case 13:
              if (wDist == 0 && wType == WALLTYPE_LAND && aboveDist != 0) { wType = WALLTYPE_FIRE; setWall = true; }
              break;
            case 14:
              if (wDist == 0 && (wType == WALLTYPE_LAND || wType == WALLTYPE_RUNWAY || wType == WALLTYPE_INDUSTRIAL) && aboveDist != 0) wType = WALLTYPE_URBAN;
              break;
            case 15:
              if (wDist == 0 && (wType == WALLTYPE_LAND || wType == WALLTYPE_URBAN || wType == WALLTYPE_INDUSTRIAL) && aboveDist != 0) wType = WALLTYPE_RUNWAY;
              break;
            case 16:
              if (wDist == 0 && (wType == WALLTYPE_LAND || wType == WALLTYPE_URBAN || wType == WALLTYPE_RUNWAY) && aboveDist != 0) wType = WALLTYPE_INDUSTRIAL;
              break;
            case 20:
              if (wDist == 0 && wType != WALLTYPE_WATER && aboveDist != 0) water.z += intensity * 10.0f;
              break;
            case 21:
              if (wDist == 0 && (wType == WALLTYPE_LAND || wType == WALLTYPE_URBAN || wType == WALLTYPE_INDUSTRIAL) && aboveDist != 0) water.w += intensity * 0.5f;
              break;
            case 22:
              if (wDist == 0 && (wType == WALLTYPE_LAND || wType == WALLTYPE_FIRE || wType == WALLTYPE_URBAN || wType == WALLTYPE_INDUSTRIAL) && aboveDist != 0) wVeg += 1;
              break;
            default: break;
          }
          if (setWall) {
            wDist = 0;
            base.w = 1000.0f;
            if (wType == WALLTYPE_LAND) water.z = 25.0f;
            else if (wType == WALLTYPE_WATER) base.w = p.waterTemperature;
          }
        } else {
          if (wDist == 0) {
            if (uit == 13) { if (wType == WALLTYPE_FIRE) wType = WALLTYPE_LAND; }
            else if (uit == 14) { if (wType == WALLTYPE_URBAN) wType = WALLTYPE_LAND; }
            else if (uit == 15) { if (wType == WALLTYPE_RUNWAY) wType = WALLTYPE_LAND; }
            else if (uit == 16) { if (wType == WALLTYPE_INDUSTRIAL) wType = WALLTYPE_LAND; }
            else if (uit == 20) { water.z += intensity * 10.0f; }
            else if (uit == 21) { water.w += intensity * 0.5f; }
            else if (uit == 22) { wVeg = max(wVeg - 1, 0); }
            else if (texCoordY > g.ltexelY) {
              wDist = 255;
              base.x = 0.0f; base.y = 0.0f; base.z = 0.0f;
              base.w = initial_T[(int)(texCoordY * (1.0f / g.ltexelY))];
              water.x = 0.0f; water.y = 0.0f; water.z = 0.0f; water.w = 0.0f;
            }
          }
        }
      }
    }
  }

  if (wDist == 0) water.x = (wType == WALLTYPE_WATER) ? 1002.0f : 1001.0f;  // :403-409

  // airplane :415-457
  const float* av = d.in.airplaneValues;
  if (av[3] < 0.0f || av[3] > 0.9f) {  // the block only has an effect in these two cases
    float px, py = av[1] - texCoordY;
    if (d.in.wrapHorizontally) px = absHorizontalDist(av[0], texCoordX);
    else px = fabsf(av[0] - texCoordX);
    px *= g.ltexelY / g.ltexelX;
    px *= g.Hf;
    py *= g.Hf;
    if (av[3] < 0.0f) { px += 0.0f; py += -1.0f; }
    float distFromPlane = glength(px, py);
    float planeInfluence = gmax(1.0f - distFromPlane, 0.0f) * 0.03f;
    if (av[3] < 0.0f) water.z += planeInfluence * 100.0f;
    if (av[3] > 0.9f && distFromPlane < 1.5f) {
      if (wDist == 0) {
        if (wType == WALLTYPE_LAND && wVert == 0) wType = WALLTYPE_FIRE;
      } else {
        base.z += 0.05f;
        base.w = CtoK(50.0f);
        water.x += 1.0f;
        water.w += 10.0f;
      }
    }
  }
}

// Generic form: every fetch goes through the context (global memory with wrap, or a checked tile).
template <bool DRY, class C>
__device__ void advection_cell(const C& c, const Geom& g, const DevParams& d, const float* __restrict__ initial_T,
                               const float* __restrict__ sndT, const float* __restrict__ sndW,
                               const float* __restrict__ sndV, int x, int y, float4& base, float4& water,
                               char4& wallOut, float& vmaxOut) {
  const int gx = global_x(g, x);
  const float fragCoordX = (float)gx + 0.5f, fragCoordY = (float)y + 0.5f;
  const float texCoordX = fragCoordX * g.texelX, texCoordY = fragCoordY * g.texelY;
  const char4 w0 = c.wall4(x, y);
  int wType = w0.x, wDist = w0.y, wVert = w0.z, wVeg = w0.w;
  const int aboveDist = c.swdist(x, y + 1);

  if (wDist != 0) {  // not wall :73
    const float vx00 = c.sbx(x, y), vy00 = c.sby(x, y);
    vmaxOut = fmaxf(vmaxOut, fmaxf(fabsf(vx00), fabsf(vy00)));
    const AdvVel a = adv_velocities(vx00, vy00, c.sbx(x - 1, y), c.sby(x, y - 1), c.sby(x + 1, y), c.sbx(x, y + 1),
                                    c.sbx(x - 1, y + 1), c.sby(x + 1, y - 1));
    {  // base[VX] = bilerp(baseTex, fragCoord - velAtVx).x  :93
      BilerpSetup b = bilerp_setup(fragCoordX - a.Vxx, fragCoordY - a.Vxy);
      const int x0 = x + (b.ix - gx), x1 = x0 + 1, y0 = b.iy, y1 = b.iy + 1;
      base.x = mix2d(c.bx(x0, y0), c.bx(x1, y0), c.bx(x0, y1), c.bx(x1, y1), b.fx, b.fx, b.fy);
    }
    {  // base[VY] = bilerp(baseTex, fragCoord - velAtVy).y  :94
      BilerpSetup b = bilerp_setup(fragCoordX - a.Vyx, fragCoordY - a.Vyy);
      const int x0 = x + (b.ix - gx), x1 = x0 + 1, y0 = b.iy, y1 = b.iy + 1;
      base.y = mix2d(c.by(x0, y0), c.by(x1, y0), c.by(x0, y1), c.by(x1, y1), b.fx, b.fx, b.fy);
    }
    const float posPx = fragCoordX - a.Px, posPy = fragCoordY - a.Py;
    {  // bilerpWall at the cell centre back-trace :96-99
      BilerpSetup b = bilerp_setup(posPx, posPy);
      const int x0 = x + (b.ix - gx), x1 = x0 + 1, y0 = b.iy, y1 = b.iy + 1;
      WallMix m = wall_mix(c.wdist(x0, y0), c.wdist(x1, y0), c.wdist(x0, y1), c.wdist(x1, y1), b.fx, b.fy);
      base.z = mix2d(c.bp(x0, y0), c.bp(x1, y0), c.bp(x0, y1), c.bp(x1, y1), m.ab, m.cd, m.abcd);
      base.w = mix2d(c.bt(x0, y0), c.bt(x1, y0), c.bt(x0, y1), c.bt(x1, y1), m.ab, m.cd, m.abcd);
      if (!DRY) {
        water.x = mix2d(c.wt0(x0, y0), c.wt0(x1, y0), c.wt0(x0, y1), c.wt0(x1, y1), m.ab, m.cd, m.abcd);
        water.y = mix2d(c.wt1(x0, y0), c.wt1(x1, y0), c.wt1(x0, y1), c.wt1(x1, y1), m.ab, m.cd, m.abcd);
        water.w = mix2d(c.wt3(x0, y0), c.wt3(x1, y0), c.wt3(x0, y1), c.wt3(x1, y1), m.ab, m.cd, m.abcd);
      }
    }
    if (DRY) {
      water = c.water4(x, y);
    } else {
      {  // precipitation visualisation, advected and pushed down :103
        BilerpSetup b = bilerp_setup(posPx + 0.0f, posPy + 0.05f);
        const int x0 = x + (b.ix - gx), x1 = x0 + 1, y0 = b.iy, y1 = b.iy + 1;
        WallMix m = wall_mix(c.wdist(x0, y0), c.wdist(x1, y0), c.wdist(x0, y1), c.wdist(x1, y1), b.fx, b.fy);
        water.z = mix2d(c.wt2(x0, y0), c.wt2(x1, y0), c.wt2(x0, y1), c.wt2(x1, y1), m.ab, m.cd, m.abcd);
      }
      adv_air_thermo(g, d, sndT, sndW, sndV, texCoordY, base, water);
    }
  } else {  // wall :189
    base = c.base4(x, y);
    water = c.water4(x, y);
    adv_wall_cell<DRY>(d, texCoordY, wType, aboveDist, c.sbt(x, y + 1), wVeg, base, water);
  }
  if (!DRY) adv_user_input(g, d, initial_T, texCoordX, texCoordY, aboveDist, wType, wDist, wVert, wVeg, base, water);
  wallOut = pack_wall(wType, wDist, wVert, wVeg);
}

// ---------------------------------------------------------------------------------------------
// lightingShader.frag:38-171 (first render target only; reflectedLight is display-only)
//   Ctx: float lightS(x,y) SUNLIGHT with x wrapped, y already clamped by the caller;
//        float lightIRdown(x,y), lightIRup(x,y).
//   T, water, wall: this cell AFTER the advection pass; TBelow: base_1 T of the cell below.
//   fragCoordX: (float)global_x(g, x) + 0.5f — a per-thread constant of the fused kernel, passed
//   in so that the periodic wrap of the column is not redone for every cell.
// ---------------------------------------------------------------------------------------------
template <class C>
__device__ float4 lighting_cell(const C& c, const Geom& g, const DevParams& d, int x, int y, float fragCoordX, float T, float4 water,
                                char4 wall, float TBelow) {
  const float fragCoordY = (float)y + 0.5f;
  if (fragCoordY >= g.Hf - 1.0f) return make_float4(d.in.sunIntensity, 0.0f, 0.0f, 0.0f);  // :40
  const float texCoordY = fragCoordY * g.texelY;
  const float cellHeightCompensation = g.cellHeightComp;  // 300.0f / g.Hf
  float sunlight;
  {  // :48-49, canonical fp32 bilinear in pixel space, wrap S = REPEAT, wrap T = CLAMP_TO_EDGE
    float px = fragCoordX + d.sinSun, py = fragCoordY + d.cosSun;
    float stx = px - 0.5f, sty = py - 0.5f;
    float flx = floorf(stx), fly = floorf(sty);
    float fx = stx - flx, fy = sty - fly;
    // column offset of the tap relative to this cell: (fragCoordX - 0.5f) is the global column, exactly
    int ix = x + (int)(flx - (fragCoordX - 0.5f)), iy = (int)fly;
    int y0 = min(max(iy, 0), g.H - 1), y1 = min(max(iy + 1, 0), g.H - 1);
    sunlight = gmix(gmix(c.lightS(ix, y0), c.lightS(ix + 1, y0), fx), gmix(c.lightS(ix, y1), c.lightS(ix + 1, y1), fx), fy);
  }
  const float realTemp = potentialToRealT(d, T, texCoordY);
  if (wall.y != 0) {
    float net_heating = 0.0f;
    if (fragCoordY < g.Hf - 2.0f) {  // :65-85
      float reflection = gmin(sqrtf(water.y * 0.0010f + water.z * 0.00020f) * cellHeightCompensation, 1.0f);
      reflection += 0.0002f;
      float absorbtion = gmin(water.w * 0.020f * cellHeightCompensation, 1.0f);
      float lightReflected = sunlight * reflection;
      float lightAbsorbed = sunlight * absorbtion;
      sunlight = gmax(0.0f, sunlight - lightReflected - lightAbsorbed);
      net_heating += lightAbsorbed * WSB_lightHeatingConst;
    }
    float IR_down = c.lightIRdown(x, min(y + 1, g.H - 1));
    float IR_up = 0.0f;
    if (wall.z == 1) {  // :90-116
      switch (wall.x) {
        case WALLTYPE_RUNWAY: case WALLTYPE_URBAN: case WALLTYPE_INDUSTRIAL: case WALLTYPE_LAND:
          IR_up = IR_emitted(realTemp);
          net_heating += (IR_down - IR_up) * WSB_lightHeatingConst;
          break;
        case WALLTYPE_WATER:
          IR_up = IR_emitted(TBelow);
          net_heating += (IR_down - IR_up) * WSB_lightHeatingConst;
          break;
        case WALLTYPE_FIRE:
          IR_up = IR_emitted(realTemp + 100.0f);
          net_heating = 0.0f;
          break;
        default: break;
      }
    } else {  // :117-144
      IR_up = c.lightIRup(x, max(y - 1, 0));
      float emissivity = d.p.greenhouseGases;
      emissivity += water.x * d.p.waterGreenHouseEffect;
      emissivity += water.y * 5.0f;
      emissivity *= cellHeightCompensation;
      emissivity = gmin(emissivity, 1.0f);
      float absorbedDown = IR_down * emissivity;
      float absorbedUp = IR_up * emissivity;
      float emitted = IR_emitted(realTemp) * emissivity;
      net_heating += (absorbedDown + absorbedUp - emitted * 2.0f) * WSB_lightHeatingConst;
      IR_down -= absorbedDown;
      IR_down += emitted;
      IR_up -= absorbedUp;
      IR_up += emitted;
    }
    net_heating *= d.p.IR_rate;
    return make_float4(sunlight, net_heating, IR_down, IR_up);
  }
  if (wall.x == WALLTYPE_WATER) return make_float4(sunlight * 0.90f, 0.0f, 0.0f, 0.0f);
  return make_float4(0.0f, 0.0f, 0.0f, 0.0f);
}

}  // namespace wsb
