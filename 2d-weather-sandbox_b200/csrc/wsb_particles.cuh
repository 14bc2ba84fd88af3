// wsb_particles.cuh — precipitation particles: precipitationShader.vert:66-293 (update + transform
// feedback), the additive point sprites it rasterises into the feedback / deposition targets
// (app.js:5938-5954), the inactive-droplet latch (app.js:5957-5967) and the lightning latch
// (lightningLocationShader.frag:24-38).
//
// Sprite rule (canonical): window centre = (pos+1)/2 * resolution; a pixel is covered
// when its centre lies in [c - size/2, c + size/2); clipped to the viewport, never wrapped; a point
// whose centre is outside the clip volume is discarded.  The reference blends sprites in droplet
// order; here the adds are L2 atomics (vector red.global.add.v4.f32 / v2.f32), so overlapping
// sprites sum in arbitrary order — equal up to fp32 rounding of the sum.
#pragma once
#include "wsb_ref_kernels.cuh"

namespace wsb {

struct DropletResult {
  float posX, posY, massW, massI, density;  // transform-feedback outputs
  float glX, glY, pointSize;                // gl_Position.xy, gl_PointSize
  float4 feedback;
  float2 deposition;
  bool inactiveMarker;  // the 1-px "still inactive" point at texel (0,0) (:154-161)
};

__device__ __forceinline__ size_t nearest_texel(const Geom& g, float tx, float ty) {  // NEAREST + REPEAT
  int ix = mod_i((int)floorf(tx * g.Wf), g.pitch), iy = mod_i((int)floorf(ty * g.Hf), g.H);
  return (size_t)iy * g.pitch + ix;
}

__device__ DropletResult droplet_update(const float* __restrict__ din, const Planes4& baseT, const Planes4& waterT,
                                        const Geom& g, const DevParams& d, float lightningStart, float inactiveDroplets) {
  const wsb_params& p = d.p;
  const float dropX = din[0], dropY = din[1], massW = din[2], massI = din[3], density = din[4];
  float newPosX = dropX, newPosY = dropY, newMassW = massW, newMassI = massI, newDensity = density;
  float fbM = 0.0f, fbH = 0.0f, fbV = 0.0f, fbI = 0.0f;  // MASS, HEAT, VAPOR, [3]; varyings zero-initialised
  float depR = 0.0f, depS = 0.0f;
  bool isActive = true, spawned = false, lightningSpawned = false, inactiveMarker = false;
  float pointSize = 1.0f, glX = 0.0f, glY = 0.0f;
  float texCoordX = 0.0f, texCoordY = 0.0f, realTemp = 0.0f;
  float4 base = make_float4(0, 0, 0, 0), water = make_float4(0, 0, 0, 0);
  const float iterNum = d.iterNum;

  if (massW < 0.0f) {  // inactive :72
    texCoordX = random2d(massW, dropX + iterNum * 0.3754f);
    texCoordY = random2d(massI, dropX + iterNum * 0.073162f);
    size_t ci = nearest_texel(g, texCoordX, texCoordY);
    base = baseT.ld(ci);
    water = waterT.ld(ci);
    realTemp = potentialToRealT(d, base.w, texCoordY);
    const float initalMass = 0.15f;
    float threshold = (realTemp > CtoK(0.0f)) ? p.aboveZeroThreshold : p.subZeroThreshold;
    if (water.y > threshold && base.w < 500.0f) {
      float spawnChance = ((water.y - threshold) / (inactiveDroplets + 10.0f)) * g.Wf * g.Hf * p.spawnChanceMult;
      float t10 = water.y * 10.0f;
      float nrmRand = gfract(t10 * t10);
      if (spawnChance > nrmRand) {
        spawned = true;
        newPosX = (texCoordX - 0.5f) * 2.0f;
        newPosY = (texCoordY - 0.5f) * 2.0f;
        if (realTemp < CtoK(0.0f)) {
          newMassW = 0.0f;
          newMassI = initalMass;
          fbH += newMassI * p.meltingHeat;
          newDensity = p.snowDensity;
          float cloudPlusPrecipDensity = water.y + water.z;
          float lightningSpawnChance = gmax((cloudPlusPrecipDensity - 2.5f) * 0.0033f, 0.0f);
          if (lightningStart < iterNum - 30.0f && random2d(base.w * 0.2324f, water.x * 7.7f) < lightningSpawnChance) {
            lightningSpawned = true;
            isActive = false;
            pointSize = 1.0f;
            fbM = texCoordX;
            fbH = texCoordY;
            fbV = iterNum;
            fbI = gclamp(cloudPlusPrecipDensity / 10.0f + (random2d(texCoordX, texCoordY) - 0.5f), 0.01f, 4.0f);
            glX = -1.0f + g.texelX * 3.0f;
            glY = -1.0f + g.texelY;
          }
        } else {
          newMassW = initalMass;
          newMassI = 0.0f;
          newDensity = 1.0f;
        }
        fbV -= initalMass;  // :146
      }
    }
    if (spawned) {
      if (!lightningSpawned) { pointSize = 1.0f; glX = newPosX; glY = newPosY; }
    } else {
      isActive = false;
      inactiveMarker = true;
      pointSize = 1.0f;
      fbM = 1.0f;
      glX = -1.0f + g.texelX;
      glY = -1.0f + g.texelY;
    }
  }
  if (isActive) {  // :164
    if (!spawned) {
      texCoordX = dropX / 2.0f + 0.5f;
      texCoordY = dropY / 2.0f + 0.5f;
      size_t ci = nearest_texel(g, texCoordX, texCoordY);
      water = waterT.ld(ci);
      base = baseT.ld(ci);
      realTemp = potentialToRealT(d, base.w, texCoordY);
    }
    float totalMass = newMassW + newMassI;
    if (totalMass < 0.04f) {  // :175
      fbH = -(totalMass * p.evapHeat);
      fbV = totalMass;
      newMassW = -2.0f - dropX;
      newMassI = dropY;
    } else if (newPosY < -1.0f || water.x > 1000.0f) {  // :183
      if (baseT.c[3][nearest_texel(g, texCoordX, texCoordY + g.texelY)] > 500.0f) newPosY += g.texelY * 1.0f;
      depR = newMassW;
      depS = newMassI;
      newMassW = -2.0f - dropX;
      newMassI = dropY;
    } else {  // :193
      float surfaceArea = wsb_cbrt(totalMass);
      float growthRate = gmax(map_range(realTemp, CtoK(0.0f), CtoK(-30.0f), p.growthRate0C, p.growthRate_30C), p.growthRate0C);
      float growth = water.y * growthRate * surfaceArea;
      if (realTemp < CtoK(0.0f) && water.y > 0.0f && density == 1.0f) growth += surfaceArea * water.z * 0.0030f;
      fbV -= growth * 1.0f;
      if (realTemp < CtoK(0.0f)) {
        newMassI += growth;
        fbH += growth * p.meltingHeat;
        float freezing = gmin((CtoK(0.0f) - realTemp) * p.freezingRate * surfaceArea, newMassW);
        newMassW -= freezing;
        newMassI += freezing;
        fbH += freezing * p.meltingHeat;
      } else {
        newMassW += growth;
        float melting = gmin((realTemp - CtoK(0.0f)) * p.meltingRate * surfaceArea, newMassI);
        newMassI -= melting;
        newMassW += melting;
        fbH -= melting * p.meltingHeat;
        newDensity = gmin(newDensity + (melting / totalMass) * 1.00f, 1.0f);
      }
      float dropletTemp = potentialToRealT(d, base.w, texCoordY);
      if (newMassI > 0.0f) dropletTemp = gmin(dropletTemp, CtoK(0.0f));
      float evapAndSubli = gmax((maxWater(dropletTemp) - water.x) * surfaceArea * p.evapRate, 0.0f);
      float evap = gmin(newMassW, evapAndSubli);
      float subli = gmin(newMassI, evapAndSubli - evap);
      newMassW -= evap;
      newMassI -= subli;
      fbV += evap;
      fbV += subli;
      fbH -= evap * p.evapHeat;
      fbH -= subli * p.evapHeat;
      fbH -= subli * p.meltingHeat;
      newPosX += base.x / g.Wf * 2.0f;
      newPosY += base.y / g.Hf * 2.0f;
      newPosY -= p.fallSpeed * newDensity * sqrtf(totalMass / surfaceArea);
      newPosX = gmod(newPosX + 1.0f, 2.0f) - 1.0f;
      fbM = totalMass;
    }
    const float pntSize = 12.0f, pntSurface = pntSize * pntSize;
    fbM /= pntSurface;
    fbH /= pntSurface;
    fbV /= pntSurface;
    depR /= pntSize;
    depS /= pntSize;
    pointSize = pntSize;
    glX = newPosX;
    glY = newPosY;
  }
  DropletResult r;
  r.posX = newPosX; r.posY = newPosY; r.massW = newMassW; r.massI = newMassI; r.density = gmax(newDensity, 0.0f);
  r.glX = glX; r.glY = glY; r.pointSize = pointSize;
  r.feedback = make_float4(fbM, fbH, fbV, fbI);
  r.deposition = make_float2(depR, depS);
  r.inactiveMarker = inactiveMarker;
  return r;
}

// One thread per droplet for the update; the sprites are then rasterised WARP-COOPERATIVELY: each
// lane that has a sprite broadcasts it in turn and all 32 lanes add one pixel each, walking the
// sprite's rows — a 12-pixel row is 192 contiguous bytes of the feedback texture, so the vector
// atomics of one instruction fall into a handful of sectors instead of 32 scattered ones.
// Inactive droplets (the majority) only count themselves: the count is reduced per block and lands
// on texel (0,0) with one atomic (sums of 1.0 are exact in fp32).
__global__ void __launch_bounds__(256) k_precipitation(const float* __restrict__ dropsIn, float* __restrict__ dropsOut,
                                                       Planes4 baseT, Planes4 waterT,
                                                       float4* __restrict__ fb, float2* __restrict__ dep,
                                                       const float* __restrict__ lightning,
                                                       const float* __restrict__ inactiveUniform, Geom g, DevParams d, int ND) {
  __shared__ int sInactive;
  if (threadIdx.x == 0) sInactive = 0;
  __syncthreads();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  bool countsInactive = false;
  // this lane's sprite: origin pixel, size (0 = none), payload
  int xs = 0, ys = 0, sz = 0;
  float4 sf = make_float4(0.f, 0.f, 0.f, 0.f);
  float2 sd = make_float2(0.f, 0.f);
  if (n < ND) {
    DropletResult r = droplet_update(dropsIn + (size_t)n * 5, baseT, waterT, g, d, lightning[2], *inactiveUniform);
    float* o = dropsOut + (size_t)n * 5;
    o[0] = r.posX; o[1] = r.posY; o[2] = r.massW; o[3] = r.massI; o[4] = r.density;
    if (r.inactiveMarker) {
      countsInactive = true;  // feedback = (1,0,0,0) on texel (0,0)
    } else if (r.glX >= -1.0f && r.glX <= 1.0f && r.glY >= -1.0f && r.glY <= 1.0f) {
      const float xw = (r.glX + 1.0f) * 0.5f * g.Wf, yw = (r.glY + 1.0f) * 0.5f * g.Hf;
      const float half = r.pointSize * 0.5f;
      xs = (int)ceilf(xw - half - 0.5f);
      ys = (int)ceilf(yw - half - 0.5f);
      sz = (int)r.pointSize;
      sf = r.feedback;
      sd = r.deposition;
    }
  }
  unsigned pending = __ballot_sync(0xffffffffu, sz > 0);
  while (pending) {
    const int src = __ffs(pending) - 1;
    pending &= pending - 1;
    const int bx = __shfl_sync(0xffffffffu, xs, src), by = __shfl_sync(0xffffffffu, ys, src), bs = __shfl_sync(0xffffffffu, sz, src);
    const float4 f = make_float4(__shfl_sync(0xffffffffu, sf.x, src), __shfl_sync(0xffffffffu, sf.y, src),
                                 __shfl_sync(0xffffffffu, sf.z, src), __shfl_sync(0xffffffffu, sf.w, src));
    const float2 dp = make_float2(__shfl_sync(0xffffffffu, sd.x, src), __shfl_sync(0xffffffffu, sd.y, src));
    const bool hasDep = (dp.x != 0.0f || dp.y != 0.0f);
    for (int p = lane; p < bs * bs; p += 32) {
      const int j = by + p / bs, i = bx + p % bs;
      if (j >= 0 && j < g.H && i >= 0 && i < g.pitch) {
        const size_t ci = (size_t)j * g.pitch + i;
        atomicAdd(&fb[ci], f);
        if (hasDep) atomicAdd(&dep[ci], dp);
      }
    }
  }
  if (countsInactive) atomicAdd(&sInactive, 1);
  __syncthreads();
  if (threadIdx.x == 0 && sInactive > 0) atomicAdd(&fb[0], make_float4((float)sInactive, 0.0f, 0.0f, 0.0f));
}

// app.js:5957-5967 (every 600th iteration) + lightningLocationShader.frag:24-38, on device.
__global__ void k_latch(const float4* __restrict__ fb, float* __restrict__ inactiveUniform, float* __restrict__ lightning,
                        float iterNum, int latchInactive) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    if (latchInactive) *inactiveUniform = fb[0].x;
    float4 nl = fb[1];
    if (!(nl.z < gmax(iterNum - 1.0f, 1.0f) || nl.z > iterNum)) {
      lightning[0] = nl.x; lightning[1] = nl.y; lightning[2] = nl.z; lightning[3] = nl.w;
    }
  }
}

}  // namespace wsb
