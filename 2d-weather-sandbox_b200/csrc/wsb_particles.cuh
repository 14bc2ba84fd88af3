// wsb_particles.cuh — precipitation particles: precipitationShader.vert:66-293 (update + transform
// feedback), the additive point sprites it rasterises into the feedback / deposition targets
// (app.js:5938-5954), the inactive-droplet latch (app.js:5957-5967) and the lightning latch
// (lightningLocationShader.frag:24-38).
//
// Sprite rule (canonical): window centre = (pos+1)/2 * resolution; a pixel is covered
// when its centre lies in [c - size/2, c + size/2); clipped to the viewport, never wrapped; a point
// whose centre is outside the clip volume is discarded.  The reference blends sprites in droplet
// order — 144 blended pixels per active droplet, its author's "HUGE PERFORMANCE BOTTLENECK"
// (app.js:5936).  Every pixel of a sprite receives the SAME value, so the sum of all 12 x 12 sprites
// is a 12 x 12 box filter over the grid of sprite ORIGINS:
//   k_precipitation  updates the droplets and adds each active droplet ONCE (one vector atomic,
//                    red.global.add.v4.f32, + one v2 for the rare deposition) to its origin cell,
//                    marking the <= 4 tiles its sprite touches in a dirty map;
//   k_boxsum         for the dirty 64 x 16 tiles only (a persistent grid walks the list of tiles the
//                    particle pass touched): origins of the tile + 11-cell apron -> shared memory,
//                    12-tap row sums, 12-tap column sums, added to feedback / deposition;
//   k_clear_origins  zeroes the origin cells of the listed tiles and resets the list;
// and k_fused_pvb reads (and clears) feedback / deposition in dirty tiles only.  Overlapping
// sprites therefore sum in an order of their own — equal to the reference's up to fp32 rounding.
// 1-pixel sprites (spawn marker, lightning bolt, inactive count) are added to the target directly.
#pragma once
#include "wsb_fused_kernels.cuh"
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

constexpr int kSpriteSize = 12;             // pntSize of precipitationShader.vert:270
constexpr int kOriginShift = kSpriteSize / 2;  // origin cell = first covered pixel + 6, in [0, W] x [0, H]
constexpr int kPTX = kTX, kPTY = kTY;       // dirty-map tiles = the tiles of k_fused_pvb

// first toucher of a tile appends it to the list
__device__ __forceinline__ void mark_dirty(const SpriteGrid& sg, int tile, int bits) {
  if ((sg.dirty[tile] & bits) == bits) return;  // already flagged (plain read: a stale 0 only costs the atomic below)
  if (atomicOr(&sg.dirty[tile], bits) == 0) sg.dirtyList[atomicAdd(sg.dirtyCount, 1)] = tile;
}

// One thread per droplet.  Inactive droplets (the majority) only count themselves: the count is
// reduced per block and lands on texel (0,0) with one atomic (sums of 1.0 are exact in fp32).
__global__ void __launch_bounds__(256) k_precipitation(const float* __restrict__ dropsIn, float* __restrict__ dropsOut,
                                                       Planes4 baseT, Planes4 waterT,
                                                       float4* __restrict__ fb, float2* __restrict__ dep, SpriteGrid sg,
                                                       const float* __restrict__ lightning,
                                                       const float* __restrict__ inactiveUniform, Geom g, DevParams d, int ND) {
  __shared__ int sInactive;
  if (threadIdx.x == 0) sInactive = 0;
  __syncthreads();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  bool countsInactive = false;
  if (n < ND) {
    DropletResult r = droplet_update(dropsIn + (size_t)n * 5, baseT, waterT, g, d, lightning[2], *inactiveUniform);
    float* o = dropsOut + (size_t)n * 5;
    o[0] = r.posX; o[1] = r.posY; o[2] = r.massW; o[3] = r.massI; o[4] = r.density;
    if (r.inactiveMarker) {
      countsInactive = true;  // feedback = (1,0,0,0) on texel (0,0)
    } else if (r.glX >= -1.0f && r.glX <= 1.0f && r.glY >= -1.0f && r.glY <= 1.0f) {
      const float xw = (r.glX + 1.0f) * 0.5f * g.Wf, yw = (r.glY + 1.0f) * 0.5f * g.Hf;
      const float half = r.pointSize * 0.5f;
      const int xs = (int)ceilf(xw - half - 0.5f), ys = (int)ceilf(yw - half - 0.5f);  // first covered pixel
      const bool hasDep = r.deposition.x != 0.0f || r.deposition.y != 0.0f;
      if ((int)r.pointSize == kSpriteSize) {
        const size_t oi = (size_t)(ys + kOriginShift) * sg.Po + (xs + kOriginShift);
        atomicAdd(&sg.org4[oi], r.feedback);
        if (hasDep) atomicAdd(&sg.org2[oi], r.deposition);
        // the sprite covers pixels [xs, xs + 12) x [ys, ys + 12), clipped: at most 2 x 2 tiles
        const int tx0 = max(xs, 0) / kPTX, tx1 = min(xs + kSpriteSize - 1, g.pitch - 1) / kPTX;
        const int ty0 = max(ys, 0) / kPTY, ty1 = min(ys + kSpriteSize - 1, g.H - 1) / kPTY;
        for (int ty = ty0; ty <= ty1; ty++)
          for (int tx = tx0; tx <= tx1; tx++) mark_dirty(sg, ty * sg.tilesX + tx, hasDep ? (kDirtyFb | kDirtyDep) : kDirtyFb);
      } else if (xs >= 0 && xs < g.pitch && ys >= 0 && ys < g.H) {  // 1-pixel sprite: spawn marker, lightning bolt
        const size_t ci = (size_t)ys * g.pitch + xs;
        atomicAdd(&fb[ci], r.feedback);
        if (hasDep) atomicAdd(&dep[ci], r.deposition);
        mark_dirty(sg, (ys / kPTY) * sg.tilesX + xs / kPTX, hasDep ? (kDirtyFb | kDirtyDep) : kDirtyFb);
      }
    }
  }
  if (countsInactive) atomicAdd(&sInactive, 1);
  __syncthreads();
  if (threadIdx.x == 0 && sInactive > 0) {
    atomicAdd(&fb[0], make_float4((float)sInactive, 0.0f, 0.0f, 0.0f));
    mark_dirty(sg, 0, kDirtyFb);
  }
}

// 12 x 12 box filter of the sprite origins of one dirty tile, added to the target texels.
//   pixel (x, y) is covered by the sprites whose first pixel lies in [x - 11, x] x [y - 11, y], i.e. whose ORIGIN
//   cell (first pixel + 6) lies in [x - 5, x + 6] x [y - 5, y + 6].
constexpr int kBoxW = kPTX + kSpriteSize - 1, kBoxH = kPTY + kSpriteSize - 1;  // 75 x 27 origins per tile
// Column sums first (27 -> 16 rows, every staged column), then row sums straight into the target: 660 work items per
// tile instead of 850 the other way round, and the intermediate is 16 rows instead of 27 (52.9 KB: 4 CTAs per SM).
// Window sharing: a thread computes 4 vertically (3 horizontally) adjacent outputs from 15 (14) inputs held in
// registers, summing the inputs all of its windows share once.  THREE per thread in the row pass because with 16-byte
// elements an odd count keeps the eight lanes of a quarter warp on different banks (four per thread put them on two:
// ncu showed 2.7 wavefronts per ideal one, profiles/r3_particles_ncu_summary.txt); 64 is not a multiple of 3, so the
// staged region is padded to 77 columns (two of zeros) and the last two of 66 row outputs are dropped.
constexpr int kRowOut = (kPTX + 2) / 3 * 3;           // 66
constexpr int kBoxPitch = kRowOut + kSpriteSize - 1;  // 77
constexpr size_t kSmemBox = (size_t)(kBoxH * kBoxPitch + kPTY * kBoxPitch) * sizeof(float4);
__device__ __forceinline__ float4 vadd(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float2 vadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ bool vnonzero(float4 a) { return a.x != 0.0f || a.y != 0.0f || a.z != 0.0f || a.w != 0.0f; }
__device__ __forceinline__ bool vnonzero(float2 a) { return a.x != 0.0f || a.y != 0.0f; }
__device__ __forceinline__ void vzero(float4& a) { a = make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void vzero(float2& a) { a = make_float2(0.f, 0.f); }

template <class V>
__device__ __forceinline__ void boxsum_tile(const V* __restrict__ org, V* __restrict__ target, int Po, int W, int H, int pitch, int X0,
                                            int Y0, V* sOrg, V* sCol) {
  static_assert(kSpriteSize == 12 && kPTY % 4 == 0, "window sharing below is written for 12 taps");
  const int tid = threadIdx.x;
  for (int i = tid; i < kBoxH * kBoxPitch; i += blockDim.x) {
    const int r = i / kBoxPitch, c = i - r * kBoxPitch;
    const int oj = Y0 - (kSpriteSize - 1 - kOriginShift) + r, oi = X0 - (kSpriteSize - 1 - kOriginShift) + c;
    V v;
    vzero(v);
    if (c < kBoxW && oj >= 0 && oj <= H && oi >= 0 && oi <= W) v = org[(size_t)oj * Po + oi];
    sOrg[i] = v;
  }
  __syncthreads();
  for (int i = tid; i < (kPTY / 4) * kBoxPitch; i += blockDim.x) {  // column sums: (rows 4q .. 4q+3, staged column c)
    const int q = i / kBoxPitch, c = i - q * kBoxPitch, ty = q * 4;
    const V* p = sOrg + ty * kBoxPitch + c;
    V in[15];
#pragma unroll
    for (int k = 0; k < 15; k++) in[k] = p[k * kBoxPitch];
    V core = in[3];
#pragma unroll
    for (int k = 4; k < 12; k++) core = vadd(core, in[k]);
    V* o = sCol + ty * kBoxPitch + c;
    o[0] = vadd(vadd(in[0], in[1]), vadd(in[2], core));
    o[kBoxPitch] = vadd(vadd(in[1], in[2]), vadd(core, in[12]));
    o[2 * kBoxPitch] = vadd(vadd(in[2], core), vadd(in[12], in[13]));
    o[3 * kBoxPitch] = vadd(vadd(core, in[12]), vadd(in[13], in[14]));
  }
  __syncthreads();
  for (int i = tid; i < kPTY * (kRowOut / 3); i += blockDim.x) {  // row sums: (row r, columns 3t .. 3t+2) -> target
    const int r = i / (kRowOut / 3), c = (i - r * (kRowOut / 3)) * 3;
    const V* p = sCol + r * kBoxPitch + c;
    V in[14];
#pragma unroll
    for (int k = 0; k < 14; k++) in[k] = p[k];
    V core = in[2];
#pragma unroll
    for (int k = 3; k < 12; k++) core = vadd(core, in[k]);
    V out[3];
    out[0] = vadd(vadd(in[0], in[1]), core);
    out[1] = vadd(vadd(in[1], core), in[12]);
    out[2] = vadd(core, vadd(in[12], in[13]));
    const int y = Y0 + r;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const int x = X0 + c + k;
      if (c + k < kPTX && x < W && y < H && vnonzero(out[k])) {
        const size_t ci = (size_t)y * pitch + x;
        target[ci] = vadd(target[ci], out[k]);  // texels hit by 1-pixel sprites already hold their value
      }
    }
  }
  __syncthreads();
}

// Persistent grid (a few CTAs per SM): CTA b takes list entries b, b + gridDim.x, ...
__global__ void __launch_bounds__(256, 4) k_boxsum(SpriteGrid sg, float4* __restrict__ fb, float2* __restrict__ dep, int W, int H, int pitch) {
  WSB_DYN_SMEM(smem_raw);
  float4* sOrg = reinterpret_cast<float4*>(smem_raw);
  float4* sRow = sOrg + kBoxH * kBoxPitch;  // the column sums
  const int n = *sg.dirtyCount;
  for (int i = blockIdx.x; i < n; i += gridDim.x) {
    const int tile = sg.dirtyList[i], bits = sg.dirty[tile];
    const int X0 = (tile % sg.tilesX) * kPTX, Y0 = (tile / sg.tilesX) * kPTY;
    if (bits & kDirtyFb) boxsum_tile<float4>(sg.org4, fb, sg.Po, W, H, pitch, X0, Y0, sOrg, sRow);
    if (bits & kDirtyDep) boxsum_tile<float2>(sg.org2, dep, sg.Po, W, H, pitch, X0, Y0, reinterpret_cast<float2*>(sOrg), reinterpret_cast<float2*>(sRow));
  }
}

// Every non-zero origin cell lies inside its own sprite, i.e. inside a listed tile (or in the extra column W / row H,
// which the last tile column / row owns): zero the origin cells of the listed tiles.  A kernel of its own because the
// box sums of neighbouring tiles read each other's origins.  The last CTA to finish resets the list.
__global__ void __launch_bounds__(256) k_clear_origins(SpriteGrid sg, int W, int H, unsigned* __restrict__ ctasDone) {
  const int n = *sg.dirtyCount;
  for (int i = blockIdx.x; i < n; i += gridDim.x) {
    const int tile = sg.dirtyList[i], bits = sg.dirty[tile];
    const int bx = tile % sg.tilesX, by = tile / sg.tilesX;
    const int X0 = bx * kPTX, Y0 = by * kPTY;
    const int X1 = (bx == sg.tilesX - 1) ? W + 1 : X0 + kPTX, Y1 = (by == sg.tilesY - 1) ? H + 1 : Y0 + kPTY;
    const int w = X1 - X0, cells = w * (Y1 - Y0);
    for (int k = threadIdx.x; k < cells; k += blockDim.x) {
      const int r = k / w, c = k - r * w;
      const size_t oi = (size_t)(Y0 + r) * sg.Po + X0 + c;
      if (bits & kDirtyFb) sg.org4[oi] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (bits & kDirtyDep) sg.org2[oi] = make_float2(0.f, 0.f);
    }
  }
  __syncthreads();  // every thread of this CTA has read the count
  if (threadIdx.x == 0 && atomicAdd(ctasDone, 1u) == gridDim.x - 1) {
    *ctasDone = 0u;
    *sg.dirtyCount = 0;
  }
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
