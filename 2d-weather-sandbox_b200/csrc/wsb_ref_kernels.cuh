// wsb_ref_kernels.cuh — REFERENCE schedule: one kernel per reference pass, every neighbour fetch
// straight from global memory (L1/L2 do the reuse).  Same order and bindings as app.js:5830-6005.
// Used for per-pass parity against the oracle and as the "recompiled baseline" the fused kernels
// are measured against; the product path is wsb_fused_kernels.cuh.
#pragma once
#include "wsb_cells.cuh"

namespace wsb {

// Fetches from the packed global arrays.  AoS float4 per cell, exactly the texture layout.
struct GlobalCtx {
  const float4* __restrict__ base;
  const float4* __restrict__ water;
  const char4* __restrict__ wall;
  const float2* __restrict__ vortf;
  const float4* __restrict__ light;
  const float4* __restrict__ fb;
  const float2* __restrict__ dep;
  Geom g;
  // any (x, y): periodic in y, periodic (single domain) or clamped (strip) in x
  __device__ __forceinline__ size_t idx(int x, int y) const { return (size_t)mod_i(y, g.H) * g.pitch + gather_x(g, x); }
  // (x, y) at most one period outside the array (fixed +-1 stencils): no integer division
  __device__ __forceinline__ size_t idx_near(int x, int y) const { return (size_t)wrap_y(y, g.H) * g.pitch + wrap_x(g, x); }
  __device__ __forceinline__ float4 base4(int x, int y) const { return base[idx(x, y)]; }
  __device__ __forceinline__ float bx(int x, int y) const { return reinterpret_cast<const float*>(base)[idx(x, y) * 4 + 0]; }
  __device__ __forceinline__ float by(int x, int y) const { return reinterpret_cast<const float*>(base)[idx(x, y) * 4 + 1]; }
  __device__ __forceinline__ float bp(int x, int y) const { return reinterpret_cast<const float*>(base)[idx(x, y) * 4 + 2]; }
  __device__ __forceinline__ float bt(int x, int y) const { return reinterpret_cast<const float*>(base)[idx(x, y) * 4 + 3]; }
  __device__ __forceinline__ float sbx(int x, int y) const { return bx(x, y); }
  __device__ __forceinline__ float sby(int x, int y) const { return by(x, y); }
  __device__ __forceinline__ float sbt(int x, int y) const { return bt(x, y); }
  __device__ __forceinline__ int swdist(int x, int y) const { return wdist(x, y); }
  __device__ __forceinline__ float4 water4(int x, int y) const { return water[idx(x, y)]; }
  __device__ __forceinline__ float wt0(int x, int y) const { return reinterpret_cast<const float*>(water)[idx(x, y) * 4 + 0]; }
  __device__ __forceinline__ float wt1(int x, int y) const { return reinterpret_cast<const float*>(water)[idx(x, y) * 4 + 1]; }
  __device__ __forceinline__ float wt2(int x, int y) const { return reinterpret_cast<const float*>(water)[idx(x, y) * 4 + 2]; }
  __device__ __forceinline__ float wt3(int x, int y) const { return reinterpret_cast<const float*>(water)[idx(x, y) * 4 + 3]; }
  __device__ __forceinline__ char4 wall4(int x, int y) const { return wall[idx(x, y)]; }
  __device__ __forceinline__ int wdist(int x, int y) const { return reinterpret_cast<const signed char*>(wall)[idx(x, y) * 4 + 1]; }
  __device__ __forceinline__ float2 vort(int x, int y) const { return vortf[idx(x, y)]; }
  // light texture: wrap S = REPEAT, wrap T = CLAMP_TO_EDGE (app.js:5276-5279)
  __device__ __forceinline__ size_t lidx(int x, int y) const { return (size_t)min(max(y, 0), g.H - 1) * g.pitch + wrap_x(g, x); }
  __device__ __forceinline__ float4 light4(int x, int y) const { return light[lidx(x, y)]; }
  __device__ __forceinline__ float lightS(int x, int y) const { return reinterpret_cast<const float*>(light)[lidx(x, y) * 4 + 0]; }
  __device__ __forceinline__ float lightIRdown(int x, int y) const { return reinterpret_cast<const float*>(light)[lidx(x, y) * 4 + 2]; }
  __device__ __forceinline__ float lightIRup(int x, int y) const { return reinterpret_cast<const float*>(light)[lidx(x, y) * 4 + 3]; }
  __device__ __forceinline__ float4 fb4(int x, int y) const { return fb[idx(x, y)]; }
  __device__ __forceinline__ float2 dep2(int x, int y) const { return dep[idx(x, y)]; }
};

// GlobalCtx positioned at a cell: the offset-style fetches boundary_cell asks for.
struct GlobalAt {
  const GlobalCtx& c;
  int x, y;
  __device__ __forceinline__ float4 base4(int dx, int dy) const { return c.base[c.idx_near(x + dx, y + dy)]; }
  __device__ __forceinline__ float bx(int dx, int dy) const { return reinterpret_cast<const float*>(c.base)[c.idx_near(x + dx, y + dy) * 4 + 0]; }
  __device__ __forceinline__ float by(int dx, int dy) const { return reinterpret_cast<const float*>(c.base)[c.idx_near(x + dx, y + dy) * 4 + 1]; }
  __device__ __forceinline__ float bt(int dx, int dy) const { return reinterpret_cast<const float*>(c.base)[c.idx_near(x + dx, y + dy) * 4 + 3]; }
  __device__ __forceinline__ float4 water4(int dx, int dy) const { return c.water[c.idx_near(x + dx, y + dy)]; }
  __device__ __forceinline__ char4 wall4(int dx, int dy) const { return c.wall[c.idx_near(x + dx, y + dy)]; }
  __device__ __forceinline__ float2 vort(int dx, int dy) const { return c.vortf[c.idx_near(x + dx, y + dy)]; }
  __device__ __forceinline__ float4 light4(int dx, int dy) const { return c.light4(x + dx, y + dy); }
  __device__ __forceinline__ float4 fb4() const { return c.fb[(size_t)y * c.g.pitch + x]; }
  __device__ __forceinline__ float2 dep2() const { return c.dep[(size_t)y * c.g.pitch + x]; }
};

#define WSB_CELL_XY                                   \
  const int x = g.cx0 + blockIdx.x * blockDim.x + threadIdx.x; \
  const int y = blockIdx.y * blockDim.y + threadIdx.y;          \
  if (x >= g.cx1 || y >= g.H) return;                          \
  const size_t ci = (size_t)y * g.pitch + x;

// pass 1 — velocityShader.frag
__global__ void k_ref_velocity(GlobalCtx c, DevParams d, float4* __restrict__ baseOut, char4* __restrict__ wallOut) {
  const Geom& g = c.g;
  WSB_CELL_XY
  float4 b = c.base[ci];
  char4 w = c.wall[ci];
  velocity_cell(d, b.x, b.y, b.z, c.bp(x + 1, y), c.bp(x, y + 1), w.y);
  baseOut[ci] = b;
  wallOut[ci] = w;
}

// pass 2 — curlShader.frag
__global__ void k_ref_curl(GlobalCtx c, float* __restrict__ curlOut) {
  const Geom& g = c.g;
  WSB_CELL_XY
  float4 b = c.base[ci];
  curlOut[ci] = curl_cell(b.x, b.y, c.bx(x, y + 1), c.by(x + 1, y));
}

// pass 3 — vorticityShader.frag
__global__ void k_ref_vorticity(Geom g, const float* __restrict__ curl, float2* __restrict__ vortOut) {
  WSB_CELL_XY
  auto at = [&](int xx, int yy) { return curl[(size_t)wrap_y(yy, g.H) * g.pitch + wrap_x(g, xx)]; };
  vortOut[ci] = vorticity_cell(curl[ci], at(x - 1, y), at(x, y - 1), at(x + 1, y), at(x, y + 1));
}

// pass 4 — boundaryShader.frag
__global__ void k_ref_boundary(GlobalCtx c, DevParams d, const float* __restrict__ initial_T,
                               float4* __restrict__ baseOut, float4* __restrict__ waterOut, char4* __restrict__ wallOut) {
  const Geom& g = c.g;
  WSB_CELL_XY
  float4 b, w;
  char4 wl;
  boundary_cell(GlobalAt{c, x, y}, g, d, initial_T, x, y, b, w, wl);
  baseOut[ci] = b;
  waterOut[ci] = w;
  wallOut[ci] = wl;
}

// pass 5 — advectionShader.frag
template <bool DRY>
__global__ void k_ref_advection(GlobalCtx c, DevParams d, const float* __restrict__ initial_T,
                                const float* __restrict__ sndT, const float* __restrict__ sndW,
                                const float* __restrict__ sndV, float4* __restrict__ baseOut,
                                float4* __restrict__ waterOut, char4* __restrict__ wallOut, unsigned* __restrict__ maxv) {
  const Geom& g = c.g;
  WSB_CELL_XY
  float4 b, w;
  char4 wl;
  float vm = 0.0f;
  advection_cell<DRY>(c, g, d, initial_T, sndT, sndW, sndV, x, y, b, w, wl, vm);
  baseOut[ci] = b;
  waterOut[ci] = w;
  wallOut[ci] = wl;
  report_vmax(vm, maxv);
}

// pass 6 — pressureShader.frag
__global__ void k_ref_pressure(GlobalCtx c, float4* __restrict__ baseOut, char4* __restrict__ wallOut) {
  const Geom& g = c.g;
  WSB_CELL_XY
  float4 b = c.base[ci];
  char4 wYm = c.wall4(x, y - 1);
  pressure_cell(b.x, b.y, b.z, b.w, c.bx(x - 1, y), c.by(x, y - 1), c.bt(x, y - 1), wYm.x, wYm.y);
  baseOut[ci] = b;
  wallOut[ci] = c.wall[ci];
}

// pass 7 — lightingShader.frag
__global__ void k_ref_lighting(GlobalCtx c, DevParams d, float4* __restrict__ lightOut) {
  const Geom& g = c.g;
  WSB_CELL_XY
  float4 b = c.base[ci];
  lightOut[ci] = lighting_cell(c, g, d, x, y, b.w, c.water[ci], c.wall[ci], c.bt(x, y - 1));
}

}  // namespace wsb
