// wsb_ref_kernels.cuh — device view of the simulation state (GlobalCtx) and the REFERENCE
// schedule: one kernel per reference pass, every neighbour fetch straight from global memory
// (L1/L2 do the reuse).  Same order and bindings as app.js:5830-6005.  Used for per-pass parity
// against the oracle and as the "recompiled baseline" the fused kernels are measured against; the
// product path is wsb_fused_kernels.cuh.
//
// HBM layout: the C ABI speaks the reference's packed RGBA texels; inside the
// library every channel of base / water / light is its own float plane [H][pitch] (wall stays one
// packed 32-bit plane).  Channel planes are what the stencils want: a warp reads 128 contiguous
// bytes per channel, kernels only touch the channels they use (the boundary pass reads 2 of the 4
// light channels, the lighting pass 3), and a TMA box of one plane lands in shared memory as the
// dense [rows][columns] float tile the gathers index — no de-interleaving on the SM.
#pragma once
#include "wsb_cells.cuh"

namespace wsb {

__device__ __forceinline__ char4 as_char4(int w) { return *reinterpret_cast<const char4*>(&w); }
__device__ __forceinline__ int as_int(char4 c) { return *reinterpret_cast<const int*>(&c); }

struct Planes4 {  // one RGBA32F texture as four channel planes
  float* c[4];
  // cell indices are 32-bit (wsb_create caps the grid at 2^30 cells): an unsigned index lets the
  // compiler address every plane as [uniform 64-bit base + 32-bit register offset]
  __device__ __forceinline__ float4 ld(size_t i) const { return make_float4(c[0][i], c[1][i], c[2][i], c[3][i]); }
  // (streaming / evict-first stores, __stcs: no difference on any kernel, profiles/r3_logs/c28_variants.log)
  __device__ __forceinline__ void st(size_t i, float4 v) const { c[0][i] = v.x; c[1][i] = v.y; c[2][i] = v.z; c[3][i] = v.w; }
};

// Read-only view of one set of bindings (which ping-pong copy plays which role is the host's choice).
struct GlobalCtx {
  Planes4 base, water, light;
  const int* __restrict__ wall;       // packed (type | dist << 8 | vert << 16 | veg << 24)
  const float2* __restrict__ vortf;   // REFERENCE schedule only
  const float4* __restrict__ fb;      // precipitation feedback (RGBA32F texels; sprites add with vector atomics)
  const float2* __restrict__ dep;     // precipitation deposition
  Geom g;
  // any (x, y): periodic in y, periodic (single domain) or clamped (strip) in x
  __device__ __forceinline__ size_t idx(int x, int y) const { return (size_t)mod_i(y, g.H) * g.pitch + gather_x(g, x); }
  // (x, y) at most one period outside the array (fixed +-1 stencils): no integer division
  __device__ __forceinline__ size_t idx_near(int x, int y) const { return (size_t)wrap_y(y, g.H) * g.pitch + wrap_x(g, x); }
  __device__ __forceinline__ float4 base4(int x, int y) const { return base.ld(idx(x, y)); }
  __device__ __forceinline__ float bx(int x, int y) const { return base.c[0][idx(x, y)]; }
  __device__ __forceinline__ float by(int x, int y) const { return base.c[1][idx(x, y)]; }
  __device__ __forceinline__ float bp(int x, int y) const { return base.c[2][idx(x, y)]; }
  __device__ __forceinline__ float bt(int x, int y) const { return base.c[3][idx(x, y)]; }
  __device__ __forceinline__ float sbx(int x, int y) const { return bx(x, y); }
  __device__ __forceinline__ float sby(int x, int y) const { return by(x, y); }
  __device__ __forceinline__ float sbt(int x, int y) const { return bt(x, y); }
  __device__ __forceinline__ int swdist(int x, int y) const { return wdist(x, y); }
  __device__ __forceinline__ float4 water4(int x, int y) const { return water.ld(idx(x, y)); }
  __device__ __forceinline__ float wt0(int x, int y) const { return water.c[0][idx(x, y)]; }
  __device__ __forceinline__ float wt1(int x, int y) const { return water.c[1][idx(x, y)]; }
  __device__ __forceinline__ float wt2(int x, int y) const { return water.c[2][idx(x, y)]; }
  __device__ __forceinline__ float wt3(int x, int y) const { return water.c[3][idx(x, y)]; }
  __device__ __forceinline__ char4 wall4(int x, int y) const { return as_char4(wall[idx(x, y)]); }
  __device__ __forceinline__ int wdist(int x, int y) const { return as_char4(wall[idx(x, y)]).y; }
  __device__ __forceinline__ float2 vort(int x, int y) const { return vortf[idx(x, y)]; }
  // light texture: wrap S = REPEAT, wrap T = CLAMP_TO_EDGE (app.js:5276-5279)
  __device__ __forceinline__ size_t lidx(int x, int y) const { return (size_t)min(max(y, 0), g.H - 1) * g.pitch + wrap_x(g, x); }
  __device__ __forceinline__ float4 light4(int x, int y) const { return light.ld(lidx(x, y)); }
  __device__ __forceinline__ float lightS(int x, int y) const { return light.c[0][lidx(x, y)]; }
  __device__ __forceinline__ float lightIRdown(int x, int y) const { return light.c[2][lidx(x, y)]; }
  __device__ __forceinline__ float lightIRup(int x, int y) const { return light.c[3][lidx(x, y)]; }
  __device__ __forceinline__ float4 fb4(int x, int y) const { return fb[idx(x, y)]; }
  __device__ __forceinline__ float2 dep2(int x, int y) const { return dep[idx(x, y)]; }
};

// GlobalCtx positioned at a cell: the offset-style fetches boundary_cell asks for.
struct GlobalAt {
  const GlobalCtx& c;
  int x, y;
  __device__ __forceinline__ size_t at(int dx, int dy) const { return c.idx_near(x + dx, y + dy); }
  __device__ __forceinline__ float4 base4(int dx, int dy) const { return c.base.ld(at(dx, dy)); }
  __device__ __forceinline__ float bx(int dx, int dy) const { return c.base.c[0][at(dx, dy)]; }
  __device__ __forceinline__ float by(int dx, int dy) const { return c.base.c[1][at(dx, dy)]; }
  __device__ __forceinline__ float bt(int dx, int dy) const { return c.base.c[3][at(dx, dy)]; }
  __device__ __forceinline__ float4 water4(int dx, int dy) const { return c.water.ld(at(dx, dy)); }
  __device__ __forceinline__ char4 wall4(int dx, int dy) const { return as_char4(c.wall[at(dx, dy)]); }
  __device__ __forceinline__ float2 vort(int dx, int dy) const { return c.vortf[at(dx, dy)]; }
  __device__ __forceinline__ float4 light4(int dx, int dy) const { return c.light4(x + dx, y + dy); }
  __device__ __forceinline__ float4 fb4() const { return c.fb[(size_t)y * c.g.pitch + x]; }
  __device__ __forceinline__ float2 dep2() const { return c.dep[(size_t)y * c.g.pitch + x]; }
};

#ifdef __CUDACC__  // kernels: device only (GlobalCtx / GlobalAt above also compile for the host, tests/host_cells)
#define WSB_CELL_XY                                   \
  const int x = g.cx0 + blockIdx.x * blockDim.x + threadIdx.x; \
  const int y = blockIdx.y * blockDim.y + threadIdx.y;          \
  if (x >= g.cx1 || y >= g.H) return;                          \
  const size_t ci = (size_t)y * g.pitch + x;

// pass 1 — velocityShader.frag
__global__ void k_ref_velocity(GlobalCtx c, DevParams d, Planes4 baseOut, int* __restrict__ wallOut) {
  const Geom& g = c.g;
  WSB_CELL_XY
  float4 b = c.base.ld(ci);
  const int w = c.wall[ci];
  velocity_cell(d, b.x, b.y, b.z, c.bp(x + 1, y), c.bp(x, y + 1), as_char4(w).y);
  baseOut.st(ci, b);
  wallOut[ci] = w;
}

// pass 2 — curlShader.frag
__global__ void k_ref_curl(GlobalCtx c, float* __restrict__ curlOut) {
  const Geom& g = c.g;
  WSB_CELL_XY
  curlOut[ci] = curl_cell(c.base.c[0][ci], c.base.c[1][ci], c.bx(x, y + 1), c.by(x + 1, y));
}

// pass 3 — vorticityShader.frag
__global__ void k_ref_vorticity(Geom g, const float* __restrict__ curl, float2* __restrict__ vortOut) {
  WSB_CELL_XY
  auto at = [&](int xx, int yy) { return curl[(size_t)wrap_y(yy, g.H) * g.pitch + wrap_x(g, xx)]; };
  vortOut[ci] = vorticity_cell(curl[ci], at(x - 1, y), at(x, y - 1), at(x + 1, y), at(x, y + 1));
}

// pass 4 — boundaryShader.frag
__global__ void k_ref_boundary(GlobalCtx c, DevParams d, const float* __restrict__ initial_T, Planes4 baseOut, Planes4 waterOut,
                               int* __restrict__ wallOut) {
  const Geom& g = c.g;
  WSB_CELL_XY
  float4 b, w;
  char4 wl;
  boundary_cell(GlobalAt{c, x, y}, g, d, initial_T, x, y, b, w, wl);
  baseOut.st(ci, b);
  waterOut.st(ci, w);
  wallOut[ci] = as_int(wl);
}

// pass 5 — advectionShader.frag
template <bool DRY>
__global__ void k_ref_advection(GlobalCtx c, DevParams d, const float* __restrict__ initial_T,
                                const float* __restrict__ sndT, const float* __restrict__ sndW,
                                const float* __restrict__ sndV, Planes4 baseOut, Planes4 waterOut, int* __restrict__ wallOut,
                                unsigned* __restrict__ maxv) {
  const Geom& g = c.g;
  WSB_CELL_XY
  float4 b, w;
  char4 wl;
  float vm = 0.0f;
  advection_cell<DRY>(c, g, d, initial_T, sndT, sndW, sndV, x, y, b, w, wl, vm);
  baseOut.st(ci, b);
  waterOut.st(ci, w);
  wallOut[ci] = as_int(wl);
  report_vmax(vm, maxv);
}

// pass 6 — pressureShader.frag
__global__ void k_ref_pressure(GlobalCtx c, Planes4 baseOut, int* __restrict__ wallOut) {
  const Geom& g = c.g;
  WSB_CELL_XY
  float4 b = c.base.ld(ci);
  char4 wYm = c.wall4(x, y - 1);
  pressure_cell(b.x, b.y, b.z, b.w, c.bx(x - 1, y), c.by(x, y - 1), c.bt(x, y - 1), wYm.x, wYm.y);
  baseOut.st(ci, b);
  wallOut[ci] = c.wall[ci];
}

// pass 7 — lightingShader.frag
__global__ void k_ref_lighting(GlobalCtx c, DevParams d, Planes4 lightOut) {
  const Geom& g = c.g;
  WSB_CELL_XY
  lightOut.st(ci, lighting_cell(c, g, d, x, y, (float)global_x(g, x) + 0.5f, c.base.c[3][ci], c.water.ld(ci), as_char4(c.wall[ci]),
                                c.bt(x, y - 1)));
}

// ---------------------------------------------------------------------------------------------
// Boundary layout <-> plane layout (wsb_upload / wsb_read_rect): packed RGBA32F texels <-> planes
// ---------------------------------------------------------------------------------------------
__global__ void k_texels_to_planes(const float4* __restrict__ src, size_t n, Planes4 dst) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst.st(i, src[i]);
}
// rectangle [x0, x0+w) x [y0, y0+h) of the planes -> dense [h][w] texels
__global__ void k_planes_to_texels(Planes4 src, int pitch, int x0, int y0, int w, int h, float4* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i < w && j < h) dst[(size_t)j * w + i] = src.ld((size_t)(y0 + j) * pitch + x0 + i);
}
#endif  // __CUDACC__

}  // namespace wsb
