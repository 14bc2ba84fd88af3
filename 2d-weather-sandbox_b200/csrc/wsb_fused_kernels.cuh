// wsb_fused_kernels.cuh — the product path: one iteration of the reference loop in two fused
// stencil kernels (+ the particle kernel), and the fused "dry sweep".
//
//   k_fused_pvb   pressure(prev iteration) -> velocity -> curl -> vorticity -> boundary
//                 reads  base_1, wall_1 (tile + 3-cell halo, staged in shared memory as SoA planes),
//                        water_1, light_0, feedback, deposition (own cell, sparse neighbours)
//                 writes base_0, water_0, wall_0                      88 (+24) B / cell
//   k_fused_adv   advection (+ condensation, forcing, wall cells, brush, airplane) -> lighting
//                 reads  base_0, water_0, wall_0 (tile + 2-cell halo in shared memory), light_src
//                 writes base_1, water_1, wall_1, light_dst           104 B / cell
//   k_fused_dry   velocity -> advection(base only) -> pressure        36 B / cell
//
// The reference's pressure pass (last grid pass of iteration i) is folded into the first kernel of
// iteration i+1, so base "after pressure" never travels through HBM; wsb_read_rect materialises it
// on demand for the rectangle being read.  curl and vortForce never leave shared memory.
//
// Tiles are loaded with coalesced 16-byte loads (one float4 cell per lane, 512 B per warp) and
// transposed into per-channel planes so that the stencil and the bilinear back-trace gathers read
// 4-byte words from conflict-free consecutive banks instead of 16-byte AoS cells (4x less
// shared-memory traffic on the channel-granular gathers).  TMA is deliberately not used: a
// tensor-map box lands in shared memory in the AoS layout and cannot apply the periodic wrap at
// the domain edge; see DESIGN.md.
#pragma once
#include "wsb_cells.cuh"
#include "wsb_ref_kernels.cuh"

namespace wsb {

constexpr int kTX = 64;   // tile width  (cells) — 2 warps wide, 1 KiB of float4 per row
constexpr int kTY = 16;   // tile height (cells)
constexpr int kNT = 256;  // threads per CTA

// ---------------------------------------------------------------------------------------------
// k_fused_pvb
// ---------------------------------------------------------------------------------------------
constexpr int kH1 = 3;                      // halo of the pressure->boundary chain
constexpr int kSW1 = kTX + 2 * kH1;         // 70
constexpr int kSH1 = kTY + 2 * kH1;         // 22
constexpr int kN1 = kSW1 * kSH1;            // 1540 cells per staged tile
constexpr size_t kSmem1 = (size_t)kN1 * 4 * 8;  // vx vy P T(/curl) T2 wall vfx vfy

struct PvbCtx {  // boundary_cell context: base / wall / vortForce from the tile, the rest from HBM
  const float *sVX, *sVY, *sP, *sT2, *sVFX, *sVFY;
  const int* sWall;
  int X0, Y0;  // cell coordinates of tile-region element (0,0)
  GlobalCtx glob;
  // feedback / deposition are read AND cleared by this kernel (own cell only): plain pointers,
  // not the read-only path.  useFb = 0: both targets are known to be all zero, skip the reads.
  const float4* fbp;
  const float2* depp;
  int useFb;
  __device__ __forceinline__ int si(int x, int y) const { return (y - Y0) * kSW1 + (x - X0); }
  __device__ __forceinline__ float4 base4(int x, int y) const { int s = si(x, y); return make_float4(sVX[s], sVY[s], sP[s], sT2[s]); }
  __device__ __forceinline__ float bx(int x, int y) const { return sVX[si(x, y)]; }
  __device__ __forceinline__ float by(int x, int y) const { return sVY[si(x, y)]; }
  __device__ __forceinline__ float bt(int x, int y) const { return sT2[si(x, y)]; }
  __device__ __forceinline__ char4 wall4(int x, int y) const {
    int v = sWall[si(x, y)];
    return *reinterpret_cast<char4*>(&v);
  }
  __device__ __forceinline__ float2 vort(int x, int y) const { int s = si(x, y); return make_float2(sVFX[s], sVFY[s]); }
  __device__ __forceinline__ float4 water4(int x, int y) const { return glob.water4(x, y); }
  __device__ __forceinline__ float4 light4(int x, int y) const { return glob.light4(x, y); }
  __device__ __forceinline__ float4 fb4(int x, int y) const {
    return useFb ? fbp[glob.idx(x, y)] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __device__ __forceinline__ float2 dep2(int x, int y) const { return useFb ? depp[glob.idx(x, y)] : make_float2(0.f, 0.f); }
};

// glob: base = base_1, wall = wall_1, water = water_1, light = light_0 (glob.fb / glob.dep unused).
// useFb: feedback / deposition hold data from the last particle pass; the kernel consumes them and
// writes the zeros of the reference's gl.clear (app.js:5933-5934) back to the cells that were hit.
__global__ void __launch_bounds__(kNT) k_fused_pvb(GlobalCtx glob, DevParams d, const float* __restrict__ initial_T,
                                                   int applyPressure, int useFb, float4* fb, float2* dep,
                                                   float4* __restrict__ baseOut, float4* __restrict__ waterOut,
                                                   char4* __restrict__ wallOut) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* sVX = reinterpret_cast<float*>(smem_raw);
  float* sVY = sVX + kN1;
  float* sP = sVY + kN1;
  float* sT = sP + kN1;    // pre-pressure T, later reused for curl
  float* sT2 = sT + kN1;   // post-pressure T
  int* sWall = reinterpret_cast<int*>(sT2 + kN1);
  float* sVFX = reinterpret_cast<float*>(sWall + kN1);
  float* sVFY = sVFX + kN1;
  float* sCurl = sT;

  const Geom& g = glob.g;
  const int tid = threadIdx.x;
  const int X0 = g.cx0 + blockIdx.x * kTX - kH1, Y0 = blockIdx.y * kTY - kH1;

  // S0: stage base_1 / wall_1 tile + halo (periodic wrap) as SoA planes
  for (int s = tid; s < kN1; s += kNT) {
    const int j = s / kSW1, i = s - j * kSW1;
    const size_t ci = (size_t)wrap_y(Y0 + j, g.H) * g.pitch + wrap_x(g, X0 + i);
    const float4 b = glob.base[ci];
    sVX[s] = b.x; sVY[s] = b.y; sP[s] = b.z; sT[s] = b.w;
    sWall[s] = reinterpret_cast<const int*>(glob.wall)[ci];
  }
  __syncthreads();

  // S1: pressure pass of the previous iteration (pressureShader.frag), valid for i,j >= 1
  for (int s = tid; s < kN1; s += kNT) {
    const int j = s / kSW1, i = s - j * kSW1;
    float P = sP[s], T = sT[s];
    if (applyPressure && i >= 1 && j >= 1) {
      const int wv = sWall[s - kSW1];
      const char4 wYm = *reinterpret_cast<const char4*>(&wv);
      pressure_cell(sVX[s], sVY[s], P, T, sVX[s - 1], sVY[s - kSW1], sT[s - kSW1], wYm.x, wYm.y);
    }
    sP[s] = P;   // P only depends on velocities: in-place is safe
    sT2[s] = T;  // T reads T(y-1) of the input: separate plane
  }
  __syncthreads();

  // S2: velocity (velocityShader.frag), needs P(i+1), P(j+1): valid for 1 <= i < SW-1, 1 <= j < SH-1
  for (int s = tid; s < kN1; s += kNT) {
    const int j = s / kSW1, i = s - j * kSW1;
    if (i < kSW1 - 1 && j < kSH1 - 1) {
      const int wv = sWall[s];
      float vx = sVX[s], vy = sVY[s];
      velocity_cell(d, vx, vy, sP[s], sP[s + 1], sP[s + kSW1], (int)(*reinterpret_cast<const char4*>(&wv)).y);
      sVX[s] = vx;
      sVY[s] = vy;
    }
  }
  __syncthreads();

  // S3: curl (curlShader.frag): valid for 1 <= i < SW-2, 1 <= j < SH-2
  for (int s = tid; s < kN1; s += kNT) {
    const int j = s / kSW1, i = s - j * kSW1;
    if (i < kSW1 - 2 && j < kSH1 - 2) sCurl[s] = curl_cell(sVX[s], sVY[s], sVX[s + kSW1], sVY[s + 1]);
  }
  __syncthreads();

  // S4: vorticity force (vorticityShader.frag): valid for 2 <= i < SW-3, 2 <= j < SH-3
  for (int s = tid; s < kN1; s += kNT) {
    const int j = s / kSW1, i = s - j * kSW1;
    if (i >= 2 && i < kSW1 - 3 && j >= 2 && j < kSH1 - 3) {
      const float2 vf = vorticity_cell(sCurl[s], sCurl[s - 1], sCurl[s - kSW1], sCurl[s + 1], sCurl[s + kSW1]);
      sVFX[s] = vf.x;
      sVFY[s] = vf.y;
    }
  }
  __syncthreads();

  // S5: boundary pass on the TX x TY interior
  PvbCtx c{sVX, sVY, sP, sT2, sVFX, sVFY, sWall, X0, Y0, glob, fb, dep, useFb};
  const int tx = tid % kTX, ty0 = tid / kTX;
#pragma unroll 1
  for (int ty = ty0; ty < kTY; ty += kNT / kTX) {
    const int x = X0 + kH1 + tx, y = Y0 + kH1 + ty;
    if (x < g.cx1 && y < g.H) {
      float4 b, w;
      char4 wl;
      boundary_cell(c, g, d, initial_T, x, y, b, w, wl);
      const size_t ci = (size_t)y * g.pitch + x;
      baseOut[ci] = b;
      waterOut[ci] = w;
      wallOut[ci] = wl;
      if (useFb) {  // sprites are sparse: only cells that were hit cost a write
        const float4 f = fb[ci];
        if (f.x != 0.0f || f.y != 0.0f || f.z != 0.0f || f.w != 0.0f) fb[ci] = make_float4(0.f, 0.f, 0.f, 0.f);
        const float2 dd = dep[ci];
        if (dd.x != 0.0f || dd.y != 0.0f) dep[ci] = make_float2(0.f, 0.f);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// k_fused_adv
// ---------------------------------------------------------------------------------------------
constexpr int kH2 = 2;                  // halo: covers every back-trace with |v| < 1
constexpr int kSW2 = kTX + 2 * kH2;     // 68
constexpr int kSH2 = kTY + 2 * kH2;     // 20
constexpr int kN2 = kSW2 * kSH2;        // 1360
constexpr size_t kSmem2 = (size_t)kN2 * 4 * 9;  // vx vy P T w0 w1 w2 w3 wall

// advection_cell context: tile planes with a bounds check and an exact HBM fallback for
// back-traces that leave the halo (|v| >= 1 cell / iteration).
struct AdvCtx {
  const float *sVX, *sVY, *sP, *sT, *sW0, *sW1, *sW2, *sW3;
  const int* sWall;
  int X0, Y0;
  GlobalCtx glob;  // base_0, water_0, wall_0, light_src
  __device__ __forceinline__ bool in(int x, int y, int& s) const {
    const unsigned i = (unsigned)(x - X0), j = (unsigned)(y - Y0);
    s = (int)(j * kSW2 + i);
    return i < (unsigned)kSW2 && j < (unsigned)kSH2;
  }
  __device__ __forceinline__ int si(int x, int y) const { return (y - Y0) * kSW2 + (x - X0); }
#define WSB_ADV_ACC(name, plane, fallback) \
  __device__ __forceinline__ float name(int x, int y) const { int s; return in(x, y, s) ? plane[s] : glob.fallback(x, y); }
  WSB_ADV_ACC(bx, sVX, bx) WSB_ADV_ACC(by, sVY, by) WSB_ADV_ACC(bp, sP, bp) WSB_ADV_ACC(bt, sT, bt)
  WSB_ADV_ACC(wt0, sW0, wt0) WSB_ADV_ACC(wt1, sW1, wt1) WSB_ADV_ACC(wt2, sW2, wt2) WSB_ADV_ACC(wt3, sW3, wt3)
#undef WSB_ADV_ACC
  __device__ __forceinline__ int wdist(int x, int y) const {
    int s;
    if (in(x, y, s)) { int v = sWall[s]; return (int)(*reinterpret_cast<char4*>(&v)).y; }
    return glob.wdist(x, y);
  }
  // fixed +-1 stencil around a cell whose own coordinates are inside the tile or its inner halo
  __device__ __forceinline__ float sbx(int x, int y) const { return bx(x, y); }
  __device__ __forceinline__ float sby(int x, int y) const { return by(x, y); }
  __device__ __forceinline__ float sbt(int x, int y) const { return bt(x, y); }
  __device__ __forceinline__ int swdist(int x, int y) const { return wdist(x, y); }
  __device__ __forceinline__ float4 base4(int x, int y) const {
    int s;
    if (in(x, y, s)) return make_float4(sVX[s], sVY[s], sP[s], sT[s]);
    return glob.base4(x, y);
  }
  __device__ __forceinline__ float4 water4(int x, int y) const {
    int s;
    if (in(x, y, s)) return make_float4(sW0[s], sW1[s], sW2[s], sW3[s]);
    return glob.water4(x, y);
  }
  __device__ __forceinline__ char4 wall4(int x, int y) const {
    int s;
    if (in(x, y, s)) { int v = sWall[s]; return *reinterpret_cast<char4*>(&v); }
    return glob.wall4(x, y);
  }
  __device__ __forceinline__ float lightS(int x, int y) const { return glob.lightS(x, y); }
  __device__ __forceinline__ float lightIRdown(int x, int y) const { return glob.lightIRdown(x, y); }
  __device__ __forceinline__ float lightIRup(int x, int y) const { return glob.lightIRup(x, y); }
};

// base_1 temperature of a cell, i.e. the advection result for that cell, computed on demand: the
// lighting pass needs it for the cell BELOW a water-surface air cell (lightingShader.frag:105).
__device__ __noinline__ float advected_T(const AdvCtx& c, const Geom& g, const DevParams& d,
                                         const float* __restrict__ initial_T, const float* __restrict__ sndT,
                                         const float* __restrict__ sndW, const float* __restrict__ sndV, int x, int y) {
  float4 b, w;
  char4 wl;
  float vm = 0.0f;
  y = wrap_y(y, g.H);
  advection_cell<false>(c, g, d, initial_T, sndT, sndW, sndV, x, y, b, w, wl, vm);
  return b.w;
}

__global__ void __launch_bounds__(kNT) k_fused_adv(GlobalCtx glob, DevParams d, const float* __restrict__ initial_T,
                                                   const float* __restrict__ sndT, const float* __restrict__ sndW,
                                                   const float* __restrict__ sndV, float4* __restrict__ baseOut,
                                                   float4* __restrict__ waterOut, char4* __restrict__ wallOut,
                                                   float4* __restrict__ lightOut, unsigned* __restrict__ maxv) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* sVX = reinterpret_cast<float*>(smem_raw);
  float* sVY = sVX + kN2;
  float* sP = sVY + kN2;
  float* sT = sP + kN2;
  float* sW0 = sT + kN2;
  float* sW1 = sW0 + kN2;
  float* sW2 = sW1 + kN2;
  float* sW3 = sW2 + kN2;
  int* sWall = reinterpret_cast<int*>(sW3 + kN2);

  const Geom& g = glob.g;
  const int tid = threadIdx.x;
  const int X0 = g.cx0 + blockIdx.x * kTX - kH2, Y0 = blockIdx.y * kTY - kH2;

  for (int s = tid; s < kN2; s += kNT) {
    const int j = s / kSW2, i = s - j * kSW2;
    const size_t ci = (size_t)wrap_y(Y0 + j, g.H) * g.pitch + wrap_x(g, X0 + i);
    const float4 b = glob.base[ci];
    const float4 w = glob.water[ci];
    sVX[s] = b.x; sVY[s] = b.y; sP[s] = b.z; sT[s] = b.w;
    sW0[s] = w.x; sW1[s] = w.y; sW2[s] = w.z; sW3[s] = w.w;
    sWall[s] = reinterpret_cast<const int*>(glob.wall)[ci];
  }
  __syncthreads();

  AdvCtx c{sVX, sVY, sP, sT, sW0, sW1, sW2, sW3, sWall, X0, Y0, glob};
  const int tx = tid % kTX, ty0 = tid / kTX;
  float vm = 0.0f;
#pragma unroll 1
  for (int ty = ty0; ty < kTY; ty += kNT / kTX) {
    const int x = X0 + kH2 + tx, y = Y0 + kH2 + ty;
    if (x < g.cx1 && y < g.H) {
      float4 b, w;
      char4 wl;
      advection_cell<false>(c, g, d, initial_T, sndT, sndW, sndV, x, y, b, w, wl, vm);
      const size_t ci = (size_t)y * g.pitch + x;
      baseOut[ci] = b;
      waterOut[ci] = w;
      wallOut[ci] = wl;
      float TBelow = 0.0f;
      if (wl.y != 0 && wl.z == 1 && wl.x == WALLTYPE_WATER && (float)y + 0.5f < g.Hf - 1.0f)
        TBelow = advected_T(c, g, d, initial_T, sndT, sndW, sndV, x, y - 1);
      lightOut[ci] = lighting_cell(c, g, d, x, y, b.w, w, wl, TBelow);
    }
  }
  report_vmax(vm, maxv);
}

// ---------------------------------------------------------------------------------------------
// k_fused_dry — velocity -> advection(base) -> pressure, BASELINE config 2 / the headline sweep
// ---------------------------------------------------------------------------------------------
constexpr int kH3 = 3;
constexpr int kSW3 = kTX + 2 * kH3;  // 70
constexpr int kSH3 = kTY + 2 * kH3;  // 22
constexpr int kN3 = kSW3 * kSH3;
constexpr int kOW3 = kTX + 1, kOH3 = kTY + 1;  // advected region: tile + one column/row on the low side
constexpr int kNO3 = kOW3 * kOH3;
constexpr size_t kSmem3 = (size_t)kN3 * 4 * 5 + (size_t)kNO3 * 4 * 4;

// HBM fallback for the dry sweep: applies the velocity pass on the fly to whatever it fetches.
struct DryGlobalCtx {
  GlobalCtx glob;  // base_0, wall_0 (pre-velocity)
  DevParams d;
  __device__ __forceinline__ float4 post_velocity(int x, int y) const {
    float4 b = glob.base4(x, y);
    velocity_cell(d, b.x, b.y, b.z, glob.bp(x + 1, y), glob.bp(x, y + 1), glob.wdist(x, y));
    return b;
  }
};

struct DryCtx {
  const float *sVX, *sVY, *sP, *sT;
  const int* sWall;
  int X0, Y0;
  DryGlobalCtx dg;
  __device__ __forceinline__ bool in(int x, int y, int& s) const {
    // the last staged row / column holds pre-velocity values (velocity needs P(i+1), P(j+1))
    const unsigned i = (unsigned)(x - X0), j = (unsigned)(y - Y0);
    s = (int)(j * kSW3 + i);
    return i < (unsigned)(kSW3 - 1) && j < (unsigned)(kSH3 - 1);
  }
#define WSB_DRY_ACC(name, plane, comp) \
  __device__ __forceinline__ float name(int x, int y) const { int s; return in(x, y, s) ? plane[s] : dg.post_velocity(x, y).comp; }
  WSB_DRY_ACC(bx, sVX, x) WSB_DRY_ACC(by, sVY, y) WSB_DRY_ACC(bp, sP, z) WSB_DRY_ACC(bt, sT, w)
#undef WSB_DRY_ACC
  __device__ __forceinline__ float sbx(int x, int y) const { return bx(x, y); }
  __device__ __forceinline__ float sby(int x, int y) const { return by(x, y); }
  __device__ __forceinline__ float sbt(int x, int y) const { return bt(x, y); }
  __device__ __forceinline__ int wdist(int x, int y) const {
    int s;
    const unsigned i = (unsigned)(x - X0), j = (unsigned)(y - Y0);
    s = (int)(j * kSW3 + i);
    if (i < (unsigned)kSW3 && j < (unsigned)kSH3) { int v = sWall[s]; return (int)(*reinterpret_cast<char4*>(&v)).y; }
    return dg.glob.wdist(x, y);
  }
  __device__ __forceinline__ int swdist(int x, int y) const { return wdist(x, y); }
  __device__ __forceinline__ float4 base4(int x, int y) const {
    int s;
    if (in(x, y, s)) return make_float4(sVX[s], sVY[s], sP[s], sT[s]);
    return dg.post_velocity(x, y);
  }
  __device__ __forceinline__ char4 wall4(int x, int y) const {
    int s;
    const unsigned i = (unsigned)(x - X0), j = (unsigned)(y - Y0);
    s = (int)(j * kSW3 + i);
    if (i < (unsigned)kSW3 && j < (unsigned)kSH3) { int v = sWall[s]; return *reinterpret_cast<char4*>(&v); }
    return dg.glob.wall4(x, y);
  }
  // the dry sweep does not touch water
  __device__ __forceinline__ float4 water4(int, int) const { return make_float4(0.f, 0.f, 0.f, 0.f); }
  __device__ __forceinline__ float wt0(int, int) const { return 0.f; }
  __device__ __forceinline__ float wt1(int, int) const { return 0.f; }
  __device__ __forceinline__ float wt2(int, int) const { return 0.f; }
  __device__ __forceinline__ float wt3(int, int) const { return 0.f; }
};

__global__ void __launch_bounds__(kNT) k_fused_dry(GlobalCtx glob, DevParams d, float4* __restrict__ baseOut,
                                                   unsigned* __restrict__ maxv) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* sVX = reinterpret_cast<float*>(smem_raw);
  float* sVY = sVX + kN3;
  float* sP = sVY + kN3;
  float* sT = sP + kN3;
  int* sWall = reinterpret_cast<int*>(sT + kN3);
  float* oVX = reinterpret_cast<float*>(sWall + kN3);
  float* oVY = oVX + kNO3;
  float* oP = oVY + kNO3;
  float* oT = oP + kNO3;

  const Geom& g = glob.g;
  const int tid = threadIdx.x;
  const int X0 = g.cx0 + blockIdx.x * kTX - kH3, Y0 = blockIdx.y * kTY - kH3;

  for (int s = tid; s < kN3; s += kNT) {
    const int j = s / kSW3, i = s - j * kSW3;
    const size_t ci = (size_t)wrap_y(Y0 + j, g.H) * g.pitch + wrap_x(g, X0 + i);
    const float4 b = glob.base[ci];
    sVX[s] = b.x; sVY[s] = b.y; sP[s] = b.z; sT[s] = b.w;
    sWall[s] = reinterpret_cast<const int*>(glob.wall)[ci];
  }
  __syncthreads();

  // velocity in place (only the cell's own velocity changes); last row / column stay pre-velocity
  for (int s = tid; s < kN3; s += kNT) {
    const int j = s / kSW3, i = s - j * kSW3;
    if (i < kSW3 - 1 && j < kSH3 - 1) {
      const int wv = sWall[s];
      float vx = sVX[s], vy = sVY[s];
      velocity_cell(d, vx, vy, sP[s], sP[s + 1], sP[s + kSW3], (int)(*reinterpret_cast<const char4*>(&wv)).y);
      sVX[s] = vx;
      sVY[s] = vy;
    }
  }
  __syncthreads();

  // advection of the base field on the tile plus one column / row on the low side
  DryCtx c{sVX, sVY, sP, sT, sWall, X0, Y0, DryGlobalCtx{glob, d}};
  float vm = 0.0f;
  for (int o = tid; o < kNO3; o += kNT) {
    const int oj = o / kOW3, oi = o - oj * kOW3;
    const int x = X0 + kH3 - 1 + oi, y = Y0 + kH3 - 1 + oj;  // unwrapped cell coordinates
    float4 b, w;
    char4 wl;
    // coordinates used for fragCoord must be the wrapped ones
    const int xw = g.wrap ? (x < 0 ? x + g.pitch : (x >= g.pitch ? x - g.pitch : x)) : x;
    const int yw = wrap_y(y, g.H);
    DryCtx cw = c;
    cw.X0 = X0 + (xw - x);
    cw.Y0 = Y0 + (yw - y);
    advection_cell<true>(cw, g, d, nullptr, nullptr, nullptr, nullptr, xw, yw, b, w, wl, vm);
    oVX[o] = b.x; oVY[o] = b.y; oP[o] = b.z; oT[o] = b.w;
  }
  __syncthreads();

  // pressure on the tile
  const int tx = tid % kTX, ty0 = tid / kTX;
  for (int ty = ty0; ty < kTY; ty += kNT / kTX) {
    const int x = X0 + kH3 + tx, y = Y0 + kH3 + ty;
    if (x < g.cx1 && y < g.H) {
      const int o = (ty + 1) * kOW3 + (tx + 1);
      const int wv = sWall[(ty + kH3 - 1) * kSW3 + (tx + kH3)];
      const char4 wYm = *reinterpret_cast<const char4*>(&wv);
      float4 b = make_float4(oVX[o], oVY[o], oP[o], oT[o]);
      pressure_cell(b.x, b.y, b.z, b.w, oVX[o - 1], oVY[o - kOW3], oT[o - kOW3], wYm.x, wYm.y);
      baseOut[(size_t)y * g.pitch + x] = b;
    }
  }
  report_vmax(vm, maxv);
}

// pressure pass on a rectangle, for readbacks of frameBuff_0's base in the fused schedule
__global__ void k_pressure_rect(GlobalCtx c, int x0, int y0, int w, int h, float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= w || j >= h) return;
  const int x = x0 + i, y = y0 + j;
  float4 b = c.base4(x, y);
  char4 wYm = c.wall4(x, y - 1);
  pressure_cell(b.x, b.y, b.z, b.w, c.bx(x - 1, y), c.by(x, y - 1), c.bt(x, y - 1), wYm.x, wYm.y);
  out[(size_t)j * w + i] = b;
}

}  // namespace wsb
