// wsb_fused_kernels.cuh — the product path: one iteration of the reference loop in two fused
// stencil kernels (+ the particle kernel), and the fused "dry sweep".
//
//   k_fused_pvb   pressure(prev iteration) -> velocity -> curl -> vorticity -> boundary
//                 reads  base_1, wall_1 (tile + 3-cell halo staged in shared memory),
//                        water_1, light_0 (own cell, prefetched; sparse neighbours), feedback,
//                        deposition (only after a particle pass)
//                 writes base_0, water_0, wall_0                      88 (+24) B / cell
//   k_fused_adv   advection (+ condensation, forcing, wall cells, brush, airplane) -> lighting
//                 reads  base_0, water_0, wall_0, light_src (tile + 2-cell halo staged)
//                 writes base_1, water_1, wall_1, light_dst           104 B / cell
//   k_fused_dry   pressure(prev iteration) -> velocity -> advection(base only)   36 B / cell
//
// The reference's pressure pass (last grid pass of iteration i) is folded into the first kernel of
// iteration i+1, so base "after pressure" never travels through HBM; wsb_read_rect materialises it
// on demand for the rectangle being read.  curl and vortForce never leave shared memory.
//
// Staging: the state lives in HBM as per-channel float planes (wsb_ref_kernels.cuh), so ONE elected
// thread per CTA issues one TMA box load (cp.async.bulk.tensor.2d -> UTMALDG) per plane: tile + halo
// land in shared memory as dense [rows][columns] float planes, signalled through an mbarrier with
// the expected byte count.  No thread spends an instruction, a register or an LSU wavefront on
// staging, and all of a CTA's HBM reads are in flight at once.  Tiles that need the periodic wrap
// of the reference's REPEAT textures (the outermost ring of tiles) — or grids whose row pitch is
// not a multiple of 16 bytes — take a register-staged fallback (stage_tile) with explicit wrap.
// Stencil passes then sweep the staged planes in place; the final per-cell pass gathers with
// plain shared-memory indices, one wavefront per gather.  A cell with a velocity component of
// 0.9 cells / iteration or more (never seen in the shipped saves, which peak at 0.36) takes an
// exact, slow global-memory path instead of the near back-trace (near_tap below).
#pragma once
#ifndef WSB_HOST_EMU
#include <cuda.h>  // CUtensorMap (type only; the encoder is resolved at run time)
#endif
// WSB_HOST_EMU: tests/host_cells/ compiles this file for the HOST (one OS thread per CUDA thread,
// TMA boxes and mbarriers emulated by emu::) so that the tile plumbing of the kernels — staging,
// halos, sweeps, indices, the near back-trace — is checked against the oracle on the CPU.  Test
// infrastructure only: the emulation header defines CUtensorMap, threadIdx, __syncthreads, ...

#include "wsb_cells.cuh"
#include "wsb_ref_kernels.cuh"

namespace wsb {

constexpr int kTX = 64;   // tile width  (cells) — 2 warps wide, 1 KiB of float4 per row
constexpr int kTY = 16;   // tile height (cells)
constexpr int kNT = 256;  // threads per CTA: thread (tx, ty0) computes rows ty0, ty0+4, ty0+8, ty0+12
constexpr int kRowStep = kNT / kTX;
// Every staged tile carries 4 extra columns on each side: the innermost TMA box coordinate must be
// 16-byte aligned (4 floats), and the stencils need 2..3 of them anyway.
constexpr int kHX = 4;

// Sprite bookkeeping of the particle pass (wsb_particles.cuh), consumed by k_fused_pvb.
// Origin grids: [H + 1][W + 1] (the sprite of a droplet at the right / top border starts at pixel W - 6 / H - 6).
struct SpriteGrid {
  float4* org4;          // feedback origins
  float2* org2;          // deposition origins
  int* dirty;            // [tilesY][tilesX] bit 0: a sprite (or a direct 1-pixel add) touched this tile's feedback texels,
                         //                  bit 1: ... its deposition texels.  Set by the particle pass, cleared by the boundary kernel
  int* dirtyList;        // the tiles whose word went 0 -> non-zero in this particle pass, in arrival order ...
  int* dirtyCount;       // ... and how many (k_boxsum / k_clear_origins walk the list; the latter resets the count)
  int Po;                // origin row pitch = W + 1
  int tilesX, tilesY;
};
enum { kDirtyFb = 1, kDirtyDep = 2 };

// ---------------------------------------------------------------------------------------------
// Tile staging
// ---------------------------------------------------------------------------------------------
// packed wall word (type | dist << 8 | vert << 16 | veg << 24)
__device__ __forceinline__ bool wl_is_wall(int w) { return (w & 0xff00) == 0; }            // DISTANCE == 0
__device__ __forceinline__ bool wl_is_land_wall(int w) { return (w & 0xffff) == 0x0001; }  // DISTANCE == 0 && TYPE == LAND

// ---- TMA: one box per plane, completion through an mbarrier -----------------------------------
template <int N>
struct TileMaps { CUtensorMap m[N]; };  // passed as a __grid_constant__ kernel parameter

#ifndef WSB_HOST_EMU
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");  // make the init visible to the async proxy
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  unsigned done;
  do {
    asm volatile(
        "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
        : "=r"(done)
        : "r"(smem_addr(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
// box with lower-left corner (x, y) of the plane described by `map` -> dense rows at `dst`;
// coordinates outside the plane are filled with zeros
__device__ __forceinline__ void tma_load_box(void* dst, const CUtensorMap* map, int x, int y, unsigned long long* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   smem_addr(dst)),
               "l"(map), "r"(x), "r"(y), "r"(smem_addr(bar))
               : "memory");
}
#define WSB_DYN_SMEM(name) extern __shared__ __align__(128) unsigned char name[]
#else  // host emulation of the five primitives above (tests/host_cells/cuda_emu.h)
inline void mbar_init(unsigned long long* bar, unsigned count) { emu::mbar_init(bar, count); }
inline void mbar_expect_tx(unsigned long long* bar, unsigned bytes) { emu::mbar_expect_tx(bar, bytes); }
inline void mbar_wait(unsigned long long* bar, unsigned parity) { emu::mbar_wait(bar, parity); }
inline void tma_load_box(void* dst, const CUtensorMap* map, int x, int y, unsigned long long* bar) { emu::tma_load_box(dst, map, x, y, bar); }
#define WSB_DYN_SMEM(name) unsigned char* name = emu::dyn_smem()
#endif
// compile-time switches for A/B timing (make variant OUT=... EXTRA="-DWSB_OPT_NEAR=0 ...", profiles/tools/)
#ifndef WSB_OPT_NEAR
#define WSB_OPT_NEAR 1      // back-trace taps relative to the own cell when every |v| < 0.9
#endif
#ifndef WSB_OPT_AIRFAST
#define WSB_OPT_AIRFAST 1   // one test "all four texels are air" short-cuts the wall-aware bilerp weights
#endif
// can this tile be staged by TMA?  rows must not wrap; columns must not wrap on a periodic domain
// (a strip's out-of-range ghost-edge columns are zero-filled garbage nobody consumes)
template <int SW, int SH>
__device__ __forceinline__ bool tile_tma_ok(const Geom& g, int useTma, int X0, int Y0) {
  return useTma && Y0 >= 0 && Y0 + SH <= g.H && (!g.wrap || (X0 >= 0 && X0 + SW <= g.pitch));
}
// shared-memory plane stride in floats: TMA destinations are 128-byte aligned
template <int N> constexpr int plane_stride() { return (N + 31) / 32 * 32; }

// Fallback: stage the region [X0, X0+SW) x [Y0, Y0+SH) THROUGH REGISTERS: load(ci, cil) returns
// the cell's registers, store(s, regs) writes them to the shared-memory planes.  G cells per thread
// are loaded back to back before the first store.
//   s   shared-memory index of the cell
//   ci  global cell index with the reference's periodic wrap (REPEAT textures)
//   cil global cell index for the light texture (wrap S = REPEAT, wrap T = CLAMP_TO_EDGE, app.js:5276-5279)
// Interior tiles walk (row, column) incrementally without divisions.
template <int SW, int SH, int G, class Load, class Store>
__device__ __forceinline__ void stage_tile(const Geom& g, int X0, int Y0, Load load, Store store) {
  constexpr int N = SW * SH, R = (N + kNT - 1) / kNT;
  const int tid = threadIdx.x;
  using Regs = decltype(load(0, 0));
  const bool interior = (X0 >= 0) && (X0 + SW <= g.pitch) && (Y0 >= 0) && (Y0 + SH <= g.H);
  constexpr int DJ = kNT / SW, DI = kNT % SW;
  int i = tid % SW;
  int ci = (Y0 + tid / SW) * g.pitch + X0 + i;
  const int stepA = DJ * g.pitch + DI, stepB = g.pitch - SW;
#pragma unroll
  for (int r0 = 0; r0 < R; r0 += G) {
    Regs regs[G];
#pragma unroll
    for (int k = 0; k < G; k++) {
      const int s = tid + (r0 + k) * kNT;
      if (r0 + k < R && s < N) {
        if (interior) {
          regs[k] = load(ci, ci);
        } else {
          const int j = s / SW, ii = s - j * SW;
          const int xx = wrap_x(g, X0 + ii);
          regs[k] = load(wrap_y(Y0 + j, g.H) * g.pitch + xx, min(max(Y0 + j, 0), g.H - 1) * g.pitch + xx);
        }
      }
      i += DI;
      ci += stepA;
      if (i >= SW) { i -= SW; ci += stepB; }
    }
#pragma unroll
    for (int k = 0; k < G; k++) {
      const int s = tid + (r0 + k) * kNT;
      if (r0 + k < R && s < N) store(s, regs[k]);
    }
  }
}

struct BaseWallRegs { float vx, vy, p, t; int w; };

// ---------------------------------------------------------------------------------------------
// Exact slow paths (global memory, any coordinates).  __noinline__ with pointer arguments so that
// the hot kernels keep their contexts in registers.
// ---------------------------------------------------------------------------------------------
// Raw state = advection output with the pressure pass pending: applies pressure (optionally) and
// velocity on the fly to whatever it fetches.
struct RawPvCtx {
  const GlobalCtx& glob;
  const DevParams& d;
  int applyPressure;
  __device__ float4 post_pressure(int x, int y) const {
    float4 b = glob.base4(x, y);
    if (applyPressure) {
      const char4 wYm = glob.wall4(x, y - 1);
      pressure_cell(b.x, b.y, b.z, b.w, glob.bx(x - 1, y), glob.by(x, y - 1), glob.bt(x, y - 1), wYm.x, wYm.y);
    }
    return b;
  }
  __device__ float4 post_pv(int x, int y) const {
    float4 b = post_pressure(x, y);
    velocity_cell(d, b.x, b.y, b.z, post_pressure(x + 1, y).z, post_pressure(x, y + 1).z, glob.wdist(x, y));
    return b;
  }
  __device__ __forceinline__ float bx(int x, int y) const { return post_pv(x, y).x; }
  __device__ __forceinline__ float by(int x, int y) const { return post_pv(x, y).y; }
  __device__ __forceinline__ float bp(int x, int y) const { return post_pv(x, y).z; }
  __device__ __forceinline__ float bt(int x, int y) const { return post_pv(x, y).w; }
  __device__ __forceinline__ float sbx(int x, int y) const { return bx(x, y); }
  __device__ __forceinline__ float sby(int x, int y) const { return by(x, y); }
  __device__ __forceinline__ float sbt(int x, int y) const { return bt(x, y); }
  __device__ __forceinline__ float4 base4(int x, int y) const { return post_pv(x, y); }
  __device__ __forceinline__ char4 wall4(int x, int y) const { return glob.wall4(x, y); }
  __device__ __forceinline__ int wdist(int x, int y) const { return glob.wdist(x, y); }
  __device__ __forceinline__ int swdist(int x, int y) const { return glob.wdist(x, y); }
  // the dry sweep does not touch water
  __device__ __forceinline__ float4 water4(int, int) const { return make_float4(0.f, 0.f, 0.f, 0.f); }
  __device__ __forceinline__ float wt0(int, int) const { return 0.f; }
  __device__ __forceinline__ float wt1(int, int) const { return 0.f; }
  __device__ __forceinline__ float wt2(int, int) const { return 0.f; }
  __device__ __forceinline__ float wt3(int, int) const { return 0.f; }
};

__device__ __noinline__ float4 dry_advect_slow(const GlobalCtx* glob, const DevParams* d, int applyPressure, int x, int y,
                                               float* vm) {
  RawPvCtx c{*glob, *d, applyPressure};
  float4 b, w;
  char4 wl;
  advection_cell<true>(c, glob->g, *d, nullptr, nullptr, nullptr, nullptr, x, y, b, w, wl, *vm);
  return b;
}

struct AdvSlowOut { float4 base, water; char4 wall; float vm; };
__device__ __noinline__ void adv_cell_slow(const GlobalCtx* glob, const DevParams* d, const float* initial_T, const float* sndT,
                                           const float* sndW, const float* sndV, int x, int y, AdvSlowOut* out) {
  out->vm = 0.0f;
  advection_cell<false>(*glob, glob->g, *d, initial_T, sndT, sndW, sndV, x, y, out->base, out->water, out->wall, out->vm);
}

// ---------------------------------------------------------------------------------------------
// Shared-memory gathers of the semi-Lagrangian back-trace
// ---------------------------------------------------------------------------------------------
// Tile-local index of a bilerp's lower-left texel; `lxBase` = tile-local column of the cell whose
// fragCoord started the trace, gx its global column, Y0 the global row of tile row 0.
template <int SW>
__device__ __forceinline__ int tile_index(const BilerpSetup& b, int lxBase, int gx, int Y0, int& lx, int& ly) {
  lx = lxBase + (b.ix - gx);
  ly = b.iy - Y0;
  return ly * SW + lx;
}
// all four texels (lx..lx+1, ly..ly+1) inside [lo, SW-1-hi] x [lo, SH-1-hi]
template <int SW, int SH>
__device__ __forceinline__ bool tile_ok(int lx, int ly, int lo, int hi) {
  return (unsigned)(lx - lo) <= (unsigned)(SW - 2 - hi - lo) && (unsigned)(ly - lo) <= (unsigned)(SH - 2 - hi - lo);
}

// Near back-trace.  With every velocity component of the cell below 0.9 cells / iteration and a
// grid of at most 65536 x 65536 cells (fp32 spacing of the coordinates <= 2^-8), the sample
// position minus one half lies in [cell - 1, cell + 1), so floor() is `cell - 1` or `cell` and one
// comparison replaces the float->int conversions, the index arithmetic and the tile bounds test:
// the tap is an OFFSET from the cell's own tile index.  fx, fy are the same fp32 subtractions
// bilerp_setup performs, so the result is bit-identical to the general path.
// The threshold travels as Geom::nearV (0.9, or -1 = never for grids beyond 2^19 cells a side).
struct NearTap { bool left, down; float fx, fy; };
__device__ __forceinline__ NearTap near_tap(float posx, float posy, float gxf, float gxm1, float gyf, float gym1) {
  const float stx = posx - 0.5f, sty = posy - 0.5f;
  NearTap n;
  n.left = stx < gxf;
  n.down = sty < gyf;
  n.fx = stx - (n.left ? gxm1 : gxf);
  n.fy = sty - (n.down ? gym1 : gyf);
  return n;
}
// lower-left texel of the tap: `own` points at the cell's own entry, `below` one tile row lower
template <class T>
__device__ __forceinline__ const T* near_ptr(const NearTap& n, const T* own, const T* below) {
  const T* q = n.down ? below : own;
  return n.left ? q - 1 : q;
}
__device__ __forceinline__ float adv_vmax(const AdvVel& a) {
  return fmaxf(fmaxf(fmaxf(fabsf(a.Vxx), fabsf(a.Vxy)), fmaxf(fabsf(a.Vyx), fabsf(a.Vyy))), fmaxf(fabsf(a.Px), fabsf(a.Py)));
}
// wall-aware bilerp weights (common.glsl:234-251) of the 2 x 2 footprint whose lower-left texel is
// sWl[l]; the usual case — all four texels are air — is decided by one test on the DISTANCE bytes
template <int SW>
__device__ __forceinline__ WallMix tile_wall_mix(const int* sWl, int l, float fx, float fy) {
  const unsigned char* wd = reinterpret_cast<const unsigned char*>(sWl + l) + 1;  // DISTANCE byte; 0 = wall
  const unsigned a = wd[0], b = wd[4], c = wd[4 * SW], dd = wd[4 * SW + 4];
  if (WSB_OPT_AIRFAST && min(min(a, b), min(c, dd)) != 0u) return WallMix{fx, fx, fy};
  return wall_mix((int)a, (int)b, (int)c, (int)dd, fx, fy);
}

// ---- stencil sweeps over the staged planes ------------------------------------------------------
// Two horizontally adjacent cells per thread, 8-byte accesses.  (A four-cell / 16-byte variant executes ~25 % fewer
// instructions but measured 7 % SLOWER on the dry sweep, profiles/r2_dry_variants.md: the sweeps are bound by
// shared-memory wavefronts and by how many warps have work between two barriers, not by issue slots.)  Row strides and
// plane sizes are multiples of 4 floats and planes are 128-byte aligned, so an even index is 8-byte aligned.
// WALLS = false: the staged region is known to hold no wall cell (k_fused_dry's tile map): no wall plane is staged,
// the land-wall rule of the pressure pass cannot fire (T' == T, nothing to write) and every cell is fluid.
__device__ __forceinline__ float2 ld2(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ int2 ld2(const int* p) { return *reinterpret_cast<const int2*>(p); }
__device__ __forceinline__ void st2(float* p, float2 v) { *reinterpret_cast<float2*>(p) = v; }

// pressure pass of the previous iteration (pressureShader.frag:24-42); valid for i >= 1, j >= 1.
// P' only reads velocities: in place.  T' reads the raw T below: written to sT2.
template <int SW, int N, bool WALLS = true>
__device__ __forceinline__ void sweep_pressure(const float* sVX, const float* sVY, float* sP, const float* sT, float* sT2,
                                               const int* sWl, int applyPressure) {
  if (!WALLS && !applyPressure) return;
  for (int s = SW + 2 * (int)threadIdx.x; s < N; s += 2 * kNT) {
    float2 T;
    if (WALLS) T = ld2(sT + s);
    if (applyPressure) {
      if (WALLS) {
        const int2 wb = ld2(sWl + s - SW);
        const float2 Tb = ld2(sT + s - SW);
        if (wl_is_land_wall(wb.x)) T.x -= Tb.x - 1000.0f;
        if (wl_is_land_wall(wb.y)) T.y -= Tb.y - 1000.0f;
      }
      const float2 vx = ld2(sVX + s), vy = ld2(sVY + s), vyb = ld2(sVY + s - SW);
      const float vxm = sVX[s - 1];
      float2 P = ld2(sP + s);
      P.x += (vxm - vx.x + vyb.x - vy.x) * 0.45f;
      P.y += (vx.x - vx.y + vyb.y - vy.y) * 0.45f;
      st2(sP + s, P);
    }
    if (WALLS) st2(sT2 + s, T);
  }
}
// velocity pass (velocityShader.frag:40-61), in place; valid for 1 <= i <= SW-2, 1 <= j <= SH-2
// sAnyWall (optional, zeroed by the caller before an earlier barrier): set to 1 when any cell of
// the swept rows is a wall, so that the per-cell pass can drop its wall tests on all-air tiles.
template <int SW, int N, bool WALLS = true>
__device__ __forceinline__ void sweep_velocity(const DevParams& d, float* sVX, float* sVY, const float* sP, const int* sWl,
                                               unsigned* sAnyWall = nullptr) {
  bool any = false;
  for (int s = SW + 2 * (int)threadIdx.x; s < N - SW; s += 2 * kNT) {
    float2 vx = ld2(sVX + s), vy = ld2(sVY + s);
    const float2 P = ld2(sP + s), Pu = ld2(sP + s + SW);
    const float Pr = sP[s + 2];
    bool wa = false, wb = false;
    if (WALLS) {
      const int2 w = ld2(sWl + s);
      wa = wl_is_wall(w.x);
      wb = wl_is_wall(w.y);
    }
    any |= wa | wb;
    velocity_cell(d, vx.x, vy.x, P.x, P.y, Pu.x, wa ? 0 : 1);
    velocity_cell(d, vx.y, vy.y, P.y, Pr, Pu.y, wb ? 0 : 1);
    st2(sVX + s, vx);
    st2(sVY + s, vy);
  }
  if (sAnyWall && __any_sync(0xffffffffu, any) && (threadIdx.x & 31) == 0) *sAnyWall = 1u;
}

// ---------------------------------------------------------------------------------------------
// k_fused_dry — pressure(prev) -> velocity -> advection(base): BASELINE config 2 / headline sweep
// ---------------------------------------------------------------------------------------------
constexpr int kHD = 2;                   // raw halo: advection +-1 of post-velocity, velocity +1, pressure -1
constexpr int kSWD = kTX + 2 * kHX;      // 72
// Tile height of the dry sweep.  A taller tile amortises the halo rows the sweeps recompute and the
// TMA fill (staged / computed cells: 1.41 at 16 rows, 1.29 at 28) until shared memory costs a
// resident CTA: measured 0.620 ms (16) / 0.585 (20) / 0.560 (24) / 0.547 (28) / 0.576 (32, 3 CTAs) /
// 0.658 (48, 2 CTAs) per launch at 16384 x 4096 (profiles/r2_dry_variants.md).  28 rows = 55.3 KB,
// the tallest tile that still fits 4 CTAs per SM.
#ifndef WSB_DRY_TY
#define WSB_DRY_TY 28
#endif
constexpr int kTYD = WSB_DRY_TY;
constexpr int kSHD = kTYD + 2 * kHD;     // 32
constexpr int kND = kSWD * kSHD;         // 2304
constexpr int kPSD = plane_stride<kND>();  // 2304 floats
constexpr int kNT0 = kTX * kTY;          // cells of a tile without halo (own-cell operand tiles)
// float planes: VX, VY, P, T raw, wall, T post-pressure; mbarrier; CTA max |v|
constexpr size_t kSmemDry = (size_t)kPSD * 4 * 6 + 16;
#ifndef WSB_DRY_CTAS
#define WSB_DRY_CTAS 4   // resident CTAs per SM the register allocation is bounded for
#endif

// Advection of the base field on one tile (the third stage of k_fused_dry).  WALLS = false: no wall cell anywhere
// in the staged region — the own cell's wall test and the wall-aware bilerp weights fall away at compile time, and T
// after the pressure pass IS the staged T (plane 3 instead of plane 5).
template <bool WALLS>
__device__ __forceinline__ float dry_advect_tile(const GlobalCtx& glob, const DevParams& d, int applyPressure, const float* sVX,
                                                 int X0, int Y0, Planes4 baseOut) {
  constexpr int SW = kSWD;
  constexpr bool walls = WALLS;
  const Geom& g = glob.g;
  const int tid = threadIdx.x;
  const float* sVY = sVX + kPSD;
  const float* sP = sVY + kPSD;
  const int* sWl = reinterpret_cast<const int*>(sVX + 4 * kPSD);
  const float* sT2 = sVX + 5 * kPSD;
  (void)sP; (void)sWl; (void)sT2;
#if WSB_OPT_NEAR
  constexpr int tPlane = WALLS ? 5 * kPSD : 3 * kPSD;   // T after the pressure pass, as a displacement from the VX plane
#else
  const float* sTpost = WALLS ? sT2 : sVX + 3 * kPSD;
#endif
  float vm = 0.0f;
  const int tx = tid % kTX, ty0 = tid / kTX;
  const int x = X0 + kHX + tx;
  if (x < g.cx1) {
    const int gx = global_x(g, x);
    const float gxf = (float)gx, gxm1 = gxf - 1.0f;
    const float fragCoordX = gxf + 0.5f;
    const int lxBase = tx + kHX;
#pragma unroll 1
    for (int ty = ty0; ty < kTYD; ty += kRowStep) {
      const int y = Y0 + kHD + ty;
      if (y >= g.H) break;
      const int c = (ty + kHD) * SW + lxBase;
      const float gyf = (float)y;
      const float fragCoordY = gyf + 0.5f;
      int w0 = 0x0100;  // an air texel (DISTANCE 1)
      if (walls) w0 = sWl[c];
      float4 base;
      if (!wl_is_wall(w0)) {
        const float vx00 = sVX[c], vy00 = sVY[c];
        vm = fmaxf(vm, fmaxf(fabsf(vx00), fabsf(vy00)));
        const AdvVel a = adv_velocities(vx00, vy00, sVX[c - 1], sVY[c - SW], sVY[c + 1], sVX[c + SW], sVX[c + SW - 1],
                                        sVY[c - SW + 1]);
#if WSB_OPT_NEAR
        if (adv_vmax(a) < g.nearV) {  // taps are offsets from the own cell; no bounds test needed (halo 2)
          const float gym1 = gyf - 1.0f;
          const NearTap n1 = near_tap(fragCoordX - a.Vxx, fragCoordY - a.Vxy, gxf, gxm1, gyf, gym1);
          const NearTap n2 = near_tap(fragCoordX - a.Vyx, fragCoordY - a.Vyy, gxf, gxm1, gyf, gym1);
          const NearTap n3 = near_tap(fragCoordX - a.Px, fragCoordY - a.Py, gxf, gxm1, gyf, gym1);
          // every plane is addressed off the VX plane's pointer: the plane displacement is an immediate
          const float* pc = sVX + c;
          const float* q1 = near_ptr(n1, pc, pc - SW);
          const float* q2 = near_ptr(n2, pc, pc - SW) + kPSD;
          const float* q3 = near_ptr(n3, pc, pc - SW);
          base.x = mix2d(q1[0], q1[1], q1[SW], q1[SW + 1], n1.fx, n1.fx, n1.fy);
          base.y = mix2d(q2[0], q2[1], q2[SW], q2[SW + 1], n2.fx, n2.fx, n2.fy);
          const WallMix m = walls ? tile_wall_mix<SW>(reinterpret_cast<const int*>(q3 + 4 * kPSD), 0, n3.fx, n3.fy)
                                  : WallMix{n3.fx, n3.fx, n3.fy};
          const float* qP = q3 + 2 * kPSD;
          const float* qT = q3 + tPlane;
          base.z = mix2d(qP[0], qP[1], qP[SW], qP[SW + 1], m.ab, m.cd, m.abcd);
          base.w = mix2d(qT[0], qT[1], qT[SW], qT[SW + 1], m.ab, m.cd, m.abcd);
        } else {  // |v| >= 0.9 cells / iteration (never seen in the shipped saves): exact global-memory path
          float vmSlow = 0.0f;  // a local of the cold branch: keeps vm itself in a register
          base = dry_advect_slow(&glob, &d, applyPressure, x, y, &vmSlow);
          vm = fmaxf(vm, vmSlow);
        }
#else
        const BilerpSetup b1 = bilerp_setup(fragCoordX - a.Vxx, fragCoordY - a.Vxy);
        const BilerpSetup b2 = bilerp_setup(fragCoordX - a.Vyx, fragCoordY - a.Vyy);
        const BilerpSetup b3 = bilerp_setup(fragCoordX - a.Px, fragCoordY - a.Py);
        int lx1, ly1, lx2, ly2, lx3, ly3;
        const int l1 = tile_index<SW>(b1, lxBase, gx, Y0, lx1, ly1);
        const int l2 = tile_index<SW>(b2, lxBase, gx, Y0, lx2, ly2);
        const int l3 = tile_index<SW>(b3, lxBase, gx, Y0, lx3, ly3);
        // post-velocity / post-pressure planes are valid on [1, SW-2] x [1, SH-2]
        if (tile_ok<kSWD, kSHD>(lx1, ly1, 1, 1) && tile_ok<kSWD, kSHD>(lx2, ly2, 1, 1) && tile_ok<kSWD, kSHD>(lx3, ly3, 1, 1)) {
          base.x = mix2d(sVX[l1], sVX[l1 + 1], sVX[l1 + SW], sVX[l1 + SW + 1], b1.fx, b1.fx, b1.fy);
          base.y = mix2d(sVY[l2], sVY[l2 + 1], sVY[l2 + SW], sVY[l2 + SW + 1], b2.fx, b2.fx, b2.fy);
          const WallMix m = walls ? tile_wall_mix<SW>(sWl, l3, b3.fx, b3.fy) : WallMix{b3.fx, b3.fx, b3.fy};
          base.z = mix2d(sP[l3], sP[l3 + 1], sP[l3 + SW], sP[l3 + SW + 1], m.ab, m.cd, m.abcd);
          base.w = mix2d(sTpost[l3], sTpost[l3 + 1], sTpost[l3 + SW], sTpost[l3 + SW + 1], m.ab, m.cd, m.abcd);
        } else {
          float vmSlow = 0.0f;
          base = dry_advect_slow(&glob, &d, applyPressure, x, y, &vmSlow);
          vm = fmaxf(vm, vmSlow);
        }
#endif
      } else {  // wall: pass-through of the post-velocity cell (advectionShader.frag:189-197)
        base = make_float4(sVX[c], sVY[c], sP[c], ((w0 & 0xff) == WALLTYPE_LAND) ? 1000.0f : sT2[c]);
      }
      baseOut.st((size_t)y * g.pitch + x, base);
    }
  }
  if (x < g.ox0 || x >= g.ox1) vm = 0.0f;  // ghost columns hold the neighbour's cells (and edge garbage)
  return vm;
}

// glob: base = base_1 (advection output, pressure pending), wall = wall_1.
// maps: TMA descriptors of glob.base.c[0..3] and glob.wall with a kSWD x kSHD box.
// (Tried and dropped, profiles/r2_dry_variants.md: results leaving through shared-memory tiles and
// TMA box stores — the elected thread's wait for the store to drain keeps the CTA's slot busy, 4 %
// slower than plain coalesced stores; a persistent grid with double-buffered TMA prefetch of the
// next tile — hides the load latency completely but fits only 3 CTAs per SM, 7 % slower: the
// kernel is bound by shared-memory wavefronts and issue slots, not by exposed HBM latency.)
__global__ void __launch_bounds__(kNT, WSB_DRY_CTAS) k_fused_dry(const __grid_constant__ GlobalCtx glob,
                                                      const __grid_constant__ DevParams d,
                                                      const __grid_constant__ TileMaps<5> maps, int useTma, int applyPressure,
                                                      const unsigned char* __restrict__ tileWalls, Planes4 baseOut,
                                                      unsigned* __restrict__ maxv) {
  WSB_DYN_SMEM(smem_raw);
  float* sVX = reinterpret_cast<float*>(smem_raw);
  float* sVY = sVX + kPSD;
  float* sP = sVY + kPSD;
  float* sT = sP + kPSD;    // raw T
  int* sWl = reinterpret_cast<int*>(sT + kPSD);
  float* sT2 = reinterpret_cast<float*>(sWl + kPSD);   // T after the pressure pass
  unsigned long long* mbar = reinterpret_cast<unsigned long long*>(sT2 + kPSD);
  unsigned* sMax = reinterpret_cast<unsigned*>(mbar + 1);  // CTA maximum of |v| (report_vmax_cta)

  const Geom& g = glob.g;
  const int tid = threadIdx.x;
  const int X0 = tile_col0(g, blockIdx.x, kTX) - kHX, Y0 = blockIdx.y * kTYD - kHD;

  // tileWalls[tile] == 0: no wall cell anywhere in this tile's staged region (k_wall_tilemap; the dry sweep never
  // changes the wall texture, so the map stays valid between wall-changing calls).  Such tiles — all of the free
  // atmosphere — do not stage the wall plane at all (16 of 20 staged bytes per cell, 32 of 36 B / cell of HBM traffic),
  // their sweeps skip the wall tests and the land-wall rule of the pressure pass (T' == T).
  const bool walls = tileWalls == nullptr || tileWalls[blockIdx.y * gridDim.x + blockIdx.x] != 0;
  if (tid == 0) *sMax = 0u;
  if (tile_tma_ok<kSWD, kSHD>(g, useTma, X0, Y0)) {
    if (tid == 0) mbar_init(mbar, 1);
    __syncthreads();
    if (tid == 0) {
      mbar_expect_tx(mbar, (walls ? 5u : 4u) * kND * 4u);
      tma_load_box(sVX, &maps.m[0], X0, Y0, mbar);
      tma_load_box(sVY, &maps.m[1], X0, Y0, mbar);
      tma_load_box(sP, &maps.m[2], X0, Y0, mbar);
      tma_load_box(sT, &maps.m[3], X0, Y0, mbar);
      if (walls) tma_load_box(sWl, &maps.m[4], X0, Y0, mbar);
    }
    mbar_wait(mbar, 0);
  } else {
    stage_tile<kSWD, kSHD, 6>(
        g, X0, Y0,
        [&](int ci, int) { return BaseWallRegs{glob.base.c[0][ci], glob.base.c[1][ci], glob.base.c[2][ci], glob.base.c[3][ci], walls ? glob.wall[ci] : 0x0100}; },
        [&](int s, const BaseWallRegs& r) {
          sVX[s] = r.vx; sVY[s] = r.vy; sP[s] = r.p; sT[s] = r.t;
          sWl[s] = r.w;
        });
    __syncthreads();
  }

  if (walls) {
    sweep_pressure<kSWD, kND, true>(sVX, sVY, sP, sT, sT2, sWl, applyPressure);
    __syncthreads();
    sweep_velocity<kSWD, kND, true>(d, sVX, sVY, sP, sWl);
  } else {
    sweep_pressure<kSWD, kND, false>(sVX, sVY, sP, sT, sT2, sWl, applyPressure);
    __syncthreads();
    sweep_velocity<kSWD, kND, false>(d, sVX, sVY, sP, sWl);
  }
  __syncthreads();
  const float vm = walls ? dry_advect_tile<true>(glob, d, applyPressure, sVX, X0, Y0, baseOut)
                         : dry_advect_tile<false>(glob, d, applyPressure, sVX, X0, Y0, baseOut);
  report_vmax_cta(vm, maxv, sMax);
}

// One byte per dry-sweep tile: does its staged region (tile + halo, periodic in x and y like the textures) hold a wall
// cell?  Recomputed (4 B / cell, once) when the wall texture may have changed: upload, any full-physics iteration.
__global__ void k_wall_tilemap(GlobalCtx glob, unsigned char* __restrict__ tileWalls) {
  const Geom& g = glob.g;
  const int X0 = blockIdx.x * kTX - kHX, Y0 = blockIdx.y * kTYD - kHD;
  bool any = false;
  for (int i = threadIdx.x; i < kND; i += blockDim.x) {
    const int j = i / kSWD, ii = i - j * kSWD;
    const int x = wrap_x(g, X0 + ii), y = wrap_y(Y0 + j, g.H);
    any |= wl_is_wall(glob.wall[(size_t)y * g.pitch + x]);
  }
  const int r = __syncthreads_or(any ? 1 : 0);
  if (threadIdx.x == 0) tileWalls[blockIdx.y * gridDim.x + blockIdx.x] = r ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------
// k_fused_pvb — pressure(prev) -> velocity -> curl -> vorticity -> boundary
// ---------------------------------------------------------------------------------------------
constexpr int kH1 = 3;                      // halo of the pressure->boundary chain
constexpr int kSW1 = kTX + 2 * kHX;         // 72
constexpr int kSH1 = kTY + 2 * kH1;         // 22
constexpr int kN1 = kSW1 * kSH1;            // 1584 cells per staged tile
constexpr int kPS1 = plane_stride<kN1>();   // 1600 floats
constexpr int kNVF = kSW1 * (kTY + 1);      // vortForce is only needed on the tile rows and the row below them
constexpr int kPSV = plane_stride<kNVF>();  // 1248 floats
// float planes: VX | VY | P | T raw, later curl | wall | T post-pressure | vortForce x | vortForce y |
//               own-cell tiles: water total, cloud, precip, smoke, light sun, light net heating ; mbarrier
constexpr size_t kSmem1 = (size_t)kPS1 * 4 * 6 + (size_t)kPSV * 4 * 2 + (size_t)kNT0 * 4 * 6 + 16;

// boundary_cell context positioned at shared-memory cell c: base / wall / vortForce from the tile,
// own-cell water / light / feedback / deposition from registers (prefetched), the sparse
// neighbour fetches of water and light (surface cells only) from HBM.
struct PvbAt {
  const float *sVX, *sVY, *sP, *sT2, *sVFX, *sVFY;
  const int* sWl;
  int c;
  const GlobalCtx& glob;
  int x, y;
  const float* sOwn;  // own-cell tiles [6][kNT0]: water x4, light sun, light net heating
  int t;              // index of the cell in the own-cell tiles
  float4 fb0;
  float2 dep0;
  __device__ __forceinline__ int si(int dx, int dy) const { return c + dy * kSW1 + dx; }
  __device__ __forceinline__ float4 base4(int dx, int dy) const {
    const int s = si(dx, dy);
    return make_float4(sVX[s], sVY[s], sP[s], sT2[s]);
  }
  __device__ __forceinline__ float bx(int dx, int dy) const { return sVX[si(dx, dy)]; }
  __device__ __forceinline__ float by(int dx, int dy) const { return sVY[si(dx, dy)]; }
  __device__ __forceinline__ float bt(int dx, int dy) const { return sT2[si(dx, dy)]; }
  __device__ __forceinline__ char4 wall4(int dx, int dy) const { return as_char4(sWl[si(dx, dy)]); }
  // vortForce planes start at staged row kH1-1
  __device__ __forceinline__ float2 vort(int dx, int dy) const {
    const int s = si(dx, dy) - (kH1 - 1) * kSW1;
    return make_float2(sVFX[s], sVFY[s]);
  }
  __device__ __forceinline__ float4 water4(int dx, int dy) const {
    if (dx == 0 && dy == 0) return make_float4(sOwn[t], sOwn[kNT0 + t], sOwn[2 * kNT0 + t], sOwn[3 * kNT0 + t]);
    return glob.water.ld(glob.idx_near(x + dx, y + dy));
  }
  __device__ __forceinline__ float4 light4(int dx, int dy) const {
    if (dx == 0 && dy == 0) return make_float4(sOwn[4 * kNT0 + t], sOwn[5 * kNT0 + t], 0.0f, 0.0f);  // SUNLIGHT, NET_HEATING only
    return glob.light4(x + dx, y + dy);
  }
  __device__ __forceinline__ float4 fb4() const { return fb0; }
  __device__ __forceinline__ float2 dep2() const { return dep0; }
};

// glob: base = base_1, wall = wall_1, water = water_1, light = light_0 (glob.fb / glob.dep unused).
// maps: TMA descriptors — [0..4] base.c[0..3], wall with a kSW1 x kSH1 box; [5..10] water.c[0..3],
// light.c[0], light.c[1] with a kTX x kTY box (the cell's own operands of the boundary pass).
// useFb: feedback / deposition hold data from the last particle pass; the kernel consumes them and
// writes the zeros of the reference's gl.clear (app.js:5933-5934) back to the cells that were hit.
__global__ void __launch_bounds__(kNT, 3) k_fused_pvb(const __grid_constant__ GlobalCtx glob,
                                                      const __grid_constant__ DevParams d,
                                                      const __grid_constant__ TileMaps<11> maps, int useTma,
                                                      const float* __restrict__ initial_T, int applyPressure, int useFb,
                                                      float4* fb, float2* dep, SpriteGrid sg, Planes4 baseOut, Planes4 waterOut,
                                                      int* __restrict__ wallOut) {
  WSB_DYN_SMEM(smem_raw);
  float* sVX = reinterpret_cast<float*>(smem_raw);
  float* sVY = sVX + kPS1;
  float* sP = sVY + kPS1;
  float* sT = sP + kPS1;     // raw T; dead after the pressure sweep ...
  float* sCurl = sT;         // ... then curl
  int* sWl = reinterpret_cast<int*>(sT + kPS1);
  float* sT2 = reinterpret_cast<float*>(sWl + kPS1);     // T after the pressure pass
  float* sVFX = sT2 + kPS1;  // staged rows kH1-1 .. kH1+kTY-1
  float* sVFY = sVFX + kPSV;
  float* sOwn = sVFY + kPSV;  // [6][kNT0]
  unsigned long long* mbar = reinterpret_cast<unsigned long long*>(sOwn + 6 * kNT0);
  constexpr int SW = kSW1;

  const Geom& g = glob.g;
  const int tid = threadIdx.x;
  const int X0 = tile_col0(g, blockIdx.x, kTX) - kHX, Y0 = blockIdx.y * kTY - kH1;
  const int tx = tid % kTX, ty0 = tid / kTX;
  const int x = X0 + kHX + tx;
  const bool colOk = x < g.cx1;
  // feedback / deposition of the last particle pass: only tiles a sprite touched hold anything
  const int dirtyIdx = blockIdx.y * sg.tilesX + (X0 + kHX) / kTX;
  const int dirtyBits = useFb ? sg.dirty[dirtyIdx] : 0;
  const bool tileFb = (dirtyBits & kDirtyFb) != 0, tileDep = (dirtyBits & kDirtyDep) != 0;

  if (tile_tma_ok<kSW1, kSH1>(g, useTma, X0, Y0)) {
    if (tid == 0) mbar_init(mbar, 1);
    __syncthreads();
    if (tid == 0) {
      mbar_expect_tx(mbar, 5u * kN1 * 4u + 6u * kNT0 * 4u);
      tma_load_box(sVX, &maps.m[0], X0, Y0, mbar);
      tma_load_box(sVY, &maps.m[1], X0, Y0, mbar);
      tma_load_box(sP, &maps.m[2], X0, Y0, mbar);
      tma_load_box(sT, &maps.m[3], X0, Y0, mbar);
      tma_load_box(sWl, &maps.m[4], X0, Y0, mbar);
#pragma unroll
      for (int k = 0; k < 6; k++) tma_load_box(sOwn + k * kNT0, &maps.m[5 + k], X0 + kHX, Y0 + kH1, mbar);
    }
    mbar_wait(mbar, 0);
  } else {
    stage_tile<kSW1, kSH1, 4>(
        g, X0, Y0,
        [&](int ci, int) { return BaseWallRegs{glob.base.c[0][ci], glob.base.c[1][ci], glob.base.c[2][ci], glob.base.c[3][ci], glob.wall[ci]}; },
        [&](int s, const BaseWallRegs& r) {
          sVX[s] = r.vx; sVY[s] = r.vy; sP[s] = r.p; sT[s] = r.t;
          sWl[s] = r.w;
        });
    for (int ty = ty0; ty < kTY; ty += kRowStep) {  // own-cell tiles
      const int y = Y0 + kH1 + ty, t = ty * kTX + tx;
      const bool ok = colOk && y < g.H;
      const int ci = ok ? y * g.pitch + x : 0;
#pragma unroll
      for (int k = 0; k < 4; k++) sOwn[k * kNT0 + t] = ok ? glob.water.c[k][ci] : 0.0f;
      sOwn[4 * kNT0 + t] = ok ? glob.light.c[0][ci] : 0.0f;
      sOwn[5 * kNT0 + t] = ok ? glob.light.c[1][ci] : 0.0f;
    }
    __syncthreads();
  }

  // S1 pressure (previous iteration), S2 velocity
  sweep_pressure<kSW1, kN1>(sVX, sVY, sP, sT, sT2, sWl, applyPressure);
  __syncthreads();
  sweep_velocity<kSW1, kN1>(d, sVX, sVY, sP, sWl);
  __syncthreads();
  // S3: curl; valid for 1 <= i <= SW-3, 1 <= j <= SH-3 (raw T plane is dead: reuse it)
  for (int s = SW + 2 * tid; s < kN1 - 2 * SW; s += 2 * kNT) {
    const float2 vx = ld2(sVX + s), vy = ld2(sVY + s), vxu = ld2(sVX + s + SW);
    const float vyr = sVY[s + 2];
    st2(sCurl + s, make_float2(curl_cell(vx.x, vy.x, vxu.x, vy.y), curl_cell(vx.y, vy.y, vxu.y, vyr)));
  }
  __syncthreads();
  // S4: vorticity force on the rows the boundary pass reads (tile rows and the row below them);
  // valid for 2 <= i <= SW-4
  for (int s = (kH1 - 1) * SW + 2 * tid; s < (kH1 + kTY) * SW; s += 2 * kNT) {
    const float2 c = ld2(sCurl + s), cd = ld2(sCurl + s - SW), cu = ld2(sCurl + s + SW);
    const float cl = sCurl[s - 1], cr = sCurl[s + 2];
    const float2 va = vorticity_cell(c.x, cl, cd.x, c.y, cu.x), vb = vorticity_cell(c.y, c.x, cd.y, cr, cu.y);
    st2(sVFX + s - (kH1 - 1) * SW, make_float2(va.x, vb.x));
    st2(sVFY + s - (kH1 - 1) * SW, make_float2(va.y, vb.y));
  }
  __syncthreads();

  // S5: boundary pass on the TX x TY interior
#pragma unroll 1
  for (int ty = ty0; ty < kTY; ty += kRowStep) {
    const int y = Y0 + kH1 + ty;
    if (colOk && y < g.H) {
      const size_t ci = (size_t)y * g.pitch + x;
      PvbAt c{sVX, sVY, sP, sT2, sVFX, sVFY, sWl, (ty + kH1) * SW + tx + kHX, glob, x, y, sOwn, ty * kTX + tx,
              make_float4(0.f, 0.f, 0.f, 0.f), make_float2(0.f, 0.f)};
      if (tileFb) c.fb0 = fb[ci];
      if (tileDep) c.dep0 = dep[ci];
      float4 b, w;
      char4 wl;
      boundary_cell(c, g, d, initial_T, x, y, b, w, wl);
      baseOut.st(ci, b);
      waterOut.st(ci, w);
      wallOut[ci] = as_int(wl);
      // the reference's gl.clear (app.js:5933-5934): only texels that were hit cost a write
      if (tileFb && (c.fb0.x != 0.0f || c.fb0.y != 0.0f || c.fb0.z != 0.0f || c.fb0.w != 0.0f)) fb[ci] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (tileDep && (c.dep0.x != 0.0f || c.dep0.y != 0.0f)) dep[ci] = make_float2(0.f, 0.f);
    }
  }
  if (tid == 0 && dirtyBits) sg.dirty[dirtyIdx] = 0;  // consumed (no other CTA of this launch looks at this tile's word)
}

// ---------------------------------------------------------------------------------------------
// k_fused_adv — advection -> lighting
// ---------------------------------------------------------------------------------------------
constexpr int kH2 = 2;                  // halo: covers every back-trace with |v| < 1 and the sun-ray fetch
constexpr int kSW2 = kTX + 2 * kHX;     // 72
constexpr int kSH2 = kTY + 2 * kH2;     // 20
constexpr int kN2 = kSW2 * kSH2;        // 1440
constexpr int kPS2 = plane_stride<kN2>();  // 1440 floats
// float planes: VX VY P T | water total, cloud, precip, smoke | wall | light: sun, IR down, IR up ; mbarrier
constexpr size_t kSmem2 = (size_t)kPS2 * 4 * 12 + 16;

// light fetches of lighting_cell from the staged light planes (rows are staged with the
// CLAMP_TO_EDGE rule, so the clamped row lighting_cell passes maps straight to a tile row)
struct TileLightCtx {
  const float *sLS, *sLD, *sLU;
  int origin;  // -(Y0 * kSW2 + X0): tile index of array cell (0, 0)
  __device__ __forceinline__ int si(int x, int y) const { return y * kSW2 + x + origin; }
  __device__ __forceinline__ float lightS(int x, int y) const { return sLS[si(x, y)]; }
  __device__ __forceinline__ float lightIRdown(int x, int y) const { return sLD[si(x, y)]; }
  __device__ __forceinline__ float lightIRup(int x, int y) const { return sLU[si(x, y)]; }
};

struct AdvRegs { float4 b, w; float ls, ld, lu; int wl; };

// glob: base_0, water_0, wall_0 (boundary output), light = light_src.
// maps: TMA descriptors (kSW2 x kSH2 box) of base.c[0..3], water.c[0..3], wall, light.c[0], light.c[2], light.c[3].
__global__ void __launch_bounds__(kNT, 3) k_fused_adv(const __grid_constant__ GlobalCtx glob,
                                                      const __grid_constant__ DevParams d,
                                                      const __grid_constant__ TileMaps<12> maps, int useTma,
                                                      const float* __restrict__ initial_T, const float* __restrict__ sndT,
                                                      const float* __restrict__ sndW, const float* __restrict__ sndV,
                                                      Planes4 baseOut, Planes4 waterOut, int* __restrict__ wallOut,
                                                      Planes4 lightOut, unsigned* __restrict__ maxv) {
  WSB_DYN_SMEM(smem_raw);
  float* sVX = reinterpret_cast<float*>(smem_raw);
  float* sVY = sVX + kPS2;
  float* sP = sVY + kPS2;
  float* sT = sP + kPS2;
  float* sW0 = sT + kPS2;
  float* sW1 = sW0 + kPS2;
  float* sW2 = sW1 + kPS2;
  float* sW3 = sW2 + kPS2;
  int* sWl = reinterpret_cast<int*>(sW3 + kPS2);
  float* sLS = reinterpret_cast<float*>(sWl + kPS2);
  float* sLD = sLS + kPS2;
  float* sLU = sLD + kPS2;
  unsigned long long* mbar = reinterpret_cast<unsigned long long*>(sLU + kPS2);
  constexpr int SW = kSW2;

  const Geom& g = glob.g;
  const int tid = threadIdx.x;
  const int X0 = tile_col0(g, blockIdx.x, kTX) - kHX, Y0 = blockIdx.y * kTY - kH2;

  // (interior rows never touch the CLAMP_TO_EDGE rule of the light texture, so one box shape serves all planes)
  unsigned* sMax = reinterpret_cast<unsigned*>(mbar + 1);  // CTA maximum of |v| (report_vmax_cta)
  if (tid == 0) *sMax = 0u;
  if (tile_tma_ok<kSW2, kSH2>(g, useTma, X0, Y0)) {
    if (tid == 0) mbar_init(mbar, 1);
    __syncthreads();
    if (tid == 0) {
      mbar_expect_tx(mbar, 12u * kN2 * 4u);
      float* dst = sVX;
#pragma unroll
      for (int k = 0; k < 12; k++) tma_load_box(dst + k * kPS2, &maps.m[k], X0, Y0, mbar);
    }
    mbar_wait(mbar, 0);
  } else {
    stage_tile<kSW2, kSH2, 2>(
        g, X0, Y0,
        [&](int ci, int cil) {
          return AdvRegs{glob.base.ld(ci), glob.water.ld(ci), glob.light.c[0][cil], glob.light.c[2][cil], glob.light.c[3][cil], glob.wall[ci]};
        },
        [&](int s, const AdvRegs& r) {
          sVX[s] = r.b.x; sVY[s] = r.b.y; sP[s] = r.b.z; sT[s] = r.b.w;
          sW0[s] = r.w.x; sW1[s] = r.w.y; sW2[s] = r.w.z; sW3[s] = r.w.w;
          sWl[s] = r.wl;
          sLS[s] = r.ls; sLD[s] = r.ld; sLU[s] = r.lu;
        });
    __syncthreads();
  }

  const TileLightCtx lc{sLS, sLD, sLU, -(Y0 * kSW2 + X0)};
  const int tx = tid % kTX, ty0 = tid / kTX;
  const int x = X0 + kHX + tx;
  float vm = 0.0f;
  if (x < g.cx1) {
    const int gx = global_x(g, x);
    const float gxf = (float)gx, gxm1 = gxf - 1.0f;
    const float fragCoordX = gxf + 0.5f;
    const float texCoordX = fragCoordX * g.texelX;
    const int lxBase = tx + kHX;
#pragma unroll 1  // (two cells in flight per thread, unroll 2: adv +8 % slower at 80 registers, profiles/r3_logs/c28_variants.log)
    for (int ty = ty0; ty < kTY; ty += kRowStep) {
      const int y = Y0 + kH2 + ty;
      if (y >= g.H) break;
      const int c = (ty + kH2) * SW + lxBase;
      const float fragCoordY = (float)y + 0.5f;
      const float texCoordY = fragCoordY * g.texelY;
      const char4 w0 = as_char4(sWl[c]);
      int wType = w0.x, wDist = w0.y, wVert = w0.z, wVeg = w0.w;
      const int aboveDist = wl_is_wall(sWl[c + SW]) ? 0 : 1;  // only ever tested against 0
      float4 base, water;
      char4 wl;
      bool done = false;
      if (wDist != 0) {  // air: semi-Lagrangian gathers from the tile
        const float vx00 = sVX[c], vy00 = sVY[c];
        const AdvVel a = adv_velocities(vx00, vy00, sVX[c - 1], sVY[c - SW], sVY[c + 1], sVX[c + SW], sVX[c + SW - 1],
                                        sVY[c - SW + 1]);
#if WSB_OPT_NEAR
        if (adv_vmax(a) < g.nearV) {  // taps are offsets from the own cell; halo 2 covers them without a bounds test
          const float gyf = (float)y, gym1 = gyf - 1.0f;
          const float posPx = fragCoordX - a.Px, posPy = fragCoordY - a.Py;
          const NearTap n1 = near_tap(fragCoordX - a.Vxx, fragCoordY - a.Vxy, gxf, gxm1, gyf, gym1);
          const NearTap n2 = near_tap(fragCoordX - a.Vyx, fragCoordY - a.Vyy, gxf, gxm1, gyf, gym1);
          const NearTap n3 = near_tap(posPx, posPy, gxf, gxm1, gyf, gym1);
          const NearTap n4 = near_tap(posPx + 0.0f, posPy + 0.05f, gxf, gxm1, gyf, gym1);  // :137 precipitation falls
          // every plane is addressed off the VX plane's pointer: the plane displacement is an immediate
          const float* pc = sVX + c;
          const float* q1 = near_ptr(n1, pc, pc - SW);
          const float* q2 = near_ptr(n2, pc, pc - SW) + kPS2;
          const float* q3 = near_ptr(n3, pc, pc - SW);
          const float* q4 = near_ptr(n4, pc, pc - SW);
          vm = fmaxf(vm, fmaxf(fabsf(vx00), fabsf(vy00)));
          base.x = mix2d(q1[0], q1[1], q1[SW], q1[SW + 1], n1.fx, n1.fx, n1.fy);
          base.y = mix2d(q2[0], q2[1], q2[SW], q2[SW + 1], n2.fx, n2.fx, n2.fy);
          {
            const WallMix m = tile_wall_mix<SW>(reinterpret_cast<const int*>(q3 + 8 * kPS2), 0, n3.fx, n3.fy);
            const float *qP = q3 + 2 * kPS2, *qT = q3 + 3 * kPS2, *qW0 = q3 + 4 * kPS2, *qW1 = q3 + 5 * kPS2, *qW3 = q3 + 7 * kPS2;
            base.z = mix2d(qP[0], qP[1], qP[SW], qP[SW + 1], m.ab, m.cd, m.abcd);
            base.w = mix2d(qT[0], qT[1], qT[SW], qT[SW + 1], m.ab, m.cd, m.abcd);
            water.x = mix2d(qW0[0], qW0[1], qW0[SW], qW0[SW + 1], m.ab, m.cd, m.abcd);
            water.y = mix2d(qW1[0], qW1[1], qW1[SW], qW1[SW + 1], m.ab, m.cd, m.abcd);
            water.w = mix2d(qW3[0], qW3[1], qW3[SW], qW3[SW + 1], m.ab, m.cd, m.abcd);
          }
          {
            const WallMix m = tile_wall_mix<SW>(reinterpret_cast<const int*>(q4 + 8 * kPS2), 0, n4.fx, n4.fy);
            const float* qW2 = q4 + 6 * kPS2;
            water.z = mix2d(qW2[0], qW2[1], qW2[SW], qW2[SW + 1], m.ab, m.cd, m.abcd);
          }
          adv_air_thermo(g, d, sndT, sndW, sndV, texCoordY, base, water);
        } else {  // |v| >= 0.9 cells / iteration: exact global-memory path for the whole cell
#else
        const BilerpSetup b1 = bilerp_setup(fragCoordX - a.Vxx, fragCoordY - a.Vxy);
        const BilerpSetup b2 = bilerp_setup(fragCoordX - a.Vyx, fragCoordY - a.Vyy);
        const float posPx = fragCoordX - a.Px, posPy = fragCoordY - a.Py;
        const BilerpSetup b3 = bilerp_setup(posPx, posPy);
        const BilerpSetup b4 = bilerp_setup(posPx + 0.0f, posPy + 0.05f);
        int lx1, ly1, lx2, ly2, lx3, ly3, lx4, ly4;
        const int l1 = tile_index<SW>(b1, lxBase, gx, Y0, lx1, ly1);
        const int l2 = tile_index<SW>(b2, lxBase, gx, Y0, lx2, ly2);
        const int l3 = tile_index<SW>(b3, lxBase, gx, Y0, lx3, ly3);
        const int l4 = tile_index<SW>(b4, lxBase, gx, Y0, lx4, ly4);
        if (tile_ok<kSW2, kSH2>(lx1, ly1, 0, 0) && tile_ok<kSW2, kSH2>(lx2, ly2, 0, 0) && tile_ok<kSW2, kSH2>(lx3, ly3, 0, 0) &&
            tile_ok<kSW2, kSH2>(lx4, ly4, 0, 0)) {
          vm = fmaxf(vm, fmaxf(fabsf(vx00), fabsf(vy00)));
          base.x = mix2d(sVX[l1], sVX[l1 + 1], sVX[l1 + SW], sVX[l1 + SW + 1], b1.fx, b1.fx, b1.fy);
          base.y = mix2d(sVY[l2], sVY[l2 + 1], sVY[l2 + SW], sVY[l2 + SW + 1], b2.fx, b2.fx, b2.fy);
          {
            const WallMix m = tile_wall_mix<SW>(sWl, l3, b3.fx, b3.fy);
            base.z = mix2d(sP[l3], sP[l3 + 1], sP[l3 + SW], sP[l3 + SW + 1], m.ab, m.cd, m.abcd);
            base.w = mix2d(sT[l3], sT[l3 + 1], sT[l3 + SW], sT[l3 + SW + 1], m.ab, m.cd, m.abcd);
            water.x = mix2d(sW0[l3], sW0[l3 + 1], sW0[l3 + SW], sW0[l3 + SW + 1], m.ab, m.cd, m.abcd);
            water.y = mix2d(sW1[l3], sW1[l3 + 1], sW1[l3 + SW], sW1[l3 + SW + 1], m.ab, m.cd, m.abcd);
            water.w = mix2d(sW3[l3], sW3[l3 + 1], sW3[l3 + SW], sW3[l3 + SW + 1], m.ab, m.cd, m.abcd);
          }
          {
            const WallMix m = tile_wall_mix<SW>(sWl, l4, b4.fx, b4.fy);
            water.z = mix2d(sW2[l4], sW2[l4 + 1], sW2[l4 + SW], sW2[l4 + SW + 1], m.ab, m.cd, m.abcd);
          }
          adv_air_thermo(g, d, sndT, sndW, sndV, texCoordY, base, water);
        } else {  // a back-trace left the halo: exact global-memory path for the whole cell
#endif
          AdvSlowOut o;
          adv_cell_slow(&glob, &d, initial_T, sndT, sndW, sndV, x, y, &o);
          base = o.base;
          water = o.water;
          wl = o.wall;
          vm = fmaxf(vm, o.vm);
          done = true;
        }
      } else {  // wall: pass-through + surface processes
        base = make_float4(sVX[c], sVY[c], sP[c], sT[c]);
        water = make_float4(sW0[c], sW1[c], sW2[c], sW3[c]);
        adv_wall_cell<false>(d, texCoordY, wType, aboveDist, sT[c + SW], wVeg, base, water);
      }
      if (!done) {
        adv_user_input(g, d, initial_T, texCoordX, texCoordY, aboveDist, wType, wDist, wVert, wVeg, base, water);
        // Only the wall tools (userInputType >= 10: distance 255, vegetation + 1) can push a component out of the
        // RGBA8I range; every other frame the texel is re-assembled from in-range bytes without the saturation
        if (d.in.userInputType >= 10) wl = pack_wall(wType, wDist, wVert, wVeg);
        else wl = as_char4(pack_wall_in_range(wType, wDist, wVert, wVeg));
      }
      const size_t ci = (size_t)y * g.pitch + x;
      baseOut.st(ci, base);
      waterOut.st(ci, water);
      wallOut[ci] = as_int(wl);

      // lighting needs base_1's temperature of the cell BELOW a water-surface air cell
      // (lightingShader.frag:105), i.e. that cell's advection result
      float TBelow = 0.0f;
      // air (DISTANCE != 0) one above (VERT_DISTANCE == 1) a WATER surface: one masked compare on the packed texel
      const int wli = as_int(wl);
      if ((wli & 0x00ff00ff) == ((1 << 16) | WALLTYPE_WATER) && (wli & 0xff00) != 0 && fragCoordY < g.Hf - 1.0f) {
        const int cb = c - SW;
        const char4 wb = as_char4(sWl[cb]);
        if (wb.y == 0) {  // the usual case: a wall cell, whose advection is a local update
          int tB = wb.x, dB = wb.y, vB = wb.w;
          float4 bb = make_float4(sVX[cb], sVY[cb], sP[cb], sT[cb]), wtb = make_float4(sW0[cb], sW1[cb], sW2[cb], sW3[cb]);
          const float texCoordYb = ((float)wrap_y(y - 1, g.H) + 0.5f) * g.texelY;
          adv_wall_cell<false>(d, texCoordYb, tB, w0.y, sT[c], vB, bb, wtb);
          adv_user_input(g, d, initial_T, texCoordX, texCoordYb, w0.y, tB, dB, wb.z, vB, bb, wtb);
          TBelow = bb.w;
        } else {
          AdvSlowOut o;
          adv_cell_slow(&glob, &d, initial_T, sndT, sndW, sndV, x, wrap_y(y - 1, g.H), &o);
          TBelow = o.base.w;
        }
      }
      lightOut.st(ci, lighting_cell(lc, g, d, x, y, fragCoordX, base.w, water, wl, TBelow));
    }
  }
  if (x < g.ox0 || x >= g.ox1) vm = 0.0f;  // ghost columns hold the neighbour's cells (and edge garbage)
  report_vmax_cta(vm, maxv, sMax);
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
