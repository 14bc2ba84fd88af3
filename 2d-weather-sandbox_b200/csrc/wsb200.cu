// wsb200.cu — host side of libwsb200.so: the C ABI of include/wsb200.h over the sm_100a kernels.
//
// One wsb_sim = one process / one GPU.  The C ABI speaks exactly the reference's packed texture
// layouts (base / water / light RGBA32F, wall RGBA8I, feedback RGBA32F, deposition RG32F, droplets
// 5 x f32; row 0 = bottom row, x fastest).  In HBM every channel of base / water / light is its own
// float plane (wsb_ref_kernels.cuh) — upload and readback transpose on the device — so that the
// fused kernels can stage tiles with one TMA box per plane.  A multi-GPU run cuts the grid into x-strips,
// one per rank, each stored with kGhost ghost columns on both sides; the ring of ranks refreshes
// the ghost columns once per iteration with ncclSend/ncclRecv over NVLink (x is periodic, so rank
// 0 and rank N-1 are neighbours).
//
// Schedules
//   WSB_SCHEDULE_REFERENCE  one kernel per reference pass, in the order of app.js:5830-6005.
//                           Canonical state after an iteration: base_0 / water_0 / wall_0.
//   WSB_SCHEDULE_FUSED      k_fused_pvb + k_fused_adv (+ k_precipitation, k_latch) per iteration.
//                           Canonical state: base_1 / water_1 / wall_1 (the advection output) with
//                           the iteration's pressure pass still pending; it runs as the first
//                           stage of the next k_fused_pvb, or on demand inside wsb_read_rect.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <unistd.h>

#include <algorithm>
#include <new>
#include <vector>

#include "../../include/wsb200.h"
#include "wsb_fused_kernels.cuh"
#include "wsb_particles.cuh"
#include "wsb_ref_kernels.cuh"

using namespace wsb;

namespace {

constexpr int kGhost = 8;  // == strips.GHOST on the Python side

thread_local char g_err[512] = "";

int fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}

#define CK(call)                                                                       \
  do {                                                                                 \
    cudaError_t e_ = (call);                                                           \
    if (e_ != cudaSuccess) return fail("%s failed: %s", #call, cudaGetErrorString(e_)); \
  } while (0)

// ---------------------------------------------------------------------------------------------
// NCCL, resolved at run time (a single-GPU process never needs it; a multi-GPU host process has
// normally loaded torch's bundled libnccl.so.2 already and dlopen returns that same image).
// ---------------------------------------------------------------------------------------------
struct Id128 { char internal[WSB_COMM_ID_BYTES]; };
struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, /*ncclUniqueId by value*/ Id128, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
constexpr int kNcclChar = 0;  // ncclInt8 / ncclChar

int load_nccl() {
  if (g_nccl.handle) return 0;
  const char* names[] = {getenv("WSB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names) {
    if (!n || !*n) continue;
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) return fail("cannot load libnccl.so.2 (%s); set WSB_NCCL_LIB", dlerror());
#define SYM(field, name)                                                  \
  *(void**)(&g_nccl.field) = dlsym(h, name);                              \
  if (!g_nccl.field) return fail("libnccl: missing symbol %s", name);
  SYM(GetUniqueId, "ncclGetUniqueId")
  SYM(CommInitRank, "ncclCommInitRank")
  SYM(CommDestroy, "ncclCommDestroy")
  SYM(Send, "ncclSend")
  SYM(Recv, "ncclRecv")
  SYM(GroupStart, "ncclGroupStart")
  SYM(GroupEnd, "ncclGroupEnd")
  SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
  g_nccl.handle = h;
  return 0;
}
#define NCK(call)                                                                          \
  do {                                                                                     \
    int r_ = (call);                                                                       \
    if (r_ != 0) return fail("%s failed: %s", #call, g_nccl.GetErrorString(r_));           \
  } while (0)

// ---------------------------------------------------------------------------------------------
// halo pack / unpack: ghost columns are strided in the x-fastest layout; they travel as
// contiguous [side][row][kGhost] blocks.
// ---------------------------------------------------------------------------------------------
constexpr int kMaxHaloPlanes = 16;
struct HaloPlanes { int* p[kMaxHaloPlanes]; int n; };  // every exchanged plane has 4-byte elements

// staging block layout: [plane][row][kGhost]
__global__ void k_pack_halo(HaloPlanes hp, int pitch, int H, int lw, int* __restrict__ toLeft, int* __restrict__ toRight) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int per = H * kGhost;
  if (t >= per * hp.n) return;
  const int k = t / per, r = t - k * per;
  const int y = r / kGhost, i = r - y * kGhost;
  const int* f = hp.p[k];
  toLeft[t] = f[(size_t)y * pitch + kGhost + i];   // my leftmost owned columns
  toRight[t] = f[(size_t)y * pitch + lw + i];      // my rightmost owned columns
}
__global__ void k_unpack_halo(HaloPlanes hp, int pitch, int H, int lw, const int* __restrict__ fromLeft,
                              const int* __restrict__ fromRight) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int per = H * kGhost;
  if (t >= per * hp.n) return;
  const int k = t / per, r = t - k * per;
  const int y = r / kGhost, i = r - y * kGhost;
  int* f = hp.p[k];
  f[(size_t)y * pitch + i] = fromLeft[t];
  f[(size_t)y * pitch + kGhost + lw + i] = fromRight[t];
}

// ---------------------------------------------------------------------------------------------
// Peer-memory ghost exchange (default transport): every rank allocates the exchanged planes in ONE
// arena and hands its neighbours an IPC handle to it (wsb_peer_info / wsb_connect_peers).  After the
// edge tiles of an iteration's advection kernel, k_push_ghosts stores this rank's outermost owned
// columns STRAIGHT into the neighbours' ghost columns over NVLink — no staging buffers, no unpack —
// and raises a sequence flag in the neighbour's arena; k_wait_ghosts (one thread, ahead of the next
// boundary kernel's edge tiles) waits for both neighbours' flags.  Flags (one 128-byte line each):
//   kFlagDataL / kFlagDataR   written by the left / right neighbour: "your ghost zone holds exchange #seq"
//   kFlagFreeL / kFlagFreeR   written by the left / right neighbour: "I have finished reading the ghost
//                             columns you filled in exchange #seq-1 (you may overwrite them)" — a rank's
//                             ghost columns are plain columns of its ping-pong planes, not a double buffer
//   kFlagBlocks               last-block-done counter of k_push_ghosts;  kFlagErr  a spin ran out
// Every wait is bounded (kSpinLimitNs): a missing neighbour becomes an error from wsb_sync, not a hang.
// ---------------------------------------------------------------------------------------------
enum { kFlagDataL = 0, kFlagDataR, kFlagFreeL, kFlagFreeR, kFlagBlocks, kFlagErr, kNumFlags };
constexpr int kFlagStride = 32;  // unsigned words: one 128-byte line per flag
constexpr unsigned long long kSpinLimitNs = 20ull * 1000 * 1000 * 1000;  // default; WSB_SPIN_LIMIT_MS overrides (tests)
constexpr int kArenaSlots = 21;  // base_0 x4 | base_1 x4 | water_1 x4 | wall_1 | light_0 x4 | light_1 x4

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// spin until *flag >= seq (sequence numbers wrap: compare as a signed difference)
__device__ unsigned g_pollNs = 100;  // back-off between two polls (WSB_DBG_POLL_NS: experiments)
__device__ __forceinline__ void wait_flag(const unsigned* flag, unsigned seq, unsigned* err, unsigned long long limitNs) {
  const unsigned long long t0 = global_ns();
  const unsigned pollNs = g_pollNs;
  while ((int)(ld_acquire_sys(flag) - seq) < 0) {
    if (global_ns() - t0 > limitNs) { atomicExch(err, 1u); return; }
    __nanosleep(pollNs);
  }
}

// Landing zones (transport "peerc"): behind the flags every arena holds two COMPACT copies of a ghost zone —
// [side: 0 = left ghost, 1 = right ghost][plane k of the exchange][row][8 columns] — 32 bytes per (plane, row), 1.7 MB
// per side at H = 4096 instead of 13 x H rows scattered over every 2-MB page of the planes.  The neighbours store
// there and the owner copies the zone into its ghost columns locally (k_unpack_zone) once the data flags are up.
constexpr size_t kZoneOffset = (size_t)kNumFlags * kFlagStride * 4;  // behind the flag lines
__host__ __device__ __forceinline__ size_t zone_bytes(int H) { return (size_t)2 * kMaxHaloPlanes * H * kGhost * 4; }
__device__ __forceinline__ int* zone_row(unsigned char* arena, unsigned long long pb, int H, int side, int k, int y) {
  return reinterpret_cast<int*>(arena + pb * kArenaSlots + kZoneOffset) + (((size_t)side * kMaxHaloPlanes + k) * H + y) * kGhost;
}

struct PushArgs {
  int compact;                          // 1: store into the neighbours' landing zones instead of their ghost columns
  unsigned char *mine, *left, *right;   // arenas: this rank's, the neighbours' (peer-mapped)
  unsigned long long pbMine, pbLeft, pbRight;  // bytes per plane slot
  int pitchMine, pitchLeft, pitchRight;
  int lw, lwLeft, H, vec;
  int n;                                // planes to exchange
  int slot[13];
  unsigned seq;
  unsigned long long spinNs;
};
__device__ __forceinline__ unsigned* flag_of(unsigned char* arena, unsigned long long pb, int f) {
  return reinterpret_cast<unsigned*>(arena + pb * kArenaSlots) + f * kFlagStride;
}

// One thread: tell both neighbours that this rank has finished reading the ghost columns of exchange seq-1
// (everything that reads them precedes this kernel in stream order), then wait for the same word from them.
// A kernel of its own so that nothing bigger than one warp ever sits on an SM while it waits (a spinning
// many-block kernel displaces a resident CTA of the advection kernel on every SM it occupies: measured
// +0.2 ms per iteration at 2 GPUs, profiles/r3_multi_gpu.md).
__global__ void k_ghosts_free(const __grid_constant__ PushArgs a) {
  unsigned* myFlags = flag_of(a.mine, a.pbMine, 0);
  st_release_sys(flag_of(a.left, a.pbLeft, kFlagFreeR), a.seq);
  st_release_sys(flag_of(a.right, a.pbRight, kFlagFreeL), a.seq);
  wait_flag(myFlags + kFlagFreeL * kFlagStride, a.seq, myFlags + kFlagErr * kFlagStride, a.spinNs);
  wait_flag(myFlags + kFlagFreeR * kFlagStride, a.seq, myFlags + kFlagErr * kFlagStride, a.spinNs);
}

// Grid-stride copy of this rank's outermost owned columns into the neighbours' ghost columns; the last block
// to finish raises the neighbours' data flags.
__global__ void __launch_bounds__(256) k_push_ghosts(const __grid_constant__ PushArgs a) {
  unsigned* myFlags = flag_of(a.mine, a.pbMine, 0);
  const int per = a.H * 2;  // (row, column quad)
  // every CTA takes a CONTIGUOUS run of (plane, row) items: consecutive rows of one plane share 2 MB pages, and the
  // peer-memory stores of a CTA then touch ~20 pages instead of all 800 (a grid-stride loop made the push 5x slower)
  const int total = per * a.n, chunk = (total + gridDim.x - 1) / gridDim.x;
  const int tEnd = min(total, ((int)blockIdx.x + 1) * chunk);
#pragma unroll 2
  for (int t = blockIdx.x * chunk + threadIdx.x; t < tEnd; t += blockDim.x) {
    const int k = t / per, r = t - k * per;
    const int y = r >> 1, q = (r & 1) * 4;
    const unsigned long long sl = (unsigned long long)a.slot[k];
    const int* src = reinterpret_cast<const int*>(a.mine + sl * a.pbMine) + (size_t)y * a.pitchMine;
    int* dl = reinterpret_cast<int*>(a.left + sl * a.pbLeft) + (size_t)y * a.pitchLeft + kGhost + a.lwLeft + q;  // its right ghost zone
    int* dr = reinterpret_cast<int*>(a.right + sl * a.pbRight) + (size_t)y * a.pitchRight + q;                   // its left ghost zone
    const int* sL = src + kGhost + q;   // my leftmost owned columns
    const int* sR = src + a.lw + q;     // my rightmost owned columns
    if (a.compact) {  // the neighbours' landing zones: always 16-byte aligned
      dl = zone_row(a.left, a.pbLeft, a.H, 1, k, y) + q;
      dr = zone_row(a.right, a.pbRight, a.H, 0, k, y) + q;
      int4 vl, vr;
      if (a.vec) {
        vl = *reinterpret_cast<const int4*>(sL);
        vr = *reinterpret_cast<const int4*>(sR);
      } else {
        vl = make_int4(sL[0], sL[1], sL[2], sL[3]);
        vr = make_int4(sR[0], sR[1], sR[2], sR[3]);
      }
      *reinterpret_cast<int4*>(dl) = vl;
      *reinterpret_cast<int4*>(dr) = vr;
    } else if (a.vec) {
      *reinterpret_cast<int4*>(dl) = *reinterpret_cast<const int4*>(sL);
      *reinterpret_cast<int4*>(dr) = *reinterpret_cast<const int4*>(sR);
    } else {
#pragma unroll
      for (int i = 0; i < 4; i++) { dl[i] = sL[i]; dr[i] = sR[i]; }
    }
  }
  // one system-scope fence per BLOCK: the barrier orders the block's stores before thread 0's fence (fence cumulativity,
  // the pattern of cooperative-groups grid sync); a fence in every thread (592 warps x membar.sys) measurably slowed
  // the advection kernel running beside this one (profiles/r3_multi_gpu.md)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    unsigned* done = myFlags + kFlagBlocks * kFlagStride;
    if (atomicAdd(done, 1u) == gridDim.x - 1) {  // last block: every block's stores are visible system-wide
      *done = 0u;
      __threadfence_system();
      st_release_sys(flag_of(a.left, a.pbLeft, kFlagDataR), a.seq);
      st_release_sys(flag_of(a.right, a.pbRight, kFlagDataL), a.seq);
    }
  }
}

__global__ void k_wait_ghosts(unsigned* myFlags, unsigned seq, unsigned long long spinNs) {
  wait_flag(myFlags + kFlagDataL * kFlagStride, seq, myFlags + kFlagErr * kFlagStride, spinNs);
  wait_flag(myFlags + kFlagDataR * kFlagStride, seq, myFlags + kFlagErr * kFlagStride, spinNs);
}

// transport "peerc": this rank's landing zones -> its ghost columns (local copy, after k_wait_ghosts in stream order)
struct UnpackArgs {
  unsigned char* mine;
  unsigned long long pb;
  int pitch, lw, H, vec, n;
  int slot[13];
};
__global__ void __launch_bounds__(256) k_unpack_zone(const __grid_constant__ UnpackArgs a) {
  const int per = a.H * 2, total = per * a.n;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const int k = t / per, r = t - k * per;
    const int y = r >> 1, q = (r & 1) * 4;
    int* row = reinterpret_cast<int*>(a.mine + (unsigned long long)a.slot[k] * a.pb) + (size_t)y * a.pitch;
    const int4 vl = *reinterpret_cast<const int4*>(zone_row(a.mine, a.pb, a.H, 0, k, y) + q);
    const int4 vr = *reinterpret_cast<const int4*>(zone_row(a.mine, a.pb, a.H, 1, k, y) + q);
    int* dl = row + q;                   // left ghost zone
    int* dr = row + kGhost + a.lw + q;   // right ghost zone
    if (a.vec) {
      *reinterpret_cast<int4*>(dl) = vl;
      *reinterpret_cast<int4*>(dr) = vr;
    } else {
      dl[0] = vl.x; dl[1] = vl.y; dl[2] = vl.z; dl[3] = vl.w;
      dr[0] = vr.x; dr[1] = vr.y; dr[2] = vr.z; dr[3] = vr.w;
    }
  }
}

// NaN / Inf scan (debug aid; the reference only has the local NaN-avoidance resets of advectionShader.frag:387-397)
__global__ void k_count_nonfinite(Planes4 a, Planes4 b, size_t n, unsigned long long* __restrict__ out) {
  unsigned c = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
#pragma unroll
    for (int k = 0; k < 4; k++) c += (isfinite(a.c[k][i]) ? 0u : 1u) + (isfinite(b.c[k][i]) ? 0u : 1u);
  }
  c = __reduce_add_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, (unsigned long long)c);
}

// n scattered texels of one RGBA32F field -> dense float4 array (weather-station style probes);
// applyPressure: the field is the fused schedule's base_1 with the pressure pass still pending
__global__ void k_gather_points(GlobalCtx c, Planes4 field, int applyPressure, int n, const int* __restrict__ xy, int lx_off,
                                float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int x = xy[2 * i] + lx_off, y = xy[2 * i + 1];
  if (x < c.g.ox0 || x >= c.g.ox1 || y < 0 || y >= c.g.H) {  // not one of this rank's own columns (ghost columns belong to the neighbour)
    out[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  float4 b = field.ld((size_t)y * c.g.pitch + x);
  if (applyPressure) {
    const char4 wYm = c.wall4(x, y - 1);
    pressure_cell(b.x, b.y, b.z, b.w, c.bx(x - 1, y), c.by(x, y - 1), c.bt(x, y - 1), wYm.x, wYm.y);
  }
  out[i] = b;
}

// cuTensorMapEncodeTiled, resolved through the runtime (no link against libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;

}  // namespace

struct wsb_sim {
  wsb_config cfg;
  int W, H, ND;
  int x_begin, lw, ghost, pitch;
  int schedule;
  Geom g;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool timed = false;

  // one RGBA32F texture = four float planes, with the TMA descriptors of each plane for the two
  // box shapes the fused kernels stage (tile + 2-cell halo: dry / advection; tile + 3: boundary)
  struct Field { Planes4 p; CUtensorMap map2[4], map3[4], map0[4], mapD[4]; int slot = -1; };  // map0: the bare tile (own-cell operands); mapD: dry-sweep box; slot: first exchange-arena slot (strips)
  Field base[2] = {}, water[2] = {}, light[2] = {};
  int* wall[2] = {};
  CUtensorMap wallMap2[2], wallMap3[2], wallMapD[2];
  bool use_tma = false;
  float4* fb = nullptr;
  float2 *dep = nullptr, *vort = nullptr;
  unsigned char* dry_tilewalls = nullptr;  // k_fused_dry: one byte per 64 x 28 tile, "a wall cell in the staged region" (k_wall_tilemap)
  bool dry_tilewalls_valid = false;        // false whenever the wall texture may have changed
  SpriteGrid sg{};  // sprite origins + dirty-tile map of the particle pass (single GPU, n_droplets > 0)
  float* curl = nullptr;
  float* drops[2] = {};
  float *initial_T = nullptr, *sndT = nullptr, *sndW = nullptr, *sndV = nullptr;
  float* lightning = nullptr;  // 4 floats: lightningDataTexture
  float* inactive = nullptr;   // the `inactiveDroplets` uniform, kept on the device
  unsigned* maxv = nullptr;
  float4* scratch = nullptr;   // read_rect staging
  size_t scratch_cells = 0;

  DevParams dp;
  long long iter = 0;
  bool even = true;
  int last_drops = 0;
  bool pressure_pending = false;  // FUSED: base_1 still needs the pressure pass
  bool fb_dirty = false;          // feedback / deposition hold non-zero data
  long long launches = 0;

  // per-kernel-class device timing (wsb_set_profiling): events around every launch of a step call
  bool profiling = false;
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_used = 0;
  struct Span { int kind; size_t e0, e1; };
  std::vector<Span> spans;

  // multi-GPU
  cudaStream_t comm_stream = nullptr;       // halo exchange runs here, beside the next boundary kernel
  cudaEvent_t evCompute = nullptr, evExch = nullptr;
  bool exch_pending = false;                // an exchange is in flight on comm_stream
  void* comm = nullptr;
  unsigned char *sendL = nullptr, *sendR = nullptr, *recvL = nullptr, *recvR = nullptr;
  size_t halo_bytes = 0;
  // peer-memory transport (strips): the exchanged planes live in one arena the neighbours map
  unsigned char* arena = nullptr;
  size_t plane_bytes = 0, arena_bytes = 0;
  int wall_slot[2] = {-1, -1};
  bool peer_mode = false;                   // wsb_connect_peers succeeded: k_push_ghosts / k_wait_ghosts instead of NCCL
  bool peer_compact = false;                // transport "peerc": the push goes through the landing zones (k_unpack_zone after the wait)
  bool pend_compact = false; int pend_n = 0; int pend_slot[13] = {0};  // the exchange in flight: how it was pushed, which planes
  unsigned char *peerL = nullptr, *peerR = nullptr;
  bool peerL_ipc = false, peerR_ipc = false;  // opened with cudaIpcOpenMemHandle (to be closed)
  size_t pbL = 0, pbR = 0;
  int pitchL = 0, pitchR = 0, lwL = 0;
  unsigned xseq = 0, pending_seq = 0;       // exchanges issued / the one in flight
  unsigned long long spin_ns = kSpinLimitNs;
  int n_sms = 148;
  int push_blocks = 16;                     // CTAs of k_push_ghosts (WSB_DBG_PUSH_BLOCKS): each displaces an advection CTA while it runs — measured 8/16/37/74/148, profiles/r3_multi_gpu.md
  bool dbg_nopush = false;                  // WSB_DBG_NOPUSH: timing experiments only — ghost columns are never refreshed
  cudaEvent_t evEdge = nullptr, evPush = nullptr, evPvbI = nullptr, evPvbE = nullptr, evAdvI = nullptr;
  bool push_pending = false;                // a k_push_ghosts is reading this rank's edge columns
  int pvbInnerEnd = 0, advEdgeStart = 0;    // tile-column split of the strip (local columns)
};

namespace {

size_t cells(const wsb_sim* s) { return (size_t)s->pitch * s->H; }

GlobalCtx make_ctx(const wsb_sim* s, int b, int w, int wl, int l) {
  GlobalCtx c;
  c.base = s->base[b].p;
  c.water = s->water[w].p;
  c.wall = s->wall[wl];
  c.vortf = s->vort;
  c.light = s->light[l].p;
  c.fb = s->fb;
  c.dep = s->dep;
  c.g = s->g;
  return c;
}

void refresh_derived(wsb_sim* s) {
  s->dp.sinSun = (float)sin((double)s->dp.in.sunAngle);
  s->dp.cosSun = (float)cos((double)s->dp.in.sunAngle);
}
void set_iter_uniform(wsb_sim* s) {
  s->dp.iterNum = (float)s->iter;
  s->dp.iterI = (int)s->dp.iterNum;
}

dim3 grid2d(const wsb_sim* s, int bx, int by) { return dim3((s->pitch + bx - 1) / bx, (s->H + by - 1) / by); }

int check_launch(wsb_sim* s, const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail("launch of %s failed: %s", what, cudaGetErrorString(e));
  s->launches++;
  return 0;
}
#define LAUNCHED(what) \
  do { if (check_launch(s, what)) return 1; } while (0)

// --- per-kernel timing -----------------------------------------------------------------------
size_t prof_mark(wsb_sim* s, cudaStream_t st) {
  if (s->ev_used == s->ev_pool.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    s->ev_pool.push_back(e);
  }
  cudaEventRecord(s->ev_pool[s->ev_used], st);
  return s->ev_used++;
}
struct ProfScope {  // brackets the launches of one kernel class on one stream
  wsb_sim* s; int kind; size_t e0; cudaStream_t st;
  ProfScope(wsb_sim* s_, int k, cudaStream_t st_ = nullptr) : s(s_), kind(k), e0(0), st(st_ ? st_ : s_->stream) {
    if (s->profiling) e0 = prof_mark(s, st);
  }
  ~ProfScope() { if (s->profiling) s->spans.push_back({kind, e0, prof_mark(s, st)}); }
};

// --- halo exchange ---------------------------------------------------------------------------
// The planes one exchange carries: pointers for the NCCL transport, arena slots for the peer transport
struct XPlanes { HaloPlanes hp{}; int slot[kMaxHaloPlanes]; };
void add_planes(XPlanes& x, const wsb_sim::Field& f) {
  for (int k = 0; k < 4; k++) {
    x.slot[x.hp.n] = f.slot + k;
    x.hp.p[x.hp.n++] = reinterpret_cast<int*>(f.p.c[k]);
  }
}
void add_wall(XPlanes& x, const wsb_sim* s, int k) {
  x.slot[x.hp.n] = s->wall_slot[k];
  x.hp.p[x.hp.n++] = s->wall[k];
}

// Peer transport: on the communication stream, after whatever the caller has ordered there — the handshake kernel,
// then the copy of this rank's outermost owned columns into the neighbours' ghost columns.
int push_ghosts(wsb_sim* s, const XPlanes& xp) {
  if (s->dbg_nopush) return 0;
  cudaStream_t cs = s->comm_stream;
  const HaloPlanes& hp = xp.hp;
  ProfScope prof(s, WSB_KERNEL_HALO, cs);
  PushArgs a{};
  a.compact = s->peer_compact ? 1 : 0;
  a.mine = s->arena; a.left = s->peerL; a.right = s->peerR;
  a.pbMine = s->plane_bytes; a.pbLeft = s->pbL; a.pbRight = s->pbR;
  a.pitchMine = s->pitch; a.pitchLeft = s->pitchL; a.pitchRight = s->pitchR;
  a.lw = s->lw; a.lwLeft = s->lwL; a.H = s->H;
  a.vec = (s->pitch % 4 == 0 && s->pitchL % 4 == 0 && s->pitchR % 4 == 0 && s->lw % 4 == 0 && s->lwL % 4 == 0) ? 1 : 0;
  a.n = hp.n;
  for (int k = 0; k < hp.n; k++) a.slot[k] = xp.slot[k];
  a.seq = ++s->xseq;
  a.spinNs = s->spin_ns;
  k_ghosts_free<<<1, 1, 0, cs>>>(a);
  LAUNCHED("k_ghosts_free");
  const int threads = 256, blocks = std::min(s->push_blocks, (hp.n * s->H * 2 + threads - 1) / threads);  // short-lived, a fraction of an SM wave
  k_push_ghosts<<<blocks, threads, 0, cs>>>(a);
  LAUNCHED("k_push_ghosts");
  CK(cudaEventRecord(s->evPush, cs));
  s->pending_seq = a.seq;
  s->pend_compact = s->peer_compact; s->pend_n = hp.n;
  for (int k = 0; k < hp.n; k++) s->pend_slot[k] = xp.slot[k];
  s->push_pending = true;
  s->exch_pending = true;
  return 0;
}

// Ghost-column exchange on comm_stream, ordered after everything enqueued on the compute stream so
// far.  It stays in flight (exch_pending) until join_exchange() orders the compute stream behind it.
int exchange(wsb_sim* s, const XPlanes& xp) {
  if (s->cfg.n_ranks <= 1) return 0;
  cudaStream_t cs = s->comm_stream;
  const HaloPlanes& hp = xp.hp;
  CK(cudaEventRecord(s->evCompute, s->stream));
  CK(cudaStreamWaitEvent(cs, s->evCompute, 0));
  if (s->peer_mode) return push_ghosts(s, xp);
  if (!s->comm) return fail("no ghost-exchange transport: call wsb_connect_peers (or pass an NCCL comm_id to wsb_create) before stepping a strip");
  {
    ProfScope prof(s, WSB_KERNEL_HALO, cs);
    const int n = s->H * kGhost * hp.n;
    const size_t bytes = (size_t)n * 4;
    if (bytes > s->halo_bytes) return fail("halo staging overflow");
    const int threads = 256, blocks = (n + threads - 1) / threads;
    k_pack_halo<<<blocks, threads, 0, cs>>>(hp, s->pitch, s->H, s->lw, (int*)s->sendL, (int*)s->sendR);
    LAUNCHED("k_pack_halo");
    const int left = (s->cfg.rank + s->cfg.n_ranks - 1) % s->cfg.n_ranks;
    const int right = (s->cfg.rank + 1) % s->cfg.n_ranks;
    // Sends go (left, right); receives are posted (right, left) so that with two ranks — where both
    // neighbours are the same peer and NCCL matches operations in call order — the block a rank
    // sends to its left neighbour lands in that neighbour's RIGHT ghost zone.
    NCK(g_nccl.GroupStart());
    NCK(g_nccl.Send(s->sendL, bytes, kNcclChar, left, s->comm, cs));
    NCK(g_nccl.Send(s->sendR, bytes, kNcclChar, right, s->comm, cs));
    NCK(g_nccl.Recv(s->recvR, bytes, kNcclChar, right, s->comm, cs));
    NCK(g_nccl.Recv(s->recvL, bytes, kNcclChar, left, s->comm, cs));
    NCK(g_nccl.GroupEnd());
    s->launches++;
    k_unpack_halo<<<blocks, threads, 0, cs>>>(hp, s->pitch, s->H, s->lw, (const int*)s->recvL, (const int*)s->recvR);
    LAUNCHED("k_unpack_halo");
  }
  CK(cudaEventRecord(s->evExch, cs));
  s->exch_pending = true;
  return 0;
}
// the neighbours' columns of the exchange in flight have arrived: one-thread kernel on stream `st`
int wait_ghosts(wsb_sim* s, cudaStream_t st) {
  ProfScope prof(s, WSB_KERNEL_WAIT, st);
  k_wait_ghosts<<<1, 1, 0, st>>>(reinterpret_cast<unsigned*>(s->arena + s->plane_bytes * kArenaSlots), s->pending_seq, s->spin_ns);
  LAUNCHED("k_wait_ghosts");
  if (s->pend_compact) {  // the neighbours' columns are in this rank's landing zones: copy them into the ghost columns
    UnpackArgs u{};
    u.mine = s->arena; u.pb = s->plane_bytes; u.pitch = s->pitch; u.lw = s->lw; u.H = s->H;
    u.vec = (s->pitch % 4 == 0 && s->lw % 4 == 0) ? 1 : 0;
    u.n = s->pend_n;
    for (int k = 0; k < u.n; k++) u.slot[k] = s->pend_slot[k];
    const int threads = 256, blocks = std::min(32, (u.n * s->H * 2 + threads - 1) / threads);
    k_unpack_zone<<<blocks, threads, 0, st>>>(u);
    LAUNCHED("k_unpack_zone");
  }
  return 0;
}
// the compute stream may touch the ghost columns again: the neighbours' data of the exchange in flight has arrived
int join_exchange(wsb_sim* s) {
  if (s->exch_pending) {
    if (s->peer_mode) {
      if (wait_ghosts(s, s->stream)) return 1;
    } else {
      CK(cudaStreamWaitEvent(s->stream, s->evExch, 0));
    }
    s->exch_pending = false;
  }
  return 0;
}
// the compute stream may overwrite this rank's edge columns again: its own push has read them
int join_push(wsb_sim* s) {
  if (s->push_pending) {
    CK(cudaStreamWaitEvent(s->stream, s->evPush, 0));
    s->push_pending = false;
  }
  return 0;
}

// --- reference schedule ----------------------------------------------------------------------
const dim3 kRefBlock(64, 4);

int ref_velocity(wsb_sim* s) {
  k_ref_velocity<<<grid2d(s, 64, 4), kRefBlock, 0, s->stream>>>(make_ctx(s, 0, 0, 0, 0), s->dp, s->base[1].p, s->wall[1]);
  LAUNCHED("k_ref_velocity");
  return 0;
}
int ref_curl(wsb_sim* s) {
  k_ref_curl<<<grid2d(s, 64, 4), kRefBlock, 0, s->stream>>>(make_ctx(s, 1, 1, 1, 0), s->curl);
  LAUNCHED("k_ref_curl");
  return 0;
}
int ref_vorticity(wsb_sim* s) {
  k_ref_vorticity<<<grid2d(s, 64, 4), kRefBlock, 0, s->stream>>>(s->g, s->curl, s->vort);
  LAUNCHED("k_ref_vorticity");
  return 0;
}
int ref_boundary(wsb_sim* s) {
  set_iter_uniform(s);
  k_ref_boundary<<<grid2d(s, 64, 4), kRefBlock, 0, s->stream>>>(make_ctx(s, 1, 1, 1, 0), s->dp, s->initial_T, s->base[0].p,
                                                                s->water[0].p, s->wall[0]);
  LAUNCHED("k_ref_boundary");
  return 0;
}
int ref_advection(wsb_sim* s, bool dry) {
  GlobalCtx c = make_ctx(s, 0, 0, 0, 0);
  if (dry)
    k_ref_advection<true><<<grid2d(s, 64, 4), kRefBlock, 0, s->stream>>>(c, s->dp, s->initial_T, s->sndT, s->sndW, s->sndV,
                                                                         s->base[1].p, s->water[1].p, s->wall[1], s->maxv);
  else
    k_ref_advection<false><<<grid2d(s, 64, 4), kRefBlock, 0, s->stream>>>(c, s->dp, s->initial_T, s->sndT, s->sndW, s->sndV,
                                                                          s->base[1].p, s->water[1].p, s->wall[1], s->maxv);
  LAUNCHED("k_ref_advection");
  return 0;
}
int ref_pressure(wsb_sim* s) {
  k_ref_pressure<<<grid2d(s, 64, 4), kRefBlock, 0, s->stream>>>(make_ctx(s, 1, 1, 1, 0), s->base[0].p, s->wall[0]);
  LAUNCHED("k_ref_pressure");
  return 0;
}
int ref_lighting(wsb_sim* s) {
  const int src = s->even ? 0 : 1, dst = s->even ? 1 : 0;  // app.js:5912-5926
  k_ref_lighting<<<grid2d(s, 64, 4), kRefBlock, 0, s->stream>>>(make_ctx(s, 1, 1, 1, src), s->dp, s->light[dst].p);
  LAUNCHED("k_ref_lighting");
  s->even = !s->even;
  return 0;
}

// feedback clear (app.js:5933-5934) as a separate step: REFERENCE schedule, and FUSED when the
// clear cannot ride on k_fused_pvb.
int clear_feedback(wsb_sim* s) {
  if (!s->fb_dirty) return 0;
  CK(cudaMemsetAsync(s->fb, 0, cells(s) * sizeof(float4), s->stream));
  CK(cudaMemsetAsync(s->dep, 0, cells(s) * sizeof(float2), s->stream));
  if (s->sg.dirty) CK(cudaMemsetAsync(s->sg.dirty, 0, (size_t)s->sg.tilesX * s->sg.tilesY * sizeof(int), s->stream));
  s->launches += 2;
  s->fb_dirty = false;
  return 0;
}

// precipitation particles + latches (app.js:5936-5983).  `even` has already been toggled by the
// lighting step: the source buffer is the one selected BEFORE the toggle.
int precipitation(wsb_sim* s) {
  if (!s->dp.p.enablePrecipitation || s->ND == 0) return 0;
  if (s->cfg.n_ranks > 1) return fail("precipitation particles are single-GPU only (SURVEY 8e)");
  const int src = s->even ? 1 : 0, dst = s->even ? 0 : 1;
  set_iter_uniform(s);
  ProfScope prof(s, WSB_KERNEL_PRECIP);
  const int threads = 256, blocks = (s->ND + threads - 1) / threads;
  k_precipitation<<<blocks, threads, 0, s->stream>>>(s->drops[src], s->drops[dst], s->base[1].p, s->water[1].p, s->fb, s->dep, s->sg,
                                                     s->lightning, s->inactive, s->g, s->dp, s->ND);
  LAUNCHED("k_precipitation");
  {  // sprites = 12 x 12 box filter of the origins, on the tiles that were touched
    ProfScope profSprites(s, WSB_KERNEL_SPRITES);
    const int ctas = std::min(s->n_sms * 4, s->sg.tilesX * s->sg.tilesY);  // persistent: the list of touched tiles is walked on the device
    k_boxsum<<<ctas, 256, kSmemBox, s->stream>>>(s->sg, s->fb, s->dep, s->W, s->H, s->pitch);
    LAUNCHED("k_boxsum");
    k_clear_origins<<<ctas, 256, 0, s->stream>>>(s->sg, s->W, s->H, reinterpret_cast<unsigned*>(s->sg.dirtyCount + 1));
    LAUNCHED("k_clear_origins");
  }
  s->fb_dirty = true;
  s->last_drops = dst;
  k_latch<<<1, 32, 0, s->stream>>>(s->fb, s->inactive, s->lightning, s->dp.iterNum, (s->iter % 600 == 0) ? 1 : 0);
  LAUNCHED("k_latch");
  return 0;
}

int ref_iteration(wsb_sim* s) {
  s->dry_tilewalls_valid = false;
  if (ref_velocity(s) || ref_curl(s) || ref_vorticity(s) || ref_boundary(s) || ref_advection(s, false) ||
      ref_pressure(s) || ref_lighting(s) || clear_feedback(s) || precipitation(s))
    return 1;
  s->iter++;
  return 0;
}

// --- fused schedule --------------------------------------------------------------------------

// tile columns [cx0, cx1) minus [gapAt, gapAt + gapLen) of one fused kernel on stream `st`
constexpr int kNoGap = 0x7fffffff;
int launch_pvb(wsb_sim* s, cudaStream_t st, int cx0, int cx1, int gapAt, int gapLen) {
  GlobalCtx c = make_ctx(s, 1, 1, 1, 0);
  TileMaps<11> maps;
  for (int k = 0; k < 4; k++) {
    maps.m[k] = s->base[1].map3[k];
    maps.m[5 + k] = s->water[1].map0[k];
  }
  maps.m[4] = s->wallMap3[1];
  maps.m[9] = s->light[0].map0[0];   // SUNLIGHT
  maps.m[10] = s->light[0].map0[1];  // NET_HEATING
  c.g.cx0 = cx0; c.g.cx1 = cx1; c.g.cxGapAt = gapAt; c.g.cxGapLen = gapLen;
  k_fused_pvb<<<dim3((cx1 - cx0 - gapLen + kTX - 1) / kTX, (s->H + kTY - 1) / kTY), kNT, kSmem1, st>>>(
      c, s->dp, maps, s->use_tma ? 1 : 0, s->initial_T, s->pressure_pending ? 1 : 0, (s->fb_dirty && s->sg.dirty) ? 1 : 0, s->fb, s->dep,
      s->sg, s->base[0].p, s->water[0].p, s->wall[0]);
  return check_launch(s, "k_fused_pvb");
}
int launch_adv(wsb_sim* s, cudaStream_t st, int src, int dst, int cx0, int cx1, int gapAt, int gapLen) {
  TileMaps<12> maps;
  for (int k = 0; k < 4; k++) {
    maps.m[k] = s->base[0].map2[k];
    maps.m[4 + k] = s->water[0].map2[k];
  }
  maps.m[8] = s->wallMap2[0];
  maps.m[9] = s->light[src].map2[0];   // SUNLIGHT
  maps.m[10] = s->light[src].map2[2];  // IR_DOWN
  maps.m[11] = s->light[src].map2[3];  // IR_UP
  GlobalCtx c = make_ctx(s, 0, 0, 0, src);
  c.g.cx0 = cx0; c.g.cx1 = cx1; c.g.cxGapAt = gapAt; c.g.cxGapLen = gapLen;
  k_fused_adv<<<dim3((cx1 - cx0 - gapLen + kTX - 1) / kTX, (s->H + kTY - 1) / kTY), kNT, kSmem2, st>>>(
      c, s->dp, maps, s->use_tma ? 1 : 0, s->initial_T, s->sndT, s->sndW, s->sndV, s->base[1].p, s->water[1].p, s->wall[1],
      s->light[dst].p, s->maxv);
  return check_launch(s, "k_fused_adv");
}
void strip_exchange_planes(const wsb_sim* s, int dst, XPlanes& xp) {
  add_planes(xp, s->base[1]);
  add_planes(xp, s->water[1]);
  add_wall(xp, s, 1);
  add_planes(xp, s->light[dst]);
}

// One iteration of a strip with the peer transport, on TWO streams.  The two edge tile columns of each fused kernel —
// the only tiles that read ghost columns (boundary kernel) or produce the columns the neighbours need (advection
// kernel) — run on the high-priority communication stream BESIDE the interior tile columns on the compute stream,
// together with the wait for the neighbours' columns and the push of this rank's:
//
//   compute stream   [wait advE(i-1)]  pvb interior          [wait pvbE]  adv interior
//   comm stream      [wait advI(i-1)]  wait ghosts(i-1), pvb edges  [wait pvbI]  adv edges, handshake, push(i)
//
// Edge launches are 2 of ~130 (2 GPUs) .. 33 (8 GPUs) tile columns; as launches of their own on the compute stream each
// costs a drain + refill of the whole GPU (measured +0.05 / +0.08 ms per iteration at 2 GPUs, profiles/r3_multi_gpu.md).
// Cross dependencies: the interior tile column next to an edge stages 4 of the edge's columns, and vice versa.
int fused_iteration_overlapped(wsb_sim* s) {
  set_iter_uniform(s);
  const int src = s->even ? 0 : 1, dst = s->even ? 1 : 0;
  cudaStream_t cs = s->comm_stream;
  CK(cudaStreamWaitEvent(s->stream, s->evEdge, 0));  // advection edges of the previous iteration
  {
    ProfScope prof(s, WSB_KERNEL_PVB);
    if (launch_pvb(s, s->stream, kTX, s->pvbInnerEnd, kNoGap, 0)) return 1;
  }
  CK(cudaEventRecord(s->evPvbI, s->stream));
  CK(cudaStreamWaitEvent(cs, s->evAdvI, 0));         // advection interior of the previous iteration
  if (s->exch_pending) {
    if (wait_ghosts(s, cs)) return 1;
    s->exch_pending = false;
  }
  {
    ProfScope prof(s, WSB_KERNEL_EDGE, cs);
    if (launch_pvb(s, cs, 0, s->pitch, kTX, s->pvbInnerEnd - kTX)) return 1;
  }
  CK(cudaEventRecord(s->evPvbE, cs));
  s->fb_dirty = false;
  CK(cudaStreamWaitEvent(s->stream, s->evPvbE, 0));
  {
    ProfScope prof(s, WSB_KERNEL_ADV);
    if (launch_adv(s, s->stream, src, dst, kTX, s->advEdgeStart, kNoGap, 0)) return 1;
  }
  CK(cudaEventRecord(s->evAdvI, s->stream));
  CK(cudaStreamWaitEvent(cs, s->evPvbI, 0));
  {
    ProfScope prof(s, WSB_KERNEL_EDGE, cs);
    if (launch_adv(s, cs, src, dst, 0, s->pitch, kTX, s->advEdgeStart - kTX)) return 1;
  }
  CK(cudaEventRecord(s->evEdge, cs));
  s->even = !s->even;
  s->pressure_pending = true;
  XPlanes xp;
  strip_exchange_planes(s, dst, xp);
  if (push_ghosts(s, xp)) return 1;
  s->iter++;
  return 0;
}

int fused_iteration(wsb_sim* s) {
  s->dry_tilewalls_valid = false;  // boundary / advection rewrite the wall texture
  if (s->peer_mode && s->pvbInnerEnd > kTX && s->advEdgeStart > kTX) return fused_iteration_overlapped(s);
  set_iter_uniform(s);
  const int src = s->even ? 0 : 1, dst = s->even ? 1 : 0;
  const bool particles = s->dp.p.enablePrecipitation && s->ND > 0;
  // pressure(previous iteration) -> velocity -> curl -> vorticity -> boundary; also clears the
  // feedback / deposition cells it has consumed (app.js:5933-5934 folded in)
  {
    ProfScope prof(s, WSB_KERNEL_PVB);
    // NCCL transport: while the previous iteration's ghost exchange is still in flight, run the tiles whose staged
    // region (tile + 4 columns) stays clear of the ghost columns; then wait for the neighbours' columns; then the two
    // edge tile columns (one launch).
    const int innerEnd = s->pvbInnerEnd;
    if (s->exch_pending && innerEnd > kTX) {
      if (launch_pvb(s, s->stream, kTX, innerEnd, kNoGap, 0)) return 1;
      if (join_exchange(s)) return 1;
      if (launch_pvb(s, s->stream, 0, s->pitch, kTX, innerEnd - kTX)) return 1;
    } else {
      if (join_exchange(s)) return 1;
      if (launch_pvb(s, s->stream, 0, s->pitch, kNoGap, 0)) return 1;
    }
  }
  s->fb_dirty = false;
  // advection (+ condensation ...) -> lighting
  {
    ProfScope prof(s, WSB_KERNEL_ADV);
    if (join_push(s)) return 1;  // the previous push has finished reading the columns this kernel overwrites
    if (launch_adv(s, s->stream, src, dst, 0, s->pitch, kNoGap, 0)) return 1;
  }
  s->even = !s->even;
  s->pressure_pending = true;
  if (particles && precipitation(s)) return 1;
  if (s->cfg.n_ranks > 1) {
    XPlanes xp;
    strip_exchange_planes(s, dst, xp);
    if (exchange(s, xp)) return 1;
  }
  s->iter++;
  return 0;
}

int dry_iteration(wsb_sim* s) {
  if (s->schedule == WSB_SCHEDULE_REFERENCE) {
    // velocity renders into frameBuff_1, advection samples frameBuff_0 (app.js:5832-5890): with the
    // boundary pass left out, velocity's output is handed over unchanged.
    if (ref_velocity(s)) return 1;
    std::swap(s->base[0], s->base[1]);
    std::swap(s->wall[0], s->wall[1]);
    if (ref_advection(s, true) || ref_pressure(s)) return 1;
  } else {
    // same canonical state as the full fused schedule: base_1 = advection output, pressure pending
    {
      const dim3 tiles((s->pitch + kTX - 1) / kTX, (s->H + kTYD - 1) / kTYD);
      if (!s->dry_tilewalls_valid) {  // the dry sweep itself never changes the wall texture
        k_wall_tilemap<<<tiles, 256, 0, s->stream>>>(make_ctx(s, 1, 1, 1, 0), s->dry_tilewalls);
        LAUNCHED("k_wall_tilemap");
        s->dry_tilewalls_valid = true;
      }
      ProfScope prof(s, WSB_KERNEL_DRY);
      TileMaps<5> maps;
      for (int k = 0; k < 4; k++) maps.m[k] = s->base[1].mapD[k];
      maps.m[4] = s->wallMapD[1];
      k_fused_dry<<<tiles, kNT, kSmemDry, s->stream>>>(make_ctx(s, 1, 1, 1, 0), s->dp, maps, s->use_tma ? 1 : 0, s->pressure_pending ? 1 : 0,
                                                       s->dry_tilewalls, s->base[0].p, s->maxv);
      LAUNCHED("k_fused_dry");
    }
    std::swap(s->base[0], s->base[1]);
    s->pressure_pending = true;
    if (s->cfg.n_ranks > 1) {
      XPlanes xp;
      add_planes(xp, s->base[1]);
      if (exchange(s, xp) || join_exchange(s) || join_push(s)) return 1;
    }
  }
  s->iter++;
  return 0;
}

// TMA descriptor of one [H][pitch] 4-byte plane with a (boxW x boxH) box; out-of-range elements read 0
int make_map(wsb_sim* s, CUtensorMap* m, void* plane, bool isInt, int boxW, int boxH) {
  const cuuint64_t dims[2] = {(cuuint64_t)s->pitch, (cuuint64_t)s->H};
  const cuuint64_t strides[1] = {(cuuint64_t)s->pitch * 4};
  const cuuint32_t box[2] = {(cuuint32_t)boxW, (cuuint32_t)boxH};
  const cuuint32_t estr[2] = {1, 1};
  // (L2 promotion none / 128 B / 256 B: no difference on any kernel, profiles/r3_logs/c28_variants.log)
  CUresult r = g_encode(m, isInt ? CU_TENSOR_MAP_DATA_TYPE_INT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, plane, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

// slot >= 0: the four planes are slots slot .. slot+3 of the exchange arena (strips)
int alloc_field(wsb_sim* s, wsb_sim::Field& f, int slot = -1) {
  const size_t n = cells(s);
  f.slot = slot;
  for (int k = 0; k < 4; k++) {
    if (slot >= 0) f.p.c[k] = reinterpret_cast<float*>(s->arena + (size_t)(slot + k) * s->plane_bytes);
    else CK(cudaMalloc(&f.p.c[k], n * sizeof(float)));
    if (s->use_tma && (make_map(s, &f.map2[k], f.p.c[k], false, kSW2, kSH2) || make_map(s, &f.map3[k], f.p.c[k], false, kSW1, kSH1) ||
                       make_map(s, &f.map0[k], f.p.c[k], false, kTX, kTY) || make_map(s, &f.mapD[k], f.p.c[k], false, kSWD, kSHD)))
      return 1;
  }
  return 0;
}

int alloc_all(wsb_sim* s) {
  const size_t n = cells(s);
  // TMA needs 16-byte row strides and at least one box per dimension; anything else is staged through registers
  s->use_tma = false;
  if (s->pitch % 4 == 0 && s->pitch >= kSW1 && s->H >= (kSH1 > kSHD ? kSH1 : kSHD)) {
    if (!g_encode) {
      void* fn = nullptr;
      cudaDriverEntryPointQueryResult q;
      if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
        g_encode = (EncodeTiledFn)fn;
      cudaGetLastError();
    }
    s->use_tma = g_encode != nullptr;
  }
  const bool strips = s->cfg.n_ranks > 1;
  if (strips) {  // everything a ghost exchange can carry lives in one arena the neighbours map (cudaIpc), flags behind it
    s->plane_bytes = (n * 4 + 255) / 256 * 256;
    s->arena_bytes = s->plane_bytes * kArenaSlots + kZoneOffset + zone_bytes(s->H);  // planes | flag lines | landing zones
    CK(cudaMalloc(&s->arena, s->arena_bytes));
    CK(cudaMemset(s->arena, 0, s->arena_bytes));
  }
  for (int k = 0; k < 2; k++) {
    if (alloc_field(s, s->base[k], strips ? 4 * k : -1) || alloc_field(s, s->water[k], strips && k == 1 ? 8 : -1) ||
        alloc_field(s, s->light[k], strips ? 13 + 4 * k : -1))
      return 1;
    if (strips && k == 1) {
      s->wall_slot[1] = 12;
      s->wall[1] = reinterpret_cast<int*>(s->arena + (size_t)12 * s->plane_bytes);
    } else {
      CK(cudaMalloc(&s->wall[k], n * sizeof(int)));
    }
    if (s->use_tma && (make_map(s, &s->wallMap2[k], s->wall[k], true, kSW2, kSH2) || make_map(s, &s->wallMap3[k], s->wall[k], true, kSW1, kSH1) ||
                       make_map(s, &s->wallMapD[k], s->wall[k], true, kSWD, kSHD)))
      return 1;
    if (s->ND) CK(cudaMalloc(&s->drops[k], (size_t)s->ND * 5 * sizeof(float)));
  }
  CK(cudaMalloc(&s->fb, n * sizeof(float4)));
  CK(cudaMalloc(&s->dep, n * sizeof(float2)));
  if (s->ND > 0) {  // particles are single-GPU: pitch == W
    SpriteGrid& sg = s->sg;
    sg.Po = s->W + 1;
    sg.tilesX = (s->W + kPTX - 1) / kPTX;
    sg.tilesY = (s->H + kPTY - 1) / kPTY;
    const size_t no = (size_t)sg.Po * (s->H + 1), nt = (size_t)sg.tilesX * sg.tilesY;
    CK(cudaMalloc(&sg.org4, no * sizeof(float4)));
    CK(cudaMalloc(&sg.org2, no * sizeof(float2)));
    CK(cudaMalloc(&sg.dirty, nt * sizeof(int)));
    CK(cudaMalloc(&sg.dirtyList, nt * sizeof(int)));
    CK(cudaMalloc(&sg.dirtyCount, 2 * sizeof(int)));  // [1]: CTA counter of k_clear_origins
  }
  if (s->schedule == WSB_SCHEDULE_REFERENCE) {
    CK(cudaMalloc(&s->curl, n * sizeof(float)));
    CK(cudaMalloc(&s->vort, n * sizeof(float2)));
  }
  const size_t np = (size_t)s->H + 2;
  CK(cudaMalloc(&s->initial_T, np * 4));
  CK(cudaMalloc(&s->sndT, np * 4));
  CK(cudaMalloc(&s->sndW, np * 4));
  CK(cudaMalloc(&s->sndV, np * 4));
  CK(cudaMemset(s->initial_T, 0, np * 4));
  CK(cudaMemset(s->sndT, 0, np * 4));
  CK(cudaMemset(s->sndW, 0, np * 4));
  CK(cudaMemset(s->sndV, 0, np * 4));
  CK(cudaMalloc(&s->dry_tilewalls, (size_t)((s->pitch + kTX - 1) / kTX) * ((s->H + kTYD - 1) / kTYD)));
  CK(cudaMalloc(&s->lightning, 16));
  CK(cudaMalloc(&s->inactive, 4));
  CK(cudaMalloc(&s->maxv, 4));
  if (s->cfg.n_ranks > 1) {
    s->halo_bytes = (size_t)s->H * kGhost * 4 * 13;  // base 4 + water 4 + wall 1 + light 4 planes
    CK(cudaMalloc(&s->sendL, s->halo_bytes));
    CK(cudaMalloc(&s->sendR, s->halo_bytes));
    CK(cudaMalloc(&s->recvL, s->halo_bytes));
    CK(cudaMalloc(&s->recvR, s->halo_bytes));
  }
  return 0;
}

int need_scratch(wsb_sim* s, size_t ncells) {
  if (ncells > s->scratch_cells) {
    CK(cudaStreamSynchronize(s->stream));
    cudaFree(s->scratch);
    s->scratch = nullptr;
    s->scratch_cells = 0;
    CK(cudaMalloc(&s->scratch, ncells * sizeof(float4)));
    s->scratch_cells = ncells;
  }
  return 0;
}

int zero_transients(wsb_sim* s) {
  const size_t n = cells(s);
  for (int k = 0; k < 2; k++)
    for (int ch = 0; ch < 4; ch++) CK(cudaMemsetAsync(s->light[k].p.c[ch], 0, n * sizeof(float), s->stream));
  CK(cudaMemsetAsync(s->fb, 0, n * sizeof(float4), s->stream));
  CK(cudaMemsetAsync(s->dep, 0, n * sizeof(float2), s->stream));
  if (s->sg.org4) {
    const size_t no = (size_t)s->sg.Po * (s->H + 1), nt = (size_t)s->sg.tilesX * s->sg.tilesY;
    CK(cudaMemsetAsync(s->sg.org4, 0, no * sizeof(float4), s->stream));
    CK(cudaMemsetAsync(s->sg.org2, 0, no * sizeof(float2), s->stream));
    CK(cudaMemsetAsync(s->sg.dirty, 0, nt * sizeof(int), s->stream));
    CK(cudaMemsetAsync(s->sg.dirtyCount, 0, 2 * sizeof(int), s->stream));
  }
  if (s->curl) CK(cudaMemsetAsync(s->curl, 0, n * sizeof(float), s->stream));
  if (s->vort) CK(cudaMemsetAsync(s->vort, 0, n * sizeof(float2), s->stream));
  CK(cudaMemsetAsync(s->lightning, 0, 16, s->stream));
  CK(cudaMemsetAsync(s->inactive, 0, 4, s->stream));
  CK(cudaMemsetAsync(s->maxv, 0, 4, s->stream));
  s->iter = 0;
  s->even = true;
  s->last_drops = 0;
  s->pressure_pending = false;
  s->fb_dirty = false;
  s->dry_tilewalls_valid = false;
  return 0;
}

// copy the columns of this rank's padded strip out of a GLOBAL [H][W] host array
int upload_field(wsb_sim* s, void* dst, const void* src, size_t elt) {
  int start = s->x_begin - s->ghost, remaining = s->pitch, off = 0;
  while (remaining > 0) {
    const int gs = ((start % s->W) + s->W) % s->W;
    const int len = std::min(remaining, s->W - gs);
    CK(cudaMemcpy2DAsync((char*)dst + (size_t)off * elt, (size_t)s->pitch * elt, (const char*)src + (size_t)gs * elt,
                         (size_t)s->W * elt, (size_t)len * elt, s->H, cudaMemcpyHostToDevice, s->stream));
    start += len;
    off += len;
    remaining -= len;
  }
  return 0;
}

// packed RGBA32F texels on the host -> channel planes of copy 0 (through the texel scratch buffer)
int upload_texels(wsb_sim* s, const Planes4& dst, const float* src, bool global_array) {
  const size_t n = cells(s);
  if (need_scratch(s, n)) return 1;
  if (global_array) {
    if (upload_field(s, s->scratch, src, 16)) return 1;
  } else {
    CK(cudaMemcpyAsync(s->scratch, src, n * 16, cudaMemcpyHostToDevice, s->stream));
  }
  k_texels_to_planes<<<(unsigned)((n + 255) / 256), 256, 0, s->stream>>>(s->scratch, n, dst);
  LAUNCHED("k_texels_to_planes");
  return 0;
}

constexpr unsigned kPeerMagic = 0x57534250u;  // "WSBP"
struct PeerInfo {
  unsigned magic;
  int pid, device, rank, pitch, lw, H, reserved;
  unsigned long long plane_bytes, arena_bytes, raw;
  cudaIpcMemHandle_t handle;
};

// Strips stay bit-identical to the single-GPU run only while one iteration's dependency radius fits
// the ghost zone: 3 (boundary kernel) + 1 + ceil(|v|) (back-trace footprint) <= kGhost, i.e. |v| <= 4
// (strips.MAX_STRIP_VELOCITY).  Checked after every synchronisation of a strip, with the bounded-spin
// error word of the peer transport.  The stream must be idle.
constexpr float kMaxStripVelocity = 4.0f;
int strip_checks(wsb_sim* s) {
  if (s->cfg.n_ranks <= 1) return 0;
  float v = 0.0f;
  CK(cudaMemcpy(&v, s->maxv, 4, cudaMemcpyDeviceToHost));
  if (!(v <= kMaxStripVelocity))
    return fail("strip %d: max |v| = %g cells/iteration exceeds the ghost-zone budget of %g; owned cells no longer match the single-GPU run",
                s->cfg.rank, v, kMaxStripVelocity);
  if (s->arena) {
    unsigned err = 0;
    CK(cudaMemcpy(&err, reinterpret_cast<unsigned*>(s->arena + s->plane_bytes * kArenaSlots) + kFlagErr * kFlagStride, 4, cudaMemcpyDeviceToHost));
    if (err) return fail("strip %d: a ghost-exchange wait ran out (a neighbour is missing or stepped a different number of iterations)", s->cfg.rank);
  }
  return 0;
}

int use_device(const wsb_sim* s) {
  CK(cudaSetDevice(s->cfg.device));
  return 0;
}

}  // namespace

// =============================================================================================
extern "C" {

const char* wsb_last_error(void) { return g_err; }

const char* wsb_build_info(void) {
#define WSB_STR2(x) #x
#define WSB_STR(x) WSB_STR2(x)
  return "libwsb200 abi 1 | sm_100a | nvcc " WSB_STR(__CUDACC_VER_MAJOR__) "." WSB_STR(__CUDACC_VER_MINOR__) "." WSB_STR(__CUDACC_VER_BUILD__)
         " | fmad=false | tile 64x16 (dry sweep 64x28), 256 threads, TMA-staged channel planes | ghost 8 | strips: peer-memory push (NCCL send/recv optional)";
}

int wsb_comm_id_create(uint8_t out[WSB_COMM_ID_BYTES]) {
  if (!out) return fail("wsb_comm_id_create: null output");
  if (load_nccl()) return 1;
  Id128 id;
  NCK(g_nccl.GetUniqueId(&id));
  memcpy(out, id.internal, WSB_COMM_ID_BYTES);
  return 0;
}

int wsb_create(const wsb_config* cfg, wsb_sim** out) {
  if (!cfg || !out) return fail("wsb_create: null argument");
  *out = nullptr;
  if (cfg->abi_version != WSB_ABI_VERSION) return fail("wsb_create: abi_version %d != %d", cfg->abi_version, WSB_ABI_VERSION);
  if (cfg->width < 32 || cfg->height < 32 || cfg->width > 65535 || cfg->height > 65535)
    return fail("wsb_create: grid %dx%d outside 32..65535 (the save format stores u16 sizes)", cfg->width, cfg->height);
  if ((long long)cfg->width * cfg->height > (1LL << 30)) return fail("wsb_create: more than 2^30 cells (32-bit cell indices)");
  if (cfg->n_droplets < 0) return fail("wsb_create: negative n_droplets");
  if (cfg->n_ranks < 1 || cfg->rank < 0 || cfg->rank >= cfg->n_ranks) return fail("wsb_create: bad rank %d / %d", cfg->rank, cfg->n_ranks);
  if (cfg->schedule != WSB_SCHEDULE_FUSED && cfg->schedule != WSB_SCHEDULE_REFERENCE) return fail("wsb_create: unknown schedule %d", cfg->schedule);
  if (cfg->n_ranks > 1 && cfg->schedule != WSB_SCHEDULE_FUSED) return fail("wsb_create: multi-GPU needs WSB_SCHEDULE_FUSED");
  if (cfg->n_ranks > 1 && cfg->n_droplets > 0) return fail("wsb_create: precipitation particles are single-GPU only; pass n_droplets = 0");
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (cfg->device < 0 || cfg->device >= ndev) return fail("wsb_create: CUDA device %d not present (%d devices)", cfg->device, ndev);

  wsb_sim* s = new (std::nothrow) wsb_sim();
  if (!s) return fail("wsb_create: out of host memory");
  s->cfg = *cfg;
  s->W = cfg->width;
  s->H = cfg->height;
  s->ND = cfg->n_droplets;
  s->schedule = cfg->schedule;
  s->x_begin = (int)(((long long)cfg->rank * s->W) / cfg->n_ranks);
  s->lw = (int)(((long long)(cfg->rank + 1) * s->W) / cfg->n_ranks) - s->x_begin;
  s->ghost = cfg->n_ranks > 1 ? kGhost : 0;
  s->pitch = s->lw + 2 * s->ghost;
  if (cfg->n_ranks > 1 && s->lw < 2 * kGhost) {
    const int lw = s->lw;
    delete s;
    return fail("wsb_create: strip of %d columns is narrower than %d", lw, 2 * kGhost);
  }

  Geom& g = s->g;
  g.Wg = s->W; g.H = s->H; g.pitch = s->pitch; g.gx0 = s->x_begin - s->ghost; g.wrap = cfg->n_ranks == 1 ? 1 : 0;
  g.cx0 = 0; g.cx1 = s->pitch; g.cxGapAt = 0x7fffffff; g.cxGapLen = 0;
  g.ox0 = s->ghost; g.ox1 = s->ghost + s->lw;
  // strips: tile columns [kTX, pvbInnerEnd) of the boundary kernel stage nothing from the ghost zones; tile columns
  // 0 and [advEdgeStart, pitch) of the advection kernel produce the columns the neighbours receive
  s->pvbInnerEnd = kTX + ((s->pitch - (kGhost + kHX) - kTX) / kTX) * kTX;
  s->advEdgeStart = (s->lw / kTX) * kTX;
  g.texelX = (float)(1.0 / (double)s->W); g.texelY = (float)(1.0 / (double)s->H);   // app.js:5436 -> uniform2f
  g.Wf = (float)s->W; g.Hf = (float)s->H;
  g.ltexelX = 1.0f / g.Wf; g.ltexelY = 1.0f / g.Hf;                                 // advectionShader.frag:69
  g.cellHeightComp = 300.0f / g.Hf;                                                 // lightingShader.frag:44 (IEEE fp32 division, as on the device)
  g.nearV = (g.Wg <= (1 << 19) && g.H <= (1 << 19)) ? 0.9f : -1.0f;                  // fp32 coordinate spacing <= 2^-5 (near_tap)
  memset(&s->dp, 0, sizeof(s->dp));
  s->dp.in.userInputType = -1;
  refresh_derived(s);

  int rc = 0;
  do {
    if ((rc = use_device(s))) break;
    cudaError_t e;
    if ((e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaEventCreate(&s->ev0)) != cudaSuccess || (e = cudaEventCreate(&s->ev1)) != cudaSuccess) {
      rc = fail("wsb_create: %s", cudaGetErrorString(e));
      break;
    }
    if ((e = cudaFuncSetAttribute(k_fused_pvb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem1)) != cudaSuccess ||
        (e = cudaFuncSetAttribute(k_fused_adv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem2)) != cudaSuccess ||
        (e = cudaFuncSetAttribute(k_fused_dry, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemDry)) != cudaSuccess ||
        (e = cudaFuncSetAttribute(k_boxsum, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBox)) != cudaSuccess) {
      rc = fail("wsb_create: kernels not loadable on this device (built for sm_100a): %s", cudaGetErrorString(e));
      break;
    }
    {  // Load every steady-state kernel now.  With CUDA's lazy module loading the FIRST launch of a kernel can
       // synchronise the context — behind a neighbour's spinning k_push_ghosts / k_wait_ghosts that would stall
       // (several strips in one process: deadlock until the bounded wait runs out).
      cudaFuncAttributes fa;
      const void* fns[] = {(const void*)k_fused_pvb, (const void*)k_fused_adv, (const void*)k_fused_dry, (const void*)k_wall_tilemap, (const void*)k_push_ghosts, (const void*)k_ghosts_free, (const void*)k_unpack_zone,
                           (const void*)k_wait_ghosts, (const void*)k_pack_halo, (const void*)k_unpack_halo, (const void*)k_precipitation, (const void*)k_boxsum, (const void*)k_clear_origins,
                           (const void*)k_latch, (const void*)k_texels_to_planes, (const void*)k_planes_to_texels, (const void*)k_pressure_rect,
                           (const void*)k_gather_points};
      for (const void* f : fns)
        if ((e = cudaFuncGetAttributes(&fa, f)) != cudaSuccess) break;
      if (e != cudaSuccess) { rc = fail("wsb_create: kernels not loadable on this device (built for sm_100a): %s", cudaGetErrorString(e)); break; }
    }
    cudaDeviceGetAttribute(&s->n_sms, cudaDevAttrMultiProcessorCount, cfg->device);
    if ((rc = alloc_all(s))) break;
    if ((rc = zero_transients(s))) break;
    if (cfg->n_ranks > 1) {
      int prLo = 0, prHi = 0;  // exchange kernels are tiny: let them cut in front of the queued stencil CTAs
      cudaDeviceGetStreamPriorityRange(&prLo, &prHi);
      if ((e = cudaStreamCreateWithPriority(&s->comm_stream, cudaStreamNonBlocking, prHi)) != cudaSuccess ||
          (e = cudaEventCreateWithFlags(&s->evCompute, cudaEventDisableTiming)) != cudaSuccess ||
          (e = cudaEventCreateWithFlags(&s->evExch, cudaEventDisableTiming)) != cudaSuccess ||
          (e = cudaEventCreateWithFlags(&s->evEdge, cudaEventDisableTiming)) != cudaSuccess ||
          (e = cudaEventCreateWithFlags(&s->evPush, cudaEventDisableTiming)) != cudaSuccess ||
          (e = cudaEventCreateWithFlags(&s->evPvbI, cudaEventDisableTiming)) != cudaSuccess ||
          (e = cudaEventCreateWithFlags(&s->evPvbE, cudaEventDisableTiming)) != cudaSuccess ||
          (e = cudaEventCreateWithFlags(&s->evAdvI, cudaEventDisableTiming)) != cudaSuccess) {
        rc = fail("wsb_create: %s", cudaGetErrorString(e));
        break;
      }
      if (const char* v = getenv("WSB_DBG_NOPUSH")) s->dbg_nopush = atoi(v) != 0;
      if (const char* v = getenv("WSB_DBG_PUSH_BLOCKS")) s->push_blocks = std::max(1, atoi(v));
      if (const char* v = getenv("WSB_DBG_POLL_NS")) {
        const unsigned ns = (unsigned)atoi(v);
        cudaMemcpyToSymbol(g_pollNs, &ns, sizeof ns);
      }
      if (const char* lim = getenv("WSB_SPIN_LIMIT_MS")) {
        const long ms = atol(lim);
        if (ms > 0) s->spin_ns = (unsigned long long)ms * 1000000ull;
      }
      // an all-zero comm_id = no NCCL communicator: the ghost exchange then needs wsb_connect_peers
      bool haveId = false;
      for (int i = 0; i < WSB_COMM_ID_BYTES; i++) haveId |= cfg->comm_id[i] != 0;
      if (haveId) {
        if ((rc = load_nccl())) break;
        Id128 id;
        memcpy(id.internal, cfg->comm_id, WSB_COMM_ID_BYTES);
        int r = g_nccl.CommInitRank(&s->comm, cfg->n_ranks, id, cfg->rank);
        if (r != 0) { rc = fail("ncclCommInitRank failed: %s", g_nccl.GetErrorString(r)); break; }
      }
    }
    cudaError_t e2 = cudaStreamSynchronize(s->stream);
    if (e2 != cudaSuccess) { rc = fail("wsb_create: %s", cudaGetErrorString(e2)); break; }
  } while (0);
  if (rc) { wsb_destroy(s); return rc; }
  *out = s;
  return 0;
}

int wsb_destroy(wsb_sim* s) {
  if (!s) return 0;
  cudaSetDevice(s->cfg.device);
  if (s->stream) cudaStreamSynchronize(s->stream);
  if (s->comm_stream) cudaStreamSynchronize(s->comm_stream);
  if (s->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(s->comm);
  if (s->evCompute) cudaEventDestroy(s->evCompute);
  if (s->evExch) cudaEventDestroy(s->evExch);
  if (s->comm_stream) cudaStreamDestroy(s->comm_stream);
  auto in_arena = [&](const void* q) { return s->arena && (const unsigned char*)q >= s->arena && (const unsigned char*)q < s->arena + s->arena_bytes; };
  auto free_plane = [&](void* q) { if (!in_arena(q)) cudaFree(q); };
  for (int k = 0; k < 2; k++) {
    for (int ch = 0; ch < 4; ch++) { free_plane(s->base[k].p.c[ch]); free_plane(s->water[k].p.c[ch]); free_plane(s->light[k].p.c[ch]); }
    free_plane(s->wall[k]); cudaFree(s->drops[k]);
  }
  if (s->peerL_ipc && s->peerL) cudaIpcCloseMemHandle(s->peerL);
  if (s->peerR_ipc && s->peerR && s->peerR != s->peerL) cudaIpcCloseMemHandle(s->peerR);
  cudaFree(s->arena);
  if (s->evEdge) cudaEventDestroy(s->evEdge);
  if (s->evPush) cudaEventDestroy(s->evPush);
  if (s->evPvbI) cudaEventDestroy(s->evPvbI);
  if (s->evPvbE) cudaEventDestroy(s->evPvbE);
  if (s->evAdvI) cudaEventDestroy(s->evAdvI);
  cudaFree(s->fb); cudaFree(s->dep); cudaFree(s->curl); cudaFree(s->vort);
  cudaFree(s->sg.org4); cudaFree(s->sg.org2); cudaFree(s->sg.dirty); cudaFree(s->sg.dirtyList); cudaFree(s->sg.dirtyCount);
  cudaFree(s->initial_T); cudaFree(s->sndT); cudaFree(s->sndW); cudaFree(s->sndV);
  cudaFree(s->dry_tilewalls);
  cudaFree(s->lightning); cudaFree(s->inactive); cudaFree(s->maxv); cudaFree(s->scratch);
  cudaFree(s->sendL); cudaFree(s->sendR); cudaFree(s->recvL); cudaFree(s->recvR);
  for (cudaEvent_t e : s->ev_pool) cudaEventDestroy(e);
  if (s->ev0) cudaEventDestroy(s->ev0);
  if (s->ev1) cudaEventDestroy(s->ev1);
  if (s->stream) cudaStreamDestroy(s->stream);
  delete s;
  return 0;
}

namespace {
// both ping-pong copies start from the same arrays (app.js:5189-5234): one host->device copy, then
// a device->device duplicate
int finish_upload(wsb_sim* s, const float* drops) {
  const size_t n = cells(s);
  for (int ch = 0; ch < 4; ch++) {
    CK(cudaMemcpyAsync(s->base[1].p.c[ch], s->base[0].p.c[ch], n * 4, cudaMemcpyDeviceToDevice, s->stream));
    CK(cudaMemcpyAsync(s->water[1].p.c[ch], s->water[0].p.c[ch], n * 4, cudaMemcpyDeviceToDevice, s->stream));
  }
  CK(cudaMemcpyAsync(s->wall[1], s->wall[0], n * 4, cudaMemcpyDeviceToDevice, s->stream));
  if (s->ND) {
    CK(cudaMemcpyAsync(s->drops[0], drops, (size_t)s->ND * 20, cudaMemcpyHostToDevice, s->stream));
    CK(cudaMemcpyAsync(s->drops[1], s->drops[0], (size_t)s->ND * 20, cudaMemcpyDeviceToDevice, s->stream));
  }
  if (zero_transients(s)) return 1;
  CK(cudaStreamSynchronize(s->stream));  // the caller may free its arrays as soon as we return
  return 0;
}
}  // namespace

int wsb_upload(wsb_sim* s, const float* base, const float* water, const int8_t* wall, const float* drops) {
  if (!s || !base || !water || !wall) return fail("wsb_upload: null argument");
  if (s->ND > 0 && !drops) return fail("wsb_upload: droplets required (n_droplets = %d)", s->ND);
  if (use_device(s) || join_exchange(s) || join_push(s)) return 1;
  if (upload_texels(s, s->base[0].p, base, true) || upload_texels(s, s->water[0].p, water, true) ||
      upload_field(s, s->wall[0], wall, 4))
    return 1;
  return finish_upload(s, drops);
}

int wsb_upload_local(wsb_sim* s, const float* base, const float* water, const int8_t* wall, const float* drops) {
  if (!s || !base || !water || !wall) return fail("wsb_upload_local: null argument");
  if (s->ND > 0 && !drops) return fail("wsb_upload_local: droplets required (n_droplets = %d)", s->ND);
  if (use_device(s) || join_exchange(s) || join_push(s)) return 1;
  const size_t n = cells(s);
  if (upload_texels(s, s->base[0].p, base, false) || upload_texels(s, s->water[0].p, water, false)) return 1;
  CK(cudaMemcpyAsync(s->wall[0], wall, n * 4, cudaMemcpyHostToDevice, s->stream));
  return finish_upload(s, drops);
}

int wsb_get_layout(wsb_sim* s, int32_t* x_begin, int32_t* local_width, int32_t* ghost) {
  if (!s || !x_begin || !local_width || !ghost) return fail("wsb_get_layout: null argument");
  *x_begin = s->x_begin;
  *local_width = s->lw;
  *ghost = s->ghost;
  return 0;
}

int wsb_set_params(wsb_sim* s, const wsb_params* p) {
  if (!s || !p) return fail("wsb_set_params: null argument");
  s->dp.p = *p;
  return 0;
}

int wsb_set_profiles(wsb_sim* s, const float* initial_T, const float* sT, const float* sW, const float* sV) {
  if (!s) return fail("wsb_set_profiles: null sim");
  if (use_device(s)) return 1;
  const size_t n = ((size_t)s->H + 1) * 4;
  // stream-ordered after the iterations already enqueued; pageable sources are staged before return
  if (initial_T) CK(cudaMemcpyAsync(s->initial_T, initial_T, n, cudaMemcpyHostToDevice, s->stream));
  const float* src[3] = {sT, sW, sV};
  float* dst[3] = {s->sndT, s->sndW, s->sndV};
  for (int k = 0; k < 3; k++) {
    if (src[k]) CK(cudaMemcpyAsync(dst[k], src[k], n, cudaMemcpyHostToDevice, s->stream));
    else CK(cudaMemsetAsync(dst[k], 0, n, s->stream));
  }
  CK(cudaStreamSynchronize(s->stream));
  return 0;
}

int wsb_set_frame_inputs(wsb_sim* s, const wsb_frame_inputs* in) {
  if (!s || !in) return fail("wsb_set_frame_inputs: null argument");
  s->dp.in = *in;
  refresh_derived(s);
  return 0;
}

int wsb_step(wsb_sim* s, int32_t n_iters) {
  if (!s) return fail("wsb_step: null sim");
  if (n_iters < 0) return fail("wsb_step: negative iteration count");
  if (use_device(s)) return 1;
  s->ev_used = 0;
  s->spans.clear();
  CK(cudaEventRecord(s->ev0, s->stream));
  for (int i = 0; i < n_iters; i++) {
    if (s->schedule == WSB_SCHEDULE_REFERENCE ? ref_iteration(s) : fused_iteration(s)) return 1;
  }
  if (join_exchange(s) || join_push(s)) return 1;  // everything enqueued so far is ordered on the compute stream again
  CK(cudaEventRecord(s->ev1, s->stream));
  s->timed = true;
  return 0;
}

int wsb_step_dry(wsb_sim* s, int32_t n_iters) {
  if (!s) return fail("wsb_step_dry: null sim");
  if (n_iters < 0) return fail("wsb_step_dry: negative iteration count");
  if (use_device(s)) return 1;
  s->ev_used = 0;
  s->spans.clear();
  CK(cudaEventRecord(s->ev0, s->stream));
  for (int i = 0; i < n_iters; i++)
    if (dry_iteration(s)) return 1;
  CK(cudaEventRecord(s->ev1, s->stream));
  s->timed = true;
  return 0;
}

int wsb_sync(wsb_sim* s) {
  if (!s) return fail("wsb_sync: null sim");
  if (use_device(s)) return 1;
  CK(cudaStreamSynchronize(s->stream));
  return strip_checks(s);
}

int wsb_set_exchange(wsb_sim* s, int32_t transport) {
  if (!s) return fail("wsb_set_exchange: null sim");
  if (s->cfg.n_ranks <= 1) return fail("wsb_set_exchange: not a strip of a multi-GPU run");
  if (use_device(s) || join_exchange(s) || join_push(s)) return 1;
  CK(cudaStreamSynchronize(s->stream));
  if (transport == WSB_EXCHANGE_PEER || transport == WSB_EXCHANGE_PEER_COMPACT) {
    if (!s->peerL || !s->peerR) return fail("wsb_set_exchange: the peer transport needs wsb_connect_peers first");
    s->peer_mode = true;
    s->peer_compact = transport == WSB_EXCHANGE_PEER_COMPACT;
  } else if (transport == WSB_EXCHANGE_NCCL) {
    if (!s->comm) return fail("wsb_set_exchange: the NCCL transport needs a comm_id at wsb_create");
    s->peer_mode = false;
  } else {
    return fail("wsb_set_exchange: unknown transport %d", transport);
  }
  return 0;
}

int wsb_peer_info(wsb_sim* s, uint8_t out[WSB_PEER_INFO_BYTES]) {
  if (!s || !out) return fail("wsb_peer_info: null argument");
  if (s->cfg.n_ranks <= 1 || !s->arena) return fail("wsb_peer_info: not a strip of a multi-GPU run");
  if (use_device(s)) return 1;
  PeerInfo pi{};
  pi.magic = kPeerMagic;
  pi.pid = (int)getpid();
  pi.device = s->cfg.device;
  pi.rank = s->cfg.rank;
  pi.pitch = s->pitch; pi.lw = s->lw; pi.H = s->H;
  pi.plane_bytes = s->plane_bytes; pi.arena_bytes = s->arena_bytes;
  pi.raw = (unsigned long long)(uintptr_t)s->arena;
  CK(cudaIpcGetMemHandle(&pi.handle, s->arena));
  static_assert(sizeof(PeerInfo) <= WSB_PEER_INFO_BYTES, "PeerInfo must fit the ABI blob");
  memset(out, 0, WSB_PEER_INFO_BYTES);
  memcpy(out, &pi, sizeof(pi));
  return 0;
}

int wsb_connect_peers(wsb_sim* s, const uint8_t* left, const uint8_t* right) {
  if (!s || !left || !right) return fail("wsb_connect_peers: null argument");
  if (s->cfg.n_ranks <= 1 || !s->arena) return fail("wsb_connect_peers: not a strip of a multi-GPU run");
  if (s->peerL || s->peerR) return fail("wsb_connect_peers: already connected");
  if (use_device(s)) return 1;
  PeerInfo L, R;
  memcpy(&L, left, sizeof(L));
  memcpy(&R, right, sizeof(R));
  const int wantL = (s->cfg.rank + s->cfg.n_ranks - 1) % s->cfg.n_ranks, wantR = (s->cfg.rank + 1) % s->cfg.n_ranks;
  if (L.magic != kPeerMagic || R.magic != kPeerMagic) return fail("wsb_connect_peers: not a wsb_peer_info blob");
  if (L.rank != wantL || R.rank != wantR) return fail("wsb_connect_peers: got ranks (%d, %d), neighbours of rank %d are (%d, %d)", L.rank, R.rank, s->cfg.rank, wantL, wantR);
  if (L.H != s->H || R.H != s->H) return fail("wsb_connect_peers: neighbour grids have a different height");
  auto open = [&](const PeerInfo& pi, unsigned char** ptr, bool* ipc) -> int {
    if (pi.pid == (int)getpid()) {  // same process (several sims in one host process): plain pointers
      if (pi.device != s->cfg.device) {
        int can = 0;
        CK(cudaDeviceCanAccessPeer(&can, s->cfg.device, pi.device));
        if (!can) return fail("wsb_connect_peers: device %d cannot access device %d", s->cfg.device, pi.device);
        cudaError_t e = cudaDeviceEnablePeerAccess(pi.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail("cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
        cudaGetLastError();
      }
      *ptr = (unsigned char*)(uintptr_t)pi.raw;
      *ipc = false;
    } else {
      void* q = nullptr;
      CK(cudaIpcOpenMemHandle(&q, pi.handle, cudaIpcMemLazyEnablePeerAccess));
      *ptr = (unsigned char*)q;
      *ipc = true;
    }
    return 0;
  };
  if (open(L, &s->peerL, &s->peerL_ipc)) return 1;
  if (R.rank == L.rank) { s->peerR = s->peerL; s->peerR_ipc = false; }  // two ranks: both neighbours are the same peer
  else if (open(R, &s->peerR, &s->peerR_ipc)) return 1;
  s->pbL = L.plane_bytes; s->pbR = R.plane_bytes;
  s->pitchL = L.pitch; s->pitchR = R.pitch; s->lwL = L.lw;
  s->peer_mode = true;
  return 0;
}

int wsb_debug_run_pass(wsb_sim* s, int32_t pass) {
  if (!s) return fail("wsb_debug_run_pass: null sim");
  s->dry_tilewalls_valid = false;
  if (s->schedule != WSB_SCHEDULE_REFERENCE) return fail("wsb_debug_run_pass needs WSB_SCHEDULE_REFERENCE");
  if (use_device(s)) return 1;
  switch (pass) {
    case WSB_PASS_VELOCITY: return ref_velocity(s);
    case WSB_PASS_CURL: return ref_curl(s);
    case WSB_PASS_VORTICITY: return ref_vorticity(s);
    case WSB_PASS_BOUNDARY: return ref_boundary(s);
    case WSB_PASS_ADVECTION: return ref_advection(s, false);
    case WSB_PASS_PRESSURE: return ref_pressure(s);
    case WSB_PASS_LIGHTING: return ref_lighting(s);
    case WSB_PASS_PRECIPITATION: return clear_feedback(s) || precipitation(s);
    case WSB_PASS_ITER_INC: s->iter++; return 0;
    case WSB_PASS_ADVECTION_DRY: return ref_advection(s, true);
    default: return fail("wsb_debug_run_pass: unknown pass %d", pass);
  }
}

int wsb_read_rect(wsb_sim* s, int32_t field, int32_t view, int32_t x, int32_t y, int32_t w, int32_t h, void* dst) {
  if (!s || !dst) return fail("wsb_read_rect: null argument");
  if (w <= 0 || h <= 0 || x < 0 || y < 0 || x + w > s->W || y + h > s->H) return fail("wsb_read_rect: rectangle (%d,%d,%d,%d) outside %dx%d", x, y, w, h, s->W, s->H);
  if (view < WSB_VIEW_FRAMEBUFF_0 || view > WSB_VIEW_LATEST) return fail("wsb_read_rect: unknown view %d", view);
  if (use_device(s)) return 1;
  const bool fused = s->schedule == WSB_SCHEDULE_FUSED;
  const void* src = nullptr;        // packed array fields
  const Planes4* planes = nullptr;  // channel-plane fields
  size_t elt = 0;
  bool pressure_on_the_fly = false;
  const int v1 = view == WSB_VIEW_FRAMEBUFF_1 ? 1 : 0;
  switch (field) {
    case WSB_FIELD_BASE:
      elt = 16;
      if (fused) { planes = &s->base[1].p; pressure_on_the_fly = (v1 == 0) && s->pressure_pending; }
      else planes = &s->base[v1].p;
      break;
    case WSB_FIELD_WATER:
      elt = 16;
      // FUSED before the first iteration: both copies still hold the upload
      planes = &s->water[v1].p;
      break;
    case WSB_FIELD_WALL:
      elt = 4;
      // frameBuff_0's wall after an iteration is the pressure pass's pass-through of wall_1
      src = fused ? s->wall[1] : s->wall[v1];
      break;
    case WSB_FIELD_LIGHT:
      elt = 16;
      planes = view == WSB_VIEW_LATEST ? &s->light[s->even ? 0 : 1].p : &s->light[v1].p;
      break;
    case WSB_FIELD_FEEDBACK: elt = 16; src = s->fb; break;
    case WSB_FIELD_DEPOSITION: elt = 8; src = s->dep; break;
    case WSB_FIELD_CURL:
      if (!s->curl) return fail("wsb_read_rect: curl is only stored by WSB_SCHEDULE_REFERENCE");
      elt = 4; src = s->curl; break;
    case WSB_FIELD_VORTFORCE:
      if (!s->vort) return fail("wsb_read_rect: vortForce is only stored by WSB_SCHEDULE_REFERENCE");
      elt = 8; src = s->vort; break;
    default: return fail("wsb_read_rect: unknown field %d", field);
  }
  // intersect with the owned strip
  const int gx0 = std::max(x, s->x_begin), gx1 = std::min(x + w, s->x_begin + s->lw);
  if (gx1 <= gx0) { CK(cudaStreamSynchronize(s->stream)); return 0; }
  const int lx0 = gx0 - s->x_begin + s->ghost, cw = gx1 - gx0;
  char* d = (char*)dst + (size_t)(gx0 - x) * elt;
  if (planes) {  // planes -> packed texels in the scratch buffer -> host
    if (need_scratch(s, (size_t)cw * h)) return 1;
    dim3 b(32, 8), gr((cw + 31) / 32, (h + 7) / 8);
    if (pressure_on_the_fly) {
      k_pressure_rect<<<gr, b, 0, s->stream>>>(make_ctx(s, 1, 1, 1, 0), lx0, y, cw, h, s->scratch);
      LAUNCHED("k_pressure_rect");
    } else {
      k_planes_to_texels<<<gr, b, 0, s->stream>>>(*planes, s->pitch, lx0, y, cw, h, s->scratch);
      LAUNCHED("k_planes_to_texels");
    }
    CK(cudaMemcpy2DAsync(d, (size_t)w * elt, s->scratch, (size_t)cw * elt, (size_t)cw * elt, h, cudaMemcpyDeviceToHost, s->stream));
  } else {
    const char* sp = (const char*)src + ((size_t)y * s->pitch + lx0) * elt;
    CK(cudaMemcpy2DAsync(d, (size_t)w * elt, sp, (size_t)s->pitch * elt, (size_t)cw * elt, h, cudaMemcpyDeviceToHost, s->stream));
  }
  CK(cudaStreamSynchronize(s->stream));
  return strip_checks(s);
}

int wsb_read_points(wsb_sim* s, int32_t field, int32_t view, int32_t n, const int32_t* xy, float* dst) {
  if (!s || !xy || !dst) return fail("wsb_read_points: null argument");
  if (n < 0) return fail("wsb_read_points: negative count");
  if (view < WSB_VIEW_FRAMEBUFF_0 || view > WSB_VIEW_LATEST) return fail("wsb_read_points: unknown view %d", view);
  if (n == 0) return 0;
  if (use_device(s)) return 1;
  for (int i = 0; i < n; i++)
    if (xy[2 * i] < 0 || xy[2 * i] >= s->W || xy[2 * i + 1] < 0 || xy[2 * i + 1] >= s->H)
      return fail("wsb_read_points: point %d (%d,%d) outside %dx%d", i, xy[2 * i], xy[2 * i + 1], s->W, s->H);
  const bool fused = s->schedule == WSB_SCHEDULE_FUSED;
  const int v1 = view == WSB_VIEW_FRAMEBUFF_1 ? 1 : 0;
  const Planes4* planes = nullptr;
  bool pressure = false;
  switch (field) {
    case WSB_FIELD_BASE:
      if (fused) { planes = &s->base[1].p; pressure = (v1 == 0) && s->pressure_pending; }
      else planes = &s->base[v1].p;
      break;
    case WSB_FIELD_WATER: planes = &s->water[v1].p; break;
    case WSB_FIELD_LIGHT: planes = view == WSB_VIEW_LATEST ? &s->light[s->even ? 0 : 1].p : &s->light[v1].p; break;
    default: return fail("wsb_read_points: field %d is not an RGBA32F simulation texture", field);
  }
  // scratch: [n float4 results][n int2 points]
  if (need_scratch(s, (size_t)n + ((size_t)n * 8 + 15) / 16)) return 1;
  int* dxy = reinterpret_cast<int*>(s->scratch + n);
  CK(cudaMemcpyAsync(dxy, xy, (size_t)n * 8, cudaMemcpyHostToDevice, s->stream));
  k_gather_points<<<(n + 127) / 128, 128, 0, s->stream>>>(make_ctx(s, 1, 1, 1, 0), *planes, pressure ? 1 : 0, n, dxy,
                                                          s->ghost - s->x_begin, s->scratch);
  LAUNCHED("k_gather_points");
  CK(cudaMemcpyAsync(dst, s->scratch, (size_t)n * 16, cudaMemcpyDeviceToHost, s->stream));
  CK(cudaStreamSynchronize(s->stream));
  return 0;
}

int wsb_read_droplets(wsb_sim* s, int32_t buffer, int32_t first, int32_t count, float* dst) {
  if (!s || !dst) return fail("wsb_read_droplets: null argument");
  if (buffer < 0 || buffer > 2) return fail("wsb_read_droplets: buffer must be 0, 1 or 2");
  if (first < 0 || count < 0 || (long long)first + count > s->ND) return fail("wsb_read_droplets: range [%d, %d) outside %d droplets", first, first + count, s->ND);
  if (use_device(s)) return 1;
  const int b = buffer == 2 ? s->last_drops : buffer;
  if (count) CK(cudaMemcpyAsync(dst, s->drops[b] + (size_t)first * 5, (size_t)count * 20, cudaMemcpyDeviceToHost, s->stream));
  CK(cudaStreamSynchronize(s->stream));
  return 0;
}

int wsb_get_inactive_droplets(wsb_sim* s, float* out) {
  if (!s || !out) return fail("wsb_get_inactive_droplets: null argument");
  if (use_device(s)) return 1;
  CK(cudaMemcpyAsync(out, s->inactive, 4, cudaMemcpyDeviceToHost, s->stream));
  CK(cudaStreamSynchronize(s->stream));
  return 0;
}

int wsb_get_lightning(wsb_sim* s, float out[4]) {
  if (!s || !out) return fail("wsb_get_lightning: null argument");
  if (use_device(s)) return 1;
  CK(cudaMemcpyAsync(out, s->lightning, 16, cudaMemcpyDeviceToHost, s->stream));
  CK(cudaStreamSynchronize(s->stream));
  return 0;
}

int wsb_get_iter(wsb_sim* s, int64_t* out) {
  if (!s || !out) return fail("wsb_get_iter: null argument");
  *out = s->iter;
  return 0;
}

int wsb_set_iter(wsb_sim* s, int64_t iter) {
  if (!s) return fail("wsb_set_iter: null sim");
  if (iter < 0) return fail("wsb_set_iter: negative iteration number");
  s->iter = iter;
  return 0;
}

int wsb_get_strip(wsb_sim* s, int32_t* x_begin, int32_t* local_width) {
  if (!s || !x_begin || !local_width) return fail("wsb_get_strip: null argument");
  *x_begin = s->x_begin;
  *local_width = s->lw;
  return 0;
}

int wsb_get_max_velocity(wsb_sim* s, float* out) {
  if (!s || !out) return fail("wsb_get_max_velocity: null argument");
  if (use_device(s)) return 1;
  unsigned u = 0;
  CK(cudaMemcpyAsync(&u, s->maxv, 4, cudaMemcpyDeviceToHost, s->stream));
  CK(cudaStreamSynchronize(s->stream));
  memcpy(out, &u, 4);
  return 0;
}

int wsb_count_nonfinite(wsb_sim* s, int64_t* out) {
  if (!s || !out) return fail("wsb_count_nonfinite: null argument");
  if (use_device(s)) return 1;
  if (need_scratch(s, 1)) return 1;
  unsigned long long* d = reinterpret_cast<unsigned long long*>(s->scratch);
  CK(cudaMemsetAsync(d, 0, 8, s->stream));
  const int b = s->schedule == WSB_SCHEDULE_FUSED ? 1 : 0;  // the canonical copy of the schedule (ghost columns included)
  k_count_nonfinite<<<s->n_sms * 8, 256, 0, s->stream>>>(s->base[b].p, s->water[b].p, cells(s), d);
  LAUNCHED("k_count_nonfinite");
  unsigned long long v = 0;
  CK(cudaMemcpyAsync(&v, d, 8, cudaMemcpyDeviceToHost, s->stream));
  CK(cudaStreamSynchronize(s->stream));
  *out = (int64_t)v;
  return 0;
}

int wsb_get_launch_count(wsb_sim* s, int64_t* out) {
  if (!s || !out) return fail("wsb_get_launch_count: null argument");
  *out = s->launches;
  return 0;
}

int wsb_set_profiling(wsb_sim* s, int32_t on) {
  if (!s) return fail("wsb_set_profiling: null sim");
  s->profiling = on != 0;
  return 0;
}

int wsb_kernel_time_ms(wsb_sim* s, int32_t kernel, float* total_ms, int32_t* launches) {
  if (!s || !total_ms || !launches) return fail("wsb_kernel_time_ms: null argument");
  if (kernel < 0 || kernel > WSB_KERNEL_SPRITES) return fail("wsb_kernel_time_ms: unknown kernel class %d", kernel);
  if (use_device(s)) return 1;
  CK(cudaStreamSynchronize(s->stream));
  double sum = 0.0;
  int n = 0;
  for (const wsb_sim::Span& sp : s->spans) {
    if (sp.kind != kernel) continue;
    float ms = 0.0f;
    CK(cudaEventElapsedTime(&ms, s->ev_pool[sp.e0], s->ev_pool[sp.e1]));
    sum += ms;
    n++;
  }
  *total_ms = (float)sum;
  *launches = n;
  return 0;
}

int wsb_last_step_ms(wsb_sim* s, float* out) {
  if (!s || !out) return fail("wsb_last_step_ms: null argument");
  if (!s->timed) return fail("wsb_last_step_ms: no wsb_step / wsb_step_dry call yet");
  if (use_device(s)) return 1;
  CK(cudaEventSynchronize(s->ev1));
  CK(cudaEventElapsedTime(out, s->ev0, s->ev1));
  return strip_checks(s);
}

}  // extern "C"
