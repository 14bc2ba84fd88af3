// cuda_emu.h — TEST INFRASTRUCTURE ONLY: just enough of the CUDA execution model on the host to run
// the fused kernels of csrc/wsb_fused_kernels.cuh unchanged (compiled with -DWSB_HOST_EMU).
//
//   * one OS thread per CUDA thread of a CTA; CTAs of a launch run one after the other;
//   * __syncthreads = a std::barrier over the CTA; warp collectives (__reduce_max_sync, __any_sync)
//     = a std::barrier over the warp plus a 32-slot scratch array;
//   * dynamic shared memory = one buffer per launch, refilled with a garbage pattern before every
//     CTA (reads of never-written shared memory do not go unnoticed);
//   * TMA: a CUtensorMap is {plane, W, H, boxW, boxH}; tma_load_box copies the box synchronously,
//     zero-filling coordinates outside the plane (the behaviour measured on the device,
//     profiles/microbench/README.md), and completes the mbarrier's byte count;
//   * mbarrier: phase bit + pending byte count in the 8-byte word.
#pragma once
#include <cuda_runtime.h>  // vector types; qualifiers vanish under g++

#include <algorithm>
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

using std::max;
using std::min;

#ifndef __noinline__
#define __noinline__ __attribute__((noinline))
#endif
#ifndef __launch_bounds__
#define __launch_bounds__(...)
#endif
#ifndef __grid_constant__
#define __grid_constant__
#endif

struct CUtensorMap {  // emulated descriptor of one [H][W] plane of 4-byte elements
  const void* base;
  int W, H, boxW, boxH;
  unsigned char pad[128 - sizeof(void*) - 4 * sizeof(int)];
};

namespace emu {
struct Warp {
  std::barrier<> bar{32};
  unsigned slot[32];
};
struct Cta {
  explicit Cta(int nthreads) : bar(nthreads), warps((nthreads + 31) / 32) {}
  std::barrier<> bar;
  std::vector<Warp> warps;
  std::vector<unsigned char> smem;
};
struct Tls {
  uint3 tid, bid;
  dim3 bdim, gdim;
  Cta* cta;
};
inline thread_local Tls tls;

inline unsigned char* dyn_smem() {
  // 128-byte aligned, like extern __shared__ __align__(128)
  auto p = reinterpret_cast<uintptr_t>(tls.cta->smem.data());
  return reinterpret_cast<unsigned char*>((p + 127) & ~uintptr_t(127));
}

// mbarrier word: bit 0 = phase, bits 1.. = pending bytes
inline std::atomic<unsigned long long>& word(unsigned long long* bar) { return *reinterpret_cast<std::atomic<unsigned long long>*>(bar); }
inline void mbar_init(unsigned long long* bar, unsigned) { word(bar).store(0); }
inline void mbar_expect_tx(unsigned long long* bar, unsigned bytes) { word(bar).fetch_add((unsigned long long)bytes << 1); }
inline void mbar_wait(unsigned long long* bar, unsigned parity) {
  while ((word(bar).load() & 1ull) == parity) std::this_thread::yield();
}
inline void tma_load_box(void* dst, const CUtensorMap* m, int x, int y, unsigned long long* bar) {
  uint32_t* d = static_cast<uint32_t*>(dst);
  const uint32_t* src = static_cast<const uint32_t*>(m->base);
  for (int j = 0; j < m->boxH; j++)
    for (int i = 0; i < m->boxW; i++) {
      const int xx = x + i, yy = y + j;
      d[j * m->boxW + i] = (xx >= 0 && xx < m->W && yy >= 0 && yy < m->H) ? src[(size_t)yy * m->W + xx] : 0u;
    }
  const unsigned long long bytes = (unsigned long long)m->boxW * m->boxH * 4;
  const unsigned long long after = word(bar).fetch_sub(bytes << 1) - (bytes << 1);
  if ((after >> 1) == 0) word(bar).fetch_xor(1ull);  // all expected bytes have arrived: the phase completes
}

// run `kernel()` as a grid of CTAs with `nthreads` threads each
inline void launch(dim3 grid, int nthreads, size_t smem_bytes, const std::function<void()>& kernel) {
  Cta cta(nthreads);
  cta.smem.assign(smem_bytes + 256, 0);
  std::vector<std::thread> pool;
  for (int t = 0; t < nthreads; t++)
    pool.emplace_back([&, t] {
      tls.tid = make_uint3(t, 0, 0);
      tls.bdim = dim3(nthreads, 1, 1);
      tls.gdim = grid;
      tls.cta = &cta;
      for (unsigned by = 0; by < grid.y; by++)
        for (unsigned bx = 0; bx < grid.x; bx++) {
          tls.bid = make_uint3(bx, by, 0);
          if (t == 0) memset(cta.smem.data(), 0xCD, cta.smem.size());  // NaN-ish garbage, never zeros
          cta.bar.arrive_and_wait();
          kernel();
          cta.bar.arrive_and_wait();
        }
    });
  for (auto& th : pool) th.join();
}
}  // namespace emu

#define threadIdx (emu::tls.tid)
#define blockIdx (emu::tls.bid)
#define blockDim (emu::tls.bdim)
#define gridDim (emu::tls.gdim)

static inline void __syncthreads() { emu::tls.cta->bar.arrive_and_wait(); }
// block-wide OR: the warp slots of warp 0 double as the CTA's scratch word
namespace emu { inline std::atomic<int> cta_or{0}; }
static inline int __syncthreads_or(int pred) {
  emu::tls.cta->bar.arrive_and_wait();
  if (emu::tls.tid.x == 0) emu::cta_or.store(0);
  emu::tls.cta->bar.arrive_and_wait();
  if (pred) emu::cta_or.store(1);
  emu::tls.cta->bar.arrive_and_wait();
  const int r = emu::cta_or.load();
  emu::tls.cta->bar.arrive_and_wait();
  return r;
}
static inline unsigned __activemask() { return 0xffffffffu; }
static inline int __ffs(unsigned v) { return v ? __builtin_ctz(v) + 1 : 0; }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline unsigned __ldcg(const unsigned* p) { return reinterpret_cast<const std::atomic<unsigned>*>(p)->load(); }
static inline unsigned atomicMax(unsigned* p, unsigned v) {
  auto& a = *reinterpret_cast<std::atomic<unsigned>*>(p);
  unsigned old = a.load();
  while (old < v && !a.compare_exchange_weak(old, v)) {}
  return old;
}
// static __shared__ variables: CTAs run one after the other, so one static object per declaration is "the CTA's"
#undef __shared__
#define __shared__ static
// vector / integer atomics on global memory: one lock (the order of the adds is as unspecified as on the device)
namespace emu { inline std::atomic_flag add_lock = ATOMIC_FLAG_INIT; }
template <class F> static inline void emu_locked(F&& f) {
  while (emu::add_lock.test_and_set(std::memory_order_acquire)) std::this_thread::yield();
  f();
  emu::add_lock.clear(std::memory_order_release);
}
static inline void atomicAdd(float4* p, float4 v) { emu_locked([&] { p->x += v.x; p->y += v.y; p->z += v.z; p->w += v.w; }); }
static inline void atomicAdd(float2* p, float2 v) { emu_locked([&] { p->x += v.x; p->y += v.y; }); }
static inline int atomicOr(int* p, int v) { int old = 0; emu_locked([&] { old = *p; *p |= v; }); return old; }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { unsigned old = 0; emu_locked([&] { old = *p; *p += v; }); return old; }
static inline int atomicAdd(int* p, int v) { int old = 0; emu_locked([&] { old = *p; *p += v; }); return old; }
// full-warp collectives (the kernels only call them with every lane of the warp present)
static inline unsigned __reduce_max_sync(unsigned, unsigned v) {
  emu::Warp& w = emu::tls.cta->warps[emu::tls.tid.x / 32];
  w.slot[emu::tls.tid.x % 32] = v;
  w.bar.arrive_and_wait();
  unsigned r = 0;
  for (int i = 0; i < 32; i++) r = std::max(r, w.slot[i]);
  w.bar.arrive_and_wait();
  return r;
}
static inline int __any_sync(unsigned, int pred) { return __reduce_max_sync(0xffffffffu, pred ? 1u : 0u) != 0u; }
template <class T> static inline T __shfl_sync(unsigned, T v, int src) {
  static_assert(sizeof(T) == 4, "32-bit shuffles only");
  emu::Warp& w = emu::tls.cta->warps[emu::tls.tid.x / 32];
  memcpy(&w.slot[emu::tls.tid.x % 32], &v, 4);
  w.bar.arrive_and_wait();
  T r;
  memcpy(&r, &w.slot[src & 31], 4);
  w.bar.arrive_and_wait();
  return r;
}
static inline unsigned __ballot_sync(unsigned, int pred) {
  emu::Warp& w = emu::tls.cta->warps[emu::tls.tid.x / 32];
  w.slot[emu::tls.tid.x % 32] = pred ? 1u : 0u;
  w.bar.arrive_and_wait();
  unsigned r = 0;
  for (int i = 0; i < 32; i++) r |= w.slot[i] << i;
  w.bar.arrive_and_wait();
  return r;
}
