// TMA tile-read throughput vs allocation and box shape (why is the dry sweep bimodal?).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_tiles tma_tiles.cu && ./tma_tiles
// Each CTA loads NP boxes of BW x BH floats (tile + halo, like k_fused_dry), sums them, writes 1 KB.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode;
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
struct Maps { CUtensorMap m[5]; };
template <int NP>
__global__ void __launch_bounds__(256) k_tiles(const __grid_constant__ Maps maps, int bw, int bh, int tx, int ty, float* out,
                                               float* w0, float* w1, float* w2, float* w3, int W, int H) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* s = reinterpret_cast<float*>(smem);
  const int n = bw * bh, ps = (n + 31) / 32 * 32;
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(s + NP * ps);
  const int X0 = blockIdx.x * tx - 4, Y0 = blockIdx.y * ty - 2;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(NP * n * 4) : "memory");
    for (int k = 0; k < NP; k++)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                       smem_addr(s + k * ps)), "l"(&maps.m[k]), "r"(X0), "r"(Y0), "r"(smem_addr(bar)) : "memory");
  }
  unsigned done;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_addr(bar)), "r"(0) : "memory");
  } while (!done);
  float acc = 0.f;
  for (int k = 0; k < NP; k++)
    for (int i = threadIdx.x; i < n; i += 256) acc += s[k * ps + i];
  if (w0) {  // the dry sweep's write stream: 4 planes, tile interior, coalesced rows
    for (int i = threadIdx.x; i < tx * ty; i += 256) {
      const int y = Y0 + 2 + i / tx, x = X0 + 4 + i % tx;
      if (x < W && y < H) {
        const size_t ci = (size_t)y * W + x;
        const float v = s[(i / tx + 2) * bw + i % tx + 4] + acc;
        w0[ci] = v; w1[ci] = v + 1.f; w2[ci] = v + 2.f; w3[ci] = v + 3.f;
      }
    }
  } else
  out[(size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 256 + threadIdx.x] = acc;
}
// the same bytes with coalesced LDG (rows of the tile, one float per thread per step)
template <int NP>
__global__ void __launch_bounds__(256) k_ldg(const float* const* planes, size_t pitch, int W, int H, int bw, int bh, int tx, int ty, float* out) {
  const int X0 = blockIdx.x * tx - 4, Y0 = blockIdx.y * ty - 2, n = bw * bh;
  float acc = 0.f;
  for (int k = 0; k < NP; k++) {
    const float* p = planes[k];
    for (int i = threadIdx.x; i < n; i += 256) {
      int y = Y0 + i / bw, x = X0 + i % bw;
      if (x >= 0 && x < W && y >= 0 && y < H) acc += __ldcg(p + (size_t)y * pitch + x);
    }
  }
  out[(size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 256 + threadIdx.x] = acc;
}
static void make_map(CUtensorMap* m, void* p, int W, int H, int bw, int bh) {
  cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)H}, strides[1] = {(cuuint64_t)W * 4};
  cuuint32_t box[2] = {(cuuint32_t)bw, (cuuint32_t)bh}, es[2] = {1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, p, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
}
int main(int argc, char** argv) {
  void* fn; cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  g_encode = (EncodeTiledFn)fn;
  const int NP = 5;
  const int H = 4096;
  float* out; cudaMalloc(&out, (size_t)64 << 20);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  struct Cfg { int W, bw, bh; const char* alloc; size_t skew; };
  // alloc: "exact" = one cudaMalloc per plane (5 read + 4 write) of W*H*4 bytes; "slab" = one cudaMalloc, plane k at
  // k * (plane + skew)
  std::vector<Cfg> cfgs;
  for (int W : {16384, 16448}) {
    cfgs.push_back({W, 72, 20, "exact", 0});
    cfgs.push_back({W, 72, 20, "slab", 0});
    cfgs.push_back({W, 72, 20, "slab", 4096 + 256});
    cfgs.push_back({W, 72, 20, "slab", 65536 + 4096 + 256});
    cfgs.push_back({W, 72, 20, "slab", (1 << 20) + 65536 + 4096 + 256});
    cfgs.push_back({W, 136, 20, "exact", 0});
    cfgs.push_back({W, 136, 20, "slab", 65536 + 4096 + 256});
  }
  for (int rep = 0; rep < 3; rep++)
    for (auto& c : cfgs) {
      const size_t plane = (size_t)c.W * H * 4;
      std::vector<void*> bases;
      float* p[9];
      std::string a = c.alloc;
      if (a == "exact") {
        for (int k = 0; k < 9; k++) { void* b; cudaMalloc(&b, plane); bases.push_back(b); p[k] = (float*)b; }
        if (rep == 1) { for (int k = 0; k < 4; k++) std::swap(p[k], p[8 - k]); }  // another assignment of blocks to roles
      } else {
        void* b; cudaMalloc(&b, 9 * (plane + c.skew)); bases.push_back(b);
        for (int k = 0; k < 9; k++) p[k] = (float*)((char*)b + k * (plane + c.skew));
      }
      for (int k = 0; k < 9; k++) cudaMemset(p[k], 0, plane);
      Maps maps;
      for (int k = 0; k < NP; k++) make_map(&maps.m[k], p[k], c.W, H, c.bw, c.bh);
      const int tx = c.bw - 8, ty = c.bh - 4;
      dim3 grid((c.W + tx - 1) / tx, (H + ty - 1) / ty);
      const int n = c.bw * c.bh, ps = (n + 31) / 32 * 32;
      size_t smem = (size_t)NP * ps * 4 + 16;
      if (smem < 51 * 1024) smem = 51 * 1024;
      cudaFuncSetAttribute(k_tiles<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      float t[2];
      for (int mode = 0; mode < 2; mode++) {
        float best = 1e9f;
        for (int it = 0; it < 6; it++) {
          cudaEventRecord(e0);
          if (mode == 0) k_tiles<NP><<<grid, 256, smem>>>(maps, c.bw, c.bh, tx, ty, out, nullptr, nullptr, nullptr, nullptr, c.W, H);
          else k_tiles<NP><<<grid, 256, smem>>>(maps, c.bw, c.bh, tx, ty, out, p[5], p[6], p[7], p[8], c.W, H);
          cudaEventRecord(e1);
          cudaEventSynchronize(e1);
          float ms; cudaEventElapsedTime(&ms, e0, e1);
          if (it >= 2) best = ms < best ? ms : best;
        }
        t[mode] = best;
      }
      cudaError_t e = cudaGetLastError();
      const double gb = (double)plane / 1e9;
      printf("rep %d alloc %-5s skew %8zu W %5d box %3dx%-2d: read-only %.3f ms %5.0f GB/s | read 5 + write 4 planes %.3f ms %5.0f GB/s  %s  p0 %p  d(k)/2MiB:",
             rep, c.alloc, c.skew, c.W, c.bw, c.bh, t[0], 5 * gb / (t[0] * 1e-3), t[1], 9 * gb / (t[1] * 1e-3), cudaGetErrorString(e), (void*)p[0]);
      for (int k = 1; k < 9; k++) printf(" %.2f", ((char*)p[k] - (char*)p[0]) / 2097152.0);
      printf("\n");
      for (void* b : bases) cudaFree(b);
    }
  return 0;
}
