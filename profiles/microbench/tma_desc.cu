// Does the number of distinct TMA descriptors a kernel uses matter?  NP planes of one slab are read
// tile by tile (72 x 20 boxes) either through NP separate 2D tensor maps, or through ONE 3D map
// (W, H, plane) with depth-1 boxes, or ONE 3D map with a single depth-NP box.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tma_desc tma_desc.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode;
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
constexpr int MAXP = 12, BW = 72, BH = 20, N = BW * BH;
struct Maps { CUtensorMap m[MAXP]; };
// mode 0: np 2D maps; mode 1: one 3D map, np depth-1 ops; mode 2: one 3D map (depth np), one op
__global__ void __launch_bounds__(256) k(const __grid_constant__ Maps maps, int mode, int np, int spin, float* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* s = reinterpret_cast<float*>(smem);
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(s + MAXP * N);
  const int X0 = blockIdx.x * 64 - 4, Y0 = blockIdx.y * 16 - 2;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(np * N * 4) : "memory");
    if (mode == 0) {
      for (int k = 0; k < np; k++)
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                         smem_addr(s + k * N)), "l"(&maps.m[k]), "r"(X0), "r"(Y0), "r"(smem_addr(bar)) : "memory");
    } else if (mode == 1) {
      for (int k = 0; k < np; k++)
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                         smem_addr(s + k * N)), "l"(&maps.m[0]), "r"(X0), "r"(Y0), "r"(k), "r"(smem_addr(bar)) : "memory");
    } else {
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                       smem_addr(s)), "l"(&maps.m[1]), "r"(X0), "r"(Y0), "r"(0), "r"(smem_addr(bar)) : "memory");
    }
  }
  unsigned done;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_addr(bar)), "r"(0) : "memory");
  } while (!done);
  float acc = 0.f;
  for (int k = 0; k < np; k++)
    for (int i = threadIdx.x; i < N; i += 256) acc += s[k * N + i];
  for (int i = 0; i < spin; i++) acc = acc * 1.0001f + 0.5f;  // stand-in for the stencil arithmetic
  out[(size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 256 + threadIdx.x] = acc;
}
static void enc(CUtensorMap* m, void* p, int rank, int W, int H, int P, size_t planeBytes, int depth) {
  cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)P}, strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)planeBytes};
  cuuint32_t box[3] = {BW, BH, (cuuint32_t)depth}, es[3] = {1, 1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, p, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d (rank %d depth %d)\n", (int)r, rank, depth); exit(1); }
}
int main() {
  void* fn; cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  g_encode = (EncodeTiledFn)fn;
  const int W = 16384, H = 4096;
  const size_t plane = (size_t)W * H * 4;
  char* slab; cudaMalloc(&slab, MAXP * plane); cudaMemset(slab, 0, MAXP * plane);
  float* out; cudaMalloc(&out, (size_t)64 << 20);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const size_t smem = (size_t)MAXP * N * 4 + 16;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int rep = 0; rep < 2; rep++)
    for (int np : {5, 9, 12})
      for (int spin : {0, 600})
        for (int mode = 0; mode < 3; mode++) {
          Maps maps;
          if (mode == 0) for (int k = 0; k < np; k++) enc(&maps.m[k], slab + k * plane, 2, W, H, 1, plane, 1);
          else { enc(&maps.m[0], slab, 3, W, H, MAXP, plane, 1); enc(&maps.m[1], slab, 3, W, H, MAXP, plane, np); }
          float best = 1e9f, worst = 0;
          for (int it = 0; it < 7; it++) {
            cudaEventRecord(e0);
            k<<<dim3(W / 64, H / 16), 256, smem>>>(maps, mode, np, spin, out);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (it >= 2) { best = ms < best ? ms : best; worst = ms > worst ? ms : worst; }
          }
          printf("rep %d planes %2d spin %3d mode %d (%s): %.3f..%.3f ms  %5.0f GB/s  %s\n", rep, np, spin, mode,
                 mode == 0 ? "np 2D maps      " : mode == 1 ? "one 3D map, np ops" : "one 3D map, 1 op ", best, worst, np * plane / 1e9 / (best * 1e-3),
                 cudaGetErrorString(cudaGetLastError()));
        }
  return 0;
}
