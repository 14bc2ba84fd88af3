// Standalone check of the TMA plane-box load used by the fused kernels.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
template <int MODE>
__global__ void k(const __grid_constant__ CUtensorMap pmap, const CUtensorMap* gmap, float* out, int bw, int bh, int x, int y) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* tile = (float*)smem;
  unsigned long long* bar = (unsigned long long*)(smem + 8192);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bw * bh * 4) : "memory");
    const CUtensorMap* m = MODE == 0 ? &pmap : gmap;
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     smem_addr(tile)), "l"(m), "r"(x), "r"(y), "r"(smem_addr(bar)) : "memory");
  }
  unsigned done;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_addr(bar)), "r"(0) : "memory");
  } while (!done);
  for (int i = threadIdx.x; i < bw * bh; i += blockDim.x) out[i] = tile[i];
}
int main(int argc, char** argv) {
  const int only_mode = argc > 1 ? atoi(argv[1]) : -1;
  const int W = 256, H = 64;
  for (int bw : {64, 68, 72}) {
    const int bh = 20;
    std::vector<float> h(W * H);
    for (int i = 0; i < W * H; i++) h[i] = (float)i;
    float *d, *out;
    cudaMalloc(&d, W * H * 4); cudaMalloc(&out, 8192);
    cudaMemcpy(d, h.data(), W * H * 4, cudaMemcpyHostToDevice);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    CUtensorMap m;
    cuuint64_t dims[2] = {W, H}, strides[1] = {W * 4};
    cuuint32_t box[2] = {(cuuint32_t)bw, (cuuint32_t)bh}, es[2] = {1, 1};
    CUresult r = ((EncodeTiledFn)fn)(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUtensorMap* gm; cudaMalloc(&gm, sizeof(m)); cudaMemcpy(gm, &m, sizeof(m), cudaMemcpyHostToDevice);
    for (int mode = 0; mode < 2; mode++) {
      if (only_mode >= 0 && mode != only_mode) continue;
      cudaMemset(out, 0, 8192);
      if (mode == 0) k<0><<<1, 256, 8192 + 64>>>(m, gm, out, bw, bh, 62, 14);
      else k<1><<<1, 256, 8192 + 64>>>(m, gm, out, bw, bh, 62, 14);
      cudaError_t e = cudaDeviceSynchronize();
      std::vector<float> o(bw * bh);
      cudaMemcpy(o.data(), out, bw * bh * 4, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int j = 0; j < bh; j++) for (int i = 0; i < bw; i++) if (o[j * bw + i] != (float)((14 + j) * W + 62 + i)) bad++;
      printf("box %dx%d encode=%d mode=%s: %s, mismatches %d\n", bw, bh, (int)r, mode ? "global-desc" : "param-desc", cudaGetErrorString(e), bad);
      if (e != cudaSuccess) return 1;
    }
  }
  return 0;
}
