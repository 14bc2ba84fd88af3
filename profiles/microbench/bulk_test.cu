// cp.async.bulk (non-tensor, UBLKCP) sanity check: copy 4 KB global -> shared through an mbarrier.
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void k(const float* src, float* out) {
  __shared__ alignas(128) float tile[1024];
  __shared__ unsigned long long bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(&bar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&bar)), "r"(4096) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(tile)), "l"(src), "r"(4096), "r"(smem_addr(&bar)) : "memory");
  }
  unsigned done;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_addr(&bar)), "r"(0) : "memory");
  } while (!done);
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) out[i] = tile[i];
}
int main() {
  std::vector<float> h(1024);
  for (int i = 0; i < 1024; i++) h[i] = i;
  float *d, *out;
  cudaMalloc(&d, 4096); cudaMalloc(&out, 4096);
  cudaMemcpy(d, h.data(), 4096, cudaMemcpyHostToDevice);
  k<<<1, 128>>>(d, out);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<float> o(1024);
  cudaMemcpy(o.data(), out, 4096, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int i = 0; i < 1024; i++) bad += o[i] != (float)i;
  printf("bulk copy: %s mismatches=%d\n", cudaGetErrorString(e), bad);
  return 0;
}
