// push_stride.cu — does the ROW STRIDE decide how long the ghost push takes?  (profiles/r3_multi_gpu.md: at 4 GPUs,
// strips of 4096 columns, the same 3.4 MB push that is harmless at 2 and 8 GPUs takes several times longer.)
// Replays the access pattern of k_push_ghosts on ONE GPU: 13 planes of [H][pitch] floats; for every row, the columns
// [8,16) and [lw, lw+8) of a source plane are copied as two 16-byte pieces to the columns [8+lw,16+lw) and [0,8) of a
// destination plane of the same shape; 16 CTAs of 256 threads, each a contiguous run of (plane, row) items.
//   nvcc -O3 -arch=sm_100a -o push_stride push_stride.cu && ./push_stride
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) k_push(const int* __restrict__ src, int* __restrict__ dst, size_t planeElems, int pitch, int lw, int H, int n) {
  const int per = H * 2, total = per * n, chunk = (total + gridDim.x - 1) / gridDim.x;
  const int tEnd = min(total, ((int)blockIdx.x + 1) * chunk);
  for (int t = blockIdx.x * chunk + threadIdx.x; t < tEnd; t += blockDim.x) {
    const int k = t / per, r = t - k * per, y = r >> 1, q = (r & 1) * 4;
    const int* s = src + (size_t)k * planeElems + (size_t)y * pitch;
    int* d = dst + (size_t)k * planeElems + (size_t)y * pitch;
    *reinterpret_cast<int4*>(d + 8 + lw + q) = *reinterpret_cast<const int4*>(s + 8 + q);
    *reinterpret_cast<int4*>(d + q) = *reinterpret_cast<const int4*>(s + lw + q);
  }
}
// a bandwidth-bound bystander, like the advection kernel next to the push
__global__ void k_stream(const float4* __restrict__ a, float4* __restrict__ b, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}

int main() {
  const int H = 4096, n = 13;
  cudaStream_t s1, s2;
  cudaStreamCreate(&s1);
  cudaStreamCreateWithPriority(&s2, cudaStreamNonBlocking, -1);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const size_t nb = (size_t)1 << 28;  // 4 GiB of float4 for the bystander
  float4 *ba, *bb;
  cudaMalloc(&ba, nb * 4);
  cudaMalloc(&bb, nb * 4);
  cudaMemset(ba, 0, nb * 4);
  printf("%8s %8s %10s | %12s %14s %16s\n", "lw", "pad", "stride B", "push alone us", "stream alone ms", "stream+push ms");
  const int lws[] = {2048, 4096, 8192};
  const int pads[] = {0, 4, 16, 32, 64};
  for (int lw : lws)
    for (int pad : pads) {
      const int pitch = lw + 16 + pad;
      const size_t planeElems = ((size_t)pitch * H + 63) / 64 * 64;
      int *src, *dst;
      cudaMalloc(&src, planeElems * n * 4);
      cudaMalloc(&dst, planeElems * n * 4);
      cudaMemset(src, 1, planeElems * n * 4);
      float push_us = 0, alone_ms = 0, both_ms = 0;
      for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0, s1);
        for (int i = 0; i < 50; i++) k_push<<<16, 256, 0, s1>>>(src, dst, planeElems, pitch, lw, H, n);
        cudaEventRecord(e1, s1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&push_us, e0, e1);
        push_us = push_us / 50 * 1000;
      }
      for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(e0, s1);
        for (int i = 0; i < 10; i++) k_stream<<<148 * 8, 256, 0, s1>>>(ba, bb, nb / 4);
        cudaEventRecord(e1, s1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&alone_ms, e0, e1);
        alone_ms /= 10;
      }
      for (int rep = 0; rep < 2; rep++) {
        cudaDeviceSynchronize();
        cudaEventRecord(e0, s1);
        for (int i = 0; i < 10; i++) {
          k_stream<<<148 * 8, 256, 0, s1>>>(ba, bb, nb / 4);
          for (int j = 0; j < 4; j++) k_push<<<16, 256, 0, s2>>>(src, dst, planeElems, pitch, lw, H, n);
        }
        cudaEventRecord(e1, s1);
        cudaDeviceSynchronize();
        cudaEventElapsedTime(&both_ms, e0, e1);
        both_ms /= 10;
      }
      printf("%8d %8d %10d | %12.1f %14.4f %16.4f\n", lw, pad, pitch * 4, push_us, alone_ms, both_ms);
      cudaFree(src);
      cudaFree(dst);
    }
  return 0;
}
