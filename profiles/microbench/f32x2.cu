// Issue-rate microbenchmark: scalar FADD/FMUL/FFMA vs packed FADD2/FMUL2/FFMA2 on sm_100a.
// Each thread runs ILP independent dependency chains; 8 warps per SMSP; reports warp-instructions / clk / SM.
#include <cuda_runtime.h>
#include <cstdio>
template <int MODE>
__global__ void k(float* out, int iters, float a, float b) {
  float2 x[8];
#pragma unroll
  for (int i = 0; i < 8; i++) x[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f + i);
  const float2 A = make_float2(a, a), B = make_float2(b, b);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (MODE == 0) { x[i].x = x[i].x + a; x[i].y = x[i].y + a; }              // 2 FADD
      if (MODE == 1) { x[i] = __fadd2_rn(x[i], A); }                            // 1 FADD2
      if (MODE == 2) { x[i].x = __fmaf_rn(x[i].x, a, b); x[i].y = __fmaf_rn(x[i].y, a, b); }  // 2 FFMA
      if (MODE == 3) { x[i] = __ffma2_rn(x[i], A, B); }                         // 1 FFMA2
      if (MODE == 4) { x[i].x = x[i].x * a; x[i].y = x[i].y * a; }              // 2 FMUL
      if (MODE == 5) { x[i] = __fmul2_rn(x[i], A); }                            // 1 FMUL2
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += x[i].x + x[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name, int lane_ops_per_iter) {
  float* out;
  cudaMalloc(&out, 148 * 4 * 1024 * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  k<MODE><<<148 * 4, 256>>>(out, 100, 1.0001f, 0.5f);
  cudaEventRecord(e0);
  k<MODE><<<148 * 4, 256>>>(out, iters, 1.0001f, 0.5f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  double laneops = (double)148 * 4 * 256 * iters * lane_ops_per_iter;
  printf("%-8s %8.3f ms  %7.2f T lane-op/s  (%.1f lane-ops/clk/SM at 1.965 GHz)\n", name, ms, laneops / ms / 1e9, laneops / (ms * 1e-3) / 148 / 1.965e9);
  cudaFree(out);
}
int main() {
  run<0>("FADD", 16); run<1>("FADD2", 16); run<2>("FFMA", 16); run<3>("FFMA2", 16); run<4>("FMUL", 16); run<5>("FMUL2", 16);
  return 0;
}
