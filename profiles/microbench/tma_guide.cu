// The CUDA programming guide's TMA example (libcu++ barrier + cuda::device::experimental), 2D tile load.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <cstdio>
#include <vector>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
constexpr int GW = 256, GH = 64, SW = 64, SH = 20;
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void kernel(const __grid_constant__ CUtensorMap tensor_map, int x, int y, int* out) {
  __shared__ alignas(128) int smem_buffer[SH][SW];
#pragma nv_diag_suppress static_var_with_dynamic_init
  __shared__ barrier bar;
  if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
  __syncthreads();
  barrier::arrival_token token;
  if (threadIdx.x == 0) {
    cde::cp_async_bulk_tensor_2d_global_to_shared(&smem_buffer, &tensor_map, x, y, bar);
    token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(smem_buffer));
  } else {
    token = bar.arrive();
  }
  bar.wait(std::move(token));
  for (int i = threadIdx.x; i < SW * SH; i += blockDim.x) out[i] = smem_buffer[i / SW][i % SW];
}
int main() {
  std::vector<int> h(GW * GH);
  for (int i = 0; i < GW * GH; i++) h[i] = i;
  int *d, *out;
  cudaMalloc(&d, GW * GH * 4); cudaMalloc(&out, SW * SH * 4);
  cudaMemcpy(d, h.data(), GW * GH * 4, cudaMemcpyHostToDevice);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  CUtensorMap m;
  cuuint64_t size[2] = {GW, GH}, stride[1] = {GW * sizeof(int)};
  cuuint32_t box[2] = {SW, SH}, es[2] = {1, 1};
  CUresult r = ((EncodeTiledFn)fn)(&m, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, d, size, stride, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  { const unsigned* w = (const unsigned*)&m; printf("ptr %p desc:", (void*)d); for (int i = 0; i < 32; i++) printf(" %08x", w[i]); printf("\n"); }
  kernel<<<1, 128>>>(m, XOFF, 14, out);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<int> o(SW * SH);
  cudaMemcpy(o.data(), out, SW * SH * 4, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int j = 0; j < SH; j++) for (int i = 0; i < SW; i++) if (o[j * SW + i] != (14 + j) * GW + XOFF + i) bad++;
  printf("guide sample: encode=%d q=%d %s mismatches=%d\n", (int)r, (int)q, cudaGetErrorString(e), bad);
  return 0;
}
