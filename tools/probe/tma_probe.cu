// Development probe: minimal 3-D TMA load (same descriptor shape as K1c) to separate
// descriptor problems from kernel problems.  nvcc -arch=sm_100a tma_probe.cu -o tma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int BOX>
__global__ void probe(const __grid_constant__ CUtensorMap map, float* out, int x, int y, int k) {
  __shared__ alignas(128) float tile[BOX];
  __shared__ alignas(8) unsigned long long mbar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)), "r"((uint32_t)(BOX * 4)) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(tile)), "l"(&map), "r"(smem_u32(&mbar)), "r"(x), "r"(y), "r"(k) : "memory");
  }
  uint32_t done = 0;
  while (!done)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.b32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(&mbar)) : "memory");
  for (int i = threadIdx.x; i < BOX; i += blockDim.x) out[i] = tile[i];
}
int main() {
  const int pitch = 1024, rows = 16, planes = 9, BOX = 256;
  std::vector<float> h((size_t)pitch * rows * planes);
  for (size_t i = 0; i < h.size(); i++) h[i] = (float)i;
  float *d, *o;
  CK(cudaMalloc(&d, h.size() * 4)); CK(cudaMalloc(&o, BOX * 4));
  CK(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  alignas(64) CUtensorMap map;
  cuuint64_t dims[3] = {(cuuint64_t)pitch, (cuuint64_t)rows, (cuuint64_t)planes};
  cuuint64_t strides[2] = {(cuuint64_t)pitch * 4, (cuuint64_t)pitch * rows * 4};
  cuuint32_t box[3] = {BOX, 1, 1}, es[3] = {1, 1, 1};
  CUresult r = ((EncodeFn)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode -> %d\n", (int)r);
  int tests[6][3] = {{0, 0, 0}, {4, 3, 2}, {-4, 3, 2}, {900, 15, 8}, {1, 7, 4}, {-1, 0, 0}};
  for (auto& t : tests) {
    probe<BOX><<<1, 128>>>(map, o, t[0], t[1], t[2]);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> res(BOX);
    if (e == cudaSuccess) CK(cudaMemcpy(res.data(), o, BOX * 4, cudaMemcpyDeviceToHost));
    const double base = (double)t[2] * pitch * rows + (double)t[1] * pitch + t[0];
    printf("coord (%d,%d,%d): %s  got[0..2] = %.0f %.0f %.0f  expect %.0f %.0f %.0f  last %.0f\n", t[0], t[1], t[2], cudaGetErrorString(e),
           res[0], res[1], res[2], base, base + 1, base + 2, res[BOX - 1]);
    if (e != cudaSuccess) { printf("(stopping: the context is dead after a fault)\n"); return 0; }
  }
  return 0;
}
