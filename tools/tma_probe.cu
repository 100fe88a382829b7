// Probe: which (box width, start coordinate) combinations of a rank-3 fp32 non-swizzled tiled TMA load work.  nvcc -arch=sm_100a -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
__global__ void k(const __grid_constant__ CUtensorMap m, float* out, int x0, int y0, int n) {
  extern __shared__ __align__(1024) uint8_t sm[];
  uint64_t* bar = (uint64_t*)(sm + 8192);
  uint32_t b = (uint32_t)__cvta_generic_to_shared(bar), d = (uint32_t)__cvta_generic_to_shared(sm);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(n * 4) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(d),
                 "l"((uint64_t)&m), "r"(b), "r"(x0), "r"(y0), "r"(0)
                 : "memory");
    uint32_t ok = 0;
    while (!ok) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p;}" : "=r"(ok) : "r"(b) : "memory");
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = ((float*)sm)[i];
}
int main() {
  const int W = 128, H = 16, C = 3;
  std::vector<float> h(W * H * C);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
  float *d, *o;
  cudaMalloc(&d, h.size() * 4); cudaMalloc(&o, 8192);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  for (int bw : {136, 132}) for (int x0 : {0, -20, 100, -16, 124}) {
    CUtensorMap m;
    cuuint64_t gd[3] = {W, H, C}, gs[2] = {W * 4, (cuuint64_t)W * H * 4};
    cuuint32_t bd[3] = {(cuuint32_t)bw, 1, C}, es[3] = {1, 1, 1};
    CUresult r = cuTensorMapEncodeTiled(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, gd, gs, bd, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("box %d x0 %d: encode failed %d\n", bw, x0, (int)r); continue; }
    k<<<1, 128, 9216>>>(m, o, x0, -1 + (x0 == 100 ? 3 : 0), bw * C);
    cudaError_t e = cudaDeviceSynchronize();
    float res[8];
    if (e == cudaSuccess) cudaMemcpy(res, o, 32, cudaMemcpyDeviceToHost);
    printf("box %d x0 %d: %s", bw, x0, cudaGetErrorString(e));
    if (e == cudaSuccess) printf("  first: %g %g %g", res[0], res[1], res[2]);
    printf("\n");
    if (e != cudaSuccess) return 1;
  }
  return 0;
}
