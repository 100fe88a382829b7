// Stand-alone probe for the tcgen05 / TMA building blocks the 3xTF32 implicit-GEMM kernels rely on:
//   * TMA 2D load of a K-major fp32 tile with SWIZZLE_128B (incl. negative row coordinates -> zero fill)
//   * UMMA shared-memory descriptors for row-shifted views of that tile (base_offset field)
//   * tcgen05.mma kind::tf32 M=128, N=32/64, K=8 steps, fp32 accumulation in TMEM; tcgen05.ld 32x32b readback
//   * what the tensor core does with the low 13 mantissa bits of fp32 operands (truncate vs round)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I lwsnet_b200/csrc -I include \
//        tools/umma_probe.cu lwsnet_b200/csrc/tma_utils.cu -o build/umma_probe
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "tma_utils.cuh"

using namespace lws;

constexpr int A_ROWS = 136, B_ROWS = 64;
constexpr int A_BYTES = A_ROWS * 128, B_BYTES = B_ROWS * 128;

__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr, uint32_t base_off) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;            // LBO (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;  // SBO: 8 rows x 128 B
  d |= (uint64_t)1 << 46;            // descriptor version (Blackwell)
  d |= (uint64_t)(base_off & 7) << 49;
  d |= (uint64_t)2 << 61;  // SWIZZLE_128B
  return d;
}

__global__ void __launch_bounds__(128) umma_probe_kernel(const __grid_constant__ CUtensorMap mapA,
                                                         const __grid_constant__ CUtensorMap mapB, float* D, int row0,
                                                         int shift, int N, int base_off_mode) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = smem + 18432;  // next 1024-aligned offset after 17408
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + 18432 + 8192);
  uint64_t* bar_mma = bar_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_full + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    mbar_init(bar_full, 1);
    mbar_init(bar_mma, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = *tmem_slot;

  if (tid == 0) {
    mbar_expect_tx(bar_full, A_BYTES + B_BYTES);
    tma_load_2d(sA, &mapA, bar_full, 0, row0);
    tma_load_2d(sB, &mapB, bar_full, 0, 0);
    mbar_wait(bar_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t a_addr = smem_u32(sA) + shift * 128, b_addr = smem_u32(sB);
    const uint32_t boff = base_off_mode ? ((a_addr >> 7) & 7) : 0;
    for (int k = 0; k < 4; ++k) {
      const uint64_t da = make_sdesc(a_addr + k * 32, boff), db = make_sdesc(b_addr + k * 32, 0);
      const uint32_t acc = k > 0;
      asm volatile(
          "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
          "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(tmem),
          "l"(da), "l"(db), "r"(idesc), "r"(acc));
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar_mma)));
  }
  mbar_wait(bar_mma, 0);
  asm volatile("tcgen05.fence::after_thread_sync;");
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t r[32];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,"
        "%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;");
    for (int j = 0; j < 32; ++j) D[(warp * 32 + lane) * N + c0 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64));
}

static float tf32_trunc(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u &= 0xFFFFE000u;
  memcpy(&x, &u, 4);
  return x;
}
static float tf32_rn(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u += 0xFFFu + ((u >> 13) & 1);
  u &= 0xFFFFE000u;
  memcpy(&x, &u, 4);
  return x;
}

int main() {
  const int NR = 200;
  std::vector<float> A(NR * 32), B(64 * 32);
  srand(1);
  for (auto& v : A) v = (float)rand() / RAND_MAX * 2.f - 1.f;
  for (auto& v : B) v = (float)rand() / RAND_MAX * 2.f - 1.f;
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4);
  cudaMalloc(&dB, B.size() * 4);
  cudaMalloc(&dD, 128 * 64 * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  CUtensorMap mA, mB;
  uint64_t dimsA[2] = {32, (uint64_t)NR}, dimsB[2] = {32, 64}, str[1] = {128};
  uint32_t boxA[2] = {32, A_ROWS}, boxB[2] = {32, B_ROWS};
  int rc = make_tensor_map_f32(&mA, dA, 2, dimsA, str, boxA, true);
  rc |= make_tensor_map_f32(&mB, dB, 2, dimsB, str, boxB, true);
  if (rc) {
    printf("tensor map encode failed %d\n", rc);
    return 1;
  }
  const size_t smem = 18432 + 8192 + 64;
  cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int fails = 0;
  for (int N : {64, 32})
    for (int row0 : {0, -3})
      for (int mode : {1, 0})
        for (int shift = 0; shift < 8; ++shift) {
          cudaMemset(dD, 0, 128 * 64 * 4);
          umma_probe_kernel<<<1, 128, smem>>>(mA, mB, dD, row0, shift, N, mode);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) {
            printf("N=%d row0=%d mode=%d shift=%d: CUDA error %s\n", N, row0, mode, shift, cudaGetErrorString(e));
            return 2;
          }
          std::vector<float> D(128 * N);
          cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
          double et = 0, er = 0, ef = 0;
          for (int i = 0; i < 128; ++i)
            for (int n = 0; n < N; ++n) {
              double st = 0, sr = 0, sf = 0;
              const int row = row0 + shift + i;
              for (int k = 0; k < 32; ++k) {
                const float a = (row >= 0 && row < NR) ? A[row * 32 + k] : 0.f, b = B[n * 32 + k];
                st += (double)tf32_trunc(a) * tf32_trunc(b);
                sr += (double)tf32_rn(a) * tf32_rn(b);
                sf += (double)a * b;
              }
              const double d = D[i * N + n];
              et = fmax(et, fabs(d - st)), er = fmax(er, fabs(d - sr)), ef = fmax(ef, fabs(d - sf));
            }
          const bool ok = fmin(et, er) < 1e-4;
          if (!ok && mode == 1) ++fails;
          printf("N=%2d row0=%2d base_off_mode=%d shift=%d: max|d-trunc|=%.3e max|d-rn|=%.3e max|d-fp32|=%.3e %s\n", N,
                 row0, mode, shift, et, er, ef, ok ? "OK" : "MISMATCH");
        }
  printf("probe %s (%d failing configs with base_off_mode=1)\n", fails ? "FAILED" : "PASSED", fails);
  return fails ? 3 : 0;
}
