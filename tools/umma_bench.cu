// Stand-alone tcgen05 probe #2: (a) issue-rate of small-N tcgen05.mma with both operands in shared memory, per kind
// (tf32 / f16), N and layout -- the question is whether the A-tile read bounds small-N MMAs; (b) functional checks of the
// two operand formats the split-fp16 kernels rely on:
//   * kind::f16, SWIZZLE_128B, K-major rows of 64 halves = [32 hi | 32 lo], K=16 steps 32 bytes apart inside the atom;
//   * kind::f16, SWIZZLE_NONE K-major with LBO = t*16 bytes and SBO = 128 bytes, i.e. *overlapping* rows: K-chunk j of
//     row r is the 16-byte voxel r + t*j -- an im2col view of a channels-last (8 halves per voxel) line for free.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 tools/umma_bench.cu -o build/umma_bench
#include <cuda_fp16.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_LOOP:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra WAIT_DONE;\nbra WAIT_LOOP;\n"
      "WAIT_DONE:\n}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// layout: 0 = SWIZZLE_NONE, 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
template <int KIND>  // 0 = f16, 2 = tf32
__device__ __forceinline__ void mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  if (KIND == 2)
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_d),
                 "l"(da), "l"(db), "r"(idesc), "r"(acc)
                 : "memory");
  else
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_d),
                 "l"(da), "l"(db), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred = 0;
  asm volatile("{\n.reg .b32 rx;\n.reg .pred px;\nelect.sync rx|px, 0xffffffff;\n@px mov.s32 %0, 1;\n}\n" : "+r"(pred));
  return pred != 0;
}
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

struct BenchCfg {
  int kind, N, layout, a_slots, nmma, two_acc;
  int a_row_shift, b_half;  // A window shifted by rows (descriptor base not 1024-aligned); B block in the second 64 B of the rows
};

// ---- (a) issue-rate ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) rate_kernel(BenchCfg c, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;               // 8 x 16 KB
  uint8_t* sB = smem + 8 * 16384;   // 32 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 8 * 16384 + 32768);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (8 * 16384 + 32768) / 16; i += 128) reinterpret_cast<float4*>(smem)[i] = make_float4(0, 0, 0, 0);
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  if (warp == 1) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)c.kind << 7) | ((uint32_t)c.kind << 10) | ((uint32_t)(c.N >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t lbo = c.layout == 0 ? 128 : 16, sbo = c.layout == 0 ? 256 : 1024;
    const uint64_t hi = make_desc(0, lbo, sbo, c.layout);
    const uint32_t a_lo = ((smem_u32(sA) & 0x3FFFF) >> 4) + c.a_row_shift * 8, b_lo = ((smem_u32(sB) & 0x3FFFF) >> 4) + c.b_half * 4;
    const uint32_t slot_step = c.a_slots > 1 ? (16384 >> 4) : 0;
    const uint32_t kstep = c.layout == 2 ? 2 : 0;
    long long t0 = 0;
    if (elect_one_sync()) {
      t0 = clock64();
      for (int i = 0; i < c.nmma; i += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const uint64_t da = hi | (uint64_t)(a_lo + u * slot_step + (u & 3) * kstep);
          const uint64_t db = hi | (uint64_t)(b_lo + (u & 3) * kstep);
          const uint32_t d = tmem + (c.two_acc >= 2 ? (u & (c.two_acc - 1)) * 64 : (c.two_acc ? (u & 1) * 256 : 0));
          if (c.kind == 2) mma<2>(d, da, db, idesc, 1);
          else mma<0>(d, da, db, idesc, 1);
        }
      }
      commit(bar);
      mbar_wait(bar, 0);
      cycles[blockIdx.x] = clock64() - t0;
    }
    __syncwarp();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// ---- (b) functional ------------------------------------------------------------------------------------------------
// mode 0: SW128 f16, A rows [128][64 halves], B rows [N][64 halves]; D = A[:, k0:k0+32] x B[:, k0:k0+32]^T (two K=16 steps)
// mode 1: no-swizzle f16 overlapping rows: A line of (128 + t + 8) voxels x 8 halves; B core-matrix layout [N][16]
__global__ void __launch_bounds__(128) func_kernel(const __half* A, const __half* B, float* D, int mode, int N, int k0, int t) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + 32768;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 65536);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (mode == 0) {
    for (int i = tid; i < 128 * 8; i += 128) {
      const int r = i >> 3, c = i & 7;
      *reinterpret_cast<uint4*>(sA + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(A + r * 64 + c * 8);
    }
    for (int i = tid; i < N * 8; i += 128) {
      const int r = i >> 3, c = i & 7;
      *reinterpret_cast<uint4*>(sB + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(B + r * 64 + c * 8);
    }
  } else {
    for (int i = tid; i < 128 + t + 8; i += 128) *reinterpret_cast<uint4*>(sA + i * 16) = *reinterpret_cast<const uint4*>(A + i * 8);
    // B: N rows x 16 halves, canonical no-swizzle K-major: core matrix (8 rows x 16 B) contiguous; K chunk j at +LBO=N*16... use
    // LBO = 128 * (N/8) (all row groups of chunk 0, then chunk 1), SBO = 128
    for (int i = tid; i < N * 2; i += 128) {
      const int r = i >> 1, j = i & 1;
      *reinterpret_cast<uint4*>(sB + j * (N * 16) + r * 16) = *reinterpret_cast<const uint4*>(B + r * 16 + j * 8);
    }
  }
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    if (mode == 0) {
      for (int k = 0; k < 2; ++k)
        mma<0>(tmem, make_desc(smem_u32(sA) + k0 * 2 + k * 32, 16, 1024, 2), make_desc(smem_u32(sB) + k0 * 2 + k * 32, 16, 1024, 2),
               idesc, k > 0);
    } else {
      mma<0>(tmem, make_desc(smem_u32(sA), 16 * t, 128, 0), make_desc(smem_u32(sB), N * 16, 128, 0), idesc, 0);
    }
    commit(bar);
  }
  mbar_wait(bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c0 = 0; c0 < N; c0 += 8) {
    uint32_t r[8];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;");
    for (int j = 0; j < 8; ++j) D[(warp * 32 + lane) * N + c0 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e = (x);                                                               \
    if (e != cudaSuccess) {                                                            \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);   \
      exit(1);                                                                         \
    }                                                                                  \
  } while (0)

int main() {
  // ---------------- functional ----------------
  int fails = 0;
  {
    std::vector<__half> A(128 * 64), B(256 * 64);
    std::vector<float> Af(A.size()), Bf(B.size());
    srand(3);
    for (size_t i = 0; i < A.size(); ++i) A[i] = __float2half((float)rand() / RAND_MAX * 2.f - 1.f), Af[i] = __half2float(A[i]);
    for (size_t i = 0; i < B.size(); ++i) B[i] = __float2half((float)rand() / RAND_MAX * 2.f - 1.f), Bf[i] = __half2float(B[i]);
    __half *dA, *dB;
    float* dD;
    CK(cudaMalloc(&dA, A.size() * 2));
    CK(cudaMalloc(&dB, B.size() * 2));
    CK(cudaMalloc(&dD, 128 * 256 * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(func_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000));
    const int Ns[] = {16, 32, 64, 96, 192};
    for (int N : Ns)
      for (int k0 : {0, 32}) {
        func_kernel<<<1, 128, 70000>>>(dA, dB, dD, 0, N, k0, 0);
        CK(cudaDeviceSynchronize());
        std::vector<float> D(128 * N);
        CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
        double mx = 0;
        for (int r = 0; r < 128; ++r)
          for (int n = 0; n < N; ++n) {
            double s = 0;
            for (int k = 0; k < 32; ++k) s += (double)Af[r * 64 + k0 + k] * Bf[n * 64 + k0 + k];
            mx = fmax(mx, fabs(s - D[r * N + n]));
          }
        printf("func SW128 f16 N=%3d k0=%2d: max|err|=%.3e %s\n", N, k0, mx, mx < 1e-4 ? "OK" : "MISMATCH");
        fails += !(mx < 1e-4);
      }
    // overlapping rows: A line of voxels (8 halves each); row r, chunk j = voxel r + t*j
    for (int N : {16, 48})
      for (int t : {1, 3, 130}) {
        func_kernel<<<1, 128, 70000>>>(dA, dB, dD, 1, N, 0, t);
        CK(cudaDeviceSynchronize());
        std::vector<float> D(128 * N);
        CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
        double mx = 0;
        for (int r = 0; r < 128; ++r)
          for (int n = 0; n < N; ++n) {
            double s = 0;
            for (int j = 0; j < 2; ++j)
              for (int k = 0; k < 8; ++k) s += (double)Af[(r + t * j) * 8 + k] * Bf[n * 16 + j * 8 + k];
            mx = fmax(mx, fabs(s - D[r * N + n]));
          }
        printf("func NOSWZ overlap f16 N=%3d t=%3d: max|err|=%.3e %s\n", N, t, mx, mx < 1e-4 ? "OK" : "MISMATCH");
        fails += !(mx < 1e-4);
      }
  }
  // ---------------- rates ----------------
  CK(cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 16384 + 32768 + 2048));
  long long* dcy;
  CK(cudaMalloc(&dcy, 148 * 8));
  const int grids[] = {1, 148};
  for (int grid : grids)
    for (int kind : {2, 0})
      for (int layout : {2, 0})
        for (int N : {16, 32, 48, 64, 96, 128, 192, 256})
          for (int slots : {1, 8}) {
            if (kind == 2 && layout == 0) continue;
            BenchCfg c{kind, N, layout, slots, 2048, 1, 0, 0};
            rate_kernel<<<grid, 128, 8 * 16384 + 32768 + 2048>>>(c, dcy);
            CK(cudaDeviceSynchronize());
            std::vector<long long> cy(grid);
            CK(cudaMemcpy(cy.data(), dcy, grid * 8, cudaMemcpyDeviceToHost));
            long long mx = 0;
            for (auto v : cy) mx = v > mx ? v : mx;
            fflush(stdout);
            printf("rate grid=%3d kind=%s layout=%s N=%3d a_slots=%d: %.1f cyc/MMA (ideal N/2=%d)\n", grid, kind == 2 ? "tf32" : "f16 ",
                   layout == 2 ? "SW128" : "NONE ", N, slots, (double)mx / c.nmma, N / 2);
          }
  for (int N : {96, 192})
    for (int shift : {0, 1, 26})
      for (int bh : {0, 1}) {
        BenchCfg c{0, N, 2, 4, 2048, 1, shift, bh};
        rate_kernel<<<148, 128, 8 * 16384 + 32768 + 2048>>>(c, dcy);
        CK(cudaDeviceSynchronize());
        std::vector<long long> cy(148);
        CK(cudaMemcpy(cy.data(), dcy, 148 * 8, cudaMemcpyDeviceToHost));
        long long mx = 0;
        for (auto v : cy) mx = v > mx ? v : mx;
        printf("rate2 f16 SW128 N=%3d a_row_shift=%2d b_half=%d: %.1f cyc/MMA\n", N, shift, bh, (double)mx / c.nmma);
      }
  // accumulate-chain latency: back-to-back MMAs into the SAME accumulator vs 2 / 4 / 8 independent accumulators
  for (int N : {16, 48, 64})
    for (int nacc : {0, 2, 4, 8}) {
      BenchCfg c{0, N, 0, 4, 2048, nacc, 0, 0};
      rate_kernel<<<148, 128, 8 * 16384 + 32768 + 2048>>>(c, dcy);
      CK(cudaDeviceSynchronize());
      std::vector<long long> cy(148);
      CK(cudaMemcpy(cy.data(), dcy, 148 * 8, cudaMemcpyDeviceToHost));
      long long mx = 0;
      for (auto v : cy) mx = v > mx ? v : mx;
      printf("rate3 f16 NONE N=%3d accumulators=%d: %.1f cyc/MMA\n", N, nacc ? nacc : 1, (double)mx / c.nmma);
    }
  printf(fails ? "probe FAILED (%d)\n" : "probe OK\n", fails);
  return fails != 0;
}
