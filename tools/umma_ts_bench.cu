// tcgen05 probe #3: is the fixed ~40-cycle cost of an SS tcgen05.mma (both operands in shared memory, tools/umma_bench.cu) the
// un-overlapped read of the 4 KB A tile -- i.e. does the TS form (A operand in tensor memory) remove it?
//   mode 0: SS   tcgen05.mma [d], a_desc, b_desc           (what the conv kernels do today)
//   mode 1: TS   tcgen05.mma [d], [a_tmem], b_desc         A loaded once, reused (upper bound of the benefit)
//   mode 2: TS + one tcgen05.cp.128x256b smem -> TMEM per MMA into a ring of A slots (what a kernel would really have to do)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 tools/umma_ts_bench.cu -o build/umma_ts_bench
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

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
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(da),
               "l"(db), "r"(idesc), "r"(1u)
               : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t db, uint32_t idesc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}" ::"r"(d),
               "r"(a_tmem), "l"(db), "r"(idesc), "r"(1u)
               : "memory");
}
__device__ __forceinline__ void cp_128x256b(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred = 0;
  asm volatile("{\n.reg .b32 rx;\n.reg .pred px;\nelect.sync rx|px, 0xffffffff;\n@px mov.s32 %0, 1;\n}\n" : "+r"(pred));
  return pred != 0;
}

struct Cfg {
  int mode, N, nmma;
};

__global__ void __launch_bounds__(128) rate_kernel(Cfg c, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;               // 8 x 16 KB A tiles (128 rows x 128 B, SWIZZLE_128B)
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
    const uint32_t idesc = (1u << 4) | ((uint32_t)(c.N >> 3) << 17) | ((128u >> 4) << 24);  // f16 x f16 -> f32, K-major, M = 128
    const uint32_t idesc_half = (1u << 4) | ((uint32_t)(c.N >> 4) << 17) | ((128u >> 4) << 24);
    const uint64_t hi = make_desc(0, 16, 1024, 2);
    const uint32_t a_lo = (smem_u32(sA) & 0x3FFFF) >> 4, b_lo = (smem_u32(sB) & 0x3FFFF) >> 4;
    const uint32_t a_tmem0 = tmem + 256;  // A slots: 8 columns each (128 lanes x 16 halves), behind the accumulator columns
    if (elect_one_sync()) {
      if (c.mode >= 1) {
        for (int u = 0; u < 8; ++u) cp_128x256b(a_tmem0 + u * 8, hi | (uint64_t)(a_lo + u * (16384 >> 4)));
      }
      const long long t0 = clock64();
      for (int i = 0; i < c.nmma; i += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const uint64_t da = hi | (uint64_t)(a_lo + u * (16384 >> 4) + (u & 3) * 2);
          const uint64_t db = hi | (uint64_t)(b_lo + (u & 3) * 2);
          if (c.mode == 0) {
            mma_ss(tmem, da, db, idesc);
          } else if (c.mode == 3) {  // SS, two disjoint accumulators alternating
            mma_ss(tmem + (u & 1) * 256, da, db, idesc);
          } else if (c.mode == 4) {  // SS, the tz_gemm pattern: 2 x N into [0, N), then 2 x N/2 accumulated onto [N/2, N)
            if ((u & 3) < 2) mma_ss(tmem, da, db, idesc);
            else mma_ss(tmem + c.N / 2, da, db, idesc_half);
          } else if (c.mode == 5) {  // SS, same shapes as mode 4 but the half-width MMAs go to a DISJOINT accumulator
            if ((u & 3) < 2) mma_ss(tmem, da, db, idesc);
            else mma_ss(tmem + 256, da, db, idesc_half);
          } else {
            if (c.mode == 2) cp_128x256b(a_tmem0 + u * 8, da);
            mma_ts(tmem, a_tmem0 + u * 8, db, idesc);
          }
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

// ---- branch-free issue loops: PATTERN is a compile-time constant, descriptors are precomputed, 8 MMAs per unrolled iteration ----
//   0: 8 x N=192 into one accumulator                       1: (N=192, N=192, N=96 -> +96, N=96 -> +96) x 2   (the tz_gemm pattern)
//   2: same shapes, the N=96 MMAs into a disjoint accumulator  3: 4 x N=192 then 4 x N=96 -> +96 (grouped by destination)
//   4: (N=64, N=64, N=32 -> +32, N=32 -> +32) x 2  (the dwsep / conv0 pattern)   5: 8 x N=96
template <int PATTERN>
__global__ void __launch_bounds__(128) pattern_kernel(int nmma, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + 8 * 16384;
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
    auto idesc = [](uint32_t N) { return (1u << 4) | ((N >> 3) << 17) | ((128u >> 4) << 24); };
    const uint32_t i192 = idesc(192), i96 = idesc(96), i64 = idesc(64), i32 = idesc(32);
    const uint64_t hi = make_desc(0, 16, 1024, 2);
    const uint32_t a_lo = (smem_u32(sA) & 0x3FFFF) >> 4, b_lo = (smem_u32(sB) & 0x3FFFF) >> 4;
    uint64_t da[8], db[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) da[u] = hi | (uint64_t)(a_lo + u * (16384 >> 4) + (u & 3) * 2), db[u] = hi | (uint64_t)(b_lo + (u & 3) * 2);
    if (elect_one_sync()) {
      const long long t0 = clock64();
      for (int i = 0; i < nmma; i += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          if (PATTERN == 0) mma_ss(tmem, da[u], db[u], i192);
          else if (PATTERN == 1) { if ((u & 3) < 2) mma_ss(tmem, da[u], db[u], i192); else mma_ss(tmem + 96, da[u], db[u], i96); }
          else if (PATTERN == 2) { if ((u & 3) < 2) mma_ss(tmem, da[u], db[u], i192); else mma_ss(tmem + 256, da[u], db[u], i96); }
          else if (PATTERN == 3) { if (u < 4) mma_ss(tmem, da[u], db[u], i192); else mma_ss(tmem + 96, da[u], db[u], i96); }
          else if (PATTERN == 4) { if ((u & 3) < 2) mma_ss(tmem, da[u], db[u], i64); else mma_ss(tmem + 32, da[u], db[u], i32); }
          else mma_ss(tmem, da[u], db[u], i96);
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

template <int PATTERN>
static void run_pattern(const char* name, double ideal, long long* d, size_t smem) {
  cudaFuncSetAttribute(pattern_kernel<PATTERN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  pattern_kernel<PATTERN><<<1, 128, smem>>>(8192, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("pattern %-64s: %6.1f cycles / MMA (math alone %.0f)%s\n", name, (double)h / 8192, ideal, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  const size_t smem = 8 * 16384 + 32768 + 64 + 1024;
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  const char* names[6] = {"SS (A, B in smem)", "TS (A resident in TMEM)", "TS + tcgen05.cp 4 KB per MMA", "SS, 2 disjoint accumulators alternating",
                          "SS, 2xN then 2xN/2 onto [N/2,N) (tz_gemm)", "SS, 2xN then 2xN/2 onto a disjoint accumulator"};
  for (int grid : {1})
    for (int mode = 0; mode < 6; ++mode)
      for (int N : {32, 64, 96, 192, 256}) {
        if (mode >= 4 && N != 192 && N != 64) continue;
        Cfg c{mode, N, 4096};
        rate_kernel<<<grid, 128, smem>>>(c, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf("mode %d N %d: %s\n", mode, N, cudaGetErrorString(e));
          return 1;
        }
        long long h[148];
        cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("grid %3d  %-48s N=%3d: %6.1f cycles / MMA (ideal N/2 = %d)\n", grid, names[mode], N, (double)mx / c.nmma, N / 2);
      }
  run_pattern<0>("8 x N=192, one accumulator", 96, d, smem);
  run_pattern<5>("8 x N=96, one accumulator", 48, d, smem);
  run_pattern<1>("(192, 192, 96 -> +96, 96 -> +96) x 2   [tz_gemm]", 72, d, smem);
  run_pattern<2>("(192, 192, 96 -> disjoint, 96 -> disjoint) x 2", 72, d, smem);
  run_pattern<3>("4 x 192 then 4 x 96 -> +96   [grouped by destination]", 72, d, smem);
  run_pattern<4>("(64, 64, 32 -> +32, 32 -> +32) x 2   [dwsep / conv0]", 24, d, smem);
  return 0;
}
