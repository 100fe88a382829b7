// K3 (C = 32): the 32 -> 32 layers of the stage-1 3D stack as an implicit GEMM on the 5th-generation tensor cores
// (tcgen05.mma kind::tf32, accumulators in TMEM, operands staged by TMA), with a 3xTF32 error-compensated split so the
// result stays at fp32 accuracy (plain TF32 operands move stage-1 disparities by up to 2.5 px).
// Replaces 4 of the 6 cuDNN conv3d launches of post_3dconvs(4, 32) (reference models/submodules.py:216-221).
//
// Layout ("CLP"): activations are channels-last with a one-voxel zero border in y and x:
//     act[b][d][y][x][32] fp32,  y in [0,H+2), x in [0,W+2)   ->  a voxel = one 128-byte row, R = D*(H+2)*(W+2) rows / b
// so that a 3x3x3 tap is a constant row offset  off = (kd-1)*(H+2)*(W+2) + (kh-1)*(W+2) + (kw-1)  and an output tile of
// 128 consecutive rows needs, per (kd,kh), one 136-row TMA box (rows outside [0,R) are zero-filled by TMA = the d
// padding); the three kw taps are the same shared-memory tile read through UMMA descriptors shifted by 0/1/2 rows
// (SWIZZLE_128B is a function of the absolute smem address, probe: tools/umma_probe.cu).
//
// 3xTF32:  x = xh + xl, w = wh + wl (xh = x with the low 13 mantissa bits cleared — exactly what the tensor core does
// to an fp32 operand, so the raw x tile *is* xh).  acc = xh*wh + xh*wl + xl*wh in fp32.  Per K=8 step two MMAs:
//     [acc_hh | acc_hl] (N=64) += x_tile  * [wh | wl]        acc_lh (N=32) += xl_tile * wh
// (N=64 so the 4 KB A-tile read, which bounds small-N tcgen05.mma, is paid twice instead of three times).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner, warps 2-5 = xl converter
// (x - trunc(x) into a second smem tile, fence.proxy.async) and, after the last tap, epilogue (tcgen05.ld, sum of the
// three accumulators, bias + ReLU, zero the border rows, 128-byte row stores).  A CTA owns G consecutive tiles whose
// accumulators live in TMEM together, so each 24 KB weight stage is fetched once per G tiles.
#include <math.h>
#include <string.h>

#include "lws_common.cuh"
#include "tma_utils.cuh"

namespace lws {

constexpr int TC_M = 128;                 // rows (voxels) per tile
constexpr int TC_AROWS = 144;             // rows per A stage: 128 + the kw halo (2 rows for the 3D conv, 16 for dilation 8)
constexpr int TC_ABYTES = TC_AROWS * 128;  // 18432 = 18 x 1024 (keeps every slot 1024-byte aligned for SWIZZLE_128B)
constexpr int TC_BROWS = 192;             // 3 kw x (32 hi + 32 lo) rows
constexpr int TC_BBYTES = TC_BROWS * 128;  // 24576
constexpr int TC_NSX = 5;                 // x-tile ring (TMA prefetch distance)
constexpr int TC_NSL = 3;                 // xl-tile ring (written by the converter warps ahead of the MMAs)
constexpr int TC_GMAX = 4;                // tiles per CTA group (4 x 96 TMEM columns)
constexpr int TC_THREADS = 192;
constexpr int TC_SMEM = (TC_NSX + TC_NSL) * TC_ABYTES + 2 * TC_BBYTES + 1024 /*align slack*/ + 256 /*barriers*/;

struct TcArgs {
  float* out;         // CLP [B][R][32]
  const float* bias;  // [32]
  int R;              // rows per batch element
  int Hp, Wp, pad;    // padded plane height / width and the border width (rows with x or y inside the border are zeroed)
  int tiles_per_b, groups_per_b, G, total_groups;
  int nstages;        // <= 9 operand stages per tile; stage s reads source st_src[s] at row offset st_off[s]
  int kw_shift;       // rows between the three kw taps inside a stage (1 for the 3D conv, 8 for the dilated 2D conv)
  int relu;
  int Hi, Wi;         // interior (un-padded) plane size: voxels with x or y outside [pad, pad+size) are written as zeros
  int kmask[3];       // per kw shift: which of the four K=8 steps carry non-zero weights (0xF for 32-channel voxels)
  int st_off[9];
  int st_src[9];
};

__device__ __forceinline__ uint64_t tc_sdesc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;            // LBO (unused: swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;  // SBO = 8 rows x 128 B
  d |= (uint64_t)1 << 46;            // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;            // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,"
      "%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
}

// CPV = channels per voxel.  32: a row is one voxel.  8: a row is 4 consecutive voxels (x fastest) and the kw taps are
// the rows g-1 / g / g+1 with block-structured weight tables (see pack_tc_table_c8): only the K=8 step holding the
// neighbouring voxel is issued for the side rows.
template <int CPV>
__global__ void __launch_bounds__(TC_THREADS, 1)
    tc_implicit_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapA1,
                         const __grid_constant__ CUtensorMap mapB, const TcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sX = smem;                                      // [NSX][18432]  x tiles (= xh as far as the MMA is concerned)
  uint8_t* sL = smem + TC_NSX * TC_ABYTES;                 // [NSL][18432]  xl tiles
  uint8_t* sB = smem + (TC_NSX + TC_NSL) * TC_ABYTES;      // [2][24576]    weight stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (TC_NSX + TC_NSL) * TC_ABYTES + 2 * TC_BBYTES);
  uint64_t* a_full = bars;                      // [NSX] TMA landed
  uint64_t* a_empty = bars + TC_NSX;            // [NSX] MMAs that read the x slot retired
  uint64_t* a_conv = bars + 2 * TC_NSX;         // [NSL] xl written
  uint64_t* l_empty = a_conv + TC_NSL;          // [NSL] MMAs that read the xl slot retired
  uint64_t* b_full = l_empty + TC_NSL;          // [2]
  uint64_t* b_empty = b_full + 2;        // [2]
  uint64_t* acc_full = b_empty + 2;      // all MMAs of the group retired
  uint64_t* acc_empty = acc_full + 1;    // epilogue drained TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < TC_NSX; ++i) mbar_init(a_full + i, 1), mbar_init(a_empty + i, 1);
    for (int i = 0; i < TC_NSL; ++i) mbar_init(a_conv + i, 4), mbar_init(l_empty + i, 1);
    for (int i = 0; i < 2; ++i) mbar_init(b_full + i, 1), mbar_init(b_empty + i, 1);
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 4);
    mbar_fence_init();
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapA1);
    tma_prefetch_desc(&mapB);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  const int nst = a.nstages;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (elect_one_sync()) {
      uint32_t it = 0, bs = 0;
      for (int grp = blockIdx.x; grp < a.total_groups; grp += gridDim.x) {
        const int b = grp / a.groups_per_b;
        const int t0 = (grp - b * a.groups_per_b) * a.G;
        const int ntile = min(a.G, a.tiles_per_b - t0);
        for (int st = 0; st < nst; ++st, ++bs) {
          mbar_wait(b_empty + (bs & 1), ((bs >> 1) & 1) ^ 1);
          mbar_expect_tx(b_full + (bs & 1), TC_BBYTES);
          tma_load_2d(sB + (bs & 1) * TC_BBYTES, &mapB, b_full + (bs & 1), 0, st * TC_BROWS);
          const int off = a.st_off[st];
          const CUtensorMap* src = a.st_src[st] ? &mapA1 : &mapA;
          for (int g = 0; g < ntile; ++g, ++it) {
            const uint32_t slot = it % TC_NSX;
            mbar_wait(a_empty + slot, ((it / TC_NSX) & 1) ^ 1);
            mbar_expect_tx(a_full + slot, TC_ABYTES);
            // 3D map {32, R, B}: rows outside [0, R) come back as zeros (the depth padding)
            asm volatile(
                "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], "
                "[%2];" ::"r"(smem_u32(sX + slot * TC_ABYTES)),
                "l"(reinterpret_cast<uint64_t>(src)), "r"(smem_u32(a_full + slot)), "r"(0),
                "r"((t0 + g) * TC_M + off), "r"(b)
                : "memory");
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    // The whole warp walks the schedule (so control flow stays converged); one elected lane issues the MMAs and commits.
    // M=128, K-major A and B, fp32 accumulate, tf32 operands; N = 64 / 32
    const uint32_t idesc64 = (1u << 4) | (2u << 7) | (2u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t idesc32 = (1u << 4) | (2u << 7) | (2u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t desc_hi = ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);  // SBO, version, SW128
    const uint32_t km0 = a.kmask[0], km1 = a.kmask[1], km2 = a.kmask[2];
    uint32_t it = 0, bs = 0, gi = 0;
    const uint32_t kwq = ((uint32_t)a.kw_shift * 128u) >> 4;  // kw shift in 16-byte descriptor units
    for (int grp = blockIdx.x; grp < a.total_groups; grp += gridDim.x, ++gi) {
      const int b = grp / a.groups_per_b;
      const int t0 = (grp - b * a.groups_per_b) * a.G;
      const int ntile = min(a.G, a.tiles_per_b - t0);
      mbar_wait(acc_empty, (gi & 1) ^ 1);  // previous group's accumulators drained
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int st = 0; st < nst; ++st, ++bs) {
        mbar_wait(b_full + (bs & 1), (bs >> 1) & 1);
        const uint32_t b_lo = ((smem_u32(sB + (bs & 1) * TC_BBYTES) & 0x3FFFF) >> 4) | (1u << 16);
        for (int g = 0; g < ntile; ++g, ++it) {
          const uint32_t slot = it % TC_NSX, lslot = it % TC_NSL;
          mbar_wait(a_full + slot, (it / TC_NSX) & 1);
          mbar_wait(a_conv + lslot, (it / TC_NSL) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (elect_one_sync()) {
            const uint32_t x_lo = ((smem_u32(sX + slot * TC_ABYTES) & 0x3FFFF) >> 4) | (1u << 16);
            const uint32_t l_lo = ((smem_u32(sL + lslot * TC_ABYTES) & 0x3FFFF) >> 4) | (1u << 16);
            const uint32_t d_hh = tmem + g * 96, d_lh = tmem + g * 96 + 64;
            uint32_t acc = st != 0;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
              const uint32_t km = kw == 0 ? km0 : (kw == 1 ? km1 : km2);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (!((km >> k) & 1)) continue;
                const uint64_t db = desc_hi | (uint64_t)(b_lo + kw * (8192 >> 4) + k * 2);
                tc_mma_tf32(d_hh, desc_hi | (uint64_t)(x_lo + kw * kwq + k * 2), db, idesc64, acc);
                tc_mma_tf32(d_lh, desc_hi | (uint64_t)(l_lo + kw * kwq + k * 2), db, idesc32, acc);
                acc = 1;
              }
            }
            tc_commit(a_empty + slot);
            tc_commit(l_empty + lslot);
          }
          __syncwarp();
        }
        if (elect_one_sync()) tc_commit(b_empty + (bs & 1));
        __syncwarp();
      }
      if (elect_one_sync()) tc_commit(acc_full);
      __syncwarp();
    }
  } else {
    // ================================ xl converter + epilogue (warps 2..5) ================================
    const int ct = tid - 64;  // 0..127
    const int q = warp & 3;   // TMEM lane quarter this warp may access
    uint32_t it = 0, gi = 0;
    float bias[CPV];
#pragma unroll
    for (int j = 0; j < CPV; ++j) bias[j] = __ldg(a.bias + j);
    for (int grp = blockIdx.x; grp < a.total_groups; grp += gridDim.x, ++gi) {
      const int b = grp / a.groups_per_b;
      const int t0 = (grp - b * a.groups_per_b) * a.G;
      const int ntile = min(a.G, a.tiles_per_b - t0);
      for (int n = 0; n < nst * ntile; ++n, ++it) {
        const uint32_t slot = it % TC_NSX, lslot = it % TC_NSL;
        mbar_wait(l_empty + lslot, ((it / TC_NSL) & 1) ^ 1);
        mbar_wait(a_full + slot, (it / TC_NSX) & 1);
        // 18432 B = 9 float4 per thread: all nine shared loads are issued before the first use (one round trip)
        const uint32_t src = smem_u32(sX + slot * TC_ABYTES) + ct * 16, dst = smem_u32(sL + lslot * TC_ABYTES) + ct * 16;
        float4 x[9];
#pragma unroll
        for (int j = 0; j < 9; ++j)
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(x[j].x), "=f"(x[j].y), "=f"(x[j].z), "=f"(x[j].w)
                       : "r"(src + j * 2048));
#pragma unroll
        for (int j = 0; j < 9; ++j) {
          float4 l;
          l.x = x[j].x - __uint_as_float(__float_as_uint(x[j].x) & 0xFFFFE000u);
          l.y = x[j].y - __uint_as_float(__float_as_uint(x[j].y) & 0xFFFFE000u);
          l.z = x[j].z - __uint_as_float(__float_as_uint(x[j].z) & 0xFFFFE000u);
          l.w = x[j].w - __uint_as_float(__float_as_uint(x[j].w) & 0xFFFFE000u);
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst + j * 2048), "f"(l.x), "f"(l.y), "f"(l.z), "f"(l.w)
                       : "memory");
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(a_conv + lslot);
      }
      // ---- epilogue of this group ----
      mbar_wait(acc_full, gi & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int g = 0; g < ntile; ++g) {
        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + g * 96;
        float acc[32], t[32];
        tc_ld32(taddr, acc);
        tc_ld32(taddr + 32, t);
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] += t[j];
        tc_ld32(taddr + 64, t);
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] += t[j];
        const int r = (t0 + g) * TC_M + q * 32 + lane;  // row inside this batch element
        if (r < a.R) {
          constexpr int VPR = 32 / CPV;  // voxels per row
          const float lo = a.relu ? 0.f : -INFINITY;
          float4* o = reinterpret_cast<float4*>(a.out + ((long long)b * a.R + r) * 32);
#pragma unroll
          for (int u = 0; u < VPR; ++u) {
            const int vi = r * VPR + u;
            const int x = vi % a.Wp, y = (vi / a.Wp) % a.Hp;
            const bool border = x < a.pad || x >= a.pad + a.Wi || y < a.pad || y >= a.pad + a.Hi;
#pragma unroll
            for (int j = 0; j < CPV / 4; ++j) {
              const int c = u * CPV + 4 * j;
              float4 v;
              v.x = border ? 0.f : fmaxf(acc[c] + bias[c % CPV], lo);
              v.y = border ? 0.f : fmaxf(acc[c + 1] + bias[(c + 1) % CPV], lo);
              v.z = border ? 0.f : fmaxf(acc[c + 2] + bias[(c + 2) % CPV], lo);
              v.w = border ? 0.f : fmaxf(acc[c + 3] + bias[(c + 3) % CPV], lo);
              o[c / 4] = v;
            }
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty);
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// ---- 1 -> C on the raw cost, writing CLP (incl. the zero border); C/8 lanes per voxel, 8 couts per lane ---------------
// CLP geometry: voxel (d, y, x) with y in [0, H+2), x in [0, Wp); interior = y in [1, H], x in [1, W]; Wp >= W + 2.
template <int C>
__global__ void __launch_bounds__(256)
    conv3d_first_clp_kernel(const float* __restrict__ cost, const float* __restrict__ w /*[27][C]*/,
                            const float* __restrict__ bias, const float* __restrict__ affine, float* __restrict__ out,
                            int D, int H, int W, int Wp, long long total_vox) {
  constexpr int LPV = C / 8;
  __shared__ __align__(16) float sW[27 * C];
  for (int i = threadIdx.x; i < 27 * C; i += blockDim.x) sW[i] = __ldg(w + i);
  __syncthreads();
  const float s0 = __ldg(affine), t0 = __ldg(affine + 1);
  const int sub = threadIdx.x % LPV;
  const int Hp = H + 2;
  const long long hw = (long long)H * W;
  float bv[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) bv[j] = __ldg(bias + sub * 8 + j);
  for (long long vox = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / LPV; vox < total_vox;
       vox += ((long long)gridDim.x * blockDim.x) / LPV) {
    const int x = (int)(vox % Wp);
    long long t = vox / Wp;
    const int y = (int)(t % Hp);
    t /= Hp;
    const int d = (int)(t % D);
    const int b = (int)(t / D);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    const bool border = x < 1 || x > W || y < 1 || y > H;
    if (!border) {
      const float* cb = cost + (long long)b * D * hw + (long long)d * hw + (long long)(y - 1) * W + (x - 1);
#pragma unroll
      for (int kd = 0; kd < 3; ++kd) {
        const bool okd = (unsigned)(d + kd - 1) < (unsigned)D;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
          const bool okh = okd && (unsigned)(y - 1 + kh - 1) < (unsigned)H;
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const bool ok = okh && (unsigned)(x - 1 + kw - 1) < (unsigned)W;
            float v = ok ? __ldg(cb + (kd - 1) * hw + (kh - 1) * W + (kw - 1)) : 0.f;
            v = ok ? fmaxf(fmaf(v, s0, t0), 0.f) : 0.f;
            const float4 wa = *reinterpret_cast<const float4*>(sW + (kd * 9 + kh * 3 + kw) * C + sub * 8);
            const float4 wb = *reinterpret_cast<const float4*>(sW + (kd * 9 + kh * 3 + kw) * C + sub * 8 + 4);
            acc[0] = fmaf(v, wa.x, acc[0]), acc[1] = fmaf(v, wa.y, acc[1]), acc[2] = fmaf(v, wa.z, acc[2]),
            acc[3] = fmaf(v, wa.w, acc[3]), acc[4] = fmaf(v, wb.x, acc[4]), acc[5] = fmaf(v, wb.y, acc[5]),
            acc[6] = fmaf(v, wb.z, acc[6]), acc[7] = fmaf(v, wb.w, acc[7]);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaxf(acc[j] + bv[j], 0.f);
    }
    float4* o = reinterpret_cast<float4*>(out + vox * C + sub * 8);
    o[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    o[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
}

// ---- C -> 1 from CLP (+ skip), NCDHW output; C/8 lanes per voxel, 8 input channels per lane --------------------------
template <int C>
__global__ void __launch_bounds__(256)
    conv3d_last_clp_kernel(const float* __restrict__ act, const float* __restrict__ w /*[C][27] = packed [Cin][27][1]*/,
                           const float* __restrict__ skip, float* __restrict__ out, int D, int H, int W, int Wp,
                           long long total_vox) {
  constexpr int LPV = C / 8;
  __shared__ __align__(16) float sW[27 * C];  // [tap][ci]
  for (int i = threadIdx.x; i < 27 * C; i += blockDim.x) sW[(i % 27) * C + i / 27] = __ldg(w + i);
  __syncthreads();
  const int sub = threadIdx.x % LPV;
  const int Hp = H + 2;
  const long long R = (long long)D * Hp * Wp;  // voxels per batch element
  for (long long vox = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / LPV; vox < total_vox;
       vox += ((long long)gridDim.x * blockDim.x) / LPV) {
    const int x = (int)(vox % W);
    long long t = vox / W;
    const int y = (int)(t % H);
    t /= H;
    const int d = (int)(t % D);
    const int b = (int)(t / D);
    const float* base = act + ((long long)b * R + ((long long)d * Hp + (y + 1)) * Wp + (x + 1)) * C + sub * 8;
    float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
    for (int kd = 0; kd < 3; ++kd) {
      if ((unsigned)(d + kd - 1) >= (unsigned)D) continue;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const float* p = base + (((long long)(kd - 1) * Hp + (kh - 1)) * Wp + (kw - 1)) * C;
          const float4 va = __ldg(reinterpret_cast<const float4*>(p)), vb = __ldg(reinterpret_cast<const float4*>(p + 4));
          const float4 wa = *reinterpret_cast<const float4*>(sW + (kd * 9 + kh * 3 + kw) * C + sub * 8);
          const float4 wb = *reinterpret_cast<const float4*>(sW + (kd * 9 + kh * 3 + kw) * C + sub * 8 + 4);
          acc0 = fmaf(va.x, wa.x, acc0), acc0 = fmaf(va.y, wa.y, acc0), acc0 = fmaf(va.z, wa.z, acc0), acc0 = fmaf(va.w, wa.w, acc0);
          acc1 = fmaf(vb.x, wb.x, acc1), acc1 = fmaf(vb.y, wb.y, acc1), acc1 = fmaf(vb.z, wb.z, acc1), acc1 = fmaf(vb.w, wb.w, acc1);
        }
      }
    }
    float acc = acc0 + acc1;
#pragma unroll
    for (int m = 1; m < LPV; m <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
    if (sub == 0) out[vox] = acc + (skip ? __ldg(skip + vox) : 0.f);
  }
}

// ---- host ---------------------------------------------------------------------------------------------------------------
static int clp_wp(int C, int W) { return C == 32 ? W + 2 : round_up(W + 2, 32 / C); }

size_t conv3d_tc_workspace_bytes(int B, int C, int D, int H, int W) {
  const size_t vox = (size_t)B * D * (H + 2) * clp_wp(C, W);
  return 2 * ((vox * C * sizeof(float) + 255) / 256 * 256);
}

// One implicit-GEMM layer on CLP tensors viewed as rows of 32 floats: out = act(sum over stages/taps of src[row + off] x W
// + bias).  src0/src1: [B][R][32]; wtc: [nstages*192][32] operand table; stage s reads source st_src[s] at row offset
// st_off[s] (the three kw taps are kw_shift rows apart inside the stage's 144-row box); Hp x Wp voxels per plane,
// interior Hi x Wi behind a `pad`-wide border; cpv = channels per voxel (32 or 8).
int launch_tc_implicit_gemm(const float* src0, const float* src1, const float* wtc, const float* bias, float* out, int B,
                            int R, int Hp, int Wp, int pad, int Hi, int Wi, int cpv, int nstages, const int* st_off,
                            const int* st_src, int kw_shift, int relu, cudaStream_t st) {
  if (nstages < 1 || nstages > 9 || 2 * kw_shift + TC_M > TC_AROWS || (cpv != 32 && cpv != 8)) return LWS_ERR_UNSUPPORTED;
  cudaError_t e = cpv == 32 ? cudaFuncSetAttribute(tc_implicit_gemm_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM)
                            : cudaFuncSetAttribute(tc_implicit_gemm_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM);
  if (e != cudaSuccess) return (int)e;
  TcArgs a;
  memset(&a, 0, sizeof(a));
  a.R = R, a.Hp = Hp, a.Wp = Wp, a.pad = pad, a.Hi = Hi, a.Wi = Wi, a.nstages = nstages, a.kw_shift = kw_shift, a.relu = relu;
  if (cpv == 32) a.kmask[0] = a.kmask[1] = a.kmask[2] = 0xF;
  else a.kmask[0] = 0x8, a.kmask[1] = 0xF, a.kmask[2] = 0x1;  // side rows: only the K=8 step of the adjacent voxel
  for (int s = 0; s < nstages; ++s) a.st_off[s] = st_off[s], a.st_src[s] = st_src[s];
  a.tiles_per_b = (R + TC_M - 1) / TC_M;
  // tiles per CTA group: the G in {2,3,4} that minimises (rounds over 148 SMs) x G, i.e. the tail of the last round
  int bestG = TC_GMAX;
  long long best = -1;
  for (int G = TC_GMAX; G >= 2; --G) {
    const long long groups = (long long)B * ((a.tiles_per_b + G - 1) / G);
    const long long cost_ = ((groups + kNumSMs - 1) / kNumSMs) * G;
    if (best < 0 || cost_ < best) best = cost_, bestG = G;
  }
  a.G = bestG;
  a.groups_per_b = (a.tiles_per_b + a.G - 1) / a.G;
  a.total_groups = B * a.groups_per_b;
  const int grid = a.total_groups < kNumSMs ? a.total_groups : kNumSMs;
  CUtensorMap mapA0, mapA1, mapB;
  const uint64_t dimsA[3] = {32, (uint64_t)R, (uint64_t)B}, strA[2] = {128, (uint64_t)R * 128};
  const uint32_t boxA[3] = {32, TC_AROWS, 1};
  int rc = make_tensor_map_f32(&mapA0, src0, 3, dimsA, strA, boxA, true);
  if (rc) return rc;
  rc = make_tensor_map_f32(&mapA1, src1, 3, dimsA, strA, boxA, true);
  if (rc) return rc;
  const uint64_t dimsB[2] = {32, (uint64_t)nstages * TC_BROWS}, strB[1] = {128};
  const uint32_t boxB[2] = {32, TC_BROWS};
  rc = make_tensor_map_f32(&mapB, wtc, 2, dimsB, strB, boxB, true);
  if (rc) return rc;
  a.out = out, a.bias = bias;
  if (cpv == 32) tc_implicit_gemm_kernel<32><<<grid, TC_THREADS, TC_SMEM, st>>>(mapA0, mapA1, mapB, a);
  else tc_implicit_gemm_kernel<8><<<grid, TC_THREADS, TC_SMEM, st>>>(mapA0, mapA1, mapB, a);
  e = cudaPeekAtLastError();
  return e == cudaSuccess ? LWS_OK : (int)e;
}

// w_first [27][C], b_first [C]; per mid layer: wtc [9*192][32] (hi/lo rows), bias [C]; w_last [C][27]
template <int C>
static int conv3d_stack_tc_impl(const float* cost, const float* affine, const float* w_first, const float* b_first,
                                const float* const* wtc, const float* const* bias_mid, int layers, const float* w_last,
                                float* out, void* ws, int B, int D, int H, int W, int add_skip, cudaStream_t st) {
  const int Hp = H + 2, Wp = clp_wp(C, W);
  const long long vox_b = (long long)D * Hp * Wp;      // voxels per batch element
  const long long R = vox_b * C / 32;                   // 128-byte rows per batch element
  if (R >= (1ll << 31) - 4096) return LWS_ERR_BAD_SHAPE;
  const size_t act_bytes = conv3d_tc_workspace_bytes(B, C, D, H, W) / 2;
  float* bufA = (float*)ws;
  float* bufB = (float*)((char*)ws + act_bytes);
  const long long nvox = (long long)B * vox_b;
  constexpr int LPV = C / 8;
  cudaError_t e;
  {
    const long long thr = nvox * LPV;
    const int blocks = (int)((thr + 255) / 256 < 148 * 16 ? (thr + 255) / 256 : 148 * 16);
    conv3d_first_clp_kernel<C><<<blocks, 256, 0, st>>>(cost, w_first, b_first, affine, bufA, D, H, W, Wp, nvox);
    if ((e = cudaPeekAtLastError()) != cudaSuccess) return (int)e;
  }
  float* cur = bufA;
  float* nxt = bufB;
  int st_off[9], st_src[9];
  const int rows_line = Wp * C / 32;  // rows per x line
  for (int s = 0; s < 9; ++s) st_off[s] = ((s / 3 - 1) * Hp + (s % 3 - 1)) * rows_line - 1, st_src[s] = 0;
  for (int l = 0; l < layers; ++l) {
    int rc = launch_tc_implicit_gemm(cur, cur, wtc[l], bias_mid[l], nxt, B, (int)R, Hp, Wp, 1, H, W, C, 9, st_off, st_src, 1,
                                     1, st);
    if (rc) return rc;
    float* t = cur;
    cur = nxt, nxt = t;
  }
  {
    const long long vox = (long long)B * D * H * W;
    const long long thr = vox * LPV;
    const int blocks = (int)((thr + 255) / 256 < 148 * 16 ? (thr + 255) / 256 : 148 * 16);
    conv3d_last_clp_kernel<C><<<blocks, 256, 0, st>>>(cur, w_last, add_skip ? cost : nullptr, out, D, H, W, Wp, vox);
    if ((e = cudaPeekAtLastError()) != cudaSuccess) return (int)e;
  }
  return LWS_OK;
}

// one C -> C tensor-core layer on its own (CLP in, CLP out): tests and per-kernel timing
int conv3d_tc_layer(int C, const float* in_clp, const float* wtc, const float* bias, float* out_clp, int B, int D, int H,
                    int W, cudaStream_t st) {
  if (C != 32 && C != 8) return LWS_ERR_UNSUPPORTED;
  const int Hp = H + 2, Wp = clp_wp(C, W);
  const long long R = (long long)D * Hp * Wp * C / 32;
  if (R >= (1ll << 31) - 4096) return LWS_ERR_BAD_SHAPE;
  int st_off[9], st_src[9];
  const int rows_line = Wp * C / 32;
  for (int s = 0; s < 9; ++s) st_off[s] = ((s / 3 - 1) * Hp + (s % 3 - 1)) * rows_line - 1, st_src[s] = 0;
  return launch_tc_implicit_gemm(in_clp, in_clp, wtc, bias, out_clp, B, (int)R, Hp, Wp, 1, H, W, C, 9, st_off, st_src, 1, 1, st);
}

int conv3d_stack_tc(int C, const float* cost, const float* affine, const float* w_first, const float* b_first,
                    const float* const* wtc, const float* const* bias_mid, int layers, const float* w_last, float* out,
                    void* ws, int B, int D, int H, int W, int add_skip, cudaStream_t st) {
  if (C == 32)
    return conv3d_stack_tc_impl<32>(cost, affine, w_first, b_first, wtc, bias_mid, layers, w_last, out, ws, B, D, H, W,
                                    add_skip, st);
  if (C == 8)
    return conv3d_stack_tc_impl<8>(cost, affine, w_first, b_first, wtc, bias_mid, layers, w_last, out, ws, B, D, H, W,
                                   add_skip, st);
  return LWS_ERR_UNSUPPORTED;
}

}  // namespace lws
