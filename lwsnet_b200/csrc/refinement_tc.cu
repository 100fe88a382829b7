// K6 on the tensor cores: the colour-guidance refinement (reference models/submodules.py:223-327, models/models.py:158-162)
// in channels-last "CLP" layout  act[b][y][x][32] fp32 with a 16-pixel zero border (the largest dilation), so every
// 32-channel pixel is one 128-byte row and every dilated tap is a constant row offset.
//   * BN-ReLU-DW(dil)-PW block  -> dwsep_tc_kernel: the depthwise 3x3 runs on the CUDA cores straight from global memory
//     (8 lanes per pixel, 128-bit coalesced loads), its result is written as the A operand (x and x - trunc(x)) into
//     SWIZZLE_128B shared-memory tiles, and the 32x32 pointwise product is 8 tcgen05.mma (3xTF32 split, see conv3d_tc.cu)
//     against a weight tile that stays resident in shared memory; accumulators ping-pong in TMEM so the MMA of tile i
//     overlaps the depthwise phase of tile i+1 and the epilogue of tile i-1.  Only 4 accumulation steps per output, so
//     the result is fp32-exact to a couple of ulps.
//   * dense 64->32 dilation-8 3x3 -> the implicit-GEMM kernel of conv3d_tc.cu with 6 stages (2 sources x 3 kh), kw taps 8
//     rows apart in the stage tile; the concat is never formed (the two refinement1 branches are the two sources).
//   * 3->32 / 1->32 first convs and the 32->1 last conv (+ skip) are small FP32 kernels reading / writing CLP.
#include <math.h>
#include <string.h>

#include "lws_common.cuh"
#include "tma_utils.cuh"

namespace lws {

int launch_tc_implicit_gemm(const float* src0, const float* src1, const float* wtc, const float* bias, float* out, int B,
                            int R, int Hp, int Wp, int pad, int Hi, int Wi, int cpv, int nstages, const int* st_off,
                            const int* st_src, int kw_shift, int relu, cudaStream_t st);

constexpr int RP = 16;  // border of the refinement CLP tensors

// ---- first convs: NCHW [B,CIN,H,W] -> CLP; 4 lanes per pixel, 8 couts per lane ----------------------------------------
template <int CIN>
__global__ void __launch_bounds__(256)
    ref_conv0_clp_kernel(const float* __restrict__ in, const float* __restrict__ w /*[CIN][9][32]*/,
                         const float* __restrict__ bias, float* __restrict__ out, int H, int W, long long total_rows) {
  __shared__ __align__(16) float sW[CIN * 9 * 32];
  for (int i = threadIdx.x; i < CIN * 9 * 32; i += blockDim.x) sW[i] = __ldg(w + i);
  __syncthreads();
  const int sub = threadIdx.x & 3;
  const int Hp = H + 2 * RP, Wp = W + 2 * RP;
  const long long hw = (long long)H * W;
  float bv[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) bv[j] = __ldg(bias + sub * 8 + j);
  for (long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 2; row < total_rows;
       row += ((long long)gridDim.x * blockDim.x) >> 2) {
    const int x = (int)(row % Wp) - RP;
    const long long t = row / Wp;
    const int y = (int)(t % Hp) - RP;
    const int b = (int)(t / Hp);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    if (x >= 0 && x < W && y >= 0 && y < H) {
      const float* ib = in + (long long)b * CIN * hw + (long long)y * W + x;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const bool oky = (unsigned)(y + ky - 1) < (unsigned)H;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const bool ok = oky && (unsigned)(x + kx - 1) < (unsigned)W;
#pragma unroll
          for (int ci = 0; ci < CIN; ++ci) {
            const float v = ok ? __ldg(ib + ci * hw + (ky - 1) * W + (kx - 1)) : 0.f;
            const float4 wa = *reinterpret_cast<const float4*>(sW + (ci * 9 + ky * 3 + kx) * 32 + sub * 8);
            const float4 wb = *reinterpret_cast<const float4*>(sW + (ci * 9 + ky * 3 + kx) * 32 + sub * 8 + 4);
            acc[0] = fmaf(v, wa.x, acc[0]), acc[1] = fmaf(v, wa.y, acc[1]), acc[2] = fmaf(v, wa.z, acc[2]),
            acc[3] = fmaf(v, wa.w, acc[3]), acc[4] = fmaf(v, wb.x, acc[4]), acc[5] = fmaf(v, wb.y, acc[5]),
            acc[6] = fmaf(v, wb.z, acc[6]), acc[7] = fmaf(v, wb.w, acc[7]);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaxf(acc[j] + bv[j], 0.f);
    }
    float4* o = reinterpret_cast<float4*>(out + row * 32 + sub * 8);
    o[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    o[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
}

// ---- last conv: CLP -> NCHW [B,1,H,W] (+ skip); 4 lanes per pixel, 8 input channels per lane ----------------------------
__global__ void __launch_bounds__(256)
    ref_last_clp_kernel(const float* __restrict__ act, const float* __restrict__ w /*[32][9]*/,
                        const float* __restrict__ skip, float* __restrict__ out, int H, int W, long long total_px) {
  __shared__ __align__(16) float sW[9 * 32];  // [tap][ci]
  for (int i = threadIdx.x; i < 9 * 32; i += blockDim.x) sW[(i % 9) * 32 + i / 9] = __ldg(w + i);
  __syncthreads();
  const int sub = threadIdx.x & 3;
  const int Hp = H + 2 * RP, Wp = W + 2 * RP;
  const long long R = (long long)Hp * Wp;
  for (long long px = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 2; px < total_px;
       px += ((long long)gridDim.x * blockDim.x) >> 2) {
    const int x = (int)(px % W);
    const long long t = px / W;
    const int y = (int)(t % H);
    const int b = (int)(t / H);
    const float* base = act + ((long long)b * R + (long long)(y + RP) * Wp + (x + RP)) * 32 + sub * 8;
    float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const float* p = base + ((long long)(ky - 1) * Wp + (kx - 1)) * 32;
        const float4 va = __ldg(reinterpret_cast<const float4*>(p)), vb = __ldg(reinterpret_cast<const float4*>(p + 4));
        const float4 wa = *reinterpret_cast<const float4*>(sW + (ky * 3 + kx) * 32 + sub * 8);
        const float4 wb = *reinterpret_cast<const float4*>(sW + (ky * 3 + kx) * 32 + sub * 8 + 4);
        acc0 = fmaf(va.x, wa.x, acc0), acc0 = fmaf(va.y, wa.y, acc0), acc0 = fmaf(va.z, wa.z, acc0), acc0 = fmaf(va.w, wa.w, acc0);
        acc1 = fmaf(vb.x, wb.x, acc1), acc1 = fmaf(vb.y, wb.y, acc1), acc1 = fmaf(vb.z, wb.z, acc1), acc1 = fmaf(vb.w, wb.w, acc1);
      }
    float acc = acc0 + acc1;
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    if (sub == 0) out[px] = acc + __ldg(skip + px);
  }
}

// ---- BN-ReLU-DW(dil)-PW block on CLP, pointwise product on tcgen05 -------------------------------------------------------
struct DwTcArgs {
  const float* in;    // CLP [B][R][32] post-activation
  float* out;         // CLP [B][R][32]
  const float* dw;    // [32][9]
  const float* pwtc;  // [64][32]: rows 0..31 = tf32-truncated folded pointwise weights (row = cout, col = cin), 32..63 = remainder
  const float* bias;  // [32]
  int R, Hp, Wp, dil, relu;
  int nxt, segs, seg_len, total_items;  // strip schedule: item = (b, x tile, row phase, segment of seg_len phase-rows)
};
constexpr int DT_THREADS = 512;
constexpr int DT_SMEM = 4 * 16384 + 8192 + 1024 + 64;

__device__ __forceinline__ uint64_t dt_sdesc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void dt_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
}

__device__ __forceinline__ void dt_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[j]);
}

__global__ void __launch_bounds__(DT_THREADS, 1) dwsep_tc_kernel(const DwTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;              // [2][16384] depthwise result (= xh for the MMA)
  uint8_t* sL = smem + 32768;      // [2][16384] x - trunc(x)
  uint8_t* sB = smem + 65536;      // [64][128 B] pointwise operand, SWIZZLE_128B
  uint64_t* mma_bar = reinterpret_cast<uint64_t*>(smem + 65536 + 8192);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_bar + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    mbar_init(mma_bar, 1);
    mbar_init(mma_bar + 1, 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int idx = tid; idx < 64 * 8; idx += DT_THREADS) {
    const int n = idx >> 3, c = idx & 7;
    *reinterpret_cast<float4*>(sB + n * 128 + ((c ^ (n & 7)) << 4)) = __ldg(reinterpret_cast<const float4*>(a.pwtc + n * 32 + c * 4));
  }
  fence_proxy_async_smem();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  const int q = tid & 7;  // channel quad of the depthwise phase
  float4 kq[9];
#pragma unroll
  for (int t = 0; t < 9; ++t)
    kq[t] = make_float4(__ldg(a.dw + (q * 4 + 0) * 9 + t), __ldg(a.dw + (q * 4 + 1) * 9 + t), __ldg(a.dw + (q * 4 + 2) * 9 + t),
                        __ldg(a.dw + (q * 4 + 3) * 9 + t));
  const int quarter = warp & 3, cgrp = warp >> 2;  // epilogue: TMEM lane quarter / group of 8 channels of this warp
  float bias[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) bias[j] = __ldg(a.bias + cgrp * 8 + j);
  const uint32_t idesc64 = (1u << 4) | (2u << 7) | (2u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
  const uint32_t idesc32 = (1u << 4) | (2u << 7) | (2u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
  const int dil = a.dil, Wp = a.Wp, Hp = a.Hp, R = a.R;
  const long long tap_step_y = (long long)dil * Wp * 32, tap_step_x = (long long)dil * 32;

  // tile = 128 consecutive pixels of one image line: rows r0 .. r0+127 of batch element b, of which the first `nval` exist
  auto epilogue = [&](int j, int b, int r0, int nval) {
    const int pbuf = j & 1;
    mbar_wait(mma_bar + pbuf, (j >> 1) & 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t taddr = tmem + ((uint32_t)(quarter * 32) << 16) + pbuf * 96 + cgrp * 8;
    float acc[8], t[8];
    dt_ld8(taddr, acc);
    dt_ld8(taddr + 32, t);
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] += t[c];
    dt_ld8(taddr + 64, t);
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] += t[c];
    const int p = quarter * 32 + lane;
    if (p < nval) {
      const int r = r0 + p;
      const int y = r / Wp, x = r - y * Wp;
      const bool border = x < RP || x >= Wp - RP || y < RP || y >= Hp - RP;
      const float lo = a.relu ? 0.f : -INFINITY;
      float4* o = reinterpret_cast<float4*>(a.out + ((long long)b * R + r) * 32 + cgrp * 8);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float4 v;
        v.x = border ? 0.f : fmaxf(acc[4 * c] + bias[4 * c], lo);
        v.y = border ? 0.f : fmaxf(acc[4 * c + 1] + bias[4 * c + 1], lo);
        v.z = border ? 0.f : fmaxf(acc[4 * c + 2] + bias[4 * c + 2], lo);
        v.w = border ? 0.f : fmaxf(acc[4 * c + 3] + bias[4 * c + 3], lo);
        o[c] = v;
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  };

  // Strip schedule: an item is a run of seg_len image lines of the same row phase (y = py + i*dil) in one 128-pixel
  // column tile, walked top to bottom, so two of the three tap rows of every tile are L1 hits from the previous tile.
  int i = 0, pb = 0, pr0 = 0, pnval = 0;
  for (int item = blockIdx.x; item < a.total_items; item += gridDim.x) {
    int t = item;
    const int seg = t % a.segs;
    t /= a.segs;
    const int py = t % dil;
    t /= dil;
    const int xt = t % a.nxt;
    const int b = t / a.nxt;
    const int nval = min(128, Wp - xt * 128);
    for (int iy = seg * a.seg_len; iy < (seg + 1) * a.seg_len; ++iy) {
      const int y = py + iy * dil;
      if (y >= Hp) break;
      const int buf = i & 1;
      const int r0 = y * Wp + xt * 128;
      const bool yin = y >= RP && y < Hp - RP;
      // ---- depthwise phase: 128 pixels x 8 channel quads, 2 items per thread; loads are unconditional (border pixels
      //      read their own row and are zeroed afterwards) ----
      float4 v[2][9];
      bool inside[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {  // issue all 18 tap loads of this thread before touching any of them: one round trip
        const int p = (tid >> 3) + 64 * j;
        const int x = xt * 128 + p;
        inside[j] = yin && x >= RP && x < Wp - RP;
        const long long sy = inside[j] ? tap_step_y : 0, sx = inside[j] ? tap_step_x : 0;
        const float* base = a.in + ((long long)b * R + (p < nval ? r0 + p : r0)) * 32 + q * 4;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx)
            v[j][ky * 3 + kx] = __ldg(reinterpret_cast<const float4*>(base + (ky - 1) * sy + (kx - 1) * sx));
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int p = (tid >> 3) + 64 * j;
        float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const float4 k = kq[t];
          d.x = fmaf(v[j][t].x, k.x, d.x), d.y = fmaf(v[j][t].y, k.y, d.y), d.z = fmaf(v[j][t].z, k.z, d.z),
          d.w = fmaf(v[j][t].w, k.w, d.w);
        }
        if (!inside[j]) d = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 l;
        l.x = d.x - __uint_as_float(__float_as_uint(d.x) & 0xFFFFE000u);
        l.y = d.y - __uint_as_float(__float_as_uint(d.y) & 0xFFFFE000u);
        l.z = d.z - __uint_as_float(__float_as_uint(d.z) & 0xFFFFE000u);
        l.w = d.w - __uint_as_float(__float_as_uint(d.w) & 0xFFFFE000u);
        const int addr = buf * 16384 + p * 128 + ((q ^ (p & 7)) << 4);
        *reinterpret_cast<float4*>(sA + addr) = d;
        *reinterpret_cast<float4*>(sL + addr) = l;
      }
      fence_proxy_async_smem();
      __syncthreads();
      if (warp == 0 && elect_one_sync()) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_addr = smem_u32(sA + buf * 16384), l_addr = smem_u32(sL + buf * 16384), b_addr = smem_u32(sB);
        const uint32_t d_hh = tmem + buf * 96, d_lh = d_hh + 64;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t db = dt_sdesc(b_addr + k * 32);
          const uint32_t accf = k > 0;
          asm volatile(
              "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
              "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(d_hh),
              "l"(dt_sdesc(a_addr + k * 32)), "l"(db), "r"(idesc64), "r"(accf)
              : "memory");
          asm volatile(
              "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
              "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}" ::"r"(d_lh),
              "l"(dt_sdesc(l_addr + k * 32)), "l"(db), "r"(idesc32), "r"(accf)
              : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mma_bar + buf))
                     : "memory");
      }
      if (i > 0) epilogue(i - 1, pb, pr0, pnval);
      pb = b, pr0 = r0, pnval = nval;
      ++i;
    }
  }
  if (i > 0) epilogue(i - 1, pb, pr0, pnval);

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

static int launch_dwsep_tc(DwTcArgs a, int B, cudaStream_t st) {
  if (a.dil < 1 || a.dil > RP) return LWS_ERR_UNSUPPORTED;
  cudaError_t e = cudaFuncSetAttribute(dwsep_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DT_SMEM);
  if (e != cudaSuccess) return (int)e;
  a.nxt = (a.Wp + 127) / 128;
  const int rows_per_phase = (a.Hp + a.dil - 1) / a.dil;
  a.seg_len = 16;
  a.segs = (rows_per_phase + a.seg_len - 1) / a.seg_len;
  a.total_items = B * a.nxt * a.dil * a.segs;
  const int grid = a.total_items < kNumSMs ? a.total_items : kNumSMs;
  dwsep_tc_kernel<<<grid, DT_THREADS, DT_SMEM, st>>>(a);
  e = cudaPeekAtLastError();
  return e == cudaSuccess ? LWS_OK : (int)e;
}

// ---- host ------------------------------------------------------------------------------------------------------------------
struct RefTcWeights {
  const float *w0[2], *b0[2];              // first convs
  const float *dw[3][4], *pwtc[3][4], *bias[3][4];  // [left, disp, r2][block]
  const float *dense_tc, *dense_bias, *last_w;
};

size_t refinement_tc_workspace_bytes(int B, int H, int W) {
  return (size_t)4 * B * (H + 2 * RP) * (W + 2 * RP) * 32 * sizeof(float);
}

int refinement_tc(const float* left, const float* pred3, const RefTcWeights& wt, float* pred4, void* ws, int B, int H,
                  int W, cudaStream_t st) {
  const int Hp = H + 2 * RP, Wp = W + 2 * RP;
  const long long R = (long long)Hp * Wp;
  if (R >= (1ll << 31) - 65536 || B * R * 32 >= (1ll << 40)) return LWS_ERR_BAD_SHAPE;
  const long long buf_floats = (long long)B * R * 32;
  float* catL = (float*)ws;
  float* catD = catL + buf_floats;
  float* ping = catD + buf_floats;
  float* pong = ping + buf_floats;
  static const int r1_dil[4] = {2, 4, 8, 16};
  static const int r2_dil[4] = {8, 4, 2, 1};
  const long long rows = (long long)B * R;
  const int cblocks = (int)((rows * 4 + 255) / 256 < 148 * 16 ? (rows * 4 + 255) / 256 : 148 * 16);
  int rc;
  cudaError_t e;
  for (int br = 0; br < 2; ++br) {
    if (br == 0) ref_conv0_clp_kernel<3><<<cblocks, 256, 0, st>>>(left, wt.w0[0], wt.b0[0], ping, H, W, rows);
    else ref_conv0_clp_kernel<1><<<cblocks, 256, 0, st>>>(pred3, wt.w0[1], wt.b0[1], ping, H, W, rows);
    if ((e = cudaPeekAtLastError()) != cudaSuccess) return (int)e;
    float* cur = ping;
    float* nxt = pong;
    for (int j = 0; j < 4; ++j) {
      DwTcArgs d;
      memset(&d, 0, sizeof(d));
      d.in = cur, d.out = j < 3 ? nxt : (br == 0 ? catL : catD);
      d.dw = wt.dw[br][j], d.pwtc = wt.pwtc[br][j], d.bias = wt.bias[br][j];
      d.R = (int)R, d.Hp = Hp, d.Wp = Wp, d.dil = r1_dil[j], d.relu = 1;
      if ((rc = launch_dwsep_tc(d, B, st))) return rc;
      float* t = cur;
      cur = nxt, nxt = t;
    }
  }
  {
    int st_off[6], st_src[6];
    for (int s = 0; s < 6; ++s) st_off[s] = (s % 3 - 1) * 8 * Wp - 8, st_src[s] = s / 3;
    if ((rc = launch_tc_implicit_gemm(catL, catD, wt.dense_tc, wt.dense_bias, ping, B, (int)R, Hp, Wp, RP, H, W, 32, 6, st_off,
                                      st_src, 8, 1, st)))
      return rc;
  }
  float* cur = ping;
  float* nxt = pong;
  for (int j = 0; j < 4; ++j) {
    DwTcArgs d;
    memset(&d, 0, sizeof(d));
    d.in = cur, d.out = nxt, d.dw = wt.dw[2][j], d.pwtc = wt.pwtc[2][j], d.bias = wt.bias[2][j];
    d.R = (int)R, d.Hp = Hp, d.Wp = Wp, d.dil = r2_dil[j], d.relu = j < 3;
    if ((rc = launch_dwsep_tc(d, B, st))) return rc;
    float* t = cur;
    cur = nxt, nxt = t;
  }
  const long long px = (long long)B * H * W;
  const int lblocks = (int)((px * 4 + 255) / 256 < 148 * 16 ? (px * 4 + 255) / 256 : 148 * 16);
  ref_last_clp_kernel<<<lblocks, 256, 0, st>>>(cur, wt.last_w, pred3, pred4, H, W, px);
  if ((e = cudaPeekAtLastError()) != cudaSuccess) return (int)e;
  return LWS_OK;
}

}  // namespace lws
