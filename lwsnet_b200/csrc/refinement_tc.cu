// K6 on the tensor cores: the colour-guidance refinement (reference models/submodules.py:223-327, models/models.py:158-162)
// in channels-last "CLP" layout  act[b][y][x][32] fp32 with a 16-pixel zero border (the largest dilation), so every
// 32-channel pixel is one 128-byte row and every dilated tap is a constant row offset.
//   * BN-ReLU-DW(dil)-PW block  -> dwsep_f16_kernel (dwsep_tc.cu): TMA-fed depthwise on the CUDA cores with the 3x3 window in
//     registers, pointwise product on tcgen05 with split-fp16 operands.
//   * dense 64->32 dilation-8 3x3 -> the implicit-GEMM kernel of conv3d_tc.cu with 6 stages (2 sources x 3 kh), kw taps 8
//     rows apart in the stage tile; the concat is never formed (the two refinement1 branches are the two sources).
//   * 3->32 / 1->32 first convs and the 32->1 last conv (+ skip) are small FP32 kernels reading / writing CLP.
#include <math.h>
#include <string.h>

#include "lws_common.cuh"
#include "tma_utils.cuh"
#include "conv3d_f16.cuh"

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

// BN-ReLU-DW(dil)-PW block on CLP (dwsep_tc.cu)
int launch_conv0_f16(const float* img, float* out, const void* wtab, const float* scales, const float* bias, int B, int CIN, int H,
                     int W, cudaStream_t st);
int launch_dwsep_f16(const float* in, float* out, const float* dw, const void* pwh, const float* scales, const float* bias, int B,
                     int H, int W, int dil, int relu, int out_split, cudaStream_t st);

// ---- host ------------------------------------------------------------------------------------------------------------------
struct RefTcWeights {
  const float *w0[2], *b0[2];              // first convs
  const float *dw[3][4], *pwtc[3][4], *bias[3][4];  // [left, disp, r2][block]
  const float *dense_tc, *dense_bias, *last_w;
  const float* w0tc[2];
  const float* last_tc;
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
    if ((rc = launch_conv0_f16(br == 0 ? left : pred3, ping, wt.w0tc[br], wt.w0tc[br] + 1024, wt.b0[br], B, br == 0 ? 3 : 1, H, W, st)))
      return rc;
    float* cur = ping;
    float* nxt = pong;
    for (int j = 0; j < 4; ++j) {
      float* dst = j < 3 ? nxt : (br == 0 ? catL : catD);
      if ((rc = launch_dwsep_f16(cur, dst, wt.dw[br][j], wt.pwtc[br][j], wt.pwtc[br][j] + 1024, wt.bias[br][j], B, H, W,
                                 r1_dil[j], 1, j == 3 /* the concat halves feed the dense conv: split-fp16 rows */, st)))
        return rc;
      float* t = cur;
      cur = nxt, nxt = t;
    }
  }
  {
    // dense 64 -> 32, dilation 8: stages = (kh, source), kw folded into N (Toeplitz shift 8 pixels), strips down the image
    TzLayer L;
    memset(&L, 0, sizeof(L));
    L.src0 = catL, L.src1 = catD, L.wtab = wt.dense_tc, L.bias = wt.dense_bias, L.out = ping, L.B = B, L.R = (int)R;
    L.n0 = Wp, L.p0 = RP, L.i0 = W, L.n1 = Hp, L.p1 = RP, L.i1 = H;
    L.tz = 8, L.nstages = 6, L.nshift = 1, L.box_rows = 128, L.G = 2, L.srow = 8 * Wp;
    for (int s = 0; s < 6; ++s) L.st_off[s] = (s / 2 - 1) * 8 * Wp, L.st_src[s] = s & 1;
    L.out_split = 0, L.relu = 1;
    if ((rc = launch_tz_gemm(L, st))) return rc;
  }
  float* cur = ping;
  float* nxt = pong;
  for (int j = 0; j < 4; ++j) {
    if ((rc = launch_dwsep_f16(cur, nxt, wt.dw[2][j], wt.pwtc[2][j], wt.pwtc[2][j] + 1024, wt.bias[2][j], B, H, W, r2_dil[j],
                               j < 3, j == 3 /* the last block feeds the closing conv: split-fp16 rows */, st)))
      return rc;
    float* t = cur;
    cur = nxt, nxt = t;
  }
  {
    // closing 32 -> 1 conv + skip (pred4 = pred3 + r): stages = kh, kw folded into N = 16, strips down the image
    TzLayer L;
    memset(&L, 0, sizeof(L));
    L.src0 = L.src1 = cur, L.wtab = wt.last_tc, L.B = B, L.R = (int)R;
    L.n0 = Wp, L.p0 = RP, L.i0 = W, L.n1 = Hp, L.p1 = RP, L.i1 = H;
    L.tz = 1, L.nstages = 3, L.nshift = 1, L.box_rows = 128, L.G = 1, L.srow = Wp;
    for (int s = 0; s < 3; ++s) L.st_off[s] = (s - 1) * Wp, L.st_src[s] = 0;
    L.last = 1, L.skip = pred3, L.out_f32 = pred4, L.out_mode = 1;
    if ((rc = launch_tz_gemm(L, st))) return rc;
  }
  return LWS_OK;
}

}  // namespace lws
