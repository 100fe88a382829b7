// K6 on the tensor cores: the colour-guidance refinement (reference models/submodules.py:223-327, models/models.py:158-162)
// in channels-last "CLP" layout  act[b][y][x][32] fp32 with a 16-pixel zero border (the largest dilation), so every
// 32-channel pixel is one 128-byte row and every dilated tap is a constant row offset.
//   * BN-ReLU-DW(dil)-PW block  -> dwsep_f16_kernel<0> (dwsep_tc.cu): TMA-fed depthwise on the CUDA cores with the 3x3 window in
//     registers, pointwise product on tcgen05 with split-fp16 operands.
//   * 3->32 / 1->32 first convs -> dwsep_f16_kernel<3 / 1>: im2col front end, the conv weight is the "pointwise" operand.
//   * dense 64->32 dilation-8 3x3 -> the Toeplitz-N GEMM of conv3d_f16.cu: 6 stages (3 kh x 2 sources), kw folded into N = 192
//     with an 8-row epilogue shift; the concat is never formed (the two refinement1 branches are the two sources).
//   * closing 32->1 conv (+ skip) -> the N = 16 "last layer" form of the same GEMM, reading the split-fp16 rows of block 4.
#include <math.h>
#include <string.h>

#include "lws_common.cuh"
#include "tma_utils.cuh"
#include "conv3d_f16.cuh"

namespace lws {

constexpr int RP = 16;  // border of the refinement CLP tensors

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
  int rc;
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
