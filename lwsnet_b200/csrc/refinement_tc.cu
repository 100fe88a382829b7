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

#include "dwsep_common.cuh"
#include "conv3d_f16.cuh"

namespace lws {

constexpr int RP = 16;  // border of the refinement CLP tensors

// ---- host ------------------------------------------------------------------------------------------------------------------
struct RefTcWeights {
  const float *w0[2], *b0[2];              // first convs
  const float *dw[3][4], *pwtc[3][4], *bias[3][4];  // [left, disp, r2][block]
  const float *dense_tc, *dense_bias, *last_w;
  const float* w0tc[2];
  const float* last_tc;
};

static const int kR1Dil[4] = {2, 4, 8, 16};
static const int kR2Dil[4] = {8, 4, 2, 1};

// Blocks per chain launch (option "refine_chain": 0 / 1 = one block per launch).  Chains pay a pipeline fill of a few bands, so a
// launch with few (pair, band) units stays on the per-block kernel.
static int chain_len(int B, int H) {
  int n = opt(OPT_REFINE_CHAIN);
  if (n < 2) return 1;
  if (n == 3) n = 2;  // four blocks split 2 + 2 or run as one chain of 4
  if ((long long)B * ((H + 63) / 64) < opt(OPT_CHAIN_MIN_BANDS)) return 1;
  return n;
}
static size_t chain_ws_bytes(int B, int H, int W) {
  size_t m = 0;
  for (int n = 2; n <= 4; n += 2)
    for (int br = 0; br < 2; ++br)
      for (int j = 0; j + n <= 4; j += n) {
        const size_t b = dwsep_chain_workspace_bytes((br ? kR2Dil : kR1Dil) + j, n, B, H, W);
        m = b > m ? b : m;
      }
  return (m + 255) / 256 * 256;
}
static size_t clp_bytes(int B, int H, int W) {
  return ((size_t)B * (H + 2 * RP) * (W + 2 * RP) * 32 * sizeof(float) + 255) / 256 * 256;
}

int refinement_tc_launches(int B, int H) { return 4 + 12 / chain_len(B, H); }

size_t refinement_tc_workspace_bytes(int B, int H, int W) { return 4 * clp_bytes(B, H, W) + chain_ws_bytes(B, H, W); }

// blocks [j0, j0 + n) of branch br (0 / 1 = refinement1_left / _disp, 2 = refinement2): src -> dst, one block per launch through
// `tmp` ping-pong buffers or as L2-resident chains
static int run_blocks(const RefTcWeights& wt, int br, const float* src, float* dst, float* tmp0, float* tmp1, void* chain_ws,
                      size_t chain_bytes, int B, int H, int W, bool last_relu, bool last_split, cudaStream_t st) {
  const int* dil = br < 2 ? kR1Dil : kR2Dil;
  const int n = chain_len(B, H);
  int rc;
  const float* cur = src;
  for (int j = 0; j < 4; j += n) {
    float* out = j + n >= 4 ? dst : (cur == tmp0 ? tmp1 : tmp0);
    if (n == 1) {
      const bool last = j == 3;
      rc = launch_dwsep_f16(cur, out, wt.dw[br][j], wt.pwtc[br][j], wt.pwtc[br][j] + 1024, wt.bias[br][j], B, H, W, dil[j],
                            last ? last_relu : 1, last ? last_split : 0, st);
    } else {
      ChainBlockDesc blk[4];
      for (int k = 0; k < n; ++k) {
        const bool last = j + k == 3;
        blk[k].dw = wt.dw[br][j + k], blk[k].pwh = wt.pwtc[br][j + k], blk[k].scales = wt.pwtc[br][j + k] + 1024;
        blk[k].bias = wt.bias[br][j + k], blk[k].dil = dil[j + k], blk[k].relu = last ? last_relu : 1;
        blk[k].out_split = last ? last_split : 0;
      }
      rc = launch_dwsep_chain(cur, out, blk, n, chain_ws, chain_bytes, B, H, W, st);
    }
    if (rc) return rc;
    cur = out;
  }
  return LWS_OK;
}

int refinement_tc(const float* left, const float* pred3, const RefTcWeights& wt, float* pred4, void* ws, int B, int H,
                  int W, cudaStream_t st) {
  const int Hp = H + 2 * RP, Wp = W + 2 * RP;
  const long long R = (long long)Hp * Wp;
  if (R >= (1ll << 31) - 65536 || B * R * 32 >= (1ll << 40)) return LWS_ERR_BAD_SHAPE;
  const size_t cb = clp_bytes(B, H, W);
  float* catL = (float*)ws;
  float* catD = (float*)((char*)ws + cb);
  float* ping = (float*)((char*)ws + 2 * cb);
  float* pong = (float*)((char*)ws + 3 * cb);
  void* chain_ws = (char*)ws + 4 * cb;
  const size_t chain_bytes = chain_ws_bytes(B, H, W);
  int rc;
  for (int br = 0; br < 2; ++br) {
    // first conv -> pong; blocks 1..4 -> the branch's half of the concat (split-fp16 rows: it feeds the dense conv)
    if ((rc = launch_conv0_f16(br == 0 ? left : pred3, pong, wt.w0tc[br], wt.w0tc[br] + 1024, wt.b0[br], B, br == 0 ? 3 : 1, H, W, st)))
      return rc;
    if ((rc = run_blocks(wt, br, pong, br == 0 ? catL : catD, ping, pong, chain_ws, chain_bytes, B, H, W, true, true, st))) return rc;
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
    L.seg_mode = (opt(OPT_TZ_STRIPS) & 3) == 2;
    if ((rc = launch_tz_gemm(L, st))) return rc;
  }
  // blocks 1..4 of refinement2: ping -> catL (free by now); the last block has no ReLU and feeds the closing conv (split-fp16 rows)
  if ((rc = run_blocks(wt, 2, ping, catL, pong, catD, chain_ws, chain_bytes, B, H, W, false, true, st))) return rc;
  float* cur = catL;
  {
    // closing 32 -> 1 conv + skip (pred4 = pred3 + r): stages = kh, kw folded into N = 16, strips down the image
    TzLayer L;
    memset(&L, 0, sizeof(L));
    L.src0 = L.src1 = cur, L.wtab = wt.last_tc, L.B = B, L.R = (int)R;
    L.n0 = Wp, L.p0 = RP, L.i0 = W, L.n1 = Hp, L.p1 = RP, L.i1 = H;
    L.tz = 1, L.nstages = 3, L.nshift = 1, L.box_rows = 128, L.G = 1, L.srow = Wp;
    for (int s = 0; s < 3; ++s) L.st_off[s] = (s - 1) * Wp, L.st_src[s] = 0;
    L.last = 1, L.skip = pred3, L.out_f32 = pred4, L.out_mode = 1;
    L.seg_mode = (opt(OPT_TZ_STRIPS) & 3) == 2;
    if ((rc = launch_tz_gemm(L, st))) return rc;
  }
  return LWS_OK;
}

}  // namespace lws
