// K6: colour-guidance refinement  pred4 = pred3 + R2(concat[R1_left(left), R1_disp(pred3)])
// (reference models/submodules.py:223-327 preconv2d / preconv2d_depthseperated / refinement1 / refinement2,
//  called at models/models.py:158-162; ~45 full-resolution cuDNN/elementwise launches in the reference, 16 here).
//
// BatchNorm is folded on the host exactly as for the 3D stack: every kernel stores ReLU(BN_next(conv(.))), the
// concat is never materialised separately (the two R1 branches write their halves of the 64-channel buffer), and the
// final interpolate (same size => identity) is dropped.
//   dense 3x3 convs  : conv2d.cuh (FP32 direct, smem tiled)
//   BN-ReLU-DW-PW    : dwsep_block_kernel below (phase-row tiling, depthwise result kept in registers)
#include <cuda_fp16.h>
#include <math.h>
#include <string.h>

#include <stdlib.h>

#include "conv2d.cuh"

namespace lws {

// refinement_tc.cu: channels-last tensor-core path
struct RefTcWeights {
  const float *w0[2], *b0[2];
  const float *dw[3][4], *pwtc[3][4], *bias[3][4];
  const float *dense_tc, *dense_bias, *last_w;
  const float* w0tc[2];
  const float* last_tc;
};
size_t refinement_tc_workspace_bytes(int B, int H, int W);
int refinement_tc_launches(int B, int H);
int refinement_tc(const float* left, const float* pred3, const RefTcWeights& wt, float* pred4, void* ws, int B, int H,
                  int W, cudaStream_t st);

struct DwsepArgs {
  const float* in;    // [B,32,H,W] post-activation
  const float* dw;    // [32][9]
  const float* pw;    // [32 ci][32 co] (next BN scale folded)
  const float* bias;  // [32]
  float* out;
  long long in_bs, out_bs;
  int H, W, dil, relu;
};

// BN-ReLU-DW(dil)-PW block.  "Phase-row" tiling: a block owns 8 output rows spaced `dil` apart (same row phase) x 128
// consecutive columns, so the dilated 3x3 needs only 10 input rows and a +-dil column halo whatever the dilation is
// (1.56x read amplification at dil 16 instead of 9x for a square tile).  A thread owns 2 vertically adjacent
// phase-rows x 1 column x all 32 output channels: the depthwise result never leaves registers, the 32x32 pointwise
// weights are warp-uniform broadcast LDS.128.  Input channels stream through shared memory in cp.async double-buffered
// chunks of 8.
constexpr int DW_TI = 8, DW_TX = 128, DW_C = 32, DW_CK = 8, DW_THREADS = 512, DW_ROWS = DW_TI + 2;
constexpr int DW_PWMAX = DW_TX + 2 * 16;

__global__ void __launch_bounds__(DW_THREADS, 1) dwsep_block_kernel(const DwsepArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* sIn = smem;                                       // [2][CK][ROWS][pw]
  const int dil = a.dil;
  const int pw = (DW_TX + 2 * dil + 3) & ~3;
  float* sPW = smem + 2 * DW_CK * DW_ROWS * DW_PWMAX;      // [32][32]
  float* sK = sPW + DW_C * DW_C;                           // [32][12]
  const int tid = threadIdx.x;
  const int tiles_x = (a.W + DW_TX - 1) / DW_TX;
  const int x0 = (blockIdx.x % tiles_x) * DW_TX;
  const int rb = blockIdx.x / tiles_x;
  const int grp = rb / dil, py = rb - grp * dil;
  const int b = blockIdx.y;
  const int H = a.H, W = a.W;
  const long long hw = (long long)H * W;
  const float* in_b = a.in + (long long)b * a.in_bs;
  const bool vec16 = ((dil & 3) == 0) && ((W & 3) == 0) && ((((uintptr_t)a.in) & 15) == 0) && ((a.in_bs & 3) == 0);

  for (int i = tid; i < DW_C * DW_C; i += DW_THREADS) sPW[i] = __ldg(a.pw + i);
  for (int i = tid; i < DW_C * 12; i += DW_THREADS) sK[i] = (i % 12 < 9) ? __ldg(a.dw + (i / 12) * 9 + (i % 12)) : 0.f;

  // one warp per (channel, row) of the chunk: the row decode happens once per row, lanes stride over the columns
  const int warp = tid >> 5, lane = tid & 31;
  auto issue = [&](int chunk, int buf) {
    float* dst = sIn + buf * DW_CK * DW_ROWS * DW_PWMAX;
    const int c0 = chunk * DW_CK;
    for (int row = warp; row < DW_CK * DW_ROWS; row += DW_THREADS / 32) {
      const int cl = row / DW_ROWS, r = row - cl * DW_ROWS;
      const int gy = (grp * DW_TI + r - 1) * dil + py;
      float* d = dst + row * DW_PWMAX;
      const bool yok = gy >= 0 && gy < H;
      const float* src = in_b + (long long)(c0 + cl) * hw + (long long)(yok ? gy : 0) * W + (x0 - dil);
      if (vec16) {
        for (int e = lane * 4; e < pw; e += 128) {
          const int gx = x0 - dil + e;
          if (yok && gx >= 0 && gx < W) cp_async_16(d + e, src + e);
          else *reinterpret_cast<float4*>(d + e) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      } else {
        for (int e = lane; e < pw; e += 32) {
          const int gx = x0 - dil + e;
          if (yok && gx >= 0 && gx < W) cp_async_4(d + e, src + e);
          else d[e] = 0.f;
        }
      }
    }
    cp_async_commit();
  };

  const int lx = tid % DW_TX;
  const int i0 = (tid / DW_TX) * 2;  // first of this thread's two phase-rows
  float acc[2][DW_C];
#pragma unroll
  for (int p = 0; p < 2; ++p)
#pragma unroll
    for (int q = 0; q < DW_C; ++q) acc[p][q] = 0.f;

  constexpr int NCH = DW_C / DW_CK;
  issue(0, 0);
  for (int ch = 0; ch < NCH; ++ch) {
    if (ch + 1 < NCH) {
      issue(ch + 1, (ch + 1) & 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float* src = sIn + (ch & 1) * DW_CK * DW_ROWS * DW_PWMAX + i0 * DW_PWMAX + lx;
#pragma unroll 2
    for (int cl = 0; cl < DW_CK; ++cl) {
      const int ci = ch * DW_CK + cl;
      const float* rowp = src + cl * DW_ROWS * DW_PWMAX;
      float v[4][3];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        v[r][0] = rowp[r * DW_PWMAX];
        v[r][1] = rowp[r * DW_PWMAX + dil];
        v[r][2] = rowp[r * DW_PWMAX + 2 * dil];
      }
      const float4 k0 = *reinterpret_cast<const float4*>(sK + ci * 12);
      const float4 k1 = *reinterpret_cast<const float4*>(sK + ci * 12 + 4);
      const float4 k2 = *reinterpret_cast<const float4*>(sK + ci * 12 + 8);
      const float k[9] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w, k2.x};
      float d0 = 0.f, d1 = 0.f;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          d0 = fmaf(v[ky][kx], k[ky * 3 + kx], d0);
          d1 = fmaf(v[ky + 1][kx], k[ky * 3 + kx], d1);
        }
      const float4* wp = reinterpret_cast<const float4*>(sPW + ci * DW_C);
#pragma unroll
      for (int q4 = 0; q4 < DW_C / 4; ++q4) {
        const float4 w4 = wp[q4];
        acc[0][q4 * 4] = fmaf(d0, w4.x, acc[0][q4 * 4]);
        acc[0][q4 * 4 + 1] = fmaf(d0, w4.y, acc[0][q4 * 4 + 1]);
        acc[0][q4 * 4 + 2] = fmaf(d0, w4.z, acc[0][q4 * 4 + 2]);
        acc[0][q4 * 4 + 3] = fmaf(d0, w4.w, acc[0][q4 * 4 + 3]);
        acc[1][q4 * 4] = fmaf(d1, w4.x, acc[1][q4 * 4]);
        acc[1][q4 * 4 + 1] = fmaf(d1, w4.y, acc[1][q4 * 4 + 1]);
        acc[1][q4 * 4 + 2] = fmaf(d1, w4.z, acc[1][q4 * 4 + 2]);
        acc[1][q4 * 4 + 3] = fmaf(d1, w4.w, acc[1][q4 * 4 + 3]);
      }
    }
    __syncthreads();
  }

  const int gx = x0 + lx;
  if (gx >= W) return;
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const int gy = (grp * DW_TI + i0 + p) * dil + py;
    if (gy >= H) continue;
    float* o = a.out + (long long)b * a.out_bs + (long long)gy * W + gx;
#pragma unroll
    for (int q = 0; q < DW_C; ++q) {
      float r = acc[p][q] + __ldg(a.bias + q);
      if (a.relu) r = fmaxf(r, 0.f);
      o[q * hw] = r;
    }
  }
}

constexpr size_t DW_SMEM = (size_t)(2 * DW_CK * DW_ROWS * DW_PWMAX + DW_C * DW_C + DW_C * 12) * sizeof(float);

static int launch_dwsep(DwsepArgs a, int B, cudaStream_t st) {
  if (a.dil < 1 || a.dil > 16) return LWS_ERR_UNSUPPORTED;
  LWS_SET_SMEM_ONCE(dwsep_block_kernel, DW_SMEM);
  cudaError_t e;
  dim3 grid(cdiv(a.W, DW_TX) * cdiv(a.H, DW_TI * a.dil) * a.dil, B);
  dwsep_block_kernel<<<grid, DW_THREADS, DW_SMEM, st>>>(a);
  e = cudaPeekAtLastError();
  return e == cudaSuccess ? LWS_OK : (int)e;
}

// ---- packed layout -------------------------------------------------------------------------------------------
struct RefLayout {
  size_t r1_w0[2], r1_b0[2];            // conv0 [Cin][9][32], bias[32]      (0: left Cin=3, 1: disp Cin=1)
  size_t r1_dw[2][4], r1_pw[2][4], r1_b[2][4];
  size_t r2_w, r2_b;                    // [64][9][32], bias[32]
  size_t r2_dw[4], r2_pw[4], r2_bb[4];
  size_t last_w;                        // [32][9][1]
  // tensor-core operand tables (refinement_tc.cu): pointwise [64][32] per block, dense conv [6*192][32]
  size_t r1_pwtc[2][4], r2_pwtc[4], r2_wtc;
  size_t last_tc;                       // closing 32 -> 1 conv as an N = 16 Toeplitz operand table (3 blocks x 16 rows x 128 B + scales)
  size_t r1_w0tc[2];                    // first convs as [64][32] split-fp16 operand tables (+ scales), dwsep_tc.cu
  size_t total;
};
static RefLayout ref_layout() {
  RefLayout L;
  size_t off = 0;
  auto take = [&](size_t n) {
    size_t o = off;
    off += (n + 3) / 4 * 4;
    return o;
  };
  for (int br = 0; br < 2; ++br) {
    L.r1_w0[br] = take((br == 0 ? 3 : 1) * 9 * 32);
    L.r1_b0[br] = take(32);
    for (int j = 0; j < 4; ++j) L.r1_dw[br][j] = take(32 * 9), L.r1_pw[br][j] = take(32 * 32), L.r1_b[br][j] = take(32);
  }
  L.r2_w = take(64 * 9 * 32);
  L.r2_b = take(32);
  for (int j = 0; j < 4; ++j) L.r2_dw[j] = take(32 * 9), L.r2_pw[j] = take(32 * 32), L.r2_bb[j] = take(32);
  L.last_w = take(32 * 9);
  for (int br = 0; br < 2; ++br)
    for (int j = 0; j < 4; ++j) L.r1_pwtc[br][j] = take(64 * 32);
  for (int j = 0; j < 4; ++j) L.r2_pwtc[j] = take(64 * 32);
  L.r2_wtc = take(6 * 192 * 32);
  for (int br = 0; br < 2; ++br) L.r1_w0tc[br] = take(64 * 32);
  L.last_tc = take(3 * 16 * 32 + 4);
  L.total = off;
  return L;
}

}  // namespace lws

extern "C" size_t lws_refinement_packed_floats(void) { return lws::ref_layout().total; }

extern "C" int lws_pack_refinement_weights(const float* const* t, int n_tensors, float eps, float* packed) {
  using namespace lws;
  if (!t || !packed) return LWS_ERR_NULL_PTR;
  if (n_tensors != 80) return LWS_ERR_BAD_SHAPE;
  for (int i = 0; i < n_tensors; ++i)
    if (!t[i]) return LWS_ERR_NULL_PTR;
  const RefLayout L = ref_layout();
  memset(packed, 0, L.total * sizeof(float));
  struct BN {
    const float *w, *b, *m, *v;
  };
  auto scale = [&](const BN& bn, int c) { return (double)bn.w[c] / sqrt((double)bn.v[c] + (double)eps); };
  auto shift = [&](const BN& bn, int c) { return (double)bn.b[c] - (double)bn.m[c] * scale(bn, c); };
  // dense conv [Cout,Cin,3,3] -> [Cin][9][Cout] scaled per cout by `next` (or 1)
  auto pack_conv = [&](const float* w, int cout, int cin, const BN* next, int next_off, float* dst_w, float* dst_b) {
    for (int co = 0; co < cout; ++co) {
      const double s = next ? scale(*next, next_off + co) : 1.0;
      if (next && dst_b) dst_b[co] = (float)shift(*next, next_off + co);
      for (int ci = 0; ci < cin; ++ci)
        for (int k = 0; k < 9; ++k) dst_w[((size_t)ci * 9 + k) * cout + co] = (float)((double)w[((size_t)co * cin + ci) * 9 + k] * s);
    }
  };
  auto pack_pw = [&](const float* w, const BN* next, int next_off, float* dst_w, float* dst_b) {
    for (int co = 0; co < 32; ++co) {
      const double s = next ? scale(*next, next_off + co) : 1.0;
      dst_b[co] = next ? (float)shift(*next, next_off + co) : 0.f;
      for (int ci = 0; ci < 32; ++ci) dst_w[ci * 32 + co] = (float)((double)w[co * 32 + ci] * s);
    }
  };
  // tensor index helpers (see lws.h for the order)
  auto r1 = [&](int br, int i) { return t[br * 25 + i]; };                // 0: conv0; block j (1..4): 1+(j-1)*6 + {0..3 BN, 4 dw, 5 pw}
  auto r1bn = [&](int br, int j) { BN bn = {r1(br, 1 + (j - 1) * 6), r1(br, 2 + (j - 1) * 6), r1(br, 3 + (j - 1) * 6), r1(br, 4 + (j - 1) * 6)}; return bn; };
  auto r2 = [&](int i) { return t[50 + i]; };                              // 0..3 BN64, 4 conv, blocks: 5+(j-1)*6.., 29 last
  auto r2bn = [&](int j) { BN bn = {r2(5 + (j - 1) * 6), r2(6 + (j - 1) * 6), r2(7 + (j - 1) * 6), r2(8 + (j - 1) * 6)}; return bn; };
  const BN bn64 = {r2(0), r2(1), r2(2), r2(3)};

  for (int br = 0; br < 2; ++br) {
    const BN b1 = r1bn(br, 1);
    pack_conv(r1(br, 0), 32, br == 0 ? 3 : 1, &b1, 0, packed + L.r1_w0[br], packed + L.r1_b0[br]);
    for (int j = 1; j <= 4; ++j) {
      memcpy(packed + L.r1_dw[br][j - 1], r1(br, 5 + (j - 1) * 6), 32 * 9 * sizeof(float));
      if (j < 4) {
        const BN nb = r1bn(br, j + 1);
        pack_pw(r1(br, 6 + (j - 1) * 6), &nb, 0, packed + L.r1_pw[br][j - 1], packed + L.r1_b[br][j - 1]);
      } else {
        pack_pw(r1(br, 6 + (j - 1) * 6), &bn64, br * 32, packed + L.r1_pw[br][j - 1], packed + L.r1_b[br][j - 1]);
      }
    }
  }
  {
    const BN b1 = r2bn(1);
    pack_conv(r2(4), 32, 64, &b1, 0, packed + L.r2_w, packed + L.r2_b);
    for (int j = 1; j <= 4; ++j) {
      memcpy(packed + L.r2_dw[j - 1], r2(9 + (j - 1) * 6), 32 * 9 * sizeof(float));
      if (j < 4) {
        const BN nb = r2bn(j + 1);
        pack_pw(r2(10 + (j - 1) * 6), &nb, 0, packed + L.r2_pw[j - 1], packed + L.r2_bb[j - 1]);
      } else {
        pack_pw(r2(10 + (j - 1) * 6), nullptr, 0, packed + L.r2_pw[j - 1], packed + L.r2_bb[j - 1]);
      }
    }
    pack_conv(r2(29), 1, 32, nullptr, 0, packed + L.last_w, nullptr);
  }
  // 3xTF32 operand tables from the folded fp32 weights: hi = low 13 mantissa bits cleared, lo = w - hi
  auto split = [](float w, float* hi, float* lo) {
    uint32_t u;
    memcpy(&u, &w, 4);
    u &= 0xFFFFE000u;
    memcpy(hi, &u, 4);
    *lo = w - *hi;
  };
  // split-fp16 pointwise operand table (dwsep_tc.cu): [64][32] halves, rows 0..31 = fp16(w * sw), rows 32..63 =
  // fp16((w * sw - hi) * 2^11), sw = the power of two that puts max|w| into [256, 512); then the two epilogue scales
  // c0 = 1 / (act_scale * sw) and c1 = c0 * 2^-11 at float offsets 1024 / 1025 of the slot
  auto pack_pwtc = [&](const float* pwf /*[ci][co]*/, float* slot /*2048 floats*/) {
    float mx = 0.f;
    for (int i = 0; i < 32 * 32; ++i) mx = fmaxf(mx, fabsf(pwf[i]));
    int e = 0;
    if (mx > 0.f) frexpf(mx, &e);
    const float sw = mx > 0.f ? ldexpf(1.f, 9 - e) : 1.f;
    __half* h = reinterpret_cast<__half*>(slot);
    for (int co = 0; co < 32; ++co)
      for (int ci = 0; ci < 32; ++ci) {
        const float w = pwf[ci * 32 + co] * sw;
        const __half hi = __float2half_rn(w);
        h[co * 32 + ci] = hi;
        h[(32 + co) * 32 + ci] = __float2half_rn((w - __half2float(hi)) * 2048.f);
      }
    slot[1024] = 1.f / (kDwsepActScale * sw);
    slot[1025] = slot[1024] / 2048.f;
  };
  {
    // closing 32 -> 1 conv (conv3d_f16.cu, LAST layer): block kh = 16 rows x 64 halves [B1 | B2]; B1 row kw = wh, row 8 + kw = wl
    // (applied to the hi halves of a pixel), B2 row 8 + kw = wh (applied to the lo halves); scales follow the 3 blocks
    const float* wl_ = packed + L.last_w;  // [32][9]
    float* tc = packed + L.last_tc;
    float mx = 0.f;
    for (int i = 0; i < 32 * 9; ++i) mx = fmaxf(mx, fabsf(wl_[i]));
    int e = 0;
    if (mx > 0.f) frexpf(mx, &e);
    const float sw = mx > 0.f ? ldexpf(1.f, 9 - e) : 1.f;
    __half* h = reinterpret_cast<__half*>(tc);
    for (int kh = 0; kh < 3; ++kh)
      for (int kw = 0; kw < 3; ++kw)
        for (int ci = 0; ci < 32; ++ci) {
          const float w = wl_[ci * 9 + kh * 3 + kw] * sw;
          const __half hi = __float2half_rn(w);
          __half* blk = h + (size_t)kh * 16 * 64;
          blk[kw * 64 + ci] = hi;
          blk[(8 + kw) * 64 + ci] = __float2half_rn((w - __half2float(hi)) * 2048.f);
          blk[(8 + kw) * 64 + 32 + ci] = hi;
        }
    tc[3 * 16 * 32] = 1.f / sw;
    tc[3 * 16 * 32 + 1] = 1.f / (sw * 2048.f);
  }
  for (int br = 0; br < 2; ++br) {
    // first conv [CIN][9][32] -> [K slot (zero padded to 32)][co]: the same table format as a pointwise conv.  Slots are ordered
    // by tap column c = ci*3 + kx with the three ky taps adjacent (slot = 3c + ky for c < 5, 16 + 3(c-5) + ky otherwise, slot 15
    // unused): the layout the im2col front end of dwsep_tc.cu slides down the image in registers
    float kxc[32 * 32];
    memset(kxc, 0, sizeof(kxc));
    const int cin0 = br == 0 ? 3 : 1;
    for (int ci = 0; ci < cin0; ++ci)
      for (int ky = 0; ky < 3; ++ky)
        for (int kx = 0; kx < 3; ++kx) {
          const int c = ci * 3 + kx, slot = c < 5 ? 3 * c + ky : 16 + 3 * (c - 5) + ky;
          for (int co = 0; co < 32; ++co) kxc[slot * 32 + co] = packed[L.r1_w0[br] + (size_t)(ci * 9 + ky * 3 + kx) * 32 + co];
        }
    pack_pwtc(kxc, packed + L.r1_w0tc[br]);
  }
  for (int br = 0; br < 2; ++br)
    for (int j = 0; j < 4; ++j) pack_pwtc(packed + L.r1_pw[br][j], packed + L.r1_pwtc[br][j]);
  for (int j = 0; j < 4; ++j) pack_pwtc(packed + L.r2_pw[j], packed + L.r2_pwtc[j]);
  {
    // dense 64 -> 32 conv as a split-fp16 Toeplitz-N operand table (conv3d_f16.cu): block = stage = kh*2 + source, two blocks
    // per 24 KB tile; tile row n = part*96 + kw*32 + co, 64 halves per row = [block 2i | block 2i+1] x ci 0..31
    const float* wf = packed + L.r2_w;  // [64][9][32]
    float* tc = packed + L.r2_wtc;
    float mx = 0.f;
    for (int i = 0; i < 64 * 9 * 32; ++i) mx = fmaxf(mx, fabsf(wf[i]));
    int e = 0;
    if (mx > 0.f) frexpf(mx, &e);
    const float sw = mx > 0.f ? ldexpf(1.f, 9 - e) : 1.f;
    __half* h = reinterpret_cast<__half*>(tc);
    for (int kh = 0; kh < 3; ++kh)
      for (int src = 0; src < 2; ++src) {
        const int blk = kh * 2 + src;
        for (int kw = 0; kw < 3; ++kw)
          for (int co = 0; co < 32; ++co)
            for (int k = 0; k < 32; ++k) {
              const float w = wf[((size_t)(src * 32 + k) * 9 + kh * 3 + kw) * 32 + co] * sw;
              const __half hi = __float2half_rn(w);
              const size_t base = ((size_t)(blk >> 1) * 192) * 64 + (blk & 1) * 32 + k;
              h[base + (size_t)(kw * 32 + co) * 64] = hi;
              h[base + (size_t)(96 + kw * 32 + co) * 64] = __float2half_rn((w - __half2float(hi)) * 2048.f);
            }
      }
    tc[3 * 192 * 32] = 1.f / sw;
    tc[3 * 192 * 32 + 1] = 1.f / (sw * 2048.f);
  }
  return LWS_OK;
}

extern "C" size_t lws_refinement_workspace_bytes(int B, int H, int W) {
  if (B <= 0 || H <= 0 || W <= 0) return 0;
  // FFMA path: concat (64 ch) + two 32-channel ping-pong buffers; tensor-core path: four bordered channels-last buffers
  const size_t ffma = (size_t)B * 128 * H * W * sizeof(float), tc = lws::refinement_tc_workspace_bytes(B, H, W);
  return ffma > tc ? ffma : tc;
}

namespace lws {
int launch_dwsep_f16(const float* in, float* out, const float* dw, const void* pwh, const float* scales, const float* bias, int B,
                     int H, int W, int dil, int relu, int out_split, cudaStream_t st);
struct ChainBlockDesc {
  const float* dw;
  const void* pwh;
  const float* scales;
  const float* bias;
  int dil, relu, out_split;
};
size_t dwsep_chain_workspace_bytes(const int* dil, int nblk, int B, int H, int W);
int launch_dwsep_chain(const float* in, float* out, const ChainBlockDesc* blocks, int nblk, void* ws, size_t ws_bytes, int B,
                       int H, int W, cudaStream_t st);
}
extern "C" int lws_refinement_launches(int B, int H, int W) {
  if (B <= 0 || H <= 0 || W <= 0) return 0;
  return lws::opt(lws::OPT_REFINE_TC) ? lws::refinement_tc_launches(B, H) : 16;
}
extern "C" size_t lws_refinement_clp_floats(int B, int H, int W) {
  if (B <= 0 || H <= 0 || W <= 0) return 0;
  return (size_t)B * (H + 32) * (W + 32) * 32;
}
extern "C" int lws_refinement_block_clp_f32(const float* in_clp, float* out_clp, const float* pk, int branch, int block, int B,
                                            int H, int W, lws_stream_t stream) {
  using namespace lws;
  LWS_CHECK_PTR(in_clp);
  LWS_CHECK_PTR(out_clp);
  LWS_CHECK_PTR(pk);
  if (B <= 0 || H <= 0 || W <= 0 || branch < 0 || branch > 2 || block < 0 || block > 3) return LWS_ERR_BAD_SHAPE;
  if ((((uintptr_t)in_clp) | ((uintptr_t)out_clp) | ((uintptr_t)pk)) & 15) return LWS_ERR_BAD_ALIGN;
  const RefLayout L = ref_layout();
  static const int r1_dil[4] = {2, 4, 8, 16}, r2_dil[4] = {8, 4, 2, 1};
  const float* dw = pk + (branch < 2 ? L.r1_dw[branch][block] : L.r2_dw[block]);
  const float* tc = pk + (branch < 2 ? L.r1_pwtc[branch][block] : L.r2_pwtc[block]);
  const float* bias = pk + (branch < 2 ? L.r1_b[branch][block] : L.r2_bb[block]);
  const int dil = branch < 2 ? r1_dil[block] : r2_dil[block];
  return launch_dwsep_f16(in_clp, out_clp, dw, tc, tc + 1024, bias, B, H, W, dil, branch < 2 || block < 3, 0, (cudaStream_t)stream);
}

static const int kBlockDil[3][4] = {{2, 4, 8, 16}, {2, 4, 8, 16}, {8, 4, 2, 1}};

extern "C" size_t lws_refinement_chain_workspace_bytes(int branch, int block0, int nblk, int B, int H, int W) {
  if (B <= 0 || H <= 0 || W <= 0 || branch < 0 || branch > 2 || block0 < 0 || nblk < 2 || block0 + nblk > 4) return 0;
  return lws::dwsep_chain_workspace_bytes(kBlockDil[branch] + block0, nblk, B, H, W);
}

extern "C" int lws_refinement_chain_clp_f32(const float* in_clp, float* out_clp, const float* pk, int branch, int block0, int nblk,
                                            void* ws, size_t ws_bytes, int B, int H, int W, lws_stream_t stream) {
  using namespace lws;
  LWS_CHECK_PTR(in_clp);
  LWS_CHECK_PTR(out_clp);
  LWS_CHECK_PTR(pk);
  LWS_CHECK_PTR(ws);
  if (B <= 0 || H <= 0 || W <= 0 || branch < 0 || branch > 2 || block0 < 0 || nblk < 2 || block0 + nblk > 4) return LWS_ERR_BAD_SHAPE;
  if ((((uintptr_t)in_clp) | ((uintptr_t)out_clp) | ((uintptr_t)pk)) & 15) return LWS_ERR_BAD_ALIGN;
  const RefLayout L = ref_layout();
  ChainBlockDesc blk[4];
  for (int k = 0; k < nblk; ++k) {
    const int j = block0 + k;
    const float* tc = pk + (branch < 2 ? L.r1_pwtc[branch][j] : L.r2_pwtc[j]);
    blk[k].dw = pk + (branch < 2 ? L.r1_dw[branch][j] : L.r2_dw[j]);
    blk[k].pwh = tc, blk[k].scales = tc + 1024;
    blk[k].bias = pk + (branch < 2 ? L.r1_b[branch][j] : L.r2_bb[j]);
    blk[k].dil = kBlockDil[branch][j], blk[k].relu = branch < 2 || j < 3, blk[k].out_split = 0;
  }
  return launch_dwsep_chain(in_clp, out_clp, blk, nblk, ws, ws_bytes, B, H, W, (cudaStream_t)stream);
}

extern "C" int lws_refinement_f32(const float* left, const float* pred3, const float* pk, float* pred4, void* ws,
                                  size_t ws_bytes, int B, int H, int W, lws_stream_t stream) {
  using namespace lws;
  LWS_CHECK_PTR(left);
  LWS_CHECK_PTR(pred3);
  LWS_CHECK_PTR(pk);
  LWS_CHECK_PTR(pred4);
  LWS_CHECK_PTR(ws);
  if (B <= 0 || H <= 0 || W <= 0 || B > 65535) return LWS_ERR_BAD_SHAPE;
  if (ws_bytes < lws_refinement_workspace_bytes(B, H, W)) return LWS_ERR_WORKSPACE_TOO_SMALL;
  if ((((uintptr_t)ws) | ((uintptr_t)pk) | ((uintptr_t)pred4)) & 15) return LWS_ERR_BAD_ALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  const RefLayout L = ref_layout();
  {
    if (opt(OPT_REFINE_TC) != 0) {
      RefTcWeights wt;
      for (int br = 0; br < 2; ++br) {
        wt.w0[br] = pk + L.r1_w0[br], wt.b0[br] = pk + L.r1_b0[br];
        for (int j = 0; j < 4; ++j)
          wt.dw[br][j] = pk + L.r1_dw[br][j], wt.pwtc[br][j] = pk + L.r1_pwtc[br][j], wt.bias[br][j] = pk + L.r1_b[br][j];
      }
      for (int j = 0; j < 4; ++j) wt.dw[2][j] = pk + L.r2_dw[j], wt.pwtc[2][j] = pk + L.r2_pwtc[j], wt.bias[2][j] = pk + L.r2_bb[j];
      wt.dense_tc = pk + L.r2_wtc, wt.dense_bias = pk + L.r2_b, wt.last_w = pk + L.last_w;
      wt.w0tc[0] = pk + L.r1_w0tc[0], wt.w0tc[1] = pk + L.r1_w0tc[1], wt.last_tc = pk + L.last_tc;
      return refinement_tc(left, pred3, wt, pred4, ws, B, H, W, st);
    }
  }
  const long long hw = (long long)H * W;
  float* cat = (float*)ws;
  float* bufA = cat + (long long)B * 64 * hw;
  float* bufB = bufA + (long long)B * 32 * hw;
  static const int r1_dil[4] = {2, 4, 8, 16};
  static const int r2_dil[4] = {8, 4, 2, 1};
  int rc;

  for (int br = 0; br < 2; ++br) {
    Conv2dArgs c;
    memset(&c, 0, sizeof(c));
    c.H = H, c.W = W;
    c.in = br == 0 ? left : pred3, c.Cin = br == 0 ? 3 : 1, c.in_bs = c.Cin * hw;
    c.w = pk + L.r1_w0[br], c.bias = pk + L.r1_b0[br], c.out = bufA, c.out_bs = 32 * hw;
    rc = br == 0 ? launch_conv2d<3, 32, 8, 1, 64, 1, EPI_BIAS_RELU>(c, B, st)
                 : launch_conv2d<1, 32, 8, 1, 64, 1, EPI_BIAS_RELU>(c, B, st);
    if (rc) return rc;
    float* cur = bufA;
    float* nxt = bufB;
    for (int j = 0; j < 4; ++j) {
      DwsepArgs d;
      d.in = cur, d.in_bs = 32 * hw, d.H = H, d.W = W, d.dil = r1_dil[j], d.relu = 1;
      d.dw = pk + L.r1_dw[br][j], d.pw = pk + L.r1_pw[br][j], d.bias = pk + L.r1_b[br][j];
      if (j < 3) d.out = nxt, d.out_bs = 32 * hw;
      else d.out = cat + (long long)br * 32 * hw, d.out_bs = 64 * hw;
      rc = launch_dwsep(d, B, st);
      if (rc) return rc;
      float* t = cur;
      cur = nxt, nxt = t;
    }
  }
  {
    Conv2dArgs c;
    memset(&c, 0, sizeof(c));
    c.H = H, c.W = W, c.in = cat, c.Cin = 64, c.in_bs = 64 * hw;
    c.w = pk + L.r2_w, c.bias = pk + L.r2_b, c.out = bufA, c.out_bs = 32 * hw;
    rc = launch_conv2d<8, 32, 8, 1, 64, 8, EPI_BIAS_RELU>(c, B, st);
    if (rc) return rc;
    float* cur = bufA;
    float* nxt = bufB;
    for (int j = 0; j < 4; ++j) {
      DwsepArgs d;
      d.in = cur, d.in_bs = 32 * hw, d.H = H, d.W = W, d.dil = r2_dil[j], d.relu = j < 3;
      d.dw = pk + L.r2_dw[j], d.pw = pk + L.r2_pw[j], d.bias = pk + L.r2_bb[j];
      d.out = nxt, d.out_bs = 32 * hw;
      rc = launch_dwsep(d, B, st);
      if (rc) return rc;
      float* t = cur;
      cur = nxt, nxt = t;
    }
    memset(&c, 0, sizeof(c));
    c.H = H, c.W = W, c.in = cur, c.Cin = 32, c.in_bs = 32 * hw;
    c.w = pk + L.last_w, c.skip = pred3, c.out = pred4, c.out_bs = hw;
    return launch_conv2d<8, 1, 1, 4, 64, 1, EPI_SKIP_ADD>(c, B, st);
  }
}

// ---- stand-alone refinement1 / refinement2 (the reference calls them as layers, models/models.py:158-160) -----------------------
// Module-local BN folding (nothing is folded across the module boundary, unlike the fused lws_refinement_f32), NCHW in / out, on
// the exact-fp32 FFMA kernels above.  These entries exist for API parity; the model's forward uses the fused path.
namespace lws {
struct PartBN {
  const float *w, *b, *m, *v;
};
static double part_scale(const PartBN& bn, int c, float eps) { return (double)bn.w[c] / sqrt((double)bn.v[c] + (double)eps); }
static double part_shift(const PartBN& bn, int c, float eps) { return (double)bn.b[c] - (double)bn.m[c] * part_scale(bn, c, eps); }
// one BN-ReLU-DW-PW block: BN of THIS block is folded into the producer; here: dw copy + pw scaled by the NEXT block's BN (or none)
static void part_pack_block(const float* dw, const float* pw, const PartBN* next, float eps, float* dst /*288 + 1024 + 32*/) {
  memcpy(dst, dw, 288 * sizeof(float));
  for (int co = 0; co < 32; ++co) {
    const double s = next ? part_scale(*next, co, eps) : 1.0;
    dst[288 + 1024 + co] = next ? (float)part_shift(*next, co, eps) : 0.f;
    for (int ci = 0; ci < 32; ++ci) dst[288 + ci * 32 + co] = (float)((double)pw[co * 32 + ci] * s);
  }
}
constexpr size_t kPartBlock = 288 + 1024 + 32;

__global__ void bnrelu_nchw_kernel(const float* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift,
                                   float* __restrict__ y, int C, long long hw, long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)((i / hw) % C);
    y[i] = fmaxf(fmaf(x[i], __ldg(scale + c), __ldg(shift + c)), 0.f);
  }
}
}  // namespace lws

extern "C" size_t lws_refinement1_packed_floats(int in_channels) {
  if (in_channels != 1 && in_channels != 3) return 0;
  return (size_t)in_channels * 288 + 32 + 4 * lws::kPartBlock;
}
// tensors (25): conv0 [32,cin,3,3]; then for blocks 1..4: BN(32) weight, bias, _mean, _variance, dw [32,1,3,3], pw [32,32,1,1]
extern "C" int lws_pack_refinement1_weights(const float* const* t, int n_tensors, int in_channels, float eps, float* packed) {
  using namespace lws;
  if (!t || !packed) return LWS_ERR_NULL_PTR;
  if (n_tensors != 25 || (in_channels != 1 && in_channels != 3)) return LWS_ERR_BAD_SHAPE;
  for (int i = 0; i < n_tensors; ++i)
    if (!t[i]) return LWS_ERR_NULL_PTR;
  auto bn = [&](int j) { PartBN b = {t[1 + (j - 1) * 6], t[2 + (j - 1) * 6], t[3 + (j - 1) * 6], t[4 + (j - 1) * 6]}; return b; };
  const PartBN b1 = bn(1);
  float* w0 = packed;
  float* bias0 = packed + (size_t)in_channels * 288;
  for (int co = 0; co < 32; ++co) {
    const double s = part_scale(b1, co, eps);
    bias0[co] = (float)part_shift(b1, co, eps);
    for (int ci = 0; ci < in_channels; ++ci)
      for (int k = 0; k < 9; ++k) w0[((size_t)ci * 9 + k) * 32 + co] = (float)((double)t[0][((size_t)co * in_channels + ci) * 9 + k] * s);
  }
  for (int j = 1; j <= 4; ++j) {
    PartBN nb;
    if (j < 4) nb = bn(j + 1);
    part_pack_block(t[5 + (j - 1) * 6], t[6 + (j - 1) * 6], j < 4 ? &nb : nullptr, eps, bias0 + 32 + (j - 1) * kPartBlock);
  }
  return LWS_OK;
}
extern "C" size_t lws_refinement1_workspace_bytes(int B, int H, int W) {
  if (B <= 0 || H <= 0 || W <= 0) return 0;
  return (size_t)2 * B * 32 * H * W * sizeof(float);
}
// refinement1(in_channels, 32)(x): x [B,in_channels,H,W] -> out [B,32,H,W]  (models/submodules.py:282-300)
extern "C" int lws_refinement1_f32(const float* x, const float* pk, float* out, void* ws, size_t ws_bytes, int B, int in_channels,
                                   int H, int W, lws_stream_t stream) {
  using namespace lws;
  LWS_CHECK_PTR(x);
  LWS_CHECK_PTR(pk);
  LWS_CHECK_PTR(out);
  LWS_CHECK_PTR(ws);
  if (B <= 0 || H <= 0 || W <= 0 || B > 65535 || (in_channels != 1 && in_channels != 3)) return LWS_ERR_BAD_SHAPE;
  if (ws_bytes < lws_refinement1_workspace_bytes(B, H, W)) return LWS_ERR_WORKSPACE_TOO_SMALL;
  if ((((uintptr_t)ws) | ((uintptr_t)pk) | ((uintptr_t)out)) & 15) return LWS_ERR_BAD_ALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  const long long hw = (long long)H * W;
  float* bufA = (float*)ws;
  float* bufB = bufA + (long long)B * 32 * hw;
  static const int dil[4] = {2, 4, 8, 16};
  Conv2dArgs c;
  memset(&c, 0, sizeof(c));
  c.H = H, c.W = W, c.in = x, c.Cin = in_channels, c.in_bs = in_channels * hw;
  c.w = pk, c.bias = pk + (size_t)in_channels * 288, c.out = bufA, c.out_bs = 32 * hw;
  int rc = in_channels == 3 ? launch_conv2d<3, 32, 8, 1, 64, 1, EPI_BIAS_RELU>(c, B, st)
                            : launch_conv2d<1, 32, 8, 1, 64, 1, EPI_BIAS_RELU>(c, B, st);
  if (rc) return rc;
  const float* blocks = pk + (size_t)in_channels * 288 + 32;
  float* cur = bufA;
  float* nxt = bufB;
  for (int j = 0; j < 4; ++j) {
    DwsepArgs d;
    d.in = cur, d.in_bs = 32 * hw, d.H = H, d.W = W, d.dil = dil[j], d.relu = j < 3;
    d.dw = blocks + j * kPartBlock, d.pw = d.dw + 288, d.bias = d.pw + 1024;
    d.out = j < 3 ? nxt : out, d.out_bs = 32 * hw;
    if ((rc = launch_dwsep(d, B, st))) return rc;
    float* tt = cur;
    cur = nxt, nxt = tt;
  }
  return LWS_OK;
}

extern "C" size_t lws_refinement2_packed_floats(void) { return 128 + 64 * 288 + 32 + 4 * lws::kPartBlock + 288; }
// tensors (30): BN(64) weight, bias, _mean, _variance; conv [32,64,3,3]; blocks 1..4 as above; conv_last [1,32,3,3]
extern "C" int lws_pack_refinement2_weights(const float* const* t, int n_tensors, float eps, float* packed) {
  using namespace lws;
  if (!t || !packed) return LWS_ERR_NULL_PTR;
  if (n_tensors != 30) return LWS_ERR_BAD_SHAPE;
  for (int i = 0; i < n_tensors; ++i)
    if (!t[i]) return LWS_ERR_NULL_PTR;
  const PartBN b64 = {t[0], t[1], t[2], t[3]};
  for (int c = 0; c < 64; ++c) packed[c] = (float)part_scale(b64, c, eps), packed[64 + c] = (float)part_shift(b64, c, eps);
  auto bn = [&](int j) { PartBN b = {t[5 + (j - 1) * 6], t[6 + (j - 1) * 6], t[7 + (j - 1) * 6], t[8 + (j - 1) * 6]}; return b; };
  const PartBN b1 = bn(1);
  float* w = packed + 128;
  float* bias = w + 64 * 288;
  for (int co = 0; co < 32; ++co) {
    const double s = part_scale(b1, co, eps);
    bias[co] = (float)part_shift(b1, co, eps);
    for (int ci = 0; ci < 64; ++ci)
      for (int k = 0; k < 9; ++k) w[((size_t)ci * 9 + k) * 32 + co] = (float)((double)t[4][((size_t)co * 64 + ci) * 9 + k] * s);
  }
  for (int j = 1; j <= 4; ++j) {
    PartBN nb;
    if (j < 4) nb = bn(j + 1);
    part_pack_block(t[9 + (j - 1) * 6], t[10 + (j - 1) * 6], j < 4 ? &nb : nullptr, eps, bias + 32 + (j - 1) * kPartBlock);
  }
  float* last = bias + 32 + 4 * kPartBlock;  // [32][9][1]
  for (int ci = 0; ci < 32; ++ci)
    for (int k = 0; k < 9; ++k) last[ci * 9 + k] = t[29][ci * 9 + k];
  return LWS_OK;
}
extern "C" size_t lws_refinement2_workspace_bytes(int B, int H, int W) {
  if (B <= 0 || H <= 0 || W <= 0) return 0;
  return (size_t)B * (64 + 32 + 32) * H * W * sizeof(float);
}
// refinement2(64, 32)(x): x [B,64,H,W] -> out [B,1,H,W]  (models/submodules.py:302-327; no skip: the caller adds it, models.py:160-162)
extern "C" int lws_refinement2_f32(const float* x, const float* pk, float* out, void* ws, size_t ws_bytes, int B, int H, int W,
                                   lws_stream_t stream) {
  using namespace lws;
  LWS_CHECK_PTR(x);
  LWS_CHECK_PTR(pk);
  LWS_CHECK_PTR(out);
  LWS_CHECK_PTR(ws);
  if (B <= 0 || H <= 0 || W <= 0 || B > 65535) return LWS_ERR_BAD_SHAPE;
  if (ws_bytes < lws_refinement2_workspace_bytes(B, H, W)) return LWS_ERR_WORKSPACE_TOO_SMALL;
  if ((((uintptr_t)ws) | ((uintptr_t)pk) | ((uintptr_t)out)) & 15) return LWS_ERR_BAD_ALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  const long long hw = (long long)H * W;
  float* act = (float*)ws;  // ReLU(BN64(x)): the conv pads the activated tensor with zeros
  float* bufA = act + (long long)B * 64 * hw;
  float* bufB = bufA + (long long)B * 32 * hw;
  const long long total = (long long)B * 64 * hw;
  bnrelu_nchw_kernel<<<(int)min((total + 255) / 256, (long long)kNumSMs * 16), 256, 0, st>>>(x, pk, pk + 64, act, 64, hw, total);
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) return (int)e;
  static const int dil[4] = {8, 4, 2, 1};
  Conv2dArgs c;
  memset(&c, 0, sizeof(c));
  c.H = H, c.W = W, c.in = act, c.Cin = 64, c.in_bs = 64 * hw;
  c.w = pk + 128, c.bias = pk + 128 + 64 * 288, c.out = bufA, c.out_bs = 32 * hw;
  int rc = launch_conv2d<8, 32, 8, 1, 64, 8, EPI_BIAS_RELU>(c, B, st);
  if (rc) return rc;
  const float* blocks = pk + 128 + 64 * 288 + 32;
  float* cur = bufA;
  float* nxt = bufB;
  for (int j = 0; j < 4; ++j) {
    DwsepArgs d;
    d.in = cur, d.in_bs = 32 * hw, d.H = H, d.W = W, d.dil = dil[j], d.relu = j < 3;
    d.dw = blocks + j * kPartBlock, d.pw = d.dw + 288, d.bias = d.pw + 1024;
    d.out = nxt, d.out_bs = 32 * hw;
    if ((rc = launch_dwsep(d, B, st))) return rc;
    float* tt = cur;
    cur = nxt, nxt = tt;
  }
  memset(&c, 0, sizeof(c));
  c.H = H, c.W = W, c.in = cur, c.Cin = 32, c.in_bs = 32 * hw;
  c.w = blocks + 4 * kPartBlock, c.skip = nullptr, c.out = out, c.out_bs = hw;
  return launch_conv2d<8, 1, 1, 4, 64, 1, EPI_SKIP_ADD>(c, B, st);
}

// ---- stand-alone disparity_regression.forward (models/models.py:167-179): out = sum_j input[:, j] * (start + j * step) ----------
namespace lws {
__global__ void weighted_sum_kernel(const float* __restrict__ p, float* __restrict__ out, int D, long long hw, long long total,
                                    float start, float step) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / hw, r = i - b * hw;
    const float* src = p + b * D * hw + r;
    float acc = 0.f;
    // the reference multiplies by the expanded arange and reduces with paddle.sum over axis 1: plain left-to-right fp32 sum
    for (int j = 0; j < D; ++j) acc = __fadd_rn(acc, __fmul_rn(__ldg(src + j * hw), __fadd_rn(start, __fmul_rn((float)j, step))));
    out[i] = acc;
  }
}
}  // namespace lws
extern "C" int lws_disparity_regression_f32(const float* prob, float* out, int B, int D, int H, int W, float start, float step,
                                            lws_stream_t stream) {
  using namespace lws;
  LWS_CHECK_PTR(prob);
  LWS_CHECK_PTR(out);
  if (B <= 0 || D <= 0 || H <= 0 || W <= 0) return LWS_ERR_BAD_SHAPE;
  const long long hw = (long long)H * W, total = B * hw;
  weighted_sum_kernel<<<(int)min((total + 255) / 256, (long long)kNumSMs * 16), 256, 0, (cudaStream_t)stream>>>(prob, out, D, hw, total,
                                                                                                              start, step);
  LWS_RETURN_LAUNCH_STATUS();
}
