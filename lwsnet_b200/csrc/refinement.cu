// K6: colour-guidance refinement  pred4 = pred3 + R2(concat[R1_left(left), R1_disp(pred3)])
// (reference models/submodules.py:223-327 preconv2d / preconv2d_depthseperated / refinement1 / refinement2,
//  called at models/models.py:158-162; ~45 full-resolution cuDNN/elementwise launches in the reference, 16 here).
//
// BatchNorm is folded on the host exactly as for the 3D stack: every kernel stores ReLU(BN_next(conv(.))), the
// concat is never materialised separately (the two R1 branches write their halves of the 64-channel buffer), and the
// final interpolate (same size => identity) is dropped.
//   dense 3x3 convs  : conv2d.cuh (FP32 direct, smem tiled)
//   BN-ReLU-DW-PW    : dwsep_block_kernel below: depthwise dilated 3x3 into shared memory, then the 32x32 pointwise
//                      product out of shared memory with a 4 px x 8 cout register tile.
#include <math.h>
#include <string.h>

#include "conv2d.cuh"

namespace lws {

struct DwsepArgs {
  const float* in;    // [B,32,H,W] post-activation
  const float* dw;    // [32][9]
  const float* pw;    // [32 ci][32 co] (next BN scale folded)
  const float* bias;  // [32]
  float* out;
  long long in_bs, out_bs;
  int H, W, dil, relu;
};

constexpr int DW_TH = 8, DW_TW = 32, DW_C = 32;

__global__ void __launch_bounds__(256, 2) dwsep_block_kernel(const DwsepArgs a) {
  __shared__ __align__(16) float sDW[DW_C][DW_TH * DW_TW];  // 32 KB
  __shared__ __align__(16) float sPW[DW_C][DW_C];           // 4 KB
  __shared__ float sK[DW_C][9];
  const int tid = threadIdx.x;
  const int tiles_w = (a.W + DW_TW - 1) / DW_TW;
  const int w0 = (blockIdx.x % tiles_w) * DW_TW, h0 = (blockIdx.x / tiles_w) * DW_TH;
  const int b = blockIdx.y;
  const int H = a.H, W = a.W, dil = a.dil;
  const long long hw = (long long)H * W;
  for (int i = tid; i < DW_C * DW_C; i += 256) sPW[i / DW_C][i % DW_C] = __ldg(a.pw + i);
  for (int i = tid; i < DW_C * 9; i += 256) sK[i / 9][i % 9] = __ldg(a.dw + i);
  __syncthreads();

  // phase 1: depthwise dilated 3x3, one pixel per thread, channels in a loop (coalesced along w)
  {
    const int lx = tid % DW_TW, ly = tid / DW_TW;
    const int gx = w0 + lx, gy = h0 + ly;
    const float* in_b = a.in + (long long)b * a.in_bs;
    bool okx[3], oky[3];
    long long off[3][3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      okx[k] = gx + (k - 1) * dil >= 0 && gx + (k - 1) * dil < W;
      oky[k] = gy + (k - 1) * dil >= 0 && gy + (k - 1) * dil < H;
    }
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx)
        off[ky][kx] = (long long)(gy + (ky - 1) * dil) * W + (gx + (kx - 1) * dil);
#pragma unroll 4
    for (int c = 0; c < DW_C; ++c) {
      const float* p = in_b + c * hw;
      float s = 0.f;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const float v = (okx[kx] && oky[ky]) ? __ldg(p + off[ky][kx]) : 0.f;
          s = fmaf(v, sK[c][ky * 3 + kx], s);
        }
      sDW[c][tid] = s;
    }
  }
  __syncthreads();

  // phase 2: pointwise 32 -> 32; thread = 4 consecutive px x 8 cout
  const int pq = tid % 64, cg = tid / 64;
  float acc[4][8];
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[p][q] = 0.f;
#pragma unroll 8
  for (int ci = 0; ci < DW_C; ++ci) {
    const float4 x = *reinterpret_cast<const float4*>(&sDW[ci][pq * 4]);
    const float4 wa = *reinterpret_cast<const float4*>(&sPW[ci][cg * 8]);
    const float4 wb = *reinterpret_cast<const float4*>(&sPW[ci][cg * 8 + 4]);
    const float xv[4] = {x.x, x.y, x.z, x.w};
    const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[p][q] = fmaf(xv[p], wv[q], acc[p][q]);
  }
  const int lx = (pq * 4) % DW_TW, ly = (pq * 4) / DW_TW;
  const int gx = w0 + lx, gy = h0 + ly;
  if (gy >= H || gx >= W) return;
  const bool vec = ((W & 3) == 0);
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int co = cg * 8 + q;
    const float bias = __ldg(a.bias + co);
    float r[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      r[p] = acc[p][q] + bias;
      if (a.relu) r[p] = fmaxf(r[p], 0.f);
    }
    float* o = a.out + (long long)b * a.out_bs + co * hw + (long long)gy * W + gx;
    if (vec) {
      *reinterpret_cast<float4*>(o) = make_float4(r[0], r[1], r[2], r[3]);
    } else {
#pragma unroll
      for (int p = 0; p < 4; ++p)
        if (gx + p < W) o[p] = r[p];
    }
  }
}

static int launch_dwsep(DwsepArgs a, int B, cudaStream_t st) {
  dim3 grid(cdiv(a.W, DW_TW) * cdiv(a.H, DW_TH), B);
  dwsep_block_kernel<<<grid, 256, 0, st>>>(a);
  cudaError_t e = cudaPeekAtLastError();
  return e == cudaSuccess ? LWS_OK : (int)e;
}

// ---- packed layout -------------------------------------------------------------------------------------------
struct RefLayout {
  size_t r1_w0[2], r1_b0[2];            // conv0 [Cin][9][32], bias[32]      (0: left Cin=3, 1: disp Cin=1)
  size_t r1_dw[2][4], r1_pw[2][4], r1_b[2][4];
  size_t r2_w, r2_b;                    // [64][9][32], bias[32]
  size_t r2_dw[4], r2_pw[4], r2_bb[4];
  size_t last_w;                        // [32][9][1]
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
  return LWS_OK;
}

extern "C" size_t lws_refinement_workspace_bytes(int B, int H, int W) {
  if (B <= 0 || H <= 0 || W <= 0) return 0;
  return (size_t)B * 128 * H * W * sizeof(float);  // concat (64 ch) + two 32-channel ping-pong buffers
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
