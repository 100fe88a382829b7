// Direct 3x3 (optionally dilated) 2D convolution on the FP32 pipes, NCHW, stride 1, zero padding = dilation.
// Same register blocking as the 3D kernel (conv3d_stack.cu): a thread owns 8 consecutive w x Q output channels,
// lanes run over h first and the shared-memory row pitch is 4 mod 8 floats so the 128-bit row loads are
// conflict-free; input channels stream through shared memory in chunks of CK.
#pragma once
#include "lws_common.cuh"

namespace lws {

enum ConvEpilogue { EPI_BIAS_RELU = 0, EPI_BIAS = 1, EPI_SKIP_ADD = 2 };

struct Conv2dArgs {
  const float* in;   // [B,Cin,H,W] (batch stride in_bs floats)
  const float* w;    // [Cin][9][Cout]
  const float* bias; // [Cout]
  const float* skip; // [B,1,H,W] (EPI_SKIP_ADD)
  float* out;        // [B,Cout,H,W] (batch stride out_bs floats)
  long long in_bs, out_bs;
  int Cin, H, W;
  int tiles_w, tiles_h;
};

template <int CK, int COUT, int Q, int RG, int TW, int DIL, int EPI>
struct Conv2dCfg {
  static constexpr int P = 8;
  static constexpr int TH = 8 * RG;
  static constexpr int WQ = TW / P;
  static constexpr int NVT = TH * WQ;
  static constexpr int NCG = COUT / Q;
  static constexpr int THREADS = NVT * NCG;
  static constexpr int PADL = DIL == 1 ? 4 : DIL;                      // interior starts 16B aligned
  static constexpr int RAW = PADL + TW + (DIL == 1 ? 0 : DIL);
  static constexpr int PITCH = RAW + ((4 - RAW % 8) + 8) % 8;           // == 4 (mod 8)
  static constexpr int ROWS = TH + 2 * DIL;
  static constexpr int IN_FLOATS = CK * ROWS * PITCH + 4;
  static constexpr int WT_STRIDE = (COUT + 3) / 4 * 4;
  static constexpr int WT_FLOATS = CK * 9 * WT_STRIDE;
  static constexpr size_t SMEM = (size_t)(IN_FLOATS + WT_FLOATS) * sizeof(float);
  static constexpr int MIN_BLOCKS = THREADS <= 256 ? 2 : 1;
  static_assert(NVT % 32 == 0, "cout group must be warp-uniform");
  static_assert(DIL == 1 || DIL % 4 == 0, "dilated taps must stay 16B aligned");
  static_assert(PITCH % 8 == 4, "pitch");
};

template <int CK, int COUT, int Q, int RG, int TW, int DIL, int EPI>
__global__ void __launch_bounds__(Conv2dCfg<CK, COUT, Q, RG, TW, DIL, EPI>::THREADS,
                                  Conv2dCfg<CK, COUT, Q, RG, TW, DIL, EPI>::MIN_BLOCKS)
    conv2d_3x3_kernel(const Conv2dArgs a) {
  using Cfg = Conv2dCfg<CK, COUT, Q, RG, TW, DIL, EPI>;
  constexpr int P = Cfg::P, PITCH = Cfg::PITCH, ROWS = Cfg::ROWS, PADL = Cfg::PADL;
  extern __shared__ __align__(16) float smem[];
  float* sIn = smem;
  float* sW = smem + Cfg::IN_FLOATS;

  const int tid = threadIdx.x;
  const int vt = tid % Cfg::NVT;
  const int cg = tid / Cfg::NVT;
  const int th = (vt % 8) + 8 * (vt / (8 * Cfg::WQ));
  const int twq = (vt / 8) % Cfg::WQ;
  const int tw_i = blockIdx.x % a.tiles_w;
  const int th_i = blockIdx.x / a.tiles_w;
  const int b = blockIdx.y;
  const int w0 = tw_i * TW, h0 = th_i * Cfg::TH;
  const int H = a.H, W = a.W;
  const long long hw = (long long)H * W;

  float acc[P][Q];
#pragma unroll
  for (int p = 0; p < P; ++p)
#pragma unroll
    for (int q = 0; q < Q; ++q) acc[p][q] = 0.f;

  const float* in_b = a.in + (long long)b * a.in_bs;
  for (int c0 = 0; c0 < a.Cin; c0 += CK) {
    constexpr int ROW_E = TW + 2 * DIL;
    // cp.async: every copy of the chunk is in flight at once (see conv3d_stack.cu)
    for (int idx = tid; idx < CK * ROWS * ROW_E; idx += Cfg::THREADS) {
      const int e = idx % ROW_E;
      const int row = idx / ROW_E;
      const int hh = row % ROWS;
      const int ci = row / ROWS;
      const int gh = h0 - DIL + hh, gw = w0 - DIL + e;
      float* dst = sIn + row * PITCH + (PADL - DIL) + e;
      if (c0 + ci < a.Cin && gh >= 0 && gh < H && gw >= 0 && gw < W)
        cp_async_4(dst, in_b + (long long)(c0 + ci) * hw + (long long)gh * W + gw);
      else
        *dst = 0.f;
    }
    {
      const float* wsrc = a.w + (long long)c0 * 9 * COUT;
      const int nvalid = min(CK, a.Cin - c0) * 9;
      if constexpr (COUT % 4 == 0) {
        for (int idx = tid; idx < CK * 9 * COUT / 4; idx += Cfg::THREADS) {
          if (idx * 4 < nvalid * COUT) cp_async_16(sW + idx * 4, wsrc + idx * 4);
          else *reinterpret_cast<float4*>(sW + idx * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      } else {
        for (int idx = tid; idx < CK * 9 * Cfg::WT_STRIDE; idx += Cfg::THREADS) {
          const int co = idx % Cfg::WT_STRIDE;
          const int ct = idx / Cfg::WT_STRIDE;
          sW[idx] = (co < COUT && ct < nvalid) ? __ldg(wsrc + ct * COUT + co) : 0.f;
        }
      }
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();

#pragma unroll 1
    for (int ci = 0; ci < CK; ++ci) {
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const float* row = sIn + (ci * ROWS + th + kh * DIL) * PITCH + twq * P;
        const float* wrow = sW + (ci * 9 + kh * 3) * Cfg::WT_STRIDE + cg * Q;
        float x[3][P];
        if constexpr (DIL == 1) {
          const float xl = row[3];
          const float4 x1 = *reinterpret_cast<const float4*>(row + 4);
          const float4 x2 = *reinterpret_cast<const float4*>(row + 8);
          const float xr = row[12];
          const float t[P + 2] = {xl, x1.x, x1.y, x1.z, x1.w, x2.x, x2.y, x2.z, x2.w, xr};
#pragma unroll
          for (int kw = 0; kw < 3; ++kw)
#pragma unroll
            for (int p = 0; p < P; ++p) x[kw][p] = t[p + kw];
        } else {
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const float4 x1 = *reinterpret_cast<const float4*>(row + PADL + (kw - 1) * DIL);
            const float4 x2 = *reinterpret_cast<const float4*>(row + PADL + (kw - 1) * DIL + 4);
            x[kw][0] = x1.x, x[kw][1] = x1.y, x[kw][2] = x1.z, x[kw][3] = x1.w;
            x[kw][4] = x2.x, x[kw][5] = x2.y, x[kw][6] = x2.z, x[kw][7] = x2.w;
          }
        }
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          float wv[Q];
          if constexpr (Q % 4 == 0) {
#pragma unroll
            for (int q4 = 0; q4 < Q / 4; ++q4) {
              const float4 t = *reinterpret_cast<const float4*>(wrow + kw * Cfg::WT_STRIDE + q4 * 4);
              wv[q4 * 4] = t.x, wv[q4 * 4 + 1] = t.y, wv[q4 * 4 + 2] = t.z, wv[q4 * 4 + 3] = t.w;
            }
          } else {
#pragma unroll
            for (int q = 0; q < Q; ++q) wv[q] = wrow[kw * Cfg::WT_STRIDE + q];
          }
#pragma unroll
          for (int p = 0; p < P; ++p)
#pragma unroll
            for (int q = 0; q < Q; ++q) acc[p][q] = fmaf(x[kw][p], wv[q], acc[p][q]);
        }
      }
    }
    __syncthreads();
  }

  const int gh = h0 + th, gw = w0 + twq * P;
  if (gh >= H || gw >= W) return;
  const long long pix = (long long)gh * W + gw;
  const int nvec = (gw + P <= W) ? (((W & 3) == 0) ? 4 : (((W & 1) == 0) ? 2 : 1)) : 1;
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    const int co = cg * Q + q;
    float r[P];
    if constexpr (EPI == EPI_SKIP_ADD) {
      const float* sk = a.skip + (long long)b * hw + pix;
#pragma unroll
      for (int p = 0; p < P; ++p) r[p] = acc[p][q] + ((a.skip && gw + p < W) ? __ldg(sk + p) : 0.f);
    } else {
      const float bias = __ldg(a.bias + co);
#pragma unroll
      for (int p = 0; p < P; ++p) {
        r[p] = acc[p][q] + bias;
        if constexpr (EPI == EPI_BIAS_RELU) r[p] = fmaxf(r[p], 0.f);
      }
    }
    float* o = a.out + (long long)b * a.out_bs + (long long)co * hw + pix;
    if (nvec == 4) {
      *reinterpret_cast<float4*>(o) = make_float4(r[0], r[1], r[2], r[3]);
      *reinterpret_cast<float4*>(o + 4) = make_float4(r[4], r[5], r[6], r[7]);
    } else if (nvec == 2) {
#pragma unroll
      for (int p = 0; p < P; p += 2) *reinterpret_cast<float2*>(o + p) = make_float2(r[p], r[p + 1]);
    } else {
#pragma unroll
      for (int p = 0; p < P; ++p)
        if (gw + p < W) o[p] = r[p];
    }
  }
}

template <int CK, int COUT, int Q, int RG, int TW, int DIL, int EPI>
static int launch_conv2d(Conv2dArgs a, int B, cudaStream_t st) {
  using Cfg = Conv2dCfg<CK, COUT, Q, RG, TW, DIL, EPI>;
  auto kern = conv2d_3x3_kernel<CK, COUT, Q, RG, TW, DIL, EPI>;
  LWS_SET_SMEM_ONCE(kern, Cfg::SMEM);  // `kern` is fixed by the template arguments: one static flag per instantiation
  a.tiles_w = cdiv(a.W, TW);
  a.tiles_h = cdiv(a.H, Cfg::TH);
  dim3 grid(a.tiles_w * a.tiles_h, B);
  kern<<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(a);
  cudaError_t e2 = cudaPeekAtLastError();
  return e2 == cudaSuccess ? LWS_OK : (int)e2;
}

}  // namespace lws
