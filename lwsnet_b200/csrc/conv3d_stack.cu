// K3: the residual 3D-conv stack  out = cost + Conv_{n-1}(ReLU(BN_{n-1}(... Conv_0(ReLU(BN_0(cost))) ...)))
// (reference models/submodules.py:190-221 post_3dconvs / batch_relu_conv3d, skip at models/models.py:136-138;
//  18 cuDNN conv3d + 18 batch_norm + 18 relu launches per forward in the reference, 6 launches per stage here).
//
// BatchNorm (inference) is folded on the host (lws_pack_conv3d_stack_weights): BN_{i+1}'s scale goes into conv i's
// output channels, its shift becomes a bias, and every kernel stores the *post-activation* tensor
// ReLU(BN_{i+1}(conv_i(.))) so that the consumer's zero padding is the reference's padding of the ReLU output
// (SURVEY.md A.5).  BN_0 is a scalar affine applied while the raw cost tile is staged.
//
// Direct convolution on the FP32 pipes, NCDHW:
//   block  : output tile TD x 8 x TW voxels, all Cout; input channels streamed in chunks of CK through shared memory
//   thread : 8 consecutive w  x  Q output channels = 64 accumulators; per (ci,kd,kh) it reads a 10-wide input row
//            segment (2 LDS.128 + 2 LDS.32) and 3*Q weights (warp-uniform broadcast LDS.128) for 24*Q FFMA.
//   lanes run over h first, rows are TW+4 floats apart (== 4 mod 32), so the 128-bit row loads are conflict-free.
// These layers are compute-bound (54-216 flop/B): the bound is the FP32 FFMA rate, not HBM.
#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "lws_common.cuh"

namespace lws {

// conv3d_c8.cu: tcgen05 split-fp16 path for C = 8
size_t conv3d_c8_workspace_bytes(int B, int D, int H, int W);
int conv3d_stack_c8(const float* cost, const float* affine, const float* w_first, const float* b_first, const float* const* wtab,
                    const float* const* bias_mid, int layers, const float* w_last, float* out, void* ws, int B, int D, int H, int W,
                    int add_skip, cudaStream_t st);
// conv3d_f16.cu: tcgen05 split-fp16 Toeplitz-N path for C = 32
size_t conv3d_f16_workspace_bytes(int B, int D, int H, int W);
int cost_first_conv_fused_supported(int Cf, int D, int H, int W);
int conv3d_stack_f16(const float* cost, const float* affine, const float* w_first, const float* b_first, const float* const* wtab,
                     const float* const* bias_mid, int layers, const float* w_last, float* out, void* ws, int B, int D, int H, int W,
                     int add_skip, cudaStream_t st, const float* featL = nullptr, const float* featR = nullptr, int Cf = 0);
constexpr int kTcLayerFloats = 9 * 192 * 32;
constexpr int kTcLastOff = 32768;  // C = 32: table of the closing 32 -> 1 conv inside the first mid layer's slot  // per 32->32 layer: [9 (kd,kh)][3 kw][32 hi + 32 lo rows][32 ci]

struct Conv3dArgs {
  const float* in;      // [B,Cin,D,H,W]  post-activation (or raw cost when FIRST)
  const float* w;       // [Cin][27][Cout] folded weights
  const float* bias;    // [Cout] (unused when LAST)
  const float* skip;    // raw cost [B,D,H,W] (LAST only)
  float* out;           // [B,Cout,D,H,W]
  int Cin, D, H, W;
  int tiles_w, tiles_h, tiles_d;
  const float* affine;  // device [s0, t0]: BN_0 scalar affine (FIRST only)
  int cout_total, cout_off;  // the launch writes channels [cout_off, cout_off + COUT) of an output with cout_total channels
};

template <int CK, int COUT, int Q, int TD, int TW, bool FIRST, bool LAST>
struct Conv3dCfg {
  static constexpr int TH = 8;
  static constexpr int P = 8;
  static constexpr int WQ = TW / P;
  static constexpr int NVT = TD * TH * WQ;
  static constexpr int NCG = COUT / Q;
  static constexpr int THREADS = NVT * NCG;
  static constexpr int PITCH = TW + 4;
  static constexpr int ROWS = (TD + 2) * (TH + 2);
  static constexpr int IN_FLOATS = CK * ROWS * PITCH + 4;  // +4: right halo of the very last row
  static constexpr int WT_STRIDE = (COUT + 3) / 4 * 4;      // floats per (ci,tap) in shared memory
  static constexpr int WT_FLOATS = CK * 27 * WT_STRIDE;
  static constexpr size_t SMEM = (size_t)(IN_FLOATS + WT_FLOATS) * sizeof(float);
  static constexpr int MIN_BLOCKS = THREADS <= 256 ? 2 : 1;
  static_assert(NVT % 32 == 0, "cout group must be warp-uniform");
  static_assert(TW % 32 == 0, "row pitch must be 4 mod 32");
};

template <int CK, int COUT, int Q, int TD, int TW, bool FIRST, bool LAST>
__global__ void __launch_bounds__(Conv3dCfg<CK, COUT, Q, TD, TW, FIRST, LAST>::THREADS,
                                  Conv3dCfg<CK, COUT, Q, TD, TW, FIRST, LAST>::MIN_BLOCKS)
    conv3d_k3_kernel(const Conv3dArgs a) {
  using Cfg = Conv3dCfg<CK, COUT, Q, TD, TW, FIRST, LAST>;
  constexpr int TH = Cfg::TH, P = Cfg::P, PITCH = Cfg::PITCH, ROWS = Cfg::ROWS;
  extern __shared__ __align__(16) float smem[];
  float* sIn = smem;
  float* sW = smem + Cfg::IN_FLOATS;

  const int tid = threadIdx.x;
  const int vt = tid % Cfg::NVT;
  const int cg = tid / Cfg::NVT;
  const int th = vt % TH;
  const int twq = (vt / TH) % Cfg::WQ;
  const int td = vt / (TH * Cfg::WQ);

  int tile = blockIdx.x;
  const int tw_i = tile % a.tiles_w;
  tile /= a.tiles_w;
  const int th_i = tile % a.tiles_h;
  const int td_i = tile / a.tiles_h;
  const int b = blockIdx.y;
  const int w0 = tw_i * TW, h0 = th_i * TH, d0 = td_i * TD;
  const int D = a.D, H = a.H, W = a.W;
  const long long hw = (long long)H * W;
  const long long dhw = (long long)D * hw;

  float acc[P][Q];
#pragma unroll
  for (int p = 0; p < P; ++p)
#pragma unroll
    for (int q = 0; q < Q; ++q) acc[p][q] = 0.f;

  float s0 = 1.f, t0 = 0.f;
  if constexpr (FIRST) s0 = __ldg(a.affine), t0 = __ldg(a.affine + 1);
  const float* in_b = a.in + (long long)b * a.Cin * dhw;
  for (int c0 = 0; c0 < a.Cin; c0 += CK) {
    // ---- stage the input tile (halo 1, zero padded) and this chunk's weights --------------------------
    // cp.async keeps every copy of the chunk in flight at once (a register-staged loop exposes one L2 round trip
    // per element); the raw-cost layer needs BN_0 + ReLU on the way in, so it batches plain loads instead.
    constexpr int ROW_E = TW + 2;
    constexpr int N_IN = CK * ROWS * ROW_E;
    if constexpr (FIRST) {
      constexpr int UNR = 4;
      for (int base = tid; base < N_IN; base += Cfg::THREADS * UNR) {
        float v[UNR];
        int dst[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          const int idx = base + u * Cfg::THREADS;
          const int e = idx % ROW_E;
          const int row = idx / ROW_E;
          const int hh = row % (TH + 2);
          const int dd = (row / (TH + 2)) % (TD + 2);
          const int gd = d0 - 1 + dd, gh = h0 - 1 + hh, gw = w0 - 1 + e;
          dst[u] = idx < N_IN ? row * PITCH + 3 + e : -1;
          const bool inb = idx < N_IN && gd >= 0 && gd < D && gh >= 0 && gh < H && gw >= 0 && gw < W;
          v[u] = inb ? fmaxf(fmaf(__ldg(in_b + (long long)gd * hw + (long long)gh * W + gw), s0, t0), 0.f) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u)
          if (dst[u] >= 0) sIn[dst[u]] = v[u];
      }
    } else {
      for (int idx = tid; idx < N_IN; idx += Cfg::THREADS) {
        const int e = idx % ROW_E;
        const int row = idx / ROW_E;
        const int hh = row % (TH + 2);
        const int dd = (row / (TH + 2)) % (TD + 2);
        const int ci = row / ((TH + 2) * (TD + 2));
        const int gd = d0 - 1 + dd, gh = h0 - 1 + hh, gw = w0 - 1 + e;
        float* dst = sIn + row * PITCH + 3 + e;
        if (c0 + ci < a.Cin && gd >= 0 && gd < D && gh >= 0 && gh < H && gw >= 0 && gw < W)
          cp_async_4(dst, in_b + (long long)(c0 + ci) * dhw + (long long)gd * hw + (long long)gh * W + gw);
        else
          *dst = 0.f;
      }
    }
    {
      const float* wsrc = a.w + (long long)c0 * 27 * COUT;
      const int nvalid = min(CK, a.Cin - c0) * 27;
      if constexpr (COUT % 4 == 0) {
        for (int idx = tid; idx < CK * 27 * COUT / 4; idx += Cfg::THREADS) {
          if (idx * 4 < nvalid * COUT) cp_async_16(sW + idx * 4, wsrc + idx * 4);
          else *reinterpret_cast<float4*>(sW + idx * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      } else {
        for (int idx = tid; idx < CK * 27 * Cfg::WT_STRIDE; idx += Cfg::THREADS) {
          const int co = idx % Cfg::WT_STRIDE;
          const int ct = idx / Cfg::WT_STRIDE;
          sW[idx] = (co < COUT && ct < nvalid) ? __ldg(wsrc + ct * COUT + co) : 0.f;
        }
      }
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();

    // ---- FFMA main loop ---------------------------------------------------------------------------------
#pragma unroll 1
    for (int ci = 0; ci < CK; ++ci) {
#pragma unroll
      for (int kd = 0; kd < 3; ++kd) {
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
          const float* row = sIn + ((ci * (TD + 2) + td + kd) * (TH + 2) + th + kh) * PITCH + twq * P;
          float x[P + 2];
          x[0] = row[3];
          const float4 x1 = *reinterpret_cast<const float4*>(row + 4);
          const float4 x2 = *reinterpret_cast<const float4*>(row + 8);
          x[1] = x1.x, x[2] = x1.y, x[3] = x1.z, x[4] = x1.w;
          x[5] = x2.x, x[6] = x2.y, x[7] = x2.z, x[8] = x2.w;
          x[9] = row[12];
          const float* wrow = sW + (ci * 27 + (kd * 3 + kh) * 3) * Cfg::WT_STRIDE + cg * Q;
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
              for (int q = 0; q < Q; ++q) acc[p][q] = fmaf(x[p + kw], wv[q], acc[p][q]);
          }
        }
      }
    }
    __syncthreads();
  }

  // ---- epilogue -----------------------------------------------------------------------------------------
  const int gd = d0 + td, gh = h0 + th, gw = w0 + twq * P;
  if (gd >= D || gh >= H || gw >= W) return;
  const long long vox = (long long)gd * hw + (long long)gh * W + gw;
  if constexpr (LAST) {
    const float* sk = a.skip + (long long)b * dhw + vox;
    float* o = a.out + (long long)b * dhw + vox;
#pragma unroll
    for (int p = 0; p < P; ++p)
      if (gw + p < W) o[p] = acc[p][0] + (a.skip ? __ldg(sk + p) : 0.f);
  } else {
    const int nvec = (gw + P <= W) ? (((W & 3) == 0) ? 4 : (((W & 1) == 0) ? 2 : 1)) : 1;
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const int co = cg * Q + q;
      const float bias = __ldg(a.bias + co);
      float* o = a.out + ((long long)b * a.cout_total + a.cout_off + co) * dhw + vox;
      float r[P];
#pragma unroll
      for (int p = 0; p < P; ++p) r[p] = fmaxf(acc[p][q] + bias, 0.f);
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
}

template <int CK, int COUT, int Q, int TD, int TW, bool FIRST, bool LAST>
static int launch_conv3d(Conv3dArgs a, int B, cudaStream_t st) {
  using Cfg = Conv3dCfg<CK, COUT, Q, TD, TW, FIRST, LAST>;
  auto kern = conv3d_k3_kernel<CK, COUT, Q, TD, TW, FIRST, LAST>;
  LWS_SET_SMEM_ONCE(kern, Cfg::SMEM);  // `kern` is fixed by the template arguments: one static flag per instantiation
  if (a.cout_total == 0) a.cout_total = COUT, a.cout_off = 0;
  a.tiles_w = cdiv(a.W, TW);
  a.tiles_h = cdiv(a.H, Cfg::TH);
  a.tiles_d = cdiv(a.D, TD);
  dim3 grid(a.tiles_w * a.tiles_h * a.tiles_d, B);
  kern<<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(a);
  cudaError_t e2 = cudaPeekAtLastError();
  return e2 == cudaSuccess ? LWS_OK : (int)e2;
}

// packed layout (floats): [s0, t0, 0, 0] then per conv i: weights [Cin_i][27][Cout_i] (rounded up to 4 floats),
// bias [Cout_i] (rounded up to 4; zeros for the last conv)
// Widths other than 8 / 16 / 32 (the reference takes any channels_3d * growth_rate, models/models.py:19-22) run on the FP32 FFMA
// kernel in groups of 8 output channels: C is padded to Cp = round_up(C, 8) with zero weights / zero bias (the padded activations
// are exactly 0), and the packed weights of a conv are stored group-major [Cp/8][Cin_p][27][8].
static bool native_width(int C) { return C == 8 || C == 16 || C == 32; }
static int padded_width(int C) { return native_width(C) ? C : round_up(C, 8); }
static size_t conv_w_floats(int cin, int cout) { return (size_t)round_up(cin * 27 * cout, 4); }
static size_t packed_offset(int C, int layers, int conv, bool bias) {
  size_t off = 4;
  const int n = layers + 2;
  C = padded_width(C);
  for (int i = 0; i < n; ++i) {
    const int cin = i == 0 ? 1 : C, cout = i == n - 1 ? 1 : C;
    if (i == conv && !bias) return off;
    off += conv_w_floats(cin, cout);
    if (i == conv && bias) return off;
    off += (size_t)round_up(cout, 4);
  }
  return off;
}
// C == 32 only: the tensor-core operand tables follow the generic layout, one per 32 -> 32 layer
static size_t packed_tc_offset(int C, int layers, int mid_layer) {
  return packed_offset(C, layers, layers + 2, false) + (size_t)mid_layer * kTcLayerFloats;
}
static bool has_tc_tables(int C) { return C == 32 || C == 8; }
static bool use_tc_path(int C) {
  if (!has_tc_tables(C)) return false;
  return opt(OPT_CONV3D_TC) != 0;
}

template <int C>
static int run_stack(const float* cost, const float* pk, float* out, float* bufA, float* bufB, int B, int D, int H,
                     int W, int layers, int add_skip, cudaStream_t st) {
  constexpr int TD = (C == 32) ? 2 : 3;
  constexpr int TW = (C == 32) ? 32 : 64;
  constexpr int CK = (C == 32) ? 8 : 4;
  constexpr int Q = 8;
  Conv3dArgs a;
  memset(&a, 0, sizeof(a));
  a.D = D, a.H = H, a.W = W;
  int rc;
  // conv 0: 1 -> C on the raw cost
  a.in = cost, a.Cin = 1, a.out = bufA, a.affine = pk;
  a.w = pk + packed_offset(C, layers, 0, false), a.bias = pk + packed_offset(C, layers, 0, true);
  rc = launch_conv3d<1, C, Q, TD, TW, true, false>(a, B, st);
  if (rc) return rc;
  float* cur = bufA;
  float* nxt = bufB;
  for (int i = 1; i <= layers; ++i) {
    a.in = cur, a.Cin = C, a.out = nxt;
    a.w = pk + packed_offset(C, layers, i, false), a.bias = pk + packed_offset(C, layers, i, true);
    rc = launch_conv3d<CK, C, Q, TD, TW, false, false>(a, B, st);
    if (rc) return rc;
    float* t = cur;
    cur = nxt, nxt = t;
  }
  a.in = cur, a.Cin = C, a.out = out, a.skip = add_skip ? cost : nullptr;
  a.w = pk + packed_offset(C, layers, layers + 1, false), a.bias = nullptr;
  return launch_conv3d<4, 1, 1, 4, 64, false, true>(a, B, st);
}

// any other width: groups of 8 output channels on the FFMA kernel (Cp = padded width)
static int run_stack_grouped(const float* cost, const float* pk, float* out, float* bufA, float* bufB, int B, int D, int H, int W,
                             int C, int layers, int add_skip, cudaStream_t st) {
  const int Cp = padded_width(C), ng = Cp / 8;
  Conv3dArgs a;
  int rc;
  for (int g = 0; g < ng; ++g) {
    memset(&a, 0, sizeof(a));
    a.D = D, a.H = H, a.W = W, a.in = cost, a.Cin = 1, a.out = bufA, a.affine = pk, a.cout_total = Cp, a.cout_off = g * 8;
    a.w = pk + packed_offset(C, layers, 0, false) + (size_t)g * 27 * 8, a.bias = pk + packed_offset(C, layers, 0, true) + g * 8;
    if ((rc = launch_conv3d<1, 8, 8, 3, 64, true, false>(a, B, st))) return rc;
  }
  float* cur = bufA;
  float* nxt = bufB;
  for (int i = 1; i <= layers; ++i) {
    for (int g = 0; g < ng; ++g) {
      memset(&a, 0, sizeof(a));
      a.D = D, a.H = H, a.W = W, a.in = cur, a.Cin = Cp, a.out = nxt, a.cout_total = Cp, a.cout_off = g * 8;
      a.w = pk + packed_offset(C, layers, i, false) + (size_t)g * Cp * 27 * 8, a.bias = pk + packed_offset(C, layers, i, true) + g * 8;
      if ((rc = launch_conv3d<4, 8, 8, 3, 64, false, false>(a, B, st))) return rc;
    }
    float* t = cur;
    cur = nxt, nxt = t;
  }
  memset(&a, 0, sizeof(a));
  a.D = D, a.H = H, a.W = W, a.in = cur, a.Cin = Cp, a.out = out, a.skip = add_skip ? cost : nullptr;
  a.w = pk + packed_offset(C, layers, layers + 1, false), a.bias = nullptr;
  return launch_conv3d<4, 1, 1, 4, 64, false, true>(a, B, st);
}

}  // namespace lws

extern "C" int lws_conv3d_stack_launches(int C, int layers) {
  if (C <= 0 || layers < 0) return 0;
  return lws::native_width(C) ? layers + 2 : (layers + 1) * (lws::padded_width(C) / 8) + 1;
}

extern "C" size_t lws_conv3d_stack_packed_floats(int C, int layers) {
  if (C <= 0 || layers < 0) return 0;
  return lws::packed_offset(C, layers, layers + 2, false) + (lws::has_tc_tables(C) ? (size_t)layers * lws::kTcLayerFloats : 0);
}

extern "C" int lws_pack_conv3d_stack_weights(const float* const* conv_w, const float* const* bn_weight,
                                             const float* const* bn_bias, const float* const* bn_mean,
                                             const float* const* bn_var, float eps, int C, int layers,
                                             float* packed) {
  using namespace lws;
  if (!conv_w || !bn_weight || !bn_bias || !bn_mean || !bn_var || !packed) return LWS_ERR_NULL_PTR;
  if (C <= 0 || layers < 0) return LWS_ERR_BAD_SHAPE;
  const int n = layers + 2;
  memset(packed, 0, lws_conv3d_stack_packed_floats(C, layers) * sizeof(float));
  auto bn_scale = [&](int i, int c) { return (double)bn_weight[i][c] / sqrt((double)bn_var[i][c] + (double)eps); };
  auto bn_shift = [&](int i, int c) { return (double)bn_bias[i][c] - (double)bn_mean[i][c] * bn_scale(i, c); };
  packed[0] = (float)bn_scale(0, 0);
  packed[1] = (float)bn_shift(0, 0);
  const bool native = native_width(C);
  const int Cp = padded_width(C);
  for (int i = 0; i < n; ++i) {
    const int cin = i == 0 ? 1 : C, cout = i == n - 1 ? 1 : C;
    const int cinp = i == 0 ? 1 : Cp;  // padded input width the kernels iterate over (padding weights stay 0)
    float* w = packed + packed_offset(C, layers, i, false);
    float* bias = packed + packed_offset(C, layers, i, true);
    for (int co = 0; co < cout; ++co) {
      const double s = (i < n - 1) ? bn_scale(i + 1, co) : 1.0;
      if (i < n - 1) bias[co] = (float)bn_shift(i + 1, co);
      for (int ci = 0; ci < cin; ++ci)
        for (int t = 0; t < 27; ++t) {
          const float v = (float)((double)conv_w[i][((size_t)co * cin + ci) * 27 + t] * s);
          if (native || cout == 1) w[((size_t)ci * 27 + t) * cout + co] = v;
          else w[(((size_t)(co >> 3) * cinp + ci) * 27 + t) * 8 + (co & 7)] = v;  // group-major [Cp/8][Cin_p][27][8]
        }
    }
  }
  if (C == 32) {
    // split-fp16 Toeplitz-N operand tables (conv3d_f16.cu), per 32 -> 32 layer: 9 blocks (kh, kw), two per 24 KB tile;
    // tile row n = part*96 + kd*32 + co (part 0: hi = fp16(w * sw), part 1: lo = fp16((w * sw - hi) * 2^11)), 64 halves per
    // row = [block 2i: ci 0..31 | block 2i+1: ci 0..31]; sw = the power of two that puts max|w| into [256, 512).  The
    // epilogue scales 1/sw and 1/(sw * 2^11) follow the 5 tiles.
    for (int l = 0; l < layers; ++l) {
      const float* wf = packed + packed_offset(C, layers, l + 1, false);  // [ci][27][co]
      float* tc = packed + packed_tc_offset(C, layers, l);
      memset(tc, 0, kTcLayerFloats * sizeof(float));
      float mx = 0.f;
      for (int i = 0; i < 32 * 27 * 32; ++i) mx = fmaxf(mx, fabsf(wf[i]));
      int e = 0;
      if (mx > 0.f) frexpf(mx, &e);
      const float sw = mx > 0.f ? ldexpf(1.f, 9 - e) : 1.f;
      __half* h = reinterpret_cast<__half*>(tc);
      for (int kh = 0; kh < 3; ++kh)
        for (int kw = 0; kw < 3; ++kw) {
          const int blk = kh * 3 + kw;
          for (int kd = 0; kd < 3; ++kd)
            for (int co = 0; co < 32; ++co)
              for (int ci = 0; ci < 32; ++ci) {
                const float w = wf[((size_t)ci * 27 + kd * 9 + kh * 3 + kw) * 32 + co] * sw;
                const __half hi = __float2half_rn(w);
                const size_t base = ((size_t)(blk >> 1) * 192) * 64 + (blk & 1) * 32 + ci;
                h[base + (size_t)(kd * 32 + co) * 64] = hi;
                h[base + (size_t)(96 + kd * 32 + co) * 64] = __float2half_rn((w - __half2float(hi)) * 2048.f);
              }
        }
      tc[5 * 192 * 32] = 1.f / sw;
      tc[5 * 192 * 32 + 1] = 1.f / (sw * 2048.f);
    }
    if (layers > 0) {
      // closing 32 -> 1 conv as a Toeplitz-N layer with N = 16, kept behind the first mid layer's table (float offset
      // kTcLastOff of its slot): block (kh, kw) = 16 rows x 64 halves [B1 | B2]; B1 row kd = wh, row 8+kd = wl (applied to
      // the hi halves of a voxel), B2 row 8+kd = wh (applied to the lo halves); scales follow the 9 blocks.
      const float* wl_ = packed + packed_offset(C, layers, layers + 1, false);  // [ci][27][1]
      float* tc = packed + packed_tc_offset(C, layers, 0) + kTcLastOff;
      float mx = 0.f;
      for (int i = 0; i < 32 * 27; ++i) mx = fmaxf(mx, fabsf(wl_[i]));
      int e = 0;
      if (mx > 0.f) frexpf(mx, &e);
      const float sw = mx > 0.f ? ldexpf(1.f, 9 - e) : 1.f;
      __half* h = reinterpret_cast<__half*>(tc);
      for (int kh = 0; kh < 3; ++kh)
        for (int kw = 0; kw < 3; ++kw)
          for (int kd = 0; kd < 3; ++kd)
            for (int ci = 0; ci < 32; ++ci) {
              const float w = wl_[ci * 27 + kd * 9 + kh * 3 + kw] * sw;
              const __half hi = __float2half_rn(w);
              __half* blk = h + (size_t)(kh * 3 + kw) * 16 * 64;
              blk[kd * 64 + ci] = hi;
              blk[(8 + kd) * 64 + ci] = __float2half_rn((w - __half2float(hi)) * 2048.f);
              blk[(8 + kd) * 64 + 32 + ci] = hi;
            }
      tc[9 * 16 * 32] = 1.f / sw;
      tc[9 * 16 * 32 + 1] = 1.f / (sw * 2048.f);
    }
  } else if (C == 8) {
    // split-fp16 operand tables (conv3d_c8.cu), per 8 -> 8 layer: 9 taps t = kd*3 + kh of 1536 bytes each, SWIZZLE_NONE K-major:
    // K chunk 0 (applied to the hi halves of the voxel) = 48 rows x 8 halves, K chunk 1 (applied to the lo halves) 768 bytes
    // further; row n = part*24 + kw*8 + co: part 0 = [wh | 0], part 1 = [wl | wh]; wh = fp16(w * sw),
    // wl = fp16((w * sw - wh) * 2^11), sw = the power of two that puts max|w| into [256, 512).  Scales 1/sw, 1/(sw*2^11) follow.
    for (int l = 0; l < layers; ++l) {
      const float* wf = packed + packed_offset(C, layers, l + 1, false);  // [ci][27][co]
      float* tc = packed + packed_tc_offset(C, layers, l);
      memset(tc, 0, kTcLayerFloats * sizeof(float));
      float mx = 0.f;
      for (int i = 0; i < 8 * 27 * 8; ++i) mx = fmaxf(mx, fabsf(wf[i]));
      int e = 0;
      if (mx > 0.f) frexpf(mx, &e);
      const float sw = mx > 0.f ? ldexpf(1.f, 9 - e) : 1.f;
      __half* h = reinterpret_cast<__half*>(tc);
      for (int kd = 0; kd < 3; ++kd)
        for (int kh = 0; kh < 3; ++kh)
          for (int kw = 0; kw < 3; ++kw)
            for (int co = 0; co < 8; ++co)
              for (int ci = 0; ci < 8; ++ci) {
                const float w = wf[((size_t)ci * 27 + kd * 9 + kh * 3 + kw) * 8 + co] * sw;
                const __half hi = __float2half_rn(w);
                const __half lo = __float2half_rn((w - __half2float(hi)) * 2048.f);
                __half* t = h + (size_t)(kd * 3 + kh) * 768;  // 1536 bytes per tap
                const int n0 = kw * 8 + co, n1 = 24 + kw * 8 + co;
                t[n0 * 8 + ci] = hi;            // chunk 0, part 0
                t[n1 * 8 + ci] = lo;            // chunk 0, part 1
                t[384 + n1 * 8 + ci] = hi;      // chunk 1 (768 bytes further), part 1
              }
      tc[9 * 384] = 1.f / sw;
      tc[9 * 384 + 1] = 1.f / (sw * 2048.f);
    }
    if (layers > 0) {
      // closing 8 -> 1 conv with N = 16, behind the first mid layer's table (float offset kTcLastOff): per tap (kd, kh) 512
      // bytes = K chunk 0 [16 rows x 8 halves] then K chunk 1; row kw = [wh | 0], row 8 + kw = [wl | wh]; scales follow
      const float* wl_ = packed + packed_offset(C, layers, layers + 1, false);  // [ci][27][1]
      float* tc = packed + packed_tc_offset(C, layers, 0) + kTcLastOff;
      float mx = 0.f;
      for (int i = 0; i < 8 * 27; ++i) mx = fmaxf(mx, fabsf(wl_[i]));
      int e = 0;
      if (mx > 0.f) frexpf(mx, &e);
      const float sw = mx > 0.f ? ldexpf(1.f, 9 - e) : 1.f;
      __half* h = reinterpret_cast<__half*>(tc);
      for (int kd = 0; kd < 3; ++kd)
        for (int kh = 0; kh < 3; ++kh)
          for (int kw = 0; kw < 3; ++kw)
            for (int ci = 0; ci < 8; ++ci) {
              const float w = wl_[ci * 27 + kd * 9 + kh * 3 + kw] * sw;
              const __half hi = __float2half_rn(w);
              __half* t = h + (size_t)(kd * 3 + kh) * 256;  // 512 bytes per tap
              t[kw * 8 + ci] = hi;
              t[(8 + kw) * 8 + ci] = __float2half_rn((w - __half2float(hi)) * 2048.f);
              t[128 + (8 + kw) * 8 + ci] = hi;
            }
      tc[9 * 128] = 1.f / sw;
      tc[9 * 128 + 1] = 1.f / (sw * 2048.f);
    }
  }
  return LWS_OK;
}

extern "C" size_t lws_conv3d_stack_workspace_bytes(int B, int D, int H, int W, int C, int layers) {
  (void)layers;
  if (B <= 0 || D <= 0 || H <= 0 || W <= 0 || C <= 0) return 0;
  const size_t act = ((size_t)B * lws::padded_width(C) * D * H * W * sizeof(float) + 255) / 256 * 256;
  size_t tc = 0;
  if (C == 32 && lws::conv3d_f16_workspace_bytes(B, D, H, W) > tc) tc = lws::conv3d_f16_workspace_bytes(B, D, H, W);
  if (C == 8 && lws::conv3d_c8_workspace_bytes(B, D, H, W) > tc) tc = lws::conv3d_c8_workspace_bytes(B, D, H, W);
  return 2 * act > tc ? 2 * act : tc;
}

extern "C" int lws_conv3d_stack_f32(const float* cost, const float* packed_weights, float* out, void* ws,
                                    size_t ws_bytes, int B, int D, int H, int W, int C, int layers, int add_skip,
                                    lws_stream_t stream) {
  using namespace lws;
  LWS_CHECK_PTR(cost);
  LWS_CHECK_PTR(packed_weights);
  LWS_CHECK_PTR(out);
  LWS_CHECK_PTR(ws);
  if (B <= 0 || D <= 0 || H <= 0 || W <= 0 || layers < 0 || B > 65535) return LWS_ERR_BAD_SHAPE;
  if (C <= 0) return LWS_ERR_BAD_SHAPE;
  if (ws_bytes < lws_conv3d_stack_workspace_bytes(B, D, H, W, C, layers)) return LWS_ERR_WORKSPACE_TOO_SMALL;
  if ((((uintptr_t)ws) | ((uintptr_t)packed_weights)) & 15) return LWS_ERR_BAD_ALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  // the tensor-core paths need at least one mid layer, C = 32 also 128 + 2 (D + 2) <= 256 box rows; otherwise the FFMA kernels run
  if (use_tc_path(C) && !(C == 32 && 128 + 2 * (D + 2) > 256) && layers > 0) {
    const float* pk = packed_weights;
    const float* wtc[16];
    const float* bmid[16];
    if (layers > 16) return LWS_ERR_UNSUPPORTED;
    for (int l = 0; l < layers; ++l) {
      wtc[l] = pk + packed_tc_offset(C, layers, l);
      bmid[l] = pk + packed_offset(C, layers, l + 1, true);
    }
    if (C == 32)
      return conv3d_stack_f16(cost, pk, pk + packed_offset(C, layers, 0, false), pk + packed_offset(C, layers, 0, true), wtc, bmid,
                              layers, pk + packed_tc_offset(C, layers, 0) + kTcLastOff, out, ws, B, D, H, W, add_skip, st);
    if (C == 8) {
      // option "c8_group" = G > 0: depth-first over groups of G pairs (all six layers of a group before the next group), so that a
      // group's ping-pong activation planes are overwritten while they are still in L2 instead of streaming through HBM
      const int G = opt(OPT_C8_GROUP) > 0 ? opt(OPT_C8_GROUP) : B;
      const long long dhw = (long long)D * H * W;
      for (int b0 = 0; b0 < B; b0 += G) {
        const int nb = B - b0 < G ? B - b0 : G;
        const int rc = conv3d_stack_c8(cost + b0 * dhw, pk, pk + packed_offset(C, layers, 0, false), pk + packed_offset(C, layers, 0, true),
                                       wtc, bmid, layers, pk + packed_tc_offset(C, layers, 0) + kTcLastOff, out + b0 * dhw, ws, nb, D, H,
                                       W, add_skip, st);
        if (rc) return rc;
      }
      return LWS_OK;
    }
    return LWS_ERR_UNSUPPORTED;
  }
  const size_t act = ((size_t)B * padded_width(C) * D * H * W * sizeof(float) + 255) / 256 * 256;
  float* bufA = (float*)ws;
  float* bufB = (float*)((char*)ws + act);
  if (!native_width(C)) return run_stack_grouped(cost, packed_weights, out, bufA, bufB, B, D, H, W, C, layers, add_skip, st);
  // BN_0's scalar affine is the first two floats of the device blob; the first conv kernel reads it from there.
  switch (C) {
    case 8:
      return run_stack<8>(cost, packed_weights, out, bufA, bufB, B, D, H, W, layers, add_skip, st);
    case 16:
      return run_stack<16>(cost, packed_weights, out, bufA, bufB, B, D, H, W, layers, add_skip, st);
    default:
      return run_stack<32>(cost, packed_weights, out, bufA, bufB, B, D, H, W, layers, add_skip, st);
  }
}

// Stage 1 of the network in one call: the L1 volume (a2) is built inside the stack's first conv kernel (north_star item 2).
// cost_out receives the raw volume (it is the skip input, models/models.py:137, and what a caller of _build_volume_2d gets).
extern "C" int lws_cost_volume_conv3d_stack_supported(int B, int Cf, int H, int W, int maxdisp, int C, int layers) {
  using namespace lws;
  if (B <= 0 || Cf <= 0 || H <= 0 || W <= 0 || maxdisp <= 0 || layers <= 0 || layers > 16 || B > 65535) return LWS_ERR_BAD_SHAPE;
  if (C != 32 || !use_tc_path(C) || 128 + 2 * (maxdisp + 2) > 256) return LWS_ERR_UNSUPPORTED;
  return cost_first_conv_fused_supported(Cf, maxdisp, H, W);
}

extern "C" int lws_cost_volume_conv3d_stack_f32(const float* L, const float* R, const float* packed_weights, float* cost_out,
                                                float* out, void* ws, size_t ws_bytes, int B, int Cf, int H, int W, int maxdisp,
                                                int C, int layers, int add_skip, lws_stream_t stream) {
  using namespace lws;
  LWS_CHECK_PTR(L);
  LWS_CHECK_PTR(R);
  LWS_CHECK_PTR(packed_weights);
  LWS_CHECK_PTR(cost_out);
  LWS_CHECK_PTR(out);
  LWS_CHECK_PTR(ws);
  const int rc = lws_cost_volume_conv3d_stack_supported(B, Cf, H, W, maxdisp, C, layers);
  if (rc) return rc;
  const int D = maxdisp;
  if (ws_bytes < lws_conv3d_stack_workspace_bytes(B, D, H, W, C, layers)) return LWS_ERR_WORKSPACE_TOO_SMALL;
  if ((((uintptr_t)ws) | ((uintptr_t)packed_weights)) & 15) return LWS_ERR_BAD_ALIGN;
  if ((((uintptr_t)L) | ((uintptr_t)R)) & 7) return LWS_ERR_BAD_ALIGN;
  const float* pk = packed_weights;
  const float* wtc[16];
  const float* bmid[16];
  for (int l = 0; l < layers; ++l) {
    wtc[l] = pk + packed_tc_offset(C, layers, l);
    bmid[l] = pk + packed_offset(C, layers, l + 1, true);
  }
  return conv3d_stack_f16(cost_out, pk, pk + packed_offset(C, layers, 0, false), pk + packed_offset(C, layers, 0, true), wtc, bmid,
                          layers, pk + packed_tc_offset(C, layers, 0) + kTcLastOff, out, ws, B, D, H, W, add_skip,
                          (cudaStream_t)stream, L, R, Cf);
}

// One BN-folded C -> C layer of the stack on its own (the kernel the stack spends its time in): used by bench.py to
// time the dominant kernel in isolation and by tests to check a single layer.
extern "C" int lws_conv3d_bnrelu_layer_f32(const float* in, const float* w_folded, const float* bias, float* out,
                                           int B, int C, int D, int H, int W, lws_stream_t stream) {
  using namespace lws;
  LWS_CHECK_PTR(in);
  LWS_CHECK_PTR(w_folded);
  LWS_CHECK_PTR(bias);
  LWS_CHECK_PTR(out);
  if (B <= 0 || D <= 0 || H <= 0 || W <= 0 || B > 65535) return LWS_ERR_BAD_SHAPE;
  Conv3dArgs a;
  memset(&a, 0, sizeof(a));
  a.in = in, a.w = w_folded, a.bias = bias, a.out = out, a.Cin = C, a.D = D, a.H = H, a.W = W;
  cudaStream_t st = (cudaStream_t)stream;
  switch (C) {
    case 8: return launch_conv3d<4, 8, 8, 3, 64, false, false>(a, B, st);
    case 16: return launch_conv3d<4, 16, 8, 3, 64, false, false>(a, B, st);
    case 32: return launch_conv3d<8, 32, 8, 2, 32, false, false>(a, B, st);
    default: return LWS_ERR_UNSUPPORTED;
  }
}
