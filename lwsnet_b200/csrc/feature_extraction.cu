// Feature pyramid (reference models/submodules.py:5-188: convbn / deconvbn / hourglass / feature_extraction).
// SURVEY.md 8(f) "next" row n1: the step in front of the hot path, 2.3 of 91.6 GFLOP per pair, 3..16 channels, so
// every layer is bandwidth / latency bound rather than FLOP bound.  One generic direct kernel covers the 12 layers:
// 3x3 conv with stride 1|2 and dilation, or 3x3 stride-2 transposed conv (pad 1, output_padding 1), BatchNorm folded
// on the host (scale into the weights, shift as bias), optional residual add and ReLU fused in the epilogue.
// fp32 end to end (TF32 feature maps move the disparities by whole pixels).
#include <math.h>
#include <string.h>

#include "lws_common.cuh"
#include "tma_utils.cuh"

namespace lws {

struct FeConvArgs {
  const float* in;    // [B,Cin,Hi,Wi]
  const float* w;     // [Cin][9][COUT] folded
  const float* bias;  // [COUT]
  const float* res;   // [B,COUT,Ho,Wo] or null
  float* out;         // [B,COUT,Ho,Wo]
  int Cin, Hi, Wi, Ho, Wo;
  int stride, dil, pad, transposed, relu;
};

// CT = output channels per thread: the 1/8-resolution layers of a few pairs are only ~14k output pixels per image, so they are
// spread over COUT / CT times more threads (blockIdx.z = channel group) instead of leaving most SMs idle.
template <int COUT, int CT>
__global__ void __launch_bounds__(128) fe_conv_kernel(const FeConvArgs a) {
  extern __shared__ __align__(16) float sW[];  // [Cin][9][COUT]
  for (int i = threadIdx.x; i < a.Cin * 9 * COUT; i += blockDim.x) sW[i] = __ldg(a.w + i);
  __syncthreads();
  // flattened (row, 4-pixel group) index so blocks stay full whatever the row length is
  const int wq = (a.Wo + 3) >> 2;
  const int item = blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= wq * a.Ho) return;
  const int yo = item / wq;
  const int b = blockIdx.y;
  const int x0 = (item - yo * wq) * 4;
  const int c0 = blockIdx.z * CT;
  const long long in_hw = (long long)a.Hi * a.Wi;
  const long long out_hw = (long long)a.Ho * a.Wo;
  const float* in_b = a.in + (long long)b * a.Cin * in_hw;

  float acc[4][CT];
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int q = 0; q < CT; ++q) acc[p][q] = 0.f;

  // input coordinates of the 3 taps per axis (-1 = contributes nothing)
  int yi[3], xi[3][4];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    if (a.transposed) {
      const int t = yo + a.pad - k;  // yo = 2*yi - pad + k
      yi[k] = (t >= 0 && (t & 1) == 0 && (t >> 1) < a.Hi) ? (t >> 1) : -1;
    } else {
      const int t = yo * a.stride - a.pad + k * a.dil;
      yi[k] = (t >= 0 && t < a.Hi) ? t : -1;
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const int xo = x0 + p;
      if (a.transposed) {
        const int t = xo + a.pad - k;
        xi[k][p] = (t >= 0 && (t & 1) == 0 && (t >> 1) < a.Wi) ? (t >> 1) : -1;
      } else {
        const int t = xo * a.stride - a.pad + k * a.dil;
        xi[k][p] = (t >= 0 && t < a.Wi) ? t : -1;
      }
    }
  }

  for (int ci = 0; ci < a.Cin; ++ci) {
    const float* plane = in_b + ci * in_hw;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      if (yi[ky] < 0) continue;
      const float* row = plane + (long long)yi[ky] * a.Wi;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        float v[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) v[p] = xi[kx][p] >= 0 ? __ldg(row + xi[kx][p]) : 0.f;
        const float* wp = sW + (ci * 9 + ky * 3 + kx) * COUT + c0;
#pragma unroll
        for (int q = 0; q < CT; q += 4) {
          const float4 w4 = *reinterpret_cast<const float4*>(wp + q);
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            acc[p][q] = fmaf(v[p], w4.x, acc[p][q]);
            acc[p][q + 1] = fmaf(v[p], w4.y, acc[p][q + 1]);
            acc[p][q + 2] = fmaf(v[p], w4.z, acc[p][q + 2]);
            acc[p][q + 3] = fmaf(v[p], w4.w, acc[p][q + 3]);
          }
        }
      }
    }
  }

  const bool vec = ((a.Wo & 3) == 0);
#pragma unroll
  for (int q = 0; q < CT; ++q) {
    const float bias = __ldg(a.bias + c0 + q);
    const long long o = ((long long)b * COUT + c0 + q) * out_hw + (long long)yo * a.Wo + x0;
    float r[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) r[p] = acc[p][q] + bias;
    if (vec) {
      if (a.res) {
        const float4 rv = *reinterpret_cast<const float4*>(a.res + o);
        r[0] += rv.x, r[1] += rv.y, r[2] += rv.z, r[3] += rv.w;
      }
      if (a.relu) {
#pragma unroll
        for (int p = 0; p < 4; ++p) r[p] = fmaxf(r[p], 0.f);
      }
      *reinterpret_cast<float4*>(a.out + o) = make_float4(r[0], r[1], r[2], r[3]);
    } else {
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        if (x0 + p < a.Wo) {
          float t = r[p] + (a.res ? a.res[o + p] : 0.f);
          a.out[o + p] = a.relu ? fmaxf(t, 0.f) : t;
        }
      }
    }
  }
}


// ---- fast path 1: 3x3 conv with dilation DIL (pad = DIL), stride 1 or 2, even row pitch; software pipelined ---------------------
// A thread owns 4 consecutive output pixels; per (ci, ky) three aligned, fully coalesced 128-bit loads [x0-4, x0+8) cover all
// three kx taps of its four pixels for DIL in {1, 2, 4} (the generic kernel issues 12 predicated scalar loads for the same data).
// The layers of the pyramid are small (a 1/8-resolution map of a few pairs is ~100k pixels), so a kernel whose threads wait for
// their loads once per input channel is latency bound however many FLOPs the machine has.  Here a thread owns 4 consecutive output
// pixels x CT output channels (blockIdx.z = channel group, so the small layers still fill the SMs), the three window rows of input
// channel ci+1 are loaded into a second register set while channel ci is being multiplied (ping-pong), and rows whose
// pitch is only 8-byte aligned (W = 154 at 1/8 of KITTI) use 64-bit loads / stores instead of falling back to scalar taps.
template <int COUT, int CT, int DIL, int STRIDE, bool V4, int WIN = 12, bool ODD = false>
__global__ void __launch_bounds__(128) fe_conv_pipe_kernel(const FeConvArgs a) {
  extern __shared__ __align__(16) float sW[];  // [Cin][9][COUT]
  for (int i = threadIdx.x; i < a.Cin * 9 * COUT; i += blockDim.x) sW[i] = __ldg(a.w + i);
  __syncthreads();
  static_assert(4 + 3 * STRIDE + DIL < WIN, "window too narrow for this stride / dilation");
  constexpr int G = V4 ? 4 : 2, NG = WIN / G;
  const int wq = (a.Wo + 3) >> 2;
  const int item = blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= wq * a.Ho) return;
  const int yo = item / wq;
  const int b = blockIdx.y, c0 = blockIdx.z * CT;
  const int x0 = (item - yo * wq) * 4;
  const long long hw = (long long)a.Hi * a.Wi, ohw = (long long)a.Ho * a.Wo;
  const int xi0 = x0 * STRIDE;  // the aligned WIN-wide window starts at input column xi0 - 4
  const float* in_b = a.in + (long long)b * a.Cin * hw + (xi0 - 4);
  if (a.res && (threadIdx.x & 7) == 0) {  // the skip tensor is read by the epilogue: request its lines into L2 now
#pragma unroll
    for (int c = 0; c < CT; ++c)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(a.res + ((long long)b * COUT + c0 + c) * ohw + (long long)yo * a.Wo + x0));
  }
  bool okc[NG], oky[3];
  int rowoff[3];
#pragma unroll
  for (int g = 0; g < NG; ++g) okc[g] = xi0 - 4 + g * G >= 0 && xi0 - 4 + g * G + G <= a.Wi;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int yi = yo * STRIDE + (ky - 1) * DIL;
    oky[ky] = (unsigned)yi < (unsigned)a.Hi;
    rowoff[ky] = yi * a.Wi;
  }
  auto load = [&](int ci, float (&w)[3][WIN]) {
    const float* plane = in_b + ci * hw;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        if constexpr (V4) {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (oky[ky] && okc[g]) v = __ldg(reinterpret_cast<const float4*>(plane + rowoff[ky]) + g);
          w[ky][4 * g] = v.x, w[ky][4 * g + 1] = v.y, w[ky][4 * g + 2] = v.z, w[ky][4 * g + 3] = v.w;
        } else {
          float2 v = make_float2(0.f, 0.f);
          if (oky[ky] && okc[g]) v = __ldg(reinterpret_cast<const float2*>(plane + rowoff[ky]) + g);
          w[ky][2 * g] = v.x, w[ky][2 * g + 1] = v.y;
        }
      }
    }
  };
  float2 acc[4][CT / 2];  // float2 pairs: FFMA2 (fma.rn.f32x2, sm_100) does two output channels per issued instruction
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int q = 0; q < CT / 2; ++q) acc[p][q] = make_float2(0.f, 0.f);
  auto mac = [&](int ci, const float (&w)[3][WIN]) {
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const float* wp = sW + (ci * 9 + ky * 3 + kx) * COUT + c0;
#pragma unroll
        for (int q = 0; q < CT; q += 4) {
          const float4 w4 = *reinterpret_cast<const float4*>(wp + q);
          const float2 w01 = make_float2(w4.x, w4.y), w23 = make_float2(w4.z, w4.w);
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            const float v = w[ky][4 + p * STRIDE + (kx - 1) * DIL];
            const float2 vv = make_float2(v, v);
            acc[p][q / 2] = __ffma2_rn(vv, w01, acc[p][q / 2]);
            acc[p][q / 2 + 1] = __ffma2_rn(vv, w23, acc[p][q / 2 + 1]);
          }
        }
      }
    }
  };
  float wa[3][WIN], wb[3][WIN];
  load(0, wa);
  if constexpr (ODD) {  // odd channel count (the RGB input layer)
    for (int ci = 0; ci < a.Cin; ci += 2) {
      if (ci + 1 < a.Cin) load(ci + 1, wb);
      mac(ci, wa);
      if (ci + 2 < a.Cin) load(ci + 2, wa);
      if (ci + 1 < a.Cin) mac(ci + 1, wb);
    }
  } else {
    for (int ci = 0; ci < a.Cin; ci += 2) {
      load(ci + 1, wb);
      mac(ci, wa);
      if (ci + 2 < a.Cin) load(ci + 2, wa);
      mac(ci + 1, wb);
    }
  }
#pragma unroll
  for (int q = 0; q < CT; ++q) {
    const float bias = __ldg(a.bias + c0 + q);
    const long long o = ((long long)b * COUT + c0 + q) * ohw + (long long)yo * a.Wo + x0;
    float r[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) r[p] = (q & 1 ? acc[p][q / 2].y : acc[p][q / 2].x) + bias;
    if constexpr (V4) {
      if (a.res) {
        const float4 rv = *reinterpret_cast<const float4*>(a.res + o);
        r[0] += rv.x, r[1] += rv.y, r[2] += rv.z, r[3] += rv.w;
      }
      if (a.relu) {
#pragma unroll
        for (int p = 0; p < 4; ++p) r[p] = fmaxf(r[p], 0.f);
      }
      *reinterpret_cast<float4*>(a.out + o) = make_float4(r[0], r[1], r[2], r[3]);
    } else {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (x0 + 2 * h < a.Wo) {  // Wo is even: a pixel pair is wholly in or out
          float u0 = r[2 * h], u1 = r[2 * h + 1];
          if (a.res) {
            const float2 rv = *reinterpret_cast<const float2*>(a.res + o + 2 * h);
            u0 += rv.x, u1 += rv.y;
          }
          if (a.relu) u0 = fmaxf(u0, 0.f), u1 = fmaxf(u1, 0.f);
          *reinterpret_cast<float2*>(a.out + o + 2 * h) = make_float2(u0, u1);
        }
      }
    }
  }
}

// ---- fast path 1b: stride-1 3x3 conv (dilation DIL, pad = DIL) with its input tile delivered by ONE TMA load ------------------------
// fe_conv_pipe_kernel above is bound by the latency of its per-thread global loads (ncu: 34 % of its stall samples on their
// scoreboard at 16 resident warps, issue slots 45 % used).  Here a block owns a TH x (4 TQ) output tile (TQ * TH = 128 threads, one
// 4-pixel quad each) and ONE tiled TMA load brings the (TH + 2 DIL) x (4 TQ + 8) window of all Cin input channels into shared memory:
// no load instructions in the threads, out-of-bounds zero fill = the conv's zero padding, and the blocks resident on an SM overlap
// each other's load.  The window starts at column x0 - 4 (the innermost start coordinate of a tiled TMA load must be 16-byte aligned,
// tools/tma_probe.cu), so the three aligned 128-bit shared-memory reads per (ci, ky) are the same 12-float register window as above.
// Needs 16-byte aligned rows (Wi % 4 == 0: the 1/2- and 1/4-resolution maps of KITTI; the 1/8 maps have W = 154).
template <int COUT, int CT, int DIL, int TQ>
__global__ void __launch_bounds__(128) fe_conv_tile_kernel(const __grid_constant__ CUtensorMap map_in, const FeConvArgs a) {
  constexpr int TH = 128 / TQ, ROWS = TH + 2 * DIL, COLS = 4 * TQ + 8, NG = COUT / CT;
  extern __shared__ __align__(128) uint8_t fe_smem[];
  float* sIn = reinterpret_cast<float*>(fe_smem);                       // [Cin][ROWS][COLS]
  float* sW = sIn + a.Cin * ROWS * COLS;                                // [Cin][9][COUT]
  uint64_t* bar = reinterpret_cast<uint64_t*>(sW + a.Cin * 9 * COUT);   // 8-byte aligned: all counts above are even
  const int tid = threadIdx.x;
  const int b = blockIdx.z / NG, c0 = (blockIdx.z - b * NG) * CT;
  const int tx0 = blockIdx.x * (4 * TQ), ty0 = blockIdx.y * TH;
  // the layer's weights ride on the same barrier as one bulk copy when they are 16-byte aligned (they are inside the packed blob)
  const bool wbulk = (reinterpret_cast<uintptr_t>(a.w) & 15) == 0;
  const uint32_t wbytes = (uint32_t)(a.Cin * 9 * COUT * 4);
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
    mbar_expect_tx(bar, (uint32_t)(a.Cin * ROWS * COLS * 4) + (wbulk ? wbytes : 0u));
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                     smem_u32(sIn)),
                 "l"(reinterpret_cast<uint64_t>(&map_in)), "r"(smem_u32(bar)), "r"(tx0 - 4), "r"(ty0 - DIL), "r"(b * a.Cin)
                 : "memory");
    if (wbulk)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sW)),
                   "l"(a.w), "r"(wbytes), "r"(smem_u32(bar))
                   : "memory");
  }
  if (!wbulk)
    for (int i = tid; i < a.Cin * 9 * COUT; i += 128) sW[i] = __ldg(a.w + i);
  __syncthreads();  // barrier initialised (and, on the fallback path, weights staged)
  const int q = tid % TQ, ty = tid / TQ;
  const int x0 = tx0 + 4 * q, yo = ty0 + ty;
  if (a.res && x0 < a.Wo && yo < a.Ho && (q & 7) == 0) {
    // the skip tensor is read by the epilogue: request its lines (128 bytes = the outputs of 8 neighbouring threads) into L2 now
#pragma unroll
    for (int c = 0; c < CT; ++c)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(a.res + ((long long)b * COUT + c0 + c) * ((long long)a.Ho * a.Wo) + (long long)yo * a.Wo + x0));
  }
  mbar_wait(bar, 0);
  if (x0 >= a.Wo || yo >= a.Ho) return;
  float2 acc[4][CT / 2];  // float2 pairs: FFMA2 (fma.rn.f32x2, sm_100) does two output channels per issued instruction
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int c = 0; c < CT / 2; ++c) acc[p][c] = make_float2(0.f, 0.f);
  const float* tile = sIn + ty * COLS + 4 * q;
  for (int ci = 0; ci < a.Cin; ++ci) {
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const float4* rp = reinterpret_cast<const float4*>(tile + (ci * ROWS + ky * DIL) * COLS);
      const float4 v0 = rp[0], v1 = rp[1], v2 = rp[2];
      const float w[12] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w};
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const float* wp = sW + (ci * 9 + ky * 3 + kx) * COUT + c0;
#pragma unroll
        for (int c = 0; c < CT; c += 4) {
          const float4 w4 = *reinterpret_cast<const float4*>(wp + c);
          const float2 w01 = make_float2(w4.x, w4.y), w23 = make_float2(w4.z, w4.w);
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            const float v = w[4 + p + (kx - 1) * DIL];
            const float2 vv = make_float2(v, v);
            acc[p][c / 2] = __ffma2_rn(vv, w01, acc[p][c / 2]);
            acc[p][c / 2 + 1] = __ffma2_rn(vv, w23, acc[p][c / 2 + 1]);
          }
        }
      }
    }
  }
  const long long ohw = (long long)a.Ho * a.Wo;
#pragma unroll
  for (int c = 0; c < CT; ++c) {
    const float bias = __ldg(a.bias + c0 + c);
    const long long o = ((long long)b * COUT + c0 + c) * ohw + (long long)yo * a.Wo + x0;
    float r[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) r[p] = (c & 1 ? acc[p][c / 2].y : acc[p][c / 2].x) + bias;
    if (a.res) {
      const float4 rv = *reinterpret_cast<const float4*>(a.res + o);
      r[0] += rv.x, r[1] += rv.y, r[2] += rv.z, r[3] += rv.w;
    }
    if (a.relu) {
#pragma unroll
      for (int p = 0; p < 4; ++p) r[p] = fmaxf(r[p], 0.f);
    }
    *reinterpret_cast<float4*>(a.out + o) = make_float4(r[0], r[1], r[2], r[3]);
  }
}

template <int COUT, int CT, int DIL, int TQ>
static int launch_tile(const FeConvArgs& a, int B, cudaStream_t st) {
  constexpr int TH = 128 / TQ, ROWS = TH + 2 * DIL, COLS = 4 * TQ + 8;
  const size_t smem = ((size_t)a.Cin * ROWS * COLS + (size_t)a.Cin * 9 * COUT) * sizeof(float) + 16;
  if (smem > 100 * 1024 || ROWS > 256 || a.Cin > 256) return LWS_ERR_UNSUPPORTED;
  CUtensorMap map;
  const uint64_t dims[3] = {(uint64_t)a.Wi, (uint64_t)a.Hi, (uint64_t)B * a.Cin};
  const uint64_t strides[2] = {(uint64_t)a.Wi * 4, (uint64_t)a.Hi * a.Wi * 4};
  const uint32_t box[3] = {(uint32_t)COLS, (uint32_t)ROWS, (uint32_t)a.Cin};
  int rc = make_tensor_map_f32(&map, a.in, 3, dims, strides, box, false);
  if (rc) return rc;
  LWS_SET_SMEM_ONCE((fe_conv_tile_kernel<COUT, CT, DIL, TQ>), 100 * 1024);
  dim3 g(cdiv(a.Wo, 4 * TQ), cdiv(a.Ho, TH), B * (COUT / CT));
  fe_conv_tile_kernel<COUT, CT, DIL, TQ><<<g, 128, smem, st>>>(map, a);
  cudaError_t e = cudaPeekAtLastError();
  return e == cudaSuccess ? LWS_OK : (int)e;
}

// ---- fast path 2: 3x3 stride-2 transposed conv (pad 1, output_padding 1 -> exact 2x upsampling), Wi % 2 == 0 ------------------
// yo = 2*yi - 1 + ky: an even output row sees only (yi = yo/2, ky = 1), an odd one (yi = (yo-1)/2, ky = 2) and (yi = (yo+1)/2,
// ky = 0); same along x.  A thread owns the 2 x 4 outputs (rows 2i, 2i+1; columns 4j .. 4j+3) of 8 output channels: 6 input
// values and 18 FMA per (ci, co) instead of the generic kernel's 72 predicated taps.
template <int COUT>
__global__ void __launch_bounds__(128) fe_deconv_kernel(const FeConvArgs a) {
  extern __shared__ __align__(16) float sW[];  // [Cin][9][COUT]
  for (int i = threadIdx.x; i < a.Cin * 9 * COUT; i += blockDim.x) sW[i] = __ldg(a.w + i);
  __syncthreads();
  constexpr int NG = COUT / 8;  // groups of 8 output channels
  const int wq = a.Wi >> 1;     // column quads per row pair (Wo / 4)
  int item = blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= wq * a.Hi * NG) return;
  const int cg = item % NG;
  item /= NG;
  const int i = item / wq, j = item - i * wq;
  const int b = blockIdx.y;
  const long long ihw = (long long)a.Hi * a.Wi, ohw = (long long)a.Ho * a.Wo;
  const float* in_b = a.in + (long long)b * a.Cin * ihw + (long long)i * a.Wi + 2 * j;
  const bool r1 = i + 1 < a.Hi, c2 = 2 * j + 2 < a.Wi;
  if (a.res && (threadIdx.x & 7) == 0) {  // the skip tensor is read by the epilogue: request its lines into L2 now
#pragma unroll
    for (int q = 0; q < 8; ++q)
#pragma unroll
      for (int r = 0; r < 2; ++r)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(a.res + ((long long)b * COUT + cg * 8 + q) * ohw + (long long)(2 * i + r) * a.Wo + 4 * j));
  }

  float2 acc[2][4][4];  // float2 pairs of output channels: FFMA2 (fma.rn.f32x2) halves the issued FMA instructions
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[r][p][q] = make_float2(0.f, 0.f);

  // the six input values of channel ci + 1 are loaded while channel ci is multiplied (ping-pong register sets)
  auto load6 = [&](int ci, float (&v0)[3], float (&v1)[3]) {
    const float* p0 = in_b + ci * ihw;
    const float2 a01 = __ldg(reinterpret_cast<const float2*>(p0));
    v0[0] = a01.x, v0[1] = a01.y, v0[2] = c2 ? __ldg(p0 + 2) : 0.f;
    v1[0] = v1[1] = v1[2] = 0.f;
    if (r1) {
      const float2 b01 = __ldg(reinterpret_cast<const float2*>(p0 + a.Wi));
      v1[0] = b01.x, v1[1] = b01.y, v1[2] = c2 ? __ldg(p0 + a.Wi + 2) : 0.f;
    }
  };
  float nx0[3], nx1[3];
  load6(0, nx0, nx1);
  for (int ci = 0; ci < a.Cin; ++ci) {
    const float in0[3] = {nx0[0], nx0[1], nx0[2]}, in1[3] = {nx1[0], nx1[1], nx1[2]};
    if (ci + 1 < a.Cin) load6(ci + 1, nx0, nx1);
    const float* wc = sW + ci * 9 * COUT + cg * 8;
#pragma unroll
    for (int q = 0; q < 8; q += 4) {
      float4 w[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) w[k] = *reinterpret_cast<const float4*>(wc + k * COUT + q);
      // x pattern for output columns 4j + {0,1,2,3}: (kx=1,c=0) | (kx=2,c=0)+(kx=0,c=1) | (kx=1,c=1) | (kx=2,c=1)+(kx=0,c=2)
#define LWS_ROW(ACC, IN, KY)                                                                                            \
  {                                                                                                                     \
    const float4 k0 = w[(KY) * 3], k1 = w[(KY) * 3 + 1], k2 = w[(KY) * 3 + 2];                                           \
    const float2 k0a = make_float2(k0.x, k0.y), k0b = make_float2(k0.z, k0.w);                                           \
    const float2 k1a = make_float2(k1.x, k1.y), k1b = make_float2(k1.z, k1.w);                                           \
    const float2 k2a = make_float2(k2.x, k2.y), k2b = make_float2(k2.z, k2.w);                                           \
    const float2 i0 = make_float2(IN[0], IN[0]), i1 = make_float2(IN[1], IN[1]), i2 = make_float2(IN[2], IN[2]);         \
    ACC[0][q / 2] = __ffma2_rn(i0, k1a, ACC[0][q / 2]), ACC[0][q / 2 + 1] = __ffma2_rn(i0, k1b, ACC[0][q / 2 + 1]);       \
    ACC[1][q / 2] = __ffma2_rn(i0, k2a, __ffma2_rn(i1, k0a, ACC[1][q / 2]));                                             \
    ACC[1][q / 2 + 1] = __ffma2_rn(i0, k2b, __ffma2_rn(i1, k0b, ACC[1][q / 2 + 1]));                                     \
    ACC[2][q / 2] = __ffma2_rn(i1, k1a, ACC[2][q / 2]), ACC[2][q / 2 + 1] = __ffma2_rn(i1, k1b, ACC[2][q / 2 + 1]);       \
    ACC[3][q / 2] = __ffma2_rn(i1, k2a, __ffma2_rn(i2, k0a, ACC[3][q / 2]));                                             \
    ACC[3][q / 2 + 1] = __ffma2_rn(i1, k2b, __ffma2_rn(i2, k0b, ACC[3][q / 2 + 1]));                                     \
  }
      LWS_ROW(acc[0], in0, 1)  // even output row: ky = 1 on input row i
      LWS_ROW(acc[1], in0, 2)  // odd output row: ky = 2 on input row i ...
      LWS_ROW(acc[1], in1, 0)  // ... and ky = 0 on input row i + 1
#undef LWS_ROW
    }
  }
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int co = cg * 8 + q;
    const float bias = __ldg(a.bias + co);
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const long long o = ((long long)b * COUT + co) * ohw + (long long)(2 * i + r) * a.Wo + 4 * j;
      float v[4];
#pragma unroll
      for (int p = 0; p < 4; ++p) v[p] = (q & 1 ? acc[r][p][q / 2].y : acc[r][p][q / 2].x) + bias;
      if (a.res) {
        const float4 rv = *reinterpret_cast<const float4*>(a.res + o);
        v[0] += rv.x, v[1] += rv.y, v[2] += rv.z, v[3] += rv.w;
      }
      if (a.relu) {
#pragma unroll
        for (int p = 0; p < 4; ++p) v[p] = fmaxf(v[p], 0.f);
      }
      *reinterpret_cast<float4*>(a.out + o) = make_float4(v[0], v[1], v[2], v[3]);
    }
  }
}

// ---- fast path 2b: the same transposed conv with its input tile delivered by ONE TMA load (Wi % 4 == 0) -------------------------------
// fe_deconv_kernel is bound by the latency of its per-thread loads (ncu: 71 % of the stall samples on their scoreboard, issue slots
// 27 % used).  A block owns 4 input rows x 64 input columns (8 x 128 outputs) of 8 output channels; one tiled TMA load brings the
// 5 x 68 window of all Cin channels into shared memory (rows / columns past the map are the TMA's zero fill = the taps that do not
// exist), the threads only read shared memory.  Same arithmetic as fe_deconv_kernel.
template <int COUT>
__global__ void __launch_bounds__(128) fe_deconv_tile_kernel(const __grid_constant__ CUtensorMap map_in, const FeConvArgs a) {
  constexpr int TI = 4, TJ = 32, ROWS = TI + 1, COLS = 2 * TJ + 4, NG = COUT / 8;
  extern __shared__ __align__(128) uint8_t fe_smem[];
  float* sIn = reinterpret_cast<float*>(fe_smem);                      // [Cin][ROWS][COLS]
  float* sW = sIn + a.Cin * ROWS * COLS;                               // [Cin][9][COUT]
  uint64_t* bar = reinterpret_cast<uint64_t*>(sW + a.Cin * 9 * COUT);
  const int tid = threadIdx.x;
  const int b = blockIdx.z / NG, cg = blockIdx.z - b * NG;
  const int i0 = blockIdx.y * TI, jx0 = blockIdx.x * (2 * TJ);  // first input row / column of the tile
  const bool wbulk = (reinterpret_cast<uintptr_t>(a.w) & 15) == 0;  // weights as one bulk copy on the same barrier
  const uint32_t wbytes = (uint32_t)(a.Cin * 9 * COUT * 4);
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
    mbar_expect_tx(bar, (uint32_t)(a.Cin * ROWS * COLS * 4) + (wbulk ? wbytes : 0u));
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                     smem_u32(sIn)),
                 "l"(reinterpret_cast<uint64_t>(&map_in)), "r"(smem_u32(bar)), "r"(jx0), "r"(i0), "r"(b * a.Cin)
                 : "memory");
    if (wbulk)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sW)),
                   "l"(a.w), "r"(wbytes), "r"(smem_u32(bar))
                   : "memory");
  }
  if (!wbulk)
    for (int k = tid; k < a.Cin * 9 * COUT; k += 128) sW[k] = __ldg(a.w + k);
  __syncthreads();
  const int tj = tid % TJ, ti = tid / TJ;
  const int i = i0 + ti, j = blockIdx.x * TJ + tj;  // input row; input column pair (columns 2j, 2j+1)
  const long long ohw = (long long)a.Ho * a.Wo;
  if (a.res && i < a.Hi && 2 * j < a.Wi && (tj & 7) == 0) {
    // the skip tensor is read by the epilogue: request its lines (128 bytes = the outputs of 8 neighbouring threads) into L2 now
#pragma unroll
    for (int q = 0; q < 8; ++q)
#pragma unroll
      for (int r = 0; r < 2; ++r)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(a.res + ((long long)b * COUT + cg * 8 + q) * ohw + (long long)(2 * i + r) * a.Wo + 4 * j));
  }
  mbar_wait(bar, 0);
  if (i >= a.Hi || 2 * j >= a.Wi) return;
  float2 acc[2][4][4];  // float2 pairs of output channels: FFMA2 (fma.rn.f32x2) halves the issued FMA instructions
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[r][p][q] = make_float2(0.f, 0.f);

  const float* tp = sIn + ti * COLS + 2 * tj;
  for (int ci = 0; ci < a.Cin; ++ci) {
    const float* p0 = tp + ci * ROWS * COLS;
    const float2 a01 = *reinterpret_cast<const float2*>(p0), b01 = *reinterpret_cast<const float2*>(p0 + COLS);
    const float in0[3] = {a01.x, a01.y, p0[2]}, in1[3] = {b01.x, b01.y, p0[COLS + 2]};
    const float* wc = sW + ci * 9 * COUT + cg * 8;
#pragma unroll
    for (int q = 0; q < 8; q += 4) {
      float4 w[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) w[k] = *reinterpret_cast<const float4*>(wc + k * COUT + q);
      // x pattern for output columns 4j + {0,1,2,3}: (kx=1,c=0) | (kx=2,c=0)+(kx=0,c=1) | (kx=1,c=1) | (kx=2,c=1)+(kx=0,c=2)
#define LWS_ROW(ACC, IN, KY)                                                                                            \
  {                                                                                                                     \
    const float4 k0 = w[(KY) * 3], k1 = w[(KY) * 3 + 1], k2 = w[(KY) * 3 + 2];                                           \
    const float2 k0a = make_float2(k0.x, k0.y), k0b = make_float2(k0.z, k0.w);                                           \
    const float2 k1a = make_float2(k1.x, k1.y), k1b = make_float2(k1.z, k1.w);                                           \
    const float2 k2a = make_float2(k2.x, k2.y), k2b = make_float2(k2.z, k2.w);                                           \
    const float2 i0 = make_float2(IN[0], IN[0]), i1 = make_float2(IN[1], IN[1]), i2 = make_float2(IN[2], IN[2]);         \
    ACC[0][q / 2] = __ffma2_rn(i0, k1a, ACC[0][q / 2]), ACC[0][q / 2 + 1] = __ffma2_rn(i0, k1b, ACC[0][q / 2 + 1]);       \
    ACC[1][q / 2] = __ffma2_rn(i0, k2a, __ffma2_rn(i1, k0a, ACC[1][q / 2]));                                             \
    ACC[1][q / 2 + 1] = __ffma2_rn(i0, k2b, __ffma2_rn(i1, k0b, ACC[1][q / 2 + 1]));                                     \
    ACC[2][q / 2] = __ffma2_rn(i1, k1a, ACC[2][q / 2]), ACC[2][q / 2 + 1] = __ffma2_rn(i1, k1b, ACC[2][q / 2 + 1]);       \
    ACC[3][q / 2] = __ffma2_rn(i1, k2a, __ffma2_rn(i2, k0a, ACC[3][q / 2]));                                             \
    ACC[3][q / 2 + 1] = __ffma2_rn(i1, k2b, __ffma2_rn(i2, k0b, ACC[3][q / 2 + 1]));                                     \
  }
      LWS_ROW(acc[0], in0, 1)  // even output row: ky = 1 on input row i
      LWS_ROW(acc[1], in0, 2)  // odd output row: ky = 2 on input row i ...
      LWS_ROW(acc[1], in1, 0)  // ... and ky = 0 on input row i + 1
#undef LWS_ROW
    }
  }
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int co = cg * 8 + q;
    const float bias = __ldg(a.bias + co);
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const long long o = ((long long)b * COUT + co) * ohw + (long long)(2 * i + r) * a.Wo + 4 * j;
      float v[4];
#pragma unroll
      for (int p = 0; p < 4; ++p) v[p] = (q & 1 ? acc[r][p][q / 2].y : acc[r][p][q / 2].x) + bias;
      if (a.res) {
        const float4 rv = *reinterpret_cast<const float4*>(a.res + o);
        v[0] += rv.x, v[1] += rv.y, v[2] += rv.z, v[3] += rv.w;
      }
      if (a.relu) {
#pragma unroll
        for (int p = 0; p < 4; ++p) v[p] = fmaxf(v[p], 0.f);
      }
      *reinterpret_cast<float4*>(a.out + o) = make_float4(v[0], v[1], v[2], v[3]);
    }
  }
}

template <int COUT, int CT, int DIL, int STRIDE, int WIN = 12, bool ODD = false>
static int launch_pipe(const FeConvArgs& a, int B, size_t smem_w, cudaStream_t st) {
  if (!ODD && (a.Cin & 1)) return LWS_ERR_UNSUPPORTED;
  dim3 g(cdiv(cdiv(a.Wo, 4) * a.Ho, 128), B, COUT / CT);
  if ((a.Wi & 3) == 0 && (a.Wo & 3) == 0) fe_conv_pipe_kernel<COUT, CT, DIL, STRIDE, true, WIN, ODD><<<g, 128, smem_w, st>>>(a);
  else fe_conv_pipe_kernel<COUT, CT, DIL, STRIDE, false, WIN, ODD><<<g, 128, smem_w, st>>>(a);
  cudaError_t e = cudaPeekAtLastError();
  return e == cudaSuccess ? LWS_OK : (int)e;
}

static int launch_fe(FeConvArgs a, int cout, int B, cudaStream_t st) {
  const size_t smem_w = (size_t)a.Cin * 9 * cout * sizeof(float);
  const bool cin_even = (a.Cin & 1) == 0;
  const bool even = (a.Wi & 1) == 0 && (a.Wo & 1) == 0 && a.pad == a.dil && !a.transposed &&
                    (((uintptr_t)a.in | (uintptr_t)a.out | (uintptr_t)a.res) & 15) == 0;
  // input tile through TMA: 16-byte aligned rows, a tile that does not waste most of its width, B * Cin within a grid dimension
  const bool tma_ok = even && a.stride == 1 && a.Wo == a.Wi && (a.Wi & 3) == 0 && a.Wi >= 64 && opt(OPT_FE_TMA) &&
                      (long long)B * cout <= 65535;
  if (tma_ok) {
    const bool wide = a.Wi % 128 == 0 || a.Wi >= 512;  // 128-pixel tiles; 64-pixel tiles for the narrower maps (308 = 4.8 x 64)
    int rc = LWS_ERR_UNSUPPORTED;
    if (cout == 16 && a.dil == 1) rc = wide ? launch_tile<16, 8, 1, 32>(a, B, st) : launch_tile<16, 8, 1, 16>(a, B, st);
    else if (cout == 8 && a.dil == 1) rc = wide ? launch_tile<8, 8, 1, 32>(a, B, st) : launch_tile<8, 8, 1, 16>(a, B, st);
    else if (cout == 8 && a.dil == 2) rc = wide ? launch_tile<8, 8, 2, 32>(a, B, st) : launch_tile<8, 8, 2, 16>(a, B, st);
    else if (cout == 8 && a.dil == 4) rc = wide ? launch_tile<8, 8, 4, 32>(a, B, st) : launch_tile<8, 8, 4, 16>(a, B, st);
    else if (cout == 4 && a.dil == 2) rc = wide ? launch_tile<4, 4, 2, 32>(a, B, st) : launch_tile<4, 4, 2, 16>(a, B, st);
    if (rc != LWS_ERR_UNSUPPORTED) return rc;
  }
  if (even && cin_even && a.stride == 1 && a.Wo == a.Wi) {
    if (cout == 16 && a.dil == 1) return launch_pipe<16, 8, 1, 1>(a, B, smem_w, st);
    if (cout == 8 && a.dil == 1) return launch_pipe<8, 8, 1, 1>(a, B, smem_w, st);
    if (cout == 8 && a.dil == 2) return launch_pipe<8, 8, 2, 1>(a, B, smem_w, st);
    if (cout == 8 && a.dil == 4) return launch_pipe<8, 8, 4, 1>(a, B, smem_w, st);
    if (cout == 4 && a.dil == 2) return launch_pipe<4, 4, 2, 1>(a, B, smem_w, st);
  }
  if (even && cin_even && a.stride == 2 && a.dil == 1 && a.Wi == 2 * a.Wo && cout == 16) return launch_pipe<16, 8, 1, 2>(a, B, smem_w, st);
  if (even && a.stride == 2 && a.dil == 2 && a.Wi == 2 * a.Wo && cout == 4) return launch_pipe<4, 4, 2, 2, 16, true>(a, B, smem_w, st);
  if (a.transposed && a.stride == 2 && a.pad == 1 && (a.Wi & 1) == 0 && a.Ho == 2 * a.Hi && a.Wo == 2 * a.Wi && (cout == 8 || cout == 16)) {
    if ((a.Wi & 3) == 0 && a.Wi >= 32 && opt(OPT_FE_TMA) && (((uintptr_t)a.in | (uintptr_t)a.out | (uintptr_t)a.res) & 15) == 0 &&
        (long long)B * cout <= 65535) {  // input tile through TMA
      constexpr int ROWS = 5, COLS = 68;
      const size_t smem = ((size_t)a.Cin * ROWS * COLS + (size_t)a.Cin * 9 * cout) * sizeof(float) + 16;
      CUtensorMap map;
      const uint64_t dims[3] = {(uint64_t)a.Wi, (uint64_t)a.Hi, (uint64_t)B * a.Cin};
      const uint64_t strides[2] = {(uint64_t)a.Wi * 4, (uint64_t)a.Hi * a.Wi * 4};
      const uint32_t box[3] = {COLS, ROWS, (uint32_t)a.Cin};
      if (smem <= 100 * 1024 && a.Cin <= 256 && make_tensor_map_f32(&map, a.in, 3, dims, strides, box, false) == LWS_OK) {
        dim3 g(cdiv(a.Wi, 64), cdiv(a.Hi, 4), B * (cout / 8));
        if (cout == 8) {
          LWS_SET_SMEM_ONCE(fe_deconv_tile_kernel<8>, 100 * 1024);
          fe_deconv_tile_kernel<8><<<g, 128, smem, st>>>(map, a);
        } else {
          LWS_SET_SMEM_ONCE(fe_deconv_tile_kernel<16>, 100 * 1024);
          fe_deconv_tile_kernel<16><<<g, 128, smem, st>>>(map, a);
        }
        cudaError_t e = cudaPeekAtLastError();
        return e == cudaSuccess ? LWS_OK : (int)e;
      }
    }
    dim3 g(cdiv((a.Wi >> 1) * a.Hi * (cout / 8), 128), B);
    if (cout == 8) fe_deconv_kernel<8><<<g, 128, smem_w, st>>>(a);
    else fe_deconv_kernel<16><<<g, 128, smem_w, st>>>(a);
    cudaError_t e = cudaPeekAtLastError();
    return e == cudaSuccess ? LWS_OK : (int)e;
  }
  dim3 grid(cdiv(cdiv(a.Wo, 4) * a.Ho, 128), B);
  const size_t smem = (size_t)a.Cin * 9 * cout * sizeof(float);
  const bool split = (long long)grid.x * B < 4 * kNumSMs;  // too few blocks for the machine: 4 output channels per thread
  switch (cout) {
    case 4: fe_conv_kernel<4, 4><<<grid, 128, smem, st>>>(a); break;
    case 8:
      if (split) grid.z = 2, fe_conv_kernel<8, 4><<<grid, 128, smem, st>>>(a);
      else fe_conv_kernel<8, 8><<<grid, 128, smem, st>>>(a);
      break;
    case 16:
      if (split) grid.z = 4, fe_conv_kernel<16, 4><<<grid, 128, smem, st>>>(a);
      else fe_conv_kernel<16, 16><<<grid, 128, smem, st>>>(a);
      break;
    default: return LWS_ERR_UNSUPPORTED;
  }
  cudaError_t e = cudaPeekAtLastError();
  return e == cudaSuccess ? LWS_OK : (int)e;
}

// layer table of feature_extraction (reference models/submodules.py:113-188 and hourglass :35-109)
struct FeLayer {
  int cin, cout, stride, dil, transposed, relu, has_bn;
};
static const FeLayer kFe[12] = {
    {3, 4, 2, 2, 0, 1, 1},    // 0 dres0.0   pad = dil
    {4, 8, 1, 4, 0, 1, 1},    // 1 dres0.2
    {8, 4, 1, 2, 0, 1, 1},    // 2 dres1.0
    {4, 8, 1, 2, 0, 0, 1},    // 3 dres1.2   + out0
    {8, 16, 2, 1, 0, 1, 1},   // 4 dres2.conv1
    {16, 16, 1, 1, 0, 1, 1},  // 5 dres2.conv2 -> pre
    {16, 16, 2, 1, 0, 1, 1},  // 6 dres2.conv3
    {16, 16, 1, 1, 0, 1, 1},  // 7 dres2.conv4 -> feature 1/8
    {16, 16, 2, 1, 1, 1, 1},  // 8 dres2.conv5 (transposed) + pre, relu -> feature 1/4
    {16, 8, 2, 1, 1, 0, 1},   // 9 dres2.conv6 (transposed) + out1
    {8, 8, 1, 1, 0, 1, 1},    // 10 classif1.0
    {8, 8, 1, 1, 0, 0, 0},    // 11 classif1.2 -> feature 1/2
};
static size_t fe_w_off(int layer, bool bias) {
  size_t off = 0;
  for (int i = 0; i < 12; ++i) {
    if (i == layer && !bias) return off;
    off += (size_t)round_up(kFe[i].cin * 9 * kFe[i].cout, 4);
    if (i == layer && bias) return off;
    off += (size_t)round_up(kFe[i].cout, 4);
  }
  return off;
}

}  // namespace lws

extern "C" size_t lws_feature_extraction_packed_floats(void) { return lws::fe_w_off(12, false); }

extern "C" int lws_pack_feature_extraction_weights(const float* const* t, int n_tensors, float eps, float* packed) {
  using namespace lws;
  if (!t || !packed) return LWS_ERR_NULL_PTR;
  if (n_tensors != 12 + 11 * 4) return LWS_ERR_BAD_SHAPE;
  for (int i = 0; i < n_tensors; ++i)
    if (!t[i]) return LWS_ERR_NULL_PTR;
  memset(packed, 0, lws_feature_extraction_packed_floats() * sizeof(float));
  int ti = 0;
  for (int l = 0; l < 12; ++l) {
    const FeLayer& L = kFe[l];
    const float* w = t[ti++];
    const float *g = nullptr, *be = nullptr, *mu = nullptr, *var = nullptr;
    if (L.has_bn) g = t[ti++], be = t[ti++], mu = t[ti++], var = t[ti++];
    float* dw = packed + fe_w_off(l, false);
    float* db = packed + fe_w_off(l, true);
    for (int co = 0; co < L.cout; ++co) {
      const double s = L.has_bn ? (double)g[co] / sqrt((double)var[co] + (double)eps) : 1.0;
      db[co] = L.has_bn ? (float)((double)be[co] - (double)mu[co] * s) : 0.f;
      for (int ci = 0; ci < L.cin; ++ci)
        for (int k = 0; k < 9; ++k) {
          // Conv2D weight [Cout,Cin,3,3]; Conv2DTranspose weight [Cin,Cout,3,3]
          const size_t src = L.transposed ? ((size_t)ci * L.cout + co) * 9 + k : ((size_t)co * L.cin + ci) * 9 + k;
          dw[((size_t)ci * 9 + k) * L.cout + co] = (float)((double)w[src] * s);
        }
    }
  }
  return LWS_OK;
}

extern "C" size_t lws_feature_extraction_workspace_bytes(int B, int H, int W) {
  if (B <= 0 || H <= 0 || W <= 0 || (H & 7) || (W & 7)) return 0;
  const size_t p2 = (size_t)(H / 2) * (W / 2), p4 = (size_t)(H / 4) * (W / 4), p8 = (size_t)(H / 8) * (W / 8);
  return (size_t)B * (40 * p2 + 32 * p4 + 16 * p8) * sizeof(float);
}

extern "C" int lws_feature_extraction_f32(const float* img, const float* pk, float* f8, float* f4, float* f2, void* ws,
                                          size_t ws_bytes, int B, int H, int W, lws_stream_t stream) {
  using namespace lws;
  LWS_CHECK_PTR(img);
  LWS_CHECK_PTR(pk);
  LWS_CHECK_PTR(f8);
  LWS_CHECK_PTR(f4);
  LWS_CHECK_PTR(f2);
  LWS_CHECK_PTR(ws);
  if (B <= 0 || H <= 0 || W <= 0 || (H & 7) || (W & 7) || B > 65535) return LWS_ERR_BAD_SHAPE;
  if (ws_bytes < lws_feature_extraction_workspace_bytes(B, H, W)) return LWS_ERR_WORKSPACE_TOO_SMALL;
  if ((((uintptr_t)ws) | ((uintptr_t)pk) | ((uintptr_t)f8) | ((uintptr_t)f4) | ((uintptr_t)f2)) & 15) return LWS_ERR_BAD_ALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  const int H2 = H / 2, W2 = W / 2, H4 = H / 4, W4 = W / 4, H8 = H / 8, W8 = W / 8;
  const size_t p2 = (size_t)H2 * W2, p4 = (size_t)H4 * W4, p8 = (size_t)H8 * W8;
  float* base = (float*)ws;
  float* t0 = base;                      // [B,4,1/2]
  float* out0 = t0 + (size_t)B * 4 * p2;   // [B,8,1/2]
  float* t2 = out0 + (size_t)B * 8 * p2;   // [B,4,1/2]
  float* out1 = t2 + (size_t)B * 4 * p2;   // [B,8,1/2]
  float* c6 = out1 + (size_t)B * 8 * p2;   // [B,8,1/2]
  float* cl = c6 + (size_t)B * 8 * p2;     // [B,8,1/2]
  float* c1 = cl + (size_t)B * 8 * p2;     // [B,16,1/4]
  float* pre = c1 + (size_t)B * 16 * p4;   // [B,16,1/4]
  float* c3 = pre + (size_t)B * 16 * p4;   // [B,16,1/8]
  (void)p8;

  struct Step {
    int layer;
    const float* in;
    int Hi, Wi;
    float* out;
    int Ho, Wo;
    const float* res;
  };
  const Step steps[12] = {
      {0, img, H, W, t0, H2, W2, nullptr},   {1, t0, H2, W2, out0, H2, W2, nullptr}, {2, out0, H2, W2, t2, H2, W2, nullptr},
      {3, t2, H2, W2, out1, H2, W2, out0},   {4, out1, H2, W2, c1, H4, W4, nullptr}, {5, c1, H4, W4, pre, H4, W4, nullptr},
      {6, pre, H4, W4, c3, H8, W8, nullptr}, {7, c3, H8, W8, f8, H8, W8, nullptr},   {8, f8, H8, W8, f4, H4, W4, pre},
      {9, f4, H4, W4, c6, H2, W2, out1},     {10, c6, H2, W2, cl, H2, W2, nullptr},  {11, cl, H2, W2, f2, H2, W2, nullptr},
  };
  for (const Step& s : steps) {
    const FeLayer& L = kFe[s.layer];
    FeConvArgs a;
    a.in = s.in, a.w = pk + fe_w_off(s.layer, false), a.bias = pk + fe_w_off(s.layer, true), a.res = s.res, a.out = s.out;
    a.Cin = L.cin, a.Hi = s.Hi, a.Wi = s.Wi, a.Ho = s.Ho, a.Wo = s.Wo;
    a.stride = L.stride, a.dil = L.dil, a.pad = L.dil > 1 ? L.dil : 1, a.transposed = L.transposed, a.relu = L.relu;
    int rc = launch_fe(a, L.cout, B, st);
    if (rc) return rc;
  }
  return LWS_OK;
}
