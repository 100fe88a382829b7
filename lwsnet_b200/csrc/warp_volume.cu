// K2: bilinear warp + residual L1 cost volume (stages 2 and 3).
// Replaces LWSNet.warp (reference models/models.py:28-55) and LWSNet._build_volume_2d3 (models/models.py:78-104).
// The reference materialises a (2m-1)x replicated batch of L, R and disp plus a normalised sampling grid and calls
// grid_sampler on it (~25x the algorithmic bytes).  Here one thread owns one pixel, computes the 2m-1 tap positions
// with the reference's exact fp32 op sequence (lws_common.cuh: warp_coord), and because the shifts are integer
// steps the 2m-1 bilinear samples of a channel come from one contiguous (2m)-wide window of the right row, which is
// loaded once per channel and reused for all shifts.  A pixel whose floor() indices do not form that contiguous run
// (possible within 1 ulp of an integer coordinate) takes the per-tap gather path so indices stay bit-exact.
#include "lws_common.cuh"

namespace lws {

__device__ __forceinline__ float ld_row(const float* row, int x, int W) {
  return (x >= 0 && x < W) ? __ldg(row + x) : 0.f;
}

// ---- a3: stand-alone warp -------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    warp_bilinear_kernel(const float* __restrict__ x, const float* __restrict__ disp, float* __restrict__ out, int C,
                         int H, int W, WarpAxis ax, WarpAxis ay) {
  const int px = blockIdx.x * blockDim.x + threadIdx.x;
  const int py = blockIdx.y;
  const int n = blockIdx.z;
  if (px >= W) return;
  const long long hw = (long long)H * W;
  const float d = __ldg(disp + (long long)n * hw + (long long)py * W + px);
  const Tap tx = make_tap(warp_coord((float)px, d, ax), W);
  const Tap ty = make_tap(warp_coord_nodisp((float)py, ay), H);
  const bool r0 = ty.i0 >= 0 && ty.i0 < H;
  const bool r1 = ty.i0 + 1 >= 0 && ty.i0 + 1 < H;
  const float w00 = tx.w0 * ty.w0, w01 = tx.w1 * ty.w0, w10 = tx.w0 * ty.w1, w11 = tx.w1 * ty.w1;
  const float* src = x + (long long)n * C * hw;
  float* dst = out + (long long)n * C * hw + (long long)py * W + px;
  for (int c = 0; c < C; ++c) {
    const float* row0 = src + c * hw + (long long)ty.i0 * W;
    const float* row1 = row0 + W;
    const float v00 = r0 ? ld_row(row0, tx.i0, W) : 0.f;
    const float v01 = r0 ? ld_row(row0, tx.i0 + 1, W) : 0.f;
    const float v10 = r1 ? ld_row(row1, tx.i0, W) : 0.f;
    const float v11 = r1 ? ld_row(row1, tx.i0 + 1, W) : 0.f;
    dst[c * hw] = v00 * w00 + v01 * w01 + v10 * w10 + v11 * w11;
  }
}

// ---- test hook: export the tap indices / weights --------------------------------------------------------
__global__ void warp_taps_kernel(const float* __restrict__ disp, float shift, int32_t* __restrict__ x0,
                                 int32_t* __restrict__ y0, float* __restrict__ wx, float* __restrict__ wy, int H,
                                 int W, WarpAxis ax, WarpAxis ay) {
  const int px = blockIdx.x * blockDim.x + threadIdx.x;
  const int py = blockIdx.y;
  const int n = blockIdx.z;
  if (px >= W) return;
  const long long i = ((long long)n * H + py) * W + px;
  const float d = __fsub_rn(__ldg(disp + i), shift);
  const Tap tx = make_tap(warp_coord((float)px, d, ax), W);
  x0[i] = tx.i0;
  wx[2 * i] = tx.w0;
  wx[2 * i + 1] = tx.w1;
  if (n == 0 && px == 0) {
    const Tap ty = make_tap(warp_coord_nodisp((float)py, ay), H);
    y0[py] = ty.i0;
    wy[2 * py] = ty.w0;
    wy[2 * py + 1] = ty.w1;
  }
}

// ---- a4: warp + residual volume ---------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(128)
    warp_residual_volume_kernel(const float* __restrict__ L, const float* __restrict__ R,
                                const float* __restrict__ disp, float* __restrict__ cost, int C, int H, int W, int m,
                                float fstride, WarpAxis ax, WarpAxis ay) {
  const int px = blockIdx.x * blockDim.x + threadIdx.x;
  const int py = blockIdx.y;
  const int b = blockIdx.z;
  if (px >= W) return;
  const long long hw = (long long)H * W;
  const long long pix = (long long)py * W + px;
  const float d = __ldg(disp + (long long)b * hw + pix);

  int xi[K];
  float wx0[K], wx1[K];
  bool run = true;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const float shift = __fmul_rn((float)(k - (m - 1)), fstride);  // batch_shift * stride (models.py:90-92)
    const Tap t = make_tap(warp_coord((float)px, __fsub_rn(d, shift), ax), W);
    xi[k] = t.i0, wx0[k] = t.w0, wx1[k] = t.w1;
    run = run && (t.i0 == xi[0] + k);
  }
  const Tap ty = make_tap(warp_coord_nodisp((float)py, ay), H);
  // a row contributes nothing when it is out of bounds; skipping an exactly-zero weight is value-identical
  const bool r0 = ty.i0 >= 0 && ty.i0 < H && ty.w0 != 0.f;
  const bool r1 = ty.i0 + 1 >= 0 && ty.i0 + 1 < H && ty.w1 != 0.f;

  float acc[K];
#pragma unroll
  for (int k = 0; k < K; ++k) acc[k] = 0.f;

  const float* Lp = L + (long long)b * C * hw + pix;
  const float* Rb = R + (long long)b * C * hw;
  if (run) {
    const int xs = xi[0];
    for (int c = 0; c < C; ++c) {
      const float l = __ldg(Lp + c * hw);
      const float* row0 = Rb + c * hw + (long long)ty.i0 * W;
      float win[K + 1];
#pragma unroll
      for (int j = 0; j <= K; ++j) win[j] = 0.f;
      if (r0) {
#pragma unroll
        for (int j = 0; j <= K; ++j) win[j] = ty.w0 * ld_row(row0, xs + j, W);
      }
      if (r1) {
#pragma unroll
        for (int j = 0; j <= K; ++j) win[j] += ty.w1 * ld_row(row0 + W, xs + j, W);
      }
#pragma unroll
      for (int k = 0; k < K; ++k) acc[k] += fabsf(l - (win[k] * wx0[k] + win[k + 1] * wx1[k]));
    }
  } else {
    for (int c = 0; c < C; ++c) {
      const float l = __ldg(Lp + c * hw);
      const float* row0 = Rb + c * hw + (long long)ty.i0 * W;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        float a = 0.f, bb = 0.f;
        if (r0) a = ty.w0 * ld_row(row0, xi[k], W), bb = ty.w0 * ld_row(row0, xi[k] + 1, W);
        if (r1) a += ty.w1 * ld_row(row0 + W, xi[k], W), bb += ty.w1 * ld_row(row0 + W, xi[k] + 1, W);
        acc[k] += fabsf(l - (a * wx0[k] + bb * wx1[k]));
      }
    }
  }
  float* o = cost + (long long)b * K * hw + pix;
#pragma unroll
  for (int k = 0; k < K; ++k) o[k * hw] = acc[k];
}

// ---- a4, staged: one block per (pair, row) ----------------------------------------------------------------------
// The per-pixel window loads of the kernel above are gathers whose start depends on the pixel's own disparity; straight from
// global memory every warp-wide load touches up to 32 different 128-byte lines and the kernel is bound by L1 wavefronts, not by
// HBM.  Here the block first stages the (vertically blended) right-feature row of all C channels into shared memory as C/4
// planes of float4 = 4 channels of one pixel ([q][x][4], 2 zero pixels left / 3 right = the clamp range of make_tap), so a
// thread's 10-pixel window of 4 channels is 10 consecutive float4: 10*C/4 LDS.128 per pixel instead of 10*C (or 20*C) scattered
// LDG.32, and neighbouring pixels with neighbouring disparities read neighbouring 16-byte words (no bank conflicts; chaotic
// disparities cost ~2.4x in conflicts).  Global traffic is the algorithmic minimum: L, R (each row read once, twice for the rows
// whose fp32 coordinate round trip lands on y - eps) and disp read once, the volume written once.
// |l - (a*w0 + b*w1)| is evaluated as |fma(-b, w1, fma(-a, w0, l))|: 3 instructions per (shift, channel), 2 roundings.
template <int C, int K>
__global__ void __launch_bounds__(256)
    warp_residual_volume_row_kernel(const float* __restrict__ L, const float* __restrict__ R, const float* __restrict__ disp,
                                    float* __restrict__ cost, int H, int W, int m, float fstride, WarpAxis ax, WarpAxis ay) {
  extern __shared__ __align__(16) float4 s_row[];  // [C/4][W + 5], pixel -2 first
  constexpr int PADL = 2, PADR = 3, Q = C / 4;
  const int pitch = W + PADL + PADR;
  const int py = blockIdx.x;
  const int b = blockIdx.y;
  const long long hw = (long long)H * W;
  const Tap ty = make_tap(warp_coord_nodisp((float)py, ay), H);
  const bool r0 = ty.i0 >= 0 && ty.i0 < H && ty.w0 != 0.f;
  const bool r1 = ty.i0 + 1 >= 0 && ty.i0 + 1 < H && ty.w1 != 0.f;
  const float* Rb = R + (long long)b * C * hw + (long long)ty.i0 * W;

  for (int i = threadIdx.x; i < (PADL + PADR) * Q; i += blockDim.x) {
    const int q = i / (PADL + PADR), j = i - q * (PADL + PADR);
    s_row[q * pitch + (j < PADL ? j : W + j)] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int p = threadIdx.x; p < W; p += blockDim.x) {
    float v[C];
#pragma unroll
    for (int c = 0; c < C; ++c) v[c] = r0 ? ty.w0 * __ldg(Rb + c * hw + p) : 0.f;
    if (r1) {
#pragma unroll
      for (int c = 0; c < C; ++c) v[c] += ty.w1 * __ldg(Rb + c * hw + W + p);
    }
#pragma unroll
    for (int q = 0; q < Q; ++q) s_row[q * pitch + p + PADL] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
  }
  __syncthreads();

  for (int px = threadIdx.x; px < W; px += blockDim.x) {
    const long long pix = (long long)py * W + px;
    const float d = __ldg(disp + (long long)b * hw + pix);
    const float* Lp = L + (long long)b * C * hw + pix;
    float l[C];
#pragma unroll
    for (int c = 0; c < C; ++c) l[c] = __ldg(Lp + c * hw);

    int xi[K];
    float wx0[K], wx1[K];
    bool run = true;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const float shift = __fmul_rn((float)(k - (m - 1)), fstride);  // batch_shift * stride (models.py:90-92)
      const Tap t = make_tap(warp_coord((float)px, __fsub_rn(d, shift), ax), W);
      xi[k] = t.i0, wx0[k] = -t.w0, wx1[k] = -t.w1;
      run = run && (t.i0 == xi[0] + k);
    }
    float acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = 0.f;
    if (run) {
      const float4* win0 = s_row + xi[0] + PADL;
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        float4 win[K + 1];
#pragma unroll
        for (int j = 0; j <= K; ++j) win[j] = win0[q * pitch + j];
#pragma unroll
        for (int k = 0; k < K; ++k) {
          acc[k] += fabsf(fmaf(win[k + 1].x, wx1[k], fmaf(win[k].x, wx0[k], l[4 * q + 0])));
          acc[k] += fabsf(fmaf(win[k + 1].y, wx1[k], fmaf(win[k].y, wx0[k], l[4 * q + 1])));
          acc[k] += fabsf(fmaf(win[k + 1].z, wx1[k], fmaf(win[k].z, wx0[k], l[4 * q + 2])));
          acc[k] += fabsf(fmaf(win[k + 1].w, wx1[k], fmaf(win[k].w, wx0[k], l[4 * q + 3])));
        }
      }
    } else {  // taps within 1 ulp of an integer coordinate: per-shift gather keeps the indices bit-exact
#pragma unroll
      for (int q = 0; q < Q; ++q) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const float4 a = s_row[q * pitch + xi[k] + PADL];
          const float4 bb = s_row[q * pitch + xi[k] + 1 + PADL];
          acc[k] += fabsf(fmaf(bb.x, wx1[k], fmaf(a.x, wx0[k], l[4 * q + 0])));
          acc[k] += fabsf(fmaf(bb.y, wx1[k], fmaf(a.y, wx0[k], l[4 * q + 1])));
          acc[k] += fabsf(fmaf(bb.z, wx1[k], fmaf(a.z, wx0[k], l[4 * q + 2])));
          acc[k] += fabsf(fmaf(bb.w, wx1[k], fmaf(a.w, wx0[k], l[4 * q + 3])));
        }
      }
    }
    float* o = cost + (long long)b * K * hw + pix;
#pragma unroll
    for (int k = 0; k < K; ++k) __stcs(o + k * hw, acc[k]);
  }
}

// any m: one thread per (pixel, shift)
__global__ void __launch_bounds__(256)
    warp_residual_volume_generic_kernel(const float* __restrict__ L, const float* __restrict__ R,
                                        const float* __restrict__ disp, float* __restrict__ cost, int C, int H, int W,
                                        int m, float fstride, WarpAxis ax, WarpAxis ay) {
  const int px = blockIdx.x * blockDim.x + threadIdx.x;
  const int py = blockIdx.y;
  const int K = 2 * m - 1;
  const int b = blockIdx.z / K, k = blockIdx.z % K;
  if (px >= W) return;
  const long long hw = (long long)H * W;
  const long long pix = (long long)py * W + px;
  const float d = __ldg(disp + (long long)b * hw + pix);
  const float shift = __fmul_rn((float)(k - (m - 1)), fstride);
  const Tap tx = make_tap(warp_coord((float)px, __fsub_rn(d, shift), ax), W);
  const Tap ty = make_tap(warp_coord_nodisp((float)py, ay), H);
  const bool r0 = ty.i0 >= 0 && ty.i0 < H && ty.w0 != 0.f;
  const bool r1 = ty.i0 + 1 >= 0 && ty.i0 + 1 < H && ty.w1 != 0.f;
  float acc = 0.f;
  for (int c = 0; c < C; ++c) {
    const float l = __ldg(L + ((long long)b * C + c) * hw + pix);
    const float* row0 = R + ((long long)b * C + c) * hw + (long long)ty.i0 * W;
    float a = 0.f, bb = 0.f;
    if (r0) a = ty.w0 * ld_row(row0, tx.i0, W), bb = ty.w0 * ld_row(row0, tx.i0 + 1, W);
    if (r1) a += ty.w1 * ld_row(row0 + W, tx.i0, W), bb += ty.w1 * ld_row(row0 + W, tx.i0 + 1, W);
    acc += fabsf(l - (a * tx.w0 + bb * tx.w1));
  }
  cost[((long long)b * K + k) * hw + pix] = acc;
}

}  // namespace lws

extern "C" int lws_warp_bilinear_f32(const float* x, const float* disp, float* out, int N, int C, int H, int W,
                                     lws_stream_t stream) {
  using namespace lws;
  LWS_CHECK_PTR(x);
  LWS_CHECK_PTR(disp);
  LWS_CHECK_PTR(out);
  if (N <= 0 || C <= 0 || H <= 0 || W <= 0 || H > 65535 || N > 65535) return LWS_ERR_BAD_SHAPE;
  dim3 grid(cdiv(W, 128), H, N);
  warp_bilinear_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(x, disp, out, C, H, W, make_warp_axis(W),
                                                               make_warp_axis(H));
  LWS_RETURN_LAUNCH_STATUS();
}

extern "C" int lws_warp_taps_f32(const float* disp, float shift, int32_t* x0, int32_t* y0, float* wx, float* wy,
                                 int N, int H, int W, lws_stream_t stream) {
  using namespace lws;
  LWS_CHECK_PTR(disp);
  LWS_CHECK_PTR(x0);
  LWS_CHECK_PTR(y0);
  LWS_CHECK_PTR(wx);
  LWS_CHECK_PTR(wy);
  if (N <= 0 || H <= 0 || W <= 0 || H > 65535 || N > 65535) return LWS_ERR_BAD_SHAPE;
  dim3 grid(cdiv(W, 128), H, N);
  warp_taps_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(disp, shift, x0, y0, wx, wy, H, W, make_warp_axis(W),
                                                           make_warp_axis(H));
  LWS_RETURN_LAUNCH_STATUS();
}

extern "C" int lws_warp_residual_volume_l1_f32(const float* L, const float* R, const float* disp, float* cost, int B,
                                               int C, int H, int W, int m, int stride, lws_stream_t stream) {
  using namespace lws;
  LWS_CHECK_PTR(L);
  LWS_CHECK_PTR(R);
  LWS_CHECK_PTR(disp);
  LWS_CHECK_PTR(cost);
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || m <= 0 || stride <= 0 || H > 65535) return LWS_ERR_BAD_SHAPE;
  if ((long long)B * (2 * m - 1) > 65535) return LWS_ERR_BAD_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  const WarpAxis ax = make_warp_axis(W), ay = make_warp_axis(H);
  const size_t row_smem = (size_t)(W + 5) * C * sizeof(float);
  if (m == 5 && (C == 8 || C == 16) && row_smem <= 96 * 1024 && B <= 65535) {
    // threads: the row in as few equal passes as possible (W = 616 -> 3 x 224, W = 308 -> 2 x 160)
    const int passes = cdiv(W, 256);
    const int threads = round_up(cdiv(W, passes), 32);
    dim3 grid(H, B);
    if (C == 8) {
      if (row_smem > 48 * 1024) LWS_SET_SMEM_ONCE((warp_residual_volume_row_kernel<8, 9>), 96 * 1024);
      warp_residual_volume_row_kernel<8, 9><<<grid, threads, row_smem, st>>>(L, R, disp, cost, H, W, m, (float)stride, ax, ay);
    } else {
      if (row_smem > 48 * 1024) LWS_SET_SMEM_ONCE((warp_residual_volume_row_kernel<16, 9>), 96 * 1024);
      warp_residual_volume_row_kernel<16, 9><<<grid, threads, row_smem, st>>>(L, R, disp, cost, H, W, m, (float)stride, ax, ay);
    }
  } else if (m == 5) {
    dim3 grid(cdiv(W, 128), H, B);
    warp_residual_volume_kernel<9><<<grid, 128, 0, st>>>(L, R, disp, cost, C, H, W, m, (float)stride, ax, ay);
  } else {
    dim3 grid(cdiv(W, 256), H, B * (2 * m - 1));
    warp_residual_volume_generic_kernel<<<grid, 256, 0, st>>>(L, R, disp, cost, C, H, W, m, (float)stride, ax, ay);
  }
  LWS_RETURN_LAUNCH_STATUS();
}
