// K4: fused softmax(-cost) + disparity regression  (reference models/models.py:142,151-152,167-179: softmax, expand,
//     multiply, reduce_sum = 4 passes over the volume; here one streaming pass).
// K5: rescale + half-pixel bilinear upsample (+ previous-stage skip)  (models/models.py:145-148,153-156).
// K2a: wflow = previous full-resolution disparity expressed at this scale (models/models.py:119-121).
// All three are pure streaming kernels: HBM-bound, every byte is touched once.
#include "lws_common.cuh"

namespace lws {

// One thread owns VEC horizontally adjacent pixels and walks the D planes in chunks of CH with a chunked online
// softmax (one rescale per chunk, so ~(1 + 1/CH) exp per element instead of 2 for the classic online form).  CH is matched to D
// (9 planes of a residual volume = one chunk, 24 planes of the stage-1 volume = two chunks of 12): on B200 the 16/clk/SM
// MUFU.EX2 rate is within 2x of what the HBM stream asks for, so masked-out lanes of a partial chunk are not free.
template <int VEC, int CH>
// resident warps, not registers, bound this stream (measured at 64 pairs: 1 -> 3 -> 4 blocks / SM = 59 % -> 79 % -> 85 % of HBM for D = 9)
__global__ void __launch_bounds__(256, CH == 9 ? 4 : 3)
    softmax_regression_kernel(const float* __restrict__ cost, float* __restrict__ low, int D, long long hw,
                              long long n_items_per_b, float start, float step) {
  const long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= n_items_per_b) return;
  const int b = blockIdx.y;
  const float* p = cost + (long long)b * D * hw + item * VEC;
  constexpr float kLog2e = 1.4426950408889634f;
  float m[VEC], s[VEC], ws[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) m[v] = -INFINITY, s[v] = 0.f, ws[v] = 0.f;

  for (int d0 = 0; d0 < D; d0 += CH) {
    float z[CH][VEC];
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      if (d0 + j < D) {
        if constexpr (VEC == 4) {
          float4 t = __ldcs(reinterpret_cast<const float4*>(p + (long long)(d0 + j) * hw));
          z[j][0] = -t.x, z[j][1] = -t.y, z[j][2] = -t.z, z[j][3] = -t.w;
        } else {
#pragma unroll
          for (int v = 0; v < VEC; ++v) z[j][v] = -__ldcs(p + (long long)(d0 + j) * hw + v);
        }
      } else {
#pragma unroll
        for (int v = 0; v < VEC; ++v) z[j][v] = -INFINITY;
      }
    }
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      float cm = z[0][v];
#pragma unroll
      for (int j = 1; j < CH; ++j) cm = fmaxf(cm, z[j][v]);
      const float nm = fmaxf(m[v], cm);
      const float r = exp2f((m[v] - nm) * kLog2e);  // exp2f(-inf) = 0 on the first chunk
      float cs = 0.f, cws = 0.f;
#pragma unroll
      for (int j = 0; j < CH; ++j) {
        const float e = exp2f((z[j][v] - nm) * kLog2e);
        cs += e;
        cws += e * (start + step * (float)(d0 + j));
      }
      s[v] = s[v] * r + cs;
      ws[v] = ws[v] * r + cws;
      m[v] = nm;
    }
  }
  float* o = low + (long long)b * hw + item * VEC;
  if constexpr (VEC == 4) {
    *reinterpret_cast<float4*>(o) = make_float4(ws[0] / s[0], ws[1] / s[1], ws[2] / s[2], ws[3] / s[3]);
  } else {
#pragma unroll
    for (int v = 0; v < VEC; ++v) o[v] = ws[v] / s[v];
  }
}

// Output-driven: one thread per 4 horizontally adjacent full-resolution pixels.
__global__ void __launch_bounds__(256)
    scale_upsample_add_kernel(const float* __restrict__ low, const float* __restrict__ prev, float* __restrict__ pred,
                              int h, int w, int H, int W, float fH, float rh, float sy, float sx) {
  const int W4 = (W + 3) >> 2;
  const int xq = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const int b = blockIdx.z;
  if (xq >= W4) return;
  const ResizeTap ty = resize_tap(y, sy, h);
  const float* l0 = low + ((long long)b * h + ty.i0) * w;
  const float* l1 = low + ((long long)b * h + ty.i1) * w;
  float out[4];
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int x = xq * 4 + p;
    const ResizeTap tx = resize_tap(min(x, W - 1), sx, w);
    // (low * float(H)) * fl32(1/h): two separately rounded multiplies, as the reference's two scale ops
    const float a00 = __fmul_rn(__fmul_rn(__ldg(l0 + tx.i0), fH), rh);
    const float a01 = __fmul_rn(__fmul_rn(__ldg(l0 + tx.i1), fH), rh);
    const float a10 = __fmul_rn(__fmul_rn(__ldg(l1 + tx.i0), fH), rh);
    const float a11 = __fmul_rn(__fmul_rn(__ldg(l1 + tx.i1), fH), rh);
    out[p] = ty.l0 * (tx.l0 * a00 + tx.l1 * a01) + ty.l1 * (tx.l0 * a10 + tx.l1 * a11);
  }
  const long long base = ((long long)b * H + y) * W + xq * 4;
  const bool vec = ((W & 3) == 0);
  if (vec) {
    float4 o = make_float4(out[0], out[1], out[2], out[3]);
    if (prev) {
      const float4 pv = __ldcs(reinterpret_cast<const float4*>(prev + base));
      o.x += pv.x, o.y += pv.y, o.z += pv.z, o.w += pv.w;
    }
    *reinterpret_cast<float4*>(pred + base) = o;
  } else {
#pragma unroll
    for (int p = 0; p < 4; ++p)
      if (xq * 4 + p < W) pred[base + p] = out[p] + (prev ? prev[base + p] : 0.f);
  }
}

// wflow[b,0,y,x] = (resize(pred_full)[y,x] * float(h)) * fl32(1/H)
__global__ void __launch_bounds__(256)
    disp_to_scale_kernel(const float* __restrict__ pred, float* __restrict__ wflow, int H, int W, int h, int w,
                         float fh, float rH, float sy, float sx) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const int b = blockIdx.z;
  if (x >= w) return;
  const ResizeTap ty = resize_tap(y, sy, H);
  const ResizeTap tx = resize_tap(x, sx, W);
  const float* p0 = pred + ((long long)b * H + ty.i0) * W;
  const float* p1 = pred + ((long long)b * H + ty.i1) * W;
  const float v = ty.l0 * (tx.l0 * __ldg(p0 + tx.i0) + tx.l1 * __ldg(p0 + tx.i1)) +
                  ty.l1 * (tx.l0 * __ldg(p1 + tx.i0) + tx.l1 * __ldg(p1 + tx.i1));
  wflow[((long long)b * h + y) * w + x] = __fmul_rn(__fmul_rn(v, fh), rH);
}

}  // namespace lws

extern "C" int lws_softmax_regression_f32(const float* cost, float* low, int B, int D, int H, int W, float start,
                                          float step, lws_stream_t stream) {
  using namespace lws;
  LWS_CHECK_PTR(cost);
  LWS_CHECK_PTR(low);
  if (B <= 0 || D <= 0 || H <= 0 || W <= 0 || B > 65535) return LWS_ERR_BAD_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  const long long hw = (long long)H * W;
  const bool vec = (hw % 4 == 0) && ((((uintptr_t)cost) | ((uintptr_t)low)) & 15) == 0;
  if (vec) {
    const long long n = hw / 4;
    // small volumes: narrower blocks so that every SM gets work
    const int threads = ((n + 255) / 256) * B >= 2 * kNumSMs ? 256 : 64;
    dim3 grid((unsigned)((n + threads - 1) / threads), B);
    if (D % 9 == 0) softmax_regression_kernel<4, 9><<<grid, threads, 0, st>>>(cost, low, D, hw, n, start, step);
    else if (D % 12 == 0) softmax_regression_kernel<4, 12><<<grid, threads, 0, st>>>(cost, low, D, hw, n, start, step);
    else softmax_regression_kernel<4, 8><<<grid, threads, 0, st>>>(cost, low, D, hw, n, start, step);
  } else {
    dim3 grid((unsigned)((hw + 255) / 256), B);
    softmax_regression_kernel<1, 8><<<grid, 256, 0, st>>>(cost, low, D, hw, hw, start, step);
  }
  LWS_RETURN_LAUNCH_STATUS();
}

extern "C" int lws_scale_upsample_add_f32(const float* low, const float* prev_or_null, float* pred, int B, int h,
                                          int w, int H, int W, lws_stream_t stream) {
  using namespace lws;
  LWS_CHECK_PTR(low);
  LWS_CHECK_PTR(pred);
  if (B <= 0 || h <= 0 || w <= 0 || H <= 0 || W <= 0 || H > 65535 || B > 65535) return LWS_ERR_BAD_SHAPE;
  if ((W & 3) == 0 && (((uintptr_t)pred | (uintptr_t)prev_or_null) & 15) != 0) return LWS_ERR_BAD_ALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  const int W4 = (W + 3) / 4;
  dim3 grid(cdiv(W4, 128), H, B);
  scale_upsample_add_kernel<<<grid, 128, 0, st>>>(low, prev_or_null, pred, h, w, H, W, (float)H,
                                                  (float)(1.0 / (double)h), (float)h / (float)H, (float)w / (float)W);
  LWS_RETURN_LAUNCH_STATUS();
}

extern "C" int lws_disp_to_scale_f32(const float* pred_full, float* wflow, int B, int H, int W, int h, int w,
                                     lws_stream_t stream) {
  using namespace lws;
  LWS_CHECK_PTR(pred_full);
  LWS_CHECK_PTR(wflow);
  if (B <= 0 || h <= 0 || w <= 0 || H <= 0 || W <= 0 || h > 65535 || B > 65535) return LWS_ERR_BAD_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(cdiv(w, 128), h, B);
  disp_to_scale_kernel<<<grid, 128, 0, st>>>(pred_full, wflow, H, W, h, w, (float)h, (float)(1.0 / (double)H),
                                             (float)H / (float)h, (float)W / (float)w);
  LWS_RETURN_LAUNCH_STATUS();
}
