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
      float zc[CH];
#pragma unroll
      for (int j = 0; j < CH; ++j) zc[j] = z[j][v];
      softmax_chunk_update<CH>(zc, d0, start, step, m[v], s[v], ws[v]);
    }
  }
  float* o = low + (long long)b * hw + item * VEC;
  if constexpr (VEC == 4) {
    *reinterpret_cast<float4*>(o) = make_float4(__fdiv_rn(ws[0], s[0]), __fdiv_rn(ws[1], s[1]), __fdiv_rn(ws[2], s[2]), __fdiv_rn(ws[3], s[3]));
  } else {
#pragma unroll
    for (int v = 0; v < VEC; ++v) o[v] = __fdiv_rn(ws[v], s[v]);
  }
}

// Output-driven: a thread owns 4 horizontally adjacent full-resolution pixels in K5_ROWS consecutive rows: the four x taps (index
// pair + weights: the expensive part, fp32 replay of the reference's half-pixel formula incl. a float->int conversion each) are
// computed once per thread instead of once per pixel, the previous-stage rows are requested up front, and the low-resolution taps
// of neighbouring rows are L1 hits.
constexpr int K5_ROWS = 8;
__global__ void __launch_bounds__(128)
    scale_upsample_add_kernel(const float* __restrict__ low, const float* __restrict__ prev, float* __restrict__ pred,
                              int h, int w, int H, int W, float fH, float rh, float sy, float sx) {
  const int W4 = (W + 3) >> 2;
  const int xq = blockIdx.x * blockDim.x + threadIdx.x;
  const int y0 = blockIdx.y * K5_ROWS;
  const int b = blockIdx.z;
  if (xq >= W4) return;
  const bool vec = ((W & 3) == 0);
  ResizeTap tx[4];
#pragma unroll
  for (int p = 0; p < 4; ++p) tx[p] = resize_tap(min(xq * 4 + p, W - 1), sx, w);
  const long long base0 = ((long long)b * H + y0) * W + xq * 4;
  float4 pv[K5_ROWS];
  if (prev && vec) {
#pragma unroll
    for (int r = 0; r < K5_ROWS; ++r)
      pv[r] = y0 + r < H ? __ldcs(reinterpret_cast<const float4*>(prev + base0 + (long long)r * W)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int r = 0; r < K5_ROWS; ++r) {
    const int y = y0 + r;
    if (y >= H) break;
    const ResizeTap ty = resize_tap(y, sy, h);
    const float* l0 = low + ((long long)b * h + ty.i0) * w;
    const float* l1 = low + ((long long)b * h + ty.i1) * w;
    float out[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      // (low * float(H)) * fl32(1/h): two separately rounded multiplies, as the reference's two scale ops
      const float a00 = __fmul_rn(__fmul_rn(__ldg(l0 + tx[p].i0), fH), rh);
      const float a01 = __fmul_rn(__fmul_rn(__ldg(l0 + tx[p].i1), fH), rh);
      const float a10 = __fmul_rn(__fmul_rn(__ldg(l1 + tx[p].i0), fH), rh);
      const float a11 = __fmul_rn(__fmul_rn(__ldg(l1 + tx[p].i1), fH), rh);
      out[p] = bilinear_blend(a00, a01, a10, a11, tx[p], ty);
    }
    const long long base = base0 + (long long)r * W;
    if (vec) {
      float4 o = make_float4(out[0], out[1], out[2], out[3]);
      if (prev) o.x = __fadd_rn(o.x, pv[r].x), o.y = __fadd_rn(o.y, pv[r].y), o.z = __fadd_rn(o.z, pv[r].z), o.w = __fadd_rn(o.w, pv[r].w);
      *reinterpret_cast<float4*>(pred + base) = o;
    } else {
#pragma unroll
      for (int p = 0; p < 4; ++p)
        if (xq * 4 + p < W) pred[base + p] = prev ? __fadd_rn(out[p], prev[base + p]) : out[p];
    }
  }
}

// Integer scales S = H/h = W/w in {2, 4, 8} (the three stages of the network): 1/S is exact in fp32, so away from the clamped image
// borders the tap pattern is periodic -- the four pixels of a thread use low-resolution columns c, c+1 (, c+2, c+3) at fixed offsets
// and the K5_ROWS rows of a block use rows r, .., r + K5_ROWS/S + 1 at fixed offsets.  The fast path loads that window once
// ((K5_ROWS/S + 2) x NC scalars instead of 16 per output row), scales every value once and shares the horizontal blends between the
// output rows of a low-resolution row pair: ~300 instead of ~1100 instructions per thread.  Every thread CHECKS the pattern against
// the generic resize_tap() results and falls back to the generic loop where it does not hold (borders), and the blend is the same
// sequence of rounded operations: bit-identical to scale_upsample_add_kernel.
template <int S>
__global__ void __launch_bounds__(128)
    scale_upsample_add_int_kernel(const float* __restrict__ low, const float* __restrict__ prev, float* __restrict__ pred, int h, int w,
                                  int H, int W, float fH, float rh, float sy, float sx) {
  constexpr int NR = K5_ROWS / S + 2;           // low-resolution rows under K5_ROWS output rows
  constexpr int NC = S == 2 ? 4 : (S == 4 ? 3 : 2);  // low-resolution columns under 4 output pixels
  const int W4 = W >> 2;  // host: W % 4 == 0
  const int xq = blockIdx.x * blockDim.x + threadIdx.x;
  const int y0 = blockIdx.y * K5_ROWS;
  const int b = blockIdx.z;
  if (xq >= W4) return;
  ResizeTap tx[4];
#pragma unroll
  for (int p = 0; p < 4; ++p) tx[p] = resize_tap(xq * 4 + p, sx, w);
  // With scale 1/S every operation of resize_tap() is exact, so src = (2 dst + 1 - S) / 2S exactly: unclamped from the second thread
  // column / block row on, tap index = floor(src), weight l1 = ((2 dst + 1 - S) mod 2S) / 2S.  For the rows dst = y0 + r with
  // y0 % 8 == 0, so index offsets AND weights are compile-time constants of r.
  const int c0 = (8 * xq + 1 - S) / (2 * S), r0 = (2 * y0 + 1 - S) / (2 * S);
  bool fast = xq >= 1 && y0 >= K5_ROWS && y0 + K5_ROWS <= H && c0 + NC - 1 < w && r0 + NR - 1 < h;
#pragma unroll
  for (int p = 0; p < 4; ++p) {  // cheap cross-check of the closed form against the generic tap (always true)
    const int off = S == 2 ? (p + 1) / 2 : (S == 4 ? p / 2 : 0);
    fast = fast && tx[p].i0 == c0 + off && tx[p].i1 == tx[p].i0 + 1;
  }
  const long long base0 = ((long long)b * H + y0) * W + xq * 4;
  if (fast) {
    // the low-resolution window first (needed first), then the previous-stage rows, then the arithmetic
    float cw[NR][NC];
    const float* lp = low + ((long long)b * h + r0) * w + c0;
#pragma unroll
    for (int i = 0; i < NR; ++i)
#pragma unroll
      for (int j = 0; j < NC; ++j) cw[i][j] = __ldg(lp + i * w + j);
    float4 pv[K5_ROWS];
    if (prev) {
#pragma unroll
      for (int r = 0; r < K5_ROWS; ++r) pv[r] = __ldcs(reinterpret_cast<const float4*>(prev + base0 + (long long)r * W));
    }
    // horizontal blends of the NR low-resolution rows for the 4 pixels
    float hb[NR][4];
#pragma unroll
    for (int i = 0; i < NR; ++i) {
      float c[NC];
#pragma unroll
      for (int j = 0; j < NC; ++j) c[j] = __fmul_rn(__fmul_rn(cw[i][j], fH), rh);
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const int off = S == 2 ? (p + 1) / 2 : (S == 4 ? p / 2 : 0);
        hb[i][p] = __fmaf_rn(tx[p].l1, c[off + 1], __fmul_rn(tx[p].l0, c[off]));
      }
    }
#pragma unroll
    for (int r = 0; r < K5_ROWS; ++r) {
      const int off = (2 * r + S) / (2 * S);  // floor((r + 0.5) / S + 0.5): rows r0, r0 + 1, .. in steps of S starting half a period in
      const float yl1 = (float)((2 * r + 1 + S) % (2 * S)) * (0.5f / S), yl0 = 1.0f - yl1;  // exact: multiples of 1/16
      float4 o;
      o.x = __fmaf_rn(yl1, hb[off + 1][0], __fmul_rn(yl0, hb[off][0]));
      o.y = __fmaf_rn(yl1, hb[off + 1][1], __fmul_rn(yl0, hb[off][1]));
      o.z = __fmaf_rn(yl1, hb[off + 1][2], __fmul_rn(yl0, hb[off][2]));
      o.w = __fmaf_rn(yl1, hb[off + 1][3], __fmul_rn(yl0, hb[off][3]));
      if (prev) o.x = __fadd_rn(o.x, pv[r].x), o.y = __fadd_rn(o.y, pv[r].y), o.z = __fadd_rn(o.z, pv[r].z), o.w = __fadd_rn(o.w, pv[r].w);
      *reinterpret_cast<float4*>(pred + base0 + (long long)r * W) = o;
    }
    return;
  }
  // generic path (image borders): as scale_upsample_add_kernel
  for (int r = 0; r < K5_ROWS; ++r) {
    const int y = y0 + r;
    if (y >= H) break;
    const ResizeTap ty = resize_tap(y, sy, h);
    const float* l0 = low + ((long long)b * h + ty.i0) * w;
    const float* l1 = low + ((long long)b * h + ty.i1) * w;
    float out[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const float a00 = __fmul_rn(__fmul_rn(__ldg(l0 + tx[p].i0), fH), rh);
      const float a01 = __fmul_rn(__fmul_rn(__ldg(l0 + tx[p].i1), fH), rh);
      const float a10 = __fmul_rn(__fmul_rn(__ldg(l1 + tx[p].i0), fH), rh);
      const float a11 = __fmul_rn(__fmul_rn(__ldg(l1 + tx[p].i1), fH), rh);
      out[p] = bilinear_blend(a00, a01, a10, a11, tx[p], ty);
    }
    float4 o = make_float4(out[0], out[1], out[2], out[3]);
    if (prev) {
      const float4 q = __ldcs(reinterpret_cast<const float4*>(prev + base0 + (long long)r * W));
      o.x = __fadd_rn(o.x, q.x), o.y = __fadd_rn(o.y, q.y), o.z = __fadd_rn(o.z, q.z), o.w = __fadd_rn(o.w, q.w);
    }
    *reinterpret_cast<float4*>(pred + base0 + (long long)r * W) = o;
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
  const float v = bilinear_blend(__ldg(p0 + tx.i0), __ldg(p0 + tx.i1), __ldg(p1 + tx.i0), __ldg(p1 + tx.i1), tx, ty);
  wflow[((long long)b * h + y) * w + x] = __fmul_rn(__fmul_rn(v, fh), rH);
}

// ---- fused tail of a stage: K4 + K5 (+ K2a of the NEXT stage) in one pass --------------------------------------------------------
// pred = upsample((softmax-regression(cost) * H) * (1/h)) (+ prev);  wflow_next = (resize(pred, hn, wn) * hn) * (1/H)
// (reference models/models.py:142-156 and, for the next iteration of the stage loop, :119-121).  The filtered volume is streamed
// once, the low-resolution disparity never exists in HBM, and the full-resolution prediction is not read back for the next wflow.
// A block owns a 32 x 256 full-resolution tile (scale s = H/h = W/w: (32/s + 2) x (256/s + 2) low-resolution pixels incl. the
// one-pixel halo the half-pixel bilinear taps reach): phase 1 regresses those pixels into shared memory with the same chunked online
// softmax as K4, phase 2 writes the tile with 128-bit stores, phase 3 decimates it to the next stage's wflow.  The arithmetic is
// that of the three stand-alone kernels (shared device functions with explicit roundings): the results are bit-identical.
constexpr int RT_TH = 32, RT_TW = 256, RT_THREADS = 256;

template <int CH>
__global__ void __launch_bounds__(RT_THREADS)
    regression_tail_kernel(const float* __restrict__ cost, const float* __restrict__ prev, float* __restrict__ pred,
                           float* __restrict__ wflow, int D, int h, int w, int H, int W, int hn, int wn, int s, float start, float step,
                           float fH, float rh, float fhn, float rH) {
  extern __shared__ float sLow[];  // [(RT_TH/s + 2)][(RT_TW/s + 2)] low-resolution disparities, already scaled by H/h
  const int lth = RT_TH / s + 2, ltw = RT_TW / s + 2;
  const int b = blockIdx.z;
  const int Y0 = blockIdx.y * RT_TH, X0 = blockIdx.x * RT_TW;
  const int ly0 = Y0 / s - 1, lx0 = X0 / s - 1;  // low-resolution origin of the shared tile
  const long long hw = (long long)h * w;
  const float* cb = cost + (long long)b * D * hw;
  // ---- phase 1: softmax regression of the tile's low-resolution pixels ----------------------------------------------------------
  for (int i = threadIdx.x; i < lth * ltw; i += RT_THREADS) {
    const int ly = ly0 + i / ltw, lx = lx0 + i % ltw;
    float v = 0.f;
    if (ly >= 0 && ly < h && lx >= 0 && lx < w) {
      const float* p = cb + (long long)ly * w + lx;
      float m = -INFINITY, sum = 0.f, ws = 0.f;
      for (int d0 = 0; d0 < D; d0 += CH) {
        float z[CH];
#pragma unroll
        for (int j = 0; j < CH; ++j) z[j] = d0 + j < D ? -__ldg(p + (long long)(d0 + j) * hw) : -INFINITY;
        softmax_chunk_update<CH>(z, d0, start, step, m, sum, ws);
      }
      v = __fmul_rn(__fmul_rn(__fdiv_rn(ws, sum), fH), rh);  // (low * float(H)) * fl32(1/h), as K5
    }
    sLow[i] = v;
  }
  __syncthreads();
  // ---- phase 2: upsample (+ prev) -> pred, 4 pixels per thread and row -------------------------------------------------------------
  const float sy = (float)h / (float)H, sx = (float)w / (float)W;
  auto pred_at = [&](int y, int x) -> float {  // the prediction at full-resolution pixel (y, x) inside this tile
    const ResizeTap ty = resize_tap(y, sy, h), tx = resize_tap(x, sx, w);
    const float* r0 = sLow + (ty.i0 - ly0) * ltw - lx0;
    const float* r1 = sLow + (ty.i1 - ly0) * ltw - lx0;
    const float up = bilinear_blend(r0[tx.i0], r0[tx.i1], r1[tx.i0], r1[tx.i1], tx, ty);
    return prev ? __fadd_rn(up, __ldg(prev + ((long long)b * H + y) * W + x)) : up;
  };
  const bool vec = (W & 3) == 0;
  for (int i = threadIdx.x; i < RT_TH * (RT_TW / 4); i += RT_THREADS) {
    const int y = Y0 + i / (RT_TW / 4), x = X0 + (i % (RT_TW / 4)) * 4;
    if (y >= H || x >= W) continue;
    const ResizeTap ty = resize_tap(y, sy, h);
    const float* r0 = sLow + (ty.i0 - ly0) * ltw - lx0;
    const float* r1 = sLow + (ty.i1 - ly0) * ltw - lx0;
    float out[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const ResizeTap tx = resize_tap(min(x + q, W - 1), sx, w);
      out[q] = bilinear_blend(r0[tx.i0], r0[tx.i1], r1[tx.i0], r1[tx.i1], tx, ty);
    }
    const long long base = ((long long)b * H + y) * W + x;
    if (vec) {
      float4 o = make_float4(out[0], out[1], out[2], out[3]);
      if (prev) {
        const float4 pv = __ldg(reinterpret_cast<const float4*>(prev + base));
        o.x = __fadd_rn(o.x, pv.x), o.y = __fadd_rn(o.y, pv.y), o.z = __fadd_rn(o.z, pv.z), o.w = __fadd_rn(o.w, pv.w);
      }
      *reinterpret_cast<float4*>(pred + base) = o;
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (x + q < W) pred[base + q] = prev ? __fadd_rn(out[q], prev[base + q]) : out[q];
    }
  }
  // ---- phase 3: the next stage's wflow from this tile's predictions (its taps never leave the tile: host-checked) -------------------
  if (wflow) {
    const int fy = H / hn, fx = W / wn;
    const float syn = (float)H / (float)hn, sxn = (float)W / (float)wn;
    const int nty = RT_TH / fy, ntx = RT_TW / fx;
    for (int i = threadIdx.x; i < nty * ntx; i += RT_THREADS) {
      const int yn = Y0 / fy + i / ntx, xn = X0 / fx + i % ntx;
      if (yn >= hn || xn >= wn) continue;
      const ResizeTap ty = resize_tap(yn, syn, H), tx = resize_tap(xn, sxn, W);
      const float v = bilinear_blend(pred_at(ty.i0, tx.i0), pred_at(ty.i0, tx.i1), pred_at(ty.i1, tx.i0), pred_at(ty.i1, tx.i1), tx, ty);
      wflow[((long long)b * hn + yn) * wn + xn] = __fmul_rn(__fmul_rn(v, fhn), rH);
    }
  }
}

}  // namespace lws

// 0: the fused kernel applies (integer scale H/h == W/w dividing the 32 x 256 tile; next-stage decimation by 1 or an even factor
// that divides the tile); otherwise LWS_ERR_UNSUPPORTED (the caller runs the three stand-alone kernels)
extern "C" int lws_regression_tail_supported(int h, int w, int H, int W, int hn, int wn) {
  using namespace lws;
  if (h <= 0 || w <= 0 || H <= 0 || W <= 0) return LWS_ERR_BAD_SHAPE;
  if (H % h || W % w || H / h != W / w) return LWS_ERR_UNSUPPORTED;
  const int s = H / h;
  if (s > 8 || RT_TH % s || RT_TW % s) return LWS_ERR_UNSUPPORTED;
  if (hn > 0 || wn > 0) {
    if (hn <= 0 || wn <= 0 || H % hn || W % wn) return LWS_ERR_UNSUPPORTED;
    const int fy = H / hn, fx = W / wn;
    if ((fy != 1 && (fy & 1)) || (fx != 1 && (fx & 1)) || RT_TH % fy || RT_TW % fx) return LWS_ERR_UNSUPPORTED;
  }
  return LWS_OK;
}

extern "C" int lws_regression_tail_f32(const float* cost, const float* prev_or_null, float* pred, float* wflow_next_or_null, int B,
                                       int D, int h, int w, int H, int W, int hn, int wn, float start, float step,
                                       lws_stream_t stream) {
  using namespace lws;
  LWS_CHECK_PTR(cost);
  LWS_CHECK_PTR(pred);
  if (B <= 0 || D <= 0 || B > 65535) return LWS_ERR_BAD_SHAPE;
  if (!wflow_next_or_null) hn = wn = 0;
  const int rc = lws_regression_tail_supported(h, w, H, W, hn, wn);
  if (rc) return rc;
  if ((W & 3) == 0 && (((uintptr_t)pred | (uintptr_t)prev_or_null) & 15) != 0) return LWS_ERR_BAD_ALIGN;
  const int s = H / h;
  dim3 grid(cdiv(W, RT_TW), cdiv(H, RT_TH), B);
  if (grid.y > 65535) return LWS_ERR_BAD_SHAPE;
  const size_t smem = (size_t)(RT_TH / s + 2) * (RT_TW / s + 2) * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  const float fH = (float)H, rh = (float)(1.0 / (double)h), fhn = (float)hn, rH = (float)(1.0 / (double)H);
  if (D % 9 == 0)
    regression_tail_kernel<9><<<grid, RT_THREADS, smem, st>>>(cost, prev_or_null, pred, wflow_next_or_null, D, h, w, H, W, hn, wn, s,
                                                              start, step, fH, rh, fhn, rH);
  else if (D % 12 == 0)
    regression_tail_kernel<12><<<grid, RT_THREADS, smem, st>>>(cost, prev_or_null, pred, wflow_next_or_null, D, h, w, H, W, hn, wn, s,
                                                               start, step, fH, rh, fhn, rH);
  else
    regression_tail_kernel<8><<<grid, RT_THREADS, smem, st>>>(cost, prev_or_null, pred, wflow_next_or_null, D, h, w, H, W, hn, wn, s,
                                                              start, step, fH, rh, fhn, rH);
  LWS_RETURN_LAUNCH_STATUS();
}

extern "C" int lws_softmax_regression_f32(const float* cost, float* low, int B, int D, int H, int W, float start,
                                          float step, lws_stream_t stream) {
  using namespace lws;
  LWS_CHECK_PTR(cost);
  LWS_CHECK_PTR(low);
  if (B <= 0 || D <= 0 || H <= 0 || W <= 0 || B > 65535) return LWS_ERR_BAD_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  const long long hw = (long long)H * W;
  // (one pixel per thread for the small 1/8- and 1/4-resolution volumes was measured slower than 4 pixels per thread: 8.4 / 10.4 us
  // against 7.2 / 9.0 us at 24 pairs)
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
    // same chunk length per D as the vector path and the fused tail: the chunking decides the rounding, and all three agree bitwise
    dim3 grid((unsigned)((hw + 255) / 256), B);
    if (D % 9 == 0) softmax_regression_kernel<1, 9><<<grid, 256, 0, st>>>(cost, low, D, hw, hw, start, step);
    else if (D % 12 == 0) softmax_regression_kernel<1, 12><<<grid, 256, 0, st>>>(cost, low, D, hw, hw, start, step);
    else softmax_regression_kernel<1, 8><<<grid, 256, 0, st>>>(cost, low, D, hw, hw, start, step);
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
  dim3 grid(cdiv(W4, 128), cdiv(H, K5_ROWS), B);
  const float fH = (float)H, rh = (float)(1.0 / (double)h), sy = (float)h / (float)H, sx = (float)w / (float)W;
  const int S = H / h;
  if ((W & 3) == 0 && H == S * h && W == S * w && (S == 2 || S == 4 || S == 8) && opt(OPT_K5_INT)) {  // the network's three stages
    if (S == 2) scale_upsample_add_int_kernel<2><<<grid, 128, 0, st>>>(low, prev_or_null, pred, h, w, H, W, fH, rh, sy, sx);
    else if (S == 4) scale_upsample_add_int_kernel<4><<<grid, 128, 0, st>>>(low, prev_or_null, pred, h, w, H, W, fH, rh, sy, sx);
    else scale_upsample_add_int_kernel<8><<<grid, 128, 0, st>>>(low, prev_or_null, pred, h, w, H, W, fH, rh, sy, sx);
    LWS_RETURN_LAUNCH_STATUS();
  }
  scale_upsample_add_kernel<<<grid, 128, 0, st>>>(low, prev_or_null, pred, h, w, H, W, fH, rh, sy, sx);
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
