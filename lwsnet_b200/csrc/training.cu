// n4: the training path of the hot-path kernels -- backward of K1 (stage-1 L1 cost volume), K2 (warp + residual
// volume) and K4 (softmax disparity regression), and the multi-stage smooth-L1 loss of the reference's train loop
// (train.py:127-166: loss = sum_s w_s * smooth_l1(pred_s[mask], gt[mask], mean), mask = gt < maxdisp).  The forward kernels are the
// inference ones; these produce the gradients torch / Paddle autograd would (|x|' = sign(x) with sign(0) = 0; bilinear sampling
// with zero padding differentiated w.r.t. the sampled tensor AND the sampling abscissa, i.e. the disparity).  The convolution
// stacks' backward is not part of this tier (SURVEY.md 8(f) ranks it last); lwsnet_b200/training.py wraps these entries as
// torch.autograd.Functions and tests/test_training_gpu.py checks them against fp64 autograd of the oracle.
#include "lws_common.cuh"

namespace lws {

__device__ __forceinline__ float sgn(float v) { return (float)(v > 0.f) - (float)(v < 0.f); }

// ---- K1 backward: gL[b,c,y,x] = sum_d g[b,d,y,x] * sign(L - Rs),  gR[b,c,y,x'] = -sum_d g[b,d,y,x'+d] * sign(L[x'+d] - R[x']) ----
__global__ void __launch_bounds__(128)
    cost_volume_l1_bwd_kernel(const float* __restrict__ L, const float* __restrict__ R, const float* __restrict__ g, float* __restrict__ gL,
                              float* __restrict__ gR, int C, int H, int W, int planes, int stride) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y, b = blockIdx.z;
  if (x >= W) return;
  const long long hw = (long long)H * W;
  const float* gb = g + (long long)b * planes * hw + (long long)y * W;
  for (int c = 0; c < C; ++c) {
    const float* Lr = L + ((long long)b * C + c) * hw + (long long)y * W;
    const float* Rr = R + ((long long)b * C + c) * hw + (long long)y * W;
    const float l = __ldg(Lr + x), r = __ldg(Rr + x);
    float al = 0.f, ar = 0.f;
    for (int p = 0; p < planes; ++p) {
      const int d = p * stride;
      // this pixel as the LEFT pixel of plane p: right sample R[x - d] (0 where x - d < 0: the reference's occlusion branch)
      const float rs = x - d >= 0 ? __ldg(Rr + x - d) : 0.f;
      al = fmaf(__ldg(gb + p * hw + x), sgn(l - rs), al);
      // this pixel as the RIGHT pixel: it is sampled by left pixel x + d
      if (x + d < W) ar = fmaf(__ldg(gb + p * hw + x + d), -sgn(__ldg(Lr + x + d) - r), ar);
    }
    gL[((long long)b * C + c) * hw + (long long)y * W + x] = al;
    gR[((long long)b * C + c) * hw + (long long)y * W + x] = ar;
  }
}

// ---- K2 backward ---------------------------------------------------------------------------------------------------------------
// gR and gdisp must be zero-filled by the caller-facing entry (gR is accumulated with atomics: several left pixels sample one right
// pixel).  kappa = d(ix)/d(x - disp) of the reference's normalise / un-normalise round trip = (2 * recip) * half.
__global__ void __launch_bounds__(128)
    warp_residual_volume_l1_bwd_kernel(const float* __restrict__ L, const float* __restrict__ R, const float* __restrict__ disp,
                                       const float* __restrict__ g, float* __restrict__ gL, float* __restrict__ gR,
                                       float* __restrict__ gdisp, int C, int H, int W, int m, float fstride, WarpAxis ax, WarpAxis ay) {
  const int px = blockIdx.x * blockDim.x + threadIdx.x;
  const int py = blockIdx.y, b = blockIdx.z;
  if (px >= W) return;
  const int K = 2 * m - 1;
  const long long hw = (long long)H * W;
  const long long pix = (long long)py * W + px;
  const float d = __ldg(disp + (long long)b * hw + pix);
  const Tap ty = make_tap(warp_coord_nodisp((float)py, ay), H);
  const bool r0 = ty.i0 >= 0 && ty.i0 < H, r1 = ty.i0 + 1 >= 0 && ty.i0 + 1 < H;
  const float kappa = ax.div ? __fmul_rn(__fdiv_rn(2.0f, ax.denom), ax.half) : __fmul_rn(__fmul_rn(2.0f, ax.recip), ax.half);
  float gd = 0.f;
  for (int c = 0; c < C; ++c) {
    const float l = __ldg(L + ((long long)b * C + c) * hw + pix);
    const float* row0 = R + ((long long)b * C + c) * hw + (long long)ty.i0 * W;
    float* grow0 = gR + ((long long)b * C + c) * hw + (long long)ty.i0 * W;
    float gl = 0.f;
    for (int k = 0; k < K; ++k) {
      const float shift = __fmul_rn((float)(k - (m - 1)), fstride);
      const Tap t = make_tap(warp_coord((float)px, __fsub_rn(d, shift), ax), W);
      const bool c0 = t.i0 >= 0 && t.i0 < W, c1 = t.i0 + 1 >= 0 && t.i0 + 1 < W;
      const float v00 = (r0 && c0) ? __ldg(row0 + t.i0) : 0.f, v01 = (r0 && c1) ? __ldg(row0 + t.i0 + 1) : 0.f;
      const float v10 = (r1 && c0) ? __ldg(row0 + W + t.i0) : 0.f, v11 = (r1 && c1) ? __ldg(row0 + W + t.i0 + 1) : 0.f;
      const float a = ty.w0 * v00 + ty.w1 * v10, bq = ty.w0 * v01 + ty.w1 * v11;  // vertically blended taps
      const float s = a * t.w0 + bq * t.w1;
      const float sg = sgn(l - s) * __ldg(g + ((long long)b * K + k) * hw + pix);
      gl += sg;
      gd = fmaf(sg * (bq - a), kappa, gd);  // d|l - s|/d disp = sign * (b - a) * kappa
      if (sg != 0.f) {
        if (r0 && c0) atomicAdd(grow0 + t.i0, -sg * t.w0 * ty.w0);
        if (r0 && c1) atomicAdd(grow0 + t.i0 + 1, -sg * t.w1 * ty.w0);
        if (r1 && c0) atomicAdd(grow0 + W + t.i0, -sg * t.w0 * ty.w1);
        if (r1 && c1) atomicAdd(grow0 + W + t.i0 + 1, -sg * t.w1 * ty.w1);
      }
    }
    gL[((long long)b * C + c) * hw + pix] = gl;
  }
  gdisp[(long long)b * hw + pix] = gd;
}

// ---- K4 backward: low = sum_j p_j v_j, p = softmax(-cost)  =>  d low / d cost_j = -p_j (v_j - low) ----------------------------------
__global__ void __launch_bounds__(256)
    softmax_regression_bwd_kernel(const float* __restrict__ cost, const float* __restrict__ glow, float* __restrict__ gcost, int D,
                                  long long hw, float start, float step) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (i >= hw) return;
  const float* p = cost + (long long)b * D * hw + i;
  float mx = -INFINITY;
  for (int j = 0; j < D; ++j) mx = fmaxf(mx, -__ldg(p + j * hw));
  float s = 0.f, ws = 0.f;
  for (int j = 0; j < D; ++j) {
    const float e = __expf(-__ldg(p + j * hw) - mx);
    s += e;
    ws = fmaf(e, fmaf(step, (float)j, start), ws);
  }
  const float inv = 1.f / s, low = ws * inv, gl = __ldg(glow + (long long)b * hw + i);
  float* o = gcost + (long long)b * D * hw + i;
  for (int j = 0; j < D; ++j) {
    const float pj = __expf(-__ldg(p + j * hw) - mx) * inv;
    o[j * hw] = -gl * pj * (fmaf(step, (float)j, start) - low);
  }
}

// ---- multi-stage smooth-L1 loss (train.py:145-155), forward + backward ------------------------------------------------------------
constexpr int LOSS_THREADS = 256, LOSS_MAXS = 4;
struct LossPtrs {
  const float* pred[LOSS_MAXS];
  float* grad[LOSS_MAXS];
  float weight[LOSS_MAXS];
};
// partial[block][stage] = sum of huber(pred_s - gt) over the block's masked elements, partial[block][4] = number of masked elements
__global__ void __launch_bounds__(LOSS_THREADS)
    smooth_l1_partial_kernel(LossPtrs P, const float* __restrict__ gt, int S, long long n, float maxdisp, float* __restrict__ partial) {
  float acc[LOSS_MAXS + 1] = {0.f, 0.f, 0.f, 0.f, 0.f};
  for (long long i = (long long)blockIdx.x * LOSS_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * LOSS_THREADS) {
    const float t = __ldg(gt + i);
    if (t < maxdisp) {
      acc[LOSS_MAXS] += 1.f;
      for (int s = 0; s < S; ++s) {
        const float dlt = fabsf(__ldg(P.pred[s] + i) - t);
        acc[s] += dlt < 1.f ? 0.5f * dlt * dlt : dlt - 0.5f;
      }
    }
  }
  __shared__ float red[LOSS_MAXS + 1][LOSS_THREADS / 32];
#pragma unroll
  for (int k = 0; k <= LOSS_MAXS; ++k) {
    float v = acc[k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x <= LOSS_MAXS) {
    float v = 0.f;
    for (int wv = 0; wv < LOSS_THREADS / 32; ++wv) v += red[threadIdx.x][wv];  // fixed order: deterministic
    partial[blockIdx.x * (LOSS_MAXS + 1) + threadIdx.x] = v;
  }
}
// out[s] = w_s * sum_s / count (0 when the mask is empty: the reference skips such batches, train.py:139-140), out[4] = count
__global__ void smooth_l1_finalize_kernel(const float* __restrict__ partial, int nblocks, LossPtrs P, int S, float* __restrict__ out) {
  const int k = threadIdx.x;
  if (k > LOSS_MAXS) return;
  double v = 0.0;
  for (int bk = 0; bk < nblocks; ++bk) v += (double)partial[bk * (LOSS_MAXS + 1) + k];  // fixed order: deterministic
  __shared__ double tot[LOSS_MAXS + 1];
  tot[k] = v;
  __syncthreads();
  if (k == LOSS_MAXS) out[LOSS_MAXS] = (float)tot[LOSS_MAXS];
  else out[k] = (k < S && tot[LOSS_MAXS] > 0.0) ? (float)((double)P.weight[k] * tot[k] / tot[LOSS_MAXS]) : 0.f;
}
// d loss / d pred_s[i] = w_s * clamp(pred_s - gt, -1, 1) / count on the mask, 0 elsewhere
__global__ void __launch_bounds__(LOSS_THREADS)
    smooth_l1_grad_kernel(LossPtrs P, const float* __restrict__ gt, int S, long long n, float maxdisp, const float* __restrict__ out) {
  const float count = __ldg(out + LOSS_MAXS);
  const float inv = count > 0.f ? 1.f / count : 0.f;
  for (long long i = (long long)blockIdx.x * LOSS_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * LOSS_THREADS) {
    const float t = __ldg(gt + i);
    const bool mk = t < maxdisp;
    for (int s = 0; s < S; ++s)
      if (P.grad[s]) P.grad[s][i] = mk ? P.weight[s] * inv * fminf(fmaxf(__ldg(P.pred[s] + i) - t, -1.f), 1.f) : 0.f;
  }
}

}  // namespace lws

extern "C" int lws_cost_volume_l1_bwd_f32(const float* L, const float* R, const float* gcost, float* gL, float* gR, int B, int C, int H,
                                          int W, int maxdisp, int stride, lws_stream_t stream) {
  using namespace lws;
  LWS_CHECK_PTR(L);
  LWS_CHECK_PTR(R);
  LWS_CHECK_PTR(gcost);
  LWS_CHECK_PTR(gL);
  LWS_CHECK_PTR(gR);
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || maxdisp <= 0 || stride <= 0 || maxdisp % stride || H > 65535 || B > 65535)
    return LWS_ERR_BAD_SHAPE;
  dim3 grid(cdiv(W, 128), H, B);
  cost_volume_l1_bwd_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(L, R, gcost, gL, gR, C, H, W, maxdisp / stride, stride);
  LWS_RETURN_LAUNCH_STATUS();
}

extern "C" int lws_warp_residual_volume_l1_bwd_f32(const float* L, const float* R, const float* disp, const float* gcost, float* gL,
                                                   float* gR, float* gdisp, int B, int C, int H, int W, int m, int stride,
                                                   lws_stream_t stream) {
  using namespace lws;
  LWS_CHECK_PTR(L);
  LWS_CHECK_PTR(R);
  LWS_CHECK_PTR(disp);
  LWS_CHECK_PTR(gcost);
  LWS_CHECK_PTR(gL);
  LWS_CHECK_PTR(gR);
  LWS_CHECK_PTR(gdisp);
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || m <= 0 || stride <= 0 || H > 65535 || B > 65535) return LWS_ERR_BAD_SHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(gR, 0, (size_t)B * C * H * W * sizeof(float), st);  // accumulated with atomics
  if (e != cudaSuccess) return (int)e;
  dim3 grid(cdiv(W, 128), H, B);
  warp_residual_volume_l1_bwd_kernel<<<grid, 128, 0, st>>>(L, R, disp, gcost, gL, gR, gdisp, C, H, W, m, (float)stride, make_warp_axis(W),
                                                           make_warp_axis(H));
  LWS_RETURN_LAUNCH_STATUS();
}

extern "C" int lws_softmax_regression_bwd_f32(const float* cost, const float* glow, float* gcost, int B, int D, int H, int W, float start,
                                              float step, lws_stream_t stream) {
  using namespace lws;
  LWS_CHECK_PTR(cost);
  LWS_CHECK_PTR(glow);
  LWS_CHECK_PTR(gcost);
  if (B <= 0 || D <= 0 || H <= 0 || W <= 0 || B > 65535) return LWS_ERR_BAD_SHAPE;
  const long long hw = (long long)H * W;
  dim3 grid((unsigned)((hw + 255) / 256), B);
  softmax_regression_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(cost, glow, gcost, D, hw, start, step);
  LWS_RETURN_LAUNCH_STATUS();
}

extern "C" size_t lws_smooth_l1_loss_workspace_bytes(long long n) {
  if (n <= 0) return 0;
  long long blocks = (n + lws::LOSS_THREADS - 1) / lws::LOSS_THREADS;
  if (blocks > 4 * lws::kNumSMs) blocks = 4 * lws::kNumSMs;
  return (size_t)blocks * (lws::LOSS_MAXS + 1) * sizeof(float);
}

// preds[s] / grads[s]: device pointers to n floats each (grads[s] may be NULL); weights: HOST array of n_stages floats; out: device
// float[5] = the n_stages weighted stage losses (0 beyond n_stages) and the number of masked pixels
extern "C" int lws_smooth_l1_multistage_loss_f32(const float* const* preds, const float* gt, const float* weights, int n_stages,
                                                 long long n, float maxdisp, float* out, float* const* grads_or_null, void* ws,
                                                 size_t ws_bytes, lws_stream_t stream) {
  using namespace lws;
  LWS_CHECK_PTR(preds);
  LWS_CHECK_PTR(gt);
  LWS_CHECK_PTR(weights);
  LWS_CHECK_PTR(out);
  LWS_CHECK_PTR(ws);
  if (n_stages <= 0 || n_stages > LOSS_MAXS || n <= 0) return LWS_ERR_BAD_SHAPE;
  if (ws_bytes < lws_smooth_l1_loss_workspace_bytes(n)) return LWS_ERR_WORKSPACE_TOO_SMALL;
  LossPtrs P;
  memset(&P, 0, sizeof(P));
  bool any_grad = false;
  for (int s = 0; s < n_stages; ++s) {
    if (!preds[s]) return LWS_ERR_NULL_PTR;
    P.pred[s] = preds[s], P.weight[s] = weights[s];
    P.grad[s] = grads_or_null ? grads_or_null[s] : nullptr;
    any_grad = any_grad || P.grad[s] != nullptr;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = (int)(lws_smooth_l1_loss_workspace_bytes(n) / ((LOSS_MAXS + 1) * sizeof(float)));
  smooth_l1_partial_kernel<<<blocks, LOSS_THREADS, 0, st>>>(P, gt, n_stages, n, maxdisp, (float*)ws);
  smooth_l1_finalize_kernel<<<1, 32, 0, st>>>((const float*)ws, blocks, P, n_stages, out);
  if (any_grad) smooth_l1_grad_kernel<<<blocks, LOSS_THREADS, 0, st>>>(P, gt, n_stages, n, maxdisp, out);
  LWS_RETURN_LAUNCH_STATUS();
}
