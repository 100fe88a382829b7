// K1: stage-1 L1 cost volume.  Replaces the Python loop of LWSNet._build_volume_2d
// (reference models/models.py:58-76: ~9 Paddle launches per disparity) with one launch.
//
//   cost[b,d,y,x] = sum_c | L[b,c,y,x] - (x-d >= 0 ? R[b,c,y,x-d] : 0) |
//
// HBM-bound by design: algorithmic bytes = (2*C + D) * 4 per pixel, but with 2*C*D FADD-class
// operations per pixel the ALU floor is only ~1.6x below the HBM floor on B200, so the kernel is organised around
// FADD issue efficiency:
//   * a block owns `nr` whole row segments; the L rows and the R rows (extended DT to the left, zero-filled where
//     x-d < 0, which reproduces the reference's "occlusion" branch) are staged in shared memory with cp.async,
//     double-buffered over channel chunks, fully coalesced whatever the row alignment is (W=154 is only 8B aligned);
//   * a thread owns 4 consecutive pixels x DT disparities: per channel it reads its 4 L values and an aligned
//     (DT+4)-wide R window with 128-bit LDS and slides it in registers: (DT+8)/4 LDS.128 for 8*DT FADDs.
#include "lws_common.cuh"

namespace lws {

__global__ void cost_volume_l1_generic_kernel(const float* __restrict__ L, const float* __restrict__ R,
                                              float* __restrict__ cost, int C, int H, int W, int planes, int stride,
                                              long long total) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  int x = (int)(idx % W);
  long long t = idx / W;
  int y = (int)(t % H);
  t /= H;
  int p = (int)(t % planes);
  int b = (int)(t / planes);
  int d = p * stride;
  const float* l = L + ((long long)b * C * H + y) * W + x;
  const float* r = R + ((long long)b * C * H + y) * W + (x - d);
  float acc = 0.f;
  long long cs = (long long)H * W;
  for (int c = 0; c < C; ++c) {
    float rv = (x - d >= 0) ? __ldg(r + c * cs) : 0.f;
    acc += fabsf(__ldg(l + c * cs) - rv);
  }
  cost[idx] = acc;
}

template <int DT>
__global__ void __launch_bounds__(256, 1)
    cost_volume_l1_tile_kernel(const float* __restrict__ L, const float* __restrict__ R, float* __restrict__ cost,
                               int C, int H, int W, int D, int n_xtiles, int txq, int nr, int ck, int n_dtiles) {
  extern __shared__ __align__(16) float smem[];
  const int tx = txq * 4;
  const int pitchR = tx + DT;
  const int stage_floats = ck * nr * (tx + pitchR);
  const int xtile = blockIdx.x % n_xtiles;
  const int rowgroup = blockIdx.x / n_xtiles;
  const int b = blockIdx.y / n_dtiles;
  const int d0 = (blockIdx.y % n_dtiles) * DT;
  const int x_begin = xtile * tx;
  const int y_begin = rowgroup * nr;
  const int rows_here = min(nr, H - y_begin);
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const long long cs = (long long)H * W;
  const float* Lb = L + (long long)b * C * cs;
  const float* Rb = R + (long long)b * C * cs;

  // stage one channel chunk: rows are decoded once per warp iteration, elements are contiguous over lanes
  auto issue = [&](int chunk, int stage) {
    float* sL = smem + stage * stage_floats;
    float* sR = sL + ck * nr * tx;
    const int c0 = chunk * ck;
    const int nrows = ck * rows_here;
    for (int item = warp; item < 2 * nrows; item += nwarps) {
      const int which = item >= nrows;  // 0: L, 1: R
      const int ri = which ? item - nrows : item;
      const int cl = ri / rows_here, r = ri - cl * rows_here;
      const float* grow = (which ? Rb : Lb) + (long long)(c0 + cl) * cs + (long long)(y_begin + r) * W;
      if (!which) {
        float* srow = sL + (cl * nr + r) * tx;
        for (int i = lane; i < tx; i += 32) {
          int x = x_begin + i;
          if (x < W) cp_async_4(srow + i, grow + x);
          else srow[i] = 0.f;
        }
      } else {
        float* srow = sR + (cl * nr + r) * pitchR;
        const int xs = x_begin - d0 - DT;  // sR[i] = R[xs + i]
        for (int i = lane; i < pitchR; i += 32) {
          int x = xs + i;
          if (x >= 0 && x < W) cp_async_4(srow + i, grow + x);
          else srow[i] = 0.f;
        }
      }
    }
    cp_async_commit();
  };

  const bool active = tid < rows_here * txq;
  const int r = active ? tid / txq : 0;
  const int xq = active ? tid - r * txq : 0;
  const int x0l = xq * 4;

  float acc[4][DT];
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int j = 0; j < DT; ++j) acc[p][j] = 0.f;

  const int nchunks = C / ck;
  issue(0, 0);
  for (int ch = 0; ch < nchunks; ++ch) {
    if (ch + 1 < nchunks) {
      issue(ch + 1, (ch + 1) & 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (active) {
      const float* sL = smem + (ch & 1) * stage_floats;
      const float* sR = sL + ck * nr * tx;
      for (int cl = 0; cl < ck; ++cl) {
        const float4 l4 = *reinterpret_cast<const float4*>(sL + (cl * nr + r) * tx + x0l);
        const float l[4] = {l4.x, l4.y, l4.z, l4.w};
        float w[DT + 4];
        const float4* wp = reinterpret_cast<const float4*>(sR + (cl * nr + r) * pitchR + x0l);
#pragma unroll
        for (int i = 0; i < (DT + 4) / 4; ++i) {
          float4 v = wp[i];
          w[4 * i] = v.x, w[4 * i + 1] = v.y, w[4 * i + 2] = v.z, w[4 * i + 3] = v.w;
        }
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
          for (int j = 0; j < DT; ++j) acc[p][j] += fabsf(l[p] - w[p - j + DT]);
      }
    }
    __syncthreads();
  }

  if (!active) return;
  const int y = y_begin + r;
  const int x0 = x_begin + x0l;
  if (x0 >= W) return;
  float* out = cost + (((long long)b * D + d0) * H + y) * W + x0;
  const bool vec = ((W & 3) == 0);
#pragma unroll
  for (int j = 0; j < DT; ++j) {
    if (d0 + j < D) {
      float* o = out + (long long)j * cs;
      if (vec) {
        *reinterpret_cast<float4*>(o) = make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
      } else {
#pragma unroll
        for (int p = 0; p < 4; ++p)
          if (x0 + p < W) o[p] = acc[p][j];
      }
    }
  }
}

template <int DT>
static int launch_tile(const float* L, const float* R, float* cost, int B, int C, int H, int W, int D,
                       cudaStream_t st) {
  const int W4 = cdiv(W, 4);
  const int txq = W4 < 256 ? W4 : 256;
  const int n_xtiles = cdiv(W4, txq);
  int nr = 256 / txq;
  if (nr < 1) nr = 1;
  if (nr > H) nr = H;
  const int ck = (C % 4 == 0) ? 4 : (C % 2 == 0 ? 2 : 1);
  const int n_dtiles = cdiv(D, DT);
  const int tx = txq * 4;
  const size_t smem = (size_t)2 * ck * nr * (tx + tx + DT) * sizeof(float);
  if (smem > 200 * 1024) return LWS_ERR_UNSUPPORTED;
  auto kern = cost_volume_l1_tile_kernel<DT>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  dim3 grid(n_xtiles * cdiv(H, nr), B * n_dtiles);
  kern<<<grid, 256, smem, st>>>(L, R, cost, C, H, W, D, n_xtiles, txq, nr, ck, n_dtiles);
  LWS_RETURN_LAUNCH_STATUS();
}

}  // namespace lws

extern "C" int lws_cost_volume_l1_f32(const float* L, const float* R, float* cost, int B, int C, int H, int W,
                                      int maxdisp, int stride, lws_stream_t stream) {
  using namespace lws;
  LWS_CHECK_PTR(L);
  LWS_CHECK_PTR(R);
  LWS_CHECK_PTR(cost);
  if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || maxdisp <= 0 || stride <= 0) return LWS_ERR_BAD_SHAPE;
  if (maxdisp % stride != 0) return LWS_ERR_BAD_SHAPE;  // reference: assert maxdisp % stride == 0 (models.py:63)
  cudaStream_t st = (cudaStream_t)stream;
  const int D = maxdisp / stride;
  const bool aligned = (((uintptr_t)cost) & 15) == 0;
  if (stride == 1 && aligned && B * (long long)cdiv(D, 8) < 65535) {
    int rc;
    if (D % 24 == 0) rc = launch_tile<24>(L, R, cost, B, C, H, W, D, st);
    else if (D % 16 == 0) rc = launch_tile<16>(L, R, cost, B, C, H, W, D, st);
    else if (D % 12 == 0) rc = launch_tile<12>(L, R, cost, B, C, H, W, D, st);
    else rc = launch_tile<8>(L, R, cost, B, C, H, W, D, st);
    if (rc != LWS_ERR_UNSUPPORTED) return rc;
  }
  const long long total = (long long)B * D * H * W;
  const int threads = 256;
  cost_volume_l1_generic_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, st>>>(L, R, cost, C, H, W,
                                                                                                 D, stride, total);
  LWS_RETURN_LAUNCH_STATUS();
}
