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
#include <stdlib.h>

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
__global__ void __launch_bounds__(128, 3)
    cost_volume_l1_tile_kernel(const float* __restrict__ L, const float* __restrict__ R, float* __restrict__ cost,
                               int C, int H, int W, int D, int n_xtiles, int txq, int nr, int ck, int n_dtiles, int vec2) {
  extern __shared__ __align__(16) float smem[];
  const int tx = txq * 4;
  const int pitchR = tx + DT;
  const int stage_floats = ck * nr * (tx + pitchR);  // one stage when ck == C
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
      // sR[i] = R[xs + i]; xs, x_begin and (with vec2) W and the row starts are even, so an element pair is wholly in or out
      float* srow = which ? sR + (cl * nr + r) * pitchR : sL + (cl * nr + r) * tx;
      const int xs = which ? x_begin - d0 - DT : x_begin;
      const int n = which ? pitchR : tx;
      if (vec2) {
        for (int i = 2 * lane; i < n; i += 64) {
          const int x = xs + i;
          if (x >= 0 && x < W) cp_async_8(srow + i, grow + x);
          else *reinterpret_cast<float2*>(srow + i) = make_float2(0.f, 0.f);
        }
      } else {
        for (int i = lane; i < n; i += 32) {
          const int x = xs + i;
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

// ---- direct kernel: no shared memory ---------------------------------------------------------------------------------------
// A thread owns 4 consecutive pixels x DT disparities and reads, per channel, its 4 L values and the aligned (DT+4)-wide R window
// [x0 - d0 - DT, x0 - d0 + 4) straight from global memory with 64-bit loads (rows of the 1/8-resolution maps are only 8-byte
// aligned: W = 154).  Neighbouring threads' windows overlap almost completely, so the loads are L1 hits and HBM sees every row
// once; the window of channel c+1 is loaded into a second register set while channel c is accumulated (ping-pong), so there is
// no staging pass, no barrier and ~12 % instruction overhead on top of the 2*C*D FADDs per pixel that the arithmetic needs.
template <int DT>
__global__ void __launch_bounds__(128)
    cost_volume_l1_direct_kernel(const float* __restrict__ L, const float* __restrict__ R, float* __restrict__ cost, int C, int H,
                                 int W, int D, int wq, int n_dtiles, int items) {
  const int item = blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= items) return;
  const int y = item / wq;
  const int x0 = (item - y * wq) * 4;
  const int b = blockIdx.y / n_dtiles;
  const int d0 = (blockIdx.y - b * n_dtiles) * DT;
  const int cs = H * W;
  const float* Lp = L + (long long)b * C * cs + y * W + x0;
  const int xs = x0 - d0 - DT;  // w[i] = R[xs + i]; R[x0 + p - (d0 + j)] = w[p - j + DT]
  const float* Rp = R + (long long)b * C * cs + y * W + xs;
  constexpr int NW = (DT + 4) / 2;
  bool okw[NW];
#pragma unroll
  for (int i = 0; i < NW; ++i) okw[i] = xs + 2 * i >= 0 && xs + 2 * i < W;  // xs and W are even: a pair is wholly in or out
  const bool okl1 = x0 + 2 < W;
  auto load = [&](int c, float (&l)[4], float (&w)[DT + 4]) {
    const float2 a = __ldg(reinterpret_cast<const float2*>(Lp + c * cs));
    const float2 bq = okl1 ? __ldg(reinterpret_cast<const float2*>(Lp + c * cs + 2)) : make_float2(0.f, 0.f);
    l[0] = a.x, l[1] = a.y, l[2] = bq.x, l[3] = bq.y;
#pragma unroll
    for (int i = 0; i < NW; ++i) {
      const float2 v = okw[i] ? __ldg(reinterpret_cast<const float2*>(Rp + c * cs) + i) : make_float2(0.f, 0.f);
      w[2 * i] = v.x, w[2 * i + 1] = v.y;
    }
  };
  float acc[4][DT];
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int j = 0; j < DT; ++j) acc[p][j] = 0.f;
  auto mac = [&](const float (&l)[4], const float (&w)[DT + 4]) {
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int j = 0; j < DT; ++j) acc[p][j] += fabsf(l[p] - w[p - j + DT]);
  };
  float la[4], lb[4], wa[DT + 4], wb[DT + 4];
  load(0, la, wa);
  for (int c = 0; c < C; c += 2) {  // C is even (host-checked)
    load(c + 1, lb, wb);
    mac(la, wa);
    if (c + 2 < C) load(c + 2, la, wa);
    mac(lb, wb);
  }
  float* out = cost + (((long long)b * D + d0) * H + y) * W + x0;
  const bool v4 = (reinterpret_cast<uintptr_t>(out) & 15) == 0 && ((cs & 3) == 0) && okl1;
#pragma unroll
  for (int j = 0; j < DT; ++j) {
    if (d0 + j < D) {
      float* o = out + (long long)j * cs;
      if (v4) {
        __stcs(reinterpret_cast<float4*>(o), make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]));
      } else {
        __stcs(reinterpret_cast<float2*>(o), make_float2(acc[0][j], acc[1][j]));
        if (okl1) __stcs(reinterpret_cast<float2*>(o + 2), make_float2(acc[2][j], acc[3][j]));
      }
    }
  }
}

template <int DT>
static int launch_direct(const float* L, const float* R, float* cost, int B, int C, int H, int W, int D, cudaStream_t st) {
  const int wq = cdiv(W, 4), items = wq * H, n_dtiles = cdiv(D, DT);
  dim3 grid(cdiv(items, 128), B * n_dtiles);
  cost_volume_l1_direct_kernel<DT><<<grid, 128, 0, st>>>(L, R, cost, C, H, W, D, wq, n_dtiles, items);
  LWS_RETURN_LAUNCH_STATUS();
}

struct TileCfg {
  int threads, txq, n_xtiles, nr, ck, n_dtiles;
  size_t smem;
  long long blocks;
};

// One block = nr row segments x DT disparities.  All C channels are staged in one shot when they fit (a single HBM round trip
// per block; the stage-1 volume of a few pairs is latency-, not bandwidth-limited), otherwise 4-channel chunks double-buffered.
static TileCfg tile_cfg(int B, int C, int H, int W, int D, int DT, int threads) {
  TileCfg t;
  const int W4 = cdiv(W, 4);
  t.threads = threads;
  t.txq = W4 < threads ? W4 : threads;
  t.n_xtiles = cdiv(W4, t.txq);
  t.nr = threads / t.txq;
  t.nr = t.nr < 1 ? 1 : (t.nr > H ? H : t.nr);
  t.n_dtiles = cdiv(D, DT);
  const int tx = t.txq * 4;
  const size_t one = (size_t)C * t.nr * (tx + tx + DT) * sizeof(float);
  if (one <= 72 * 1024) {
    t.ck = C, t.smem = one;
  } else {
    t.ck = (C % 4 == 0) ? 4 : (C % 2 == 0 ? 2 : 1);
    t.smem = (size_t)2 * t.ck * t.nr * (tx + tx + DT) * sizeof(float);
  }
  t.blocks = (long long)t.n_xtiles * cdiv(H, t.nr) * B * t.n_dtiles;
  return t;
}

template <int DT>
static int launch_tile(const float* L, const float* R, float* cost, int B, int C, int H, int W, int D, const TileCfg& t,
                       cudaStream_t st) {
  if (t.smem > 200 * 1024) return LWS_ERR_UNSUPPORTED;
  auto kern = cost_volume_l1_tile_kernel<DT>;
  if (t.smem > 48 * 1024) LWS_SET_SMEM_ONCE(kern, 200 * 1024);  // the cap checked above, once per DT instantiation
  const int vec2 = (W % 2 == 0) && ((((uintptr_t)L) | ((uintptr_t)R)) & 7) == 0;
  dim3 grid(t.n_xtiles * cdiv(H, t.nr), B * t.n_dtiles);
  kern<<<grid, t.threads, t.smem, st>>>(L, R, cost, C, H, W, D, t.n_xtiles, t.txq, t.nr, t.ck, t.n_dtiles, vec2);
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
    // direct kernel: even W and C, 8-byte aligned tensors
    const bool direct = (W % 2 == 0) && (C % 2 == 0) && (long long)C * H * W < (1ll << 31) &&
                        ((((uintptr_t)L) | ((uintptr_t)R) | ((uintptr_t)cost)) & 7) == 0 && (long long)B * cdiv(D, 8) <= 65535;
    if (direct) {
      // DT = 8 measured fastest at every batch size (B = 64: 37 us against 41 us for DT = 12 and 55 us for DT = 24): the kernel is
      // bound by resident warps (96 registers -> 20 warps / SM against 8 at DT = 24) and L1 wavefronts, not by window re-reads
      const int force = opt(OPT_K1_DT);  // lws_set_option("k1_dt", 8 | 12 | 24)
      if (force == 24 && D % 24 == 0) return launch_direct<24>(L, R, cost, B, C, H, W, D, st);
      if (force == 12 && D % 12 == 0) return launch_direct<12>(L, R, cost, B, C, H, W, D, st);
      return launch_direct<8>(L, R, cost, B, C, H, W, D, st);
    }
    // otherwise the shared-memory tile kernel: widest disparity tile that still puts >= 3 blocks on every SM
    int rc = LWS_ERR_UNSUPPORTED;
    const long long want = 3 * kNumSMs;
    const int dts[4] = {24, 16, 12, 8};
    int pick = 8, threads = 128;
    bool found = false;
    for (int i = 0; i < 4 && !found; ++i)
      if ((D % dts[i] == 0 || dts[i] == 8) && tile_cfg(B, C, H, W, D, dts[i], 128).blocks >= want) pick = dts[i], found = true;
    if (!found && tile_cfg(B, C, H, W, D, 8, 128).blocks < 2 * kNumSMs) threads = 64;
    const TileCfg t = tile_cfg(B, C, H, W, D, pick, threads);
    if (pick == 24) rc = launch_tile<24>(L, R, cost, B, C, H, W, D, t, st);
    else if (pick == 16) rc = launch_tile<16>(L, R, cost, B, C, H, W, D, t, st);
    else if (pick == 12) rc = launch_tile<12>(L, R, cost, B, C, H, W, D, t, st);
    else rc = launch_tile<8>(L, R, cost, B, C, H, W, D, t, st);
    if (rc != LWS_ERR_UNSUPPORTED) return rc;
  }
  const long long total = (long long)B * D * H * W;
  const int threads = 256;
  cost_volume_l1_generic_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, st>>>(L, R, cost, C, H, W,
                                                                                                 D, stride, total);
  LWS_RETURN_LAUNCH_STATUS();
}
