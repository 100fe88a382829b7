// Shared helpers for the lws_b200 kernels (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "lws.h"

#define LWS_CHECK_PTR(p) \
  do {                   \
    if ((p) == nullptr) return LWS_ERR_NULL_PTR; \
  } while (0)

#define LWS_RETURN_LAUNCH_STATUS()            \
  do {                                        \
    cudaError_t e__ = cudaPeekAtLastError();  \
    return e__ == cudaSuccess ? LWS_OK : (int)e__; \
  } while (0)

namespace lws {

// explicit library options (lws_set_option / lws_get_option, lws_api.cu); the library never reads the environment
enum {
  OPT_CONV3D_TC = 0,
  OPT_REFINE_TC,
  OPT_C8_V1,
  OPT_C8_CHUNK,
  OPT_K1_DT,
  OPT_REFINE_CHAIN,
  OPT_CHAIN_SEP_ITEMS,
  OPT_WARP_DIV_MODE,
  OPT_CHAIN_MIN_BANDS,
  OPT_CHAIN_DEBUG,
  OPT_FIRST_CONV,
  OPT_FUSED_TAIL,
  OPT_C8_GROUP,
  OPT_FUSE_VOLUME,
  OPT_TZ_STRIPS,
  OPT_TZ_DEBUG,
  OPT_K5_INT,
  OPT_FE_TMA,
  OPT_COUNT
};
int opt(int id);

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per kernel and device context instead of once per launch
#define LWS_SET_SMEM_ONCE(kernel, bytes)                                                                        \
  do {                                                                                                          \
    static int done__[16];                                                                                      \
    int dev__ = 0;                                                                                              \
    cudaGetDevice(&dev__);                                                                                      \
    if (dev__ < 0 || dev__ >= 16 || !done__[dev__]) {                                                           \
      cudaError_t e__ = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)); \
      if (e__ != cudaSuccess) return (int)e__;                                                                  \
      if (dev__ >= 0 && dev__ < 16) done__[dev__] = 1;                                                          \
    }                                                                                                           \
  } while (0)

constexpr int kNumSMs = 148;  // B200
// split-fp16 tensor-core operands (x = hi + lo * 2^-11, both fp16): activations are pre-scaled by this power of two so that
// values up to 4.19e6 stay finite in fp16; weights carry their own per-layer power-of-two scale (see the pack functions)
constexpr float kDwsepActScale = 0.015625f;  // 2^-6

__host__ __device__ inline int cdiv(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline int round_up(int a, int b) { return cdiv(a, b) * b; }

// ---- division by a run-time constant without the ~100-cycle integer-divide sequence (dividend in [0, 2^31)) ----------
struct FastDiv {
  uint32_t mul, shr;
  int d;
};
__host__ inline FastDiv make_fastdiv(int d) {
  FastDiv f;
  f.d = d;
  if (d <= 1) {
    f.mul = 0, f.shr = 0;
    return f;
  }
  int lg = 0;
  while ((1ll << lg) < d) ++lg;  // ceil(log2 d)
  const unsigned p = 31 + lg;
  f.mul = (uint32_t)(((1ull << p) + (uint64_t)d - 1) / (uint64_t)d);
  f.shr = p - 32;
  return f;
}
__device__ __forceinline__ int fdiv(int n, const FastDiv& f) { return f.d == 1 ? n : (int)(__umulhi((uint32_t)n, f.mul) >> f.shr); }
__device__ __forceinline__ void fdivmod(int n, const FastDiv& f, int& q, int& r) {
  q = fdiv(n, f);
  r = n - q * f.d;
}

// ---- cp.async (LDGSTS) ------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async_4(void* smem_dst, const void* gmem_src) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_8(void* smem_dst, const void* gmem_src) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gmem_src) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// streaming (evict-first) 128-bit store for write-once outputs
__device__ __forceinline__ void st_cs_f4(float* p, float4 v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w));
}

// fp32 pair -> fp16 pair, round to nearest, SATURATING to +-65504 instead of overflowing to inf (F2FP.SATFINITE: the same single
// instruction as the plain conversion).  Split-fp16 operands carry activations * 2^-6, so values beyond +-4.19e6 saturate: the
// tensor-core paths can never produce inf / NaN from finite inputs (finite but clipped there; options conv3d_tc = 0 / refine_tc = 0
// select the exact-fp32 kernels, which take the whole fp32 range).
__device__ __forceinline__ __half2 f2h2_sat(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return *reinterpret_cast<__half2*>(&r);
}

// low half of the split-fp16 operand of a channel pair: (v - hi) * 2^11 with FADD2 + FMUL2 (sm_100 packed fp32; same roundings as
// the scalar form)
__device__ __forceinline__ float2 split_lo2(float v0, float v1, float2 hi_as_float) {
  return __fmul2_rn(__fadd2_rn(make_float2(v0, v1), make_float2(-hi_as_float.x, -hi_as_float.y)), make_float2(2048.f, 2048.f));
}

// ---- reference coordinate replay (models/models.py:44-53) ------------------------
// Every operation is an explicitly rounded intrinsic so ptxas can never contract them into FMAs: the reference
// issues one Paddle operator per arithmetic step and the tap indices must match it bit for bit.
// The normalisation `2*v / max(size-1, 1)` of models/models.py:44-45 is a tensor divided by a Python scalar.  Paddle 2.0 dygraph
// lowers that to a `scale` op, x * fl32(1/c) (the survey's least verifiable assumption); later Paddle
// versions run a true elementwise division.  Both are implemented; option "warp_div_mode" selects (0 = reciprocal multiply).
struct WarpAxis {
  float recip;  // fl32(1 / max(size-1, 1))
  float half;   // (size-1) * 0.5
  float denom;  // max(size-1, 1)
  int div;      // 1: IEEE division by denom instead of the multiplication by recip
};

__host__ inline WarpAxis make_warp_axis(int size) {
  WarpAxis a;
  a.recip = (float)(1.0 / (double)(size - 1 > 1 ? size - 1 : 1));
  a.half = (float)((double)(size - 1) * 0.5);
  a.denom = (float)(size - 1 > 1 ? size - 1 : 1);
  a.div = opt(OPT_WARP_DIV_MODE);
  return a;
}

__device__ __forceinline__ float warp_normalise(float two_v, WarpAxis a) {
  return a.div ? __fdiv_rn(two_v, a.denom) : __fmul_rn(two_v, a.recip);  // warp-uniform branch
}
// un-normalised sampling coordinate for pixel coordinate `pix` displaced by `dsp`
__device__ __forceinline__ float warp_coord(float pix, float dsp, WarpAxis a) {
  float v = __fsub_rn(pix, dsp);
  float g = __fsub_rn(warp_normalise(__fmul_rn(2.0f, v), a), 1.0f);
  return __fmul_rn(__fadd_rn(g, 1.0f), a.half);
}
__device__ __forceinline__ float warp_coord_nodisp(float pix, WarpAxis a) {
  float g = __fsub_rn(warp_normalise(__fmul_rn(2.0f, pix), a), 1.0f);
  return __fmul_rn(__fadd_rn(g, 1.0f), a.half);
}

struct Tap {
  int i0;      // floor(coord), clamped to [-2, size+1]
  float w0;    // (i0+1) - coord : weight of tap i0
  float w1;    // coord - i0     : weight of tap i0+1
};
__device__ __forceinline__ Tap make_tap(float coord, int size) {
  Tap t;
  float f = floorf(coord);
  t.w1 = __fsub_rn(coord, f);
  t.w0 = __fsub_rn(__fadd_rn(f, 1.0f), coord);
  f = fminf(fmaxf(f, -2.0f), (float)(size + 1));
  t.i0 = (int)f;
  return t;
}

// ---- half-pixel bilinear resize index (paddle F.interpolate defaults) -------------------
struct ResizeTap {
  int i0, i1;
  float l0, l1;
};
__device__ __forceinline__ ResizeTap resize_tap(int dst, float scale, int in_size) {
  float src = __fsub_rn(__fmul_rn(scale, __fadd_rn((float)dst, 0.5f)), 0.5f);
  src = fmaxf(src, 0.0f);
  ResizeTap t;
  t.i0 = min((int)src, in_size - 1);
  t.i1 = min(t.i0 + 1, in_size - 1);
  t.l1 = __fsub_rn(src, (float)t.i0);
  t.l0 = __fsub_rn(1.0f, t.l1);
  return t;
}


// half-pixel bilinear blend of the four taps (a00 a01 / a10 a11) with explicitly rounded operations, so that every kernel that
// upsamples (K5, K2a, the fused regression tail) produces the same bits for the same inputs whatever the surrounding code is
__device__ __forceinline__ float bilinear_blend(float a00, float a01, float a10, float a11, const ResizeTap& tx, const ResizeTap& ty) {
  const float top = __fmaf_rn(tx.l1, a01, __fmul_rn(tx.l0, a00));
  const float bot = __fmaf_rn(tx.l1, a11, __fmul_rn(tx.l0, a10));
  return __fmaf_rn(ty.l1, bot, __fmul_rn(ty.l0, top));
}

// chunked online softmax regression over the disparity axis (K4 and the fused tail share it: same bits).  z[j] = -cost of plane
// d0 + j (-inf beyond D); state (m, s, ws) = running max, sum of exp, sum of exp * disparity.
template <int CH>
__device__ __forceinline__ void softmax_chunk_update(const float (&z)[CH], int d0, float start, float step, float& m, float& s,
                                                     float& ws) {
  constexpr float kLog2e = 1.4426950408889634f;
  float cm = z[0];
#pragma unroll
  for (int j = 1; j < CH; ++j) cm = fmaxf(cm, z[j]);
  const float nm = fmaxf(m, cm);
  const float r = exp2f(__fmul_rn(__fsub_rn(m, nm), kLog2e));  // exp2f(-inf) = 0 on the first chunk
  float cs = 0.f, cws = 0.f;
#pragma unroll
  for (int j = 0; j < CH; ++j) {
    const float e = exp2f(__fmul_rn(__fsub_rn(z[j], nm), kLog2e));
    cs = __fadd_rn(cs, e);
    cws = __fmaf_rn(e, __fmaf_rn(step, (float)(d0 + j), start), cws);
  }
  s = __fmaf_rn(s, r, cs);
  ws = __fmaf_rn(ws, r, cws);
  m = nm;
}

}  // namespace lws
