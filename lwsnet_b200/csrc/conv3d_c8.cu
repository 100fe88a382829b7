// K3 (C = 8): the residual 3D-conv stacks of stages 2 and 3, post_3dconvs(4, 8) (reference models/submodules.py:190-221,
// models/models.py:136-138), on tcgen05 with split-fp16 operands.  HBM/latency bound by nature (8 channels: 64 B of traffic and
// 3456 flop per voxel and layer); the design minimises the number of MMAs (an SS MMA at small N is bound by the shared-memory port:
// max(N/2, (4096 + 32 N) / 128) cycles, tools/umma_ts_bench.cu).
//
// Layout: two planes per tensor, hi and lo, each  act[b][d+2][y+2][x+2][8] fp16  (one voxel = one 16-byte row; zero border in
// all three axes; x*2^-6 = hi + lo*2^-11, see conv3d_f16.cu for the numerics).  With SWIZZLE_NONE K-major UMMA descriptors a
// "core matrix" is 8 rows x 16 bytes stored contiguously, so a run of 128 consecutive voxels IS a 128 x 8 operand tile, and
// the leading-dimension offset selects where the second K chunk comes from: K = 16 = [8 hi channels | 8 lo channels] with
// LBO = distance between the hi and the lo box.  One tcgen05.mma (M=128, N=48, K=16) per (kd, kh) tap:
//     B rows  0..23 = [wh(kw, co) | 0     ]  -> acc_main = xh*wh
//     B rows 24..47 = [wl(kw, co) | wh    ]  -> acc_corr = xh*wl + xl*wh
// and the three kw taps are Toeplitz column blocks: out[r] = E0[r-1] + E1[r] + E2[r+1], E = main/sw + corr/(sw*2^11).
// 9 MMAs per 126 output voxels.  Operand boxes are plain 2 KB bulk copies (cp.async.bulk), one mbarrier per group of six
// (3 kd x hi/lo) = one (y+kh) line position; a CTA walks down y so two of the three groups of a tile are already in the ring.
// Warps: 0 = producer, 1 = MMA issuer, 2.. = four epilogue groups of four warps that take tiles round robin (accumulator
// ti % 4): the epilogue of a tile is a ~1500-cycle dependent chain, a tile's MMAs only ~500 cycles.
#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "lws_common.cuh"
#include "tma_utils.cuh"

namespace lws {

constexpr int C8_NGRP = 4;                  // epilogue groups = TMEM accumulators; group g handles the tiles with ti % NGRP == g
constexpr int C8_THREADS = 64 + C8_NGRP * 128;
constexpr int C8_NG = 8;                    // ring of line groups
constexpr int C8_GBYTES = 6 * 2048;         // 3 kd x (hi, lo) x 128 voxels x 16 B
constexpr int C8_BBYTES = 9 * 1536;         // 9 taps x 48 rows x 32 B
constexpr int C8_OFF_RING = 14336;
constexpr int C8_OFF_STAGE = C8_OFF_RING + C8_NG * C8_GBYTES;  // [NGRP groups][hi 2048 | lo 2048]
constexpr int C8_OFF_XCH = C8_OFF_STAGE + C8_NGRP * 4096;      // [NGRP groups][4 quarters][3 rows][8] floats
constexpr int C8_OFF_BAR = C8_OFF_XCH + C8_NGRP * 4 * 3 * 8 * 4;
constexpr int C8_SMEM = C8_OFF_BAR + 256 + 128;

struct C8Args {
  const uint8_t* in_hi;   // plane base (voxel 0 of batch element 0); `slack` voxels before it are readable
  const uint8_t* in_lo;
  uint8_t* out_hi;
  uint8_t* out_lo;
  const uint8_t* wtab;    // device: 9 x 1536 B operand table, then scales[2] (float): 1/sw, 1/(sw * 2^11)
  const float* bias;      // [8]
  const float* skip;      // LAST: raw cost [B,D,H,W] or null
  float* out_f32;         // LAST: [B,D,H,W]
  long long vox_b;        // voxels per batch element (Dp * Hp * Wp)
  int Hp, Wp, H, W, D;    // padded / interior plane size, interior depth
  FastDiv fWp, fHp;
  int line0, strip_len;   // first computed line (= Hp: plane 1) and number of lines (D * Hp)
  int total_tiles;        // B * ct_per_line * strip_len, split evenly (contiguously) over the CTAs
  FastDiv ct_per_line;
};

struct C8Item {
  int b, row0, ntiles;  // row0: first output voxel (inside the batch element) of tile 0; tile n is Wp voxels further
};
// Schedule: a strip = (b, column tile ct) walked down all strip_len lines; the B * ct * strip_len tiles are numbered strip-major
// and every CTA takes one contiguous range (balanced to one tile; two warm-up line groups per partial strip only).
struct C8Sched {
  int g, g1;
  __device__ __forceinline__ C8Sched(const C8Args& a) {
    g = (int)((long long)a.total_tiles * blockIdx.x / gridDim.x);
    g1 = (int)((long long)a.total_tiles * (blockIdx.x + 1) / gridDim.x);
  }
  __device__ __forceinline__ bool next(const C8Args& a, C8Item& it) {
    if (g >= g1) return false;
    const int strip = g / a.strip_len, n0 = g - strip * a.strip_len;
    int ct;
    fdivmod(strip, a.ct_per_line, it.b, ct);
    it.ntiles = min(a.strip_len - n0, g1 - g);
    it.row0 = (a.line0 + n0) * a.Wp + ct * 126;
    g += it.ntiles;
    return true;
  }
};

__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(reinterpret_cast<uint64_t>(dst)),
               "r"(smem_u32(src_smem)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void c8_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[j]);
}

// LAST: the closing 8 -> 1 conv: N = 16 (row kw = [wh | 0], row 8 + kw = [wl | wh]), fp32 NCDHW output (+ skip).
template <bool LAST>
__global__ void __launch_bounds__(C8_THREADS, 1) conv3d_c8_kernel(const C8Args a) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sB = smem;
  uint8_t* sRing = smem + C8_OFF_RING;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C8_OFF_BAR);
  uint64_t* g_full = bars;              // [NG]
  uint64_t* g_empty = g_full + C8_NG;   // [NG]
  uint64_t* t_full = g_empty + C8_NG;       // [NGRP]
  uint64_t* t_empty = t_full + C8_NGRP;     // [NGRP]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + C8_NGRP);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < C8_NG; ++i) mbar_init(g_full + i, 1), mbar_init(g_empty + i, 1);
    for (int i = 0; i < C8_NGRP; ++i) mbar_init(t_full + i, 1), mbar_init(t_empty + i, 4);
    mbar_fence_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(C8_NGRP * 64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int i = tid; i < (LAST ? 9 * 512 : C8_BBYTES) / 16; i += C8_THREADS)
    reinterpret_cast<uint4*>(sB)[i] = __ldg(reinterpret_cast<const uint4*>(a.wtab) + i);
  fence_proxy_async_smem();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ================================ producer ================================
    // lanes 0..5 each issue one of the six 2 KB bulk copies of a group (kd = lane / 2, hi / lo = lane % 2).  (A 4D TMA tensor
    // box with a 16-byte inner extent moves the same 12 KB in one operation but runs ~1.5x slower.)
    const int kd = (lane >> 1) % 3, part = lane & 1;
    const uint8_t* src_plane = part ? a.in_lo : a.in_hi;
    const long long plane = (long long)a.Hp * a.Wp;  // voxels per d plane
    uint32_t slot = 0, ph = 0;
    C8Sched sched(a);
      C8Item w;
      while (sched.next(a, w)) {
      const long long base = (long long)w.b * a.vox_b + w.row0 - 1;  // GEMM row 0 of tile 0 (Toeplitz shift 1)
      for (int n = 0; n < w.ntiles; ++n) {
        for (int kh = n == 0 ? 0 : 2; kh < 3; ++kh) {  // later tiles of a strip only need the line below
          mbar_wait(g_empty + slot, ph ^ 1);
          if (lane == 0) mbar_expect_tx(g_full + slot, C8_GBYTES);
          __syncwarp();
          if (lane < 6) {
            const long long v = base + (long long)(n + kh - 1) * a.Wp + (kd - 1) * plane;
            bulk_load(sRing + slot * C8_GBYTES + kd * 4096 + part * 2048, src_plane + v * 16, 2048, g_full + slot);
          }
          if (++slot == C8_NG) slot = 0, ph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    constexpr uint32_t NN = LAST ? 16 : 48;
    const uint32_t idesc = (1u << 4) | ((NN >> 3) << 17) | ((128u >> 4) << 24);  // f16 x f16 -> f32, K-major, M = 128, N = 48 / 16
    // SWIZZLE_NONE K-major: LBO = byte distance between the two K chunks, SBO = 128 B between 8-row groups
    const uint64_t a_hi = ((uint64_t)(2048 >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46);
    const uint64_t b_hi = ((uint64_t)((NN * 16) >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46);
    const uint32_t ring_lo = (smem_u32(sRing) & 0x3FFFF) >> 4, b_lo = (smem_u32(sB) & 0x3FFFF) >> 4;
    // the whole schedule runs inside one elected lane: a tile is only nine MMAs (~500 cycles of tensor time), so per-tile
    // elect / reconverge / __syncwarp overhead in the issuing warp would be the kernel's critical path
    if (elect_one_sync()) {
      uint32_t bslot = 0, bph = 0, ti = 0;
      C8Sched sched(a);
        C8Item w;
        while (sched.next(a, w)) {
        for (int n = 0; n < w.ntiles; ++n, ++ti) {
          const bool last = n == w.ntiles - 1;
          const uint32_t tb = ti % C8_NGRP;
          mbar_wait(t_empty + tb, ((ti / C8_NGRP) & 1) ^ 1);
          uint32_t slot = bslot, ph = bph;
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
            if (n == 0 || kh == 2) mbar_wait(g_full + slot, ph);  // the other groups were waited for by the previous tile
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t g_lo = ring_lo + slot * (C8_GBYTES >> 4);
#pragma unroll
            for (int kd = 0; kd < 3; ++kd) {
              const uint64_t da = a_hi | (uint64_t)(g_lo + kd * (4096 >> 4));
              const uint64_t db = b_hi | (uint64_t)(b_lo + (kd * 3 + kh) * ((NN * 32) >> 4));
              const uint32_t acc = (kh | kd) == 0 ? 0u : 1u;
              asm volatile(
                  "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                  "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tmem + tb * 64),
                  "l"(da), "l"(db), "r"(idesc), "r"(acc)
                  : "memory");
            }
            if (kh == 0 || last)
              asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(g_empty + slot))
                           : "memory");
            if (kh == 2)
              asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(t_full + tb))
                           : "memory");
            if (++slot == C8_NG) slot = 0, ph ^= 1;
          }
          if (++bslot == C8_NG) bslot = 0, bph ^= 1;
        }
        bslot += 2;  // the strip's last tile consumed its remaining two groups
        if (bslot >= C8_NG) bslot -= C8_NG, bph ^= 1;
      }
    }
    __syncwarp();
  } else {
    // ================================ epilogue: group g handles the tiles with ti & 1 == g ================================
    const int g = (warp - 2) >> 2;
    const int q = warp & 3;            // TMEM lane quarter
    const int j = q * 32 + lane;       // GEMM row; this thread produces output voxel j (staging row j), valid for j < 126
    const bool issuer = ((warp - 2) & 3) == 0 && lane == 0;
    const float* scl = reinterpret_cast<const float*>(a.wtab + (LAST ? 9 * 512 : C8_BBYTES));
    const float c0 = __ldg(scl) * (LAST ? 1.f / kDwsepActScale : 1.f), c1 = __ldg(scl + 1) * (LAST ? 1.f / kDwsepActScale : 1.f);
    float bias[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) bias[c] = __ldg(a.bias + c) * kDwsepActScale;
    uint8_t* stage = smem + C8_OFF_STAGE + g * 4096;
    float* xch = reinterpret_cast<float*>(smem + C8_OFF_XCH) + g * (4 * 3 * 8);
    float* xq = xch + q * 24;                    // this quarter publishes e1 of lane 0, e2 of lanes 0 and 1
    const float* xn = xch + ((q + 1) & 3) * 24;  // next quarter's
    const int bar_id = 1 + g;
    uint32_t ti = 0;
    C8Sched sched(a);
      C8Item w;
      while (sched.next(a, w)) {
      for (int n = 0; n < w.ntiles; ++n, ++ti) {
        if ((int)(ti % C8_NGRP) != g) continue;
        mbar_wait(t_full + g, (ti / C8_NGRP) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + g * 64;
        if (LAST) {
          float m[8], k[8];
          c8_ld8(taddr, m);
          c8_ld8(taddr + 8, k);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(t_empty + g);
          const float e0 = fmaf(k[0], c1, m[0] * c0), e1 = fmaf(k[1], c1, m[1] * c0), e2 = fmaf(k[2], c1, m[2] * c0);
          const float s1 = __shfl_down_sync(0xffffffffu, e1, 1), s2 = __shfl_down_sync(0xffffffffu, e2, 2);
          float v = e0 + (lane < 31 ? s1 : 0.f) + (lane < 30 ? s2 : 0.f);
          float* xp = xch + ((ti / C8_NGRP) & 1) * 12;  // this group's exchange rows, double-buffered by its tile parity
          if (lane == 0) xp[q * 3] = e1;
          if (lane < 2) xp[q * 3 + 1 + lane] = e2;
          named_bar_sync(bar_id, 128);
          if (q < 3 && lane >= 30) {
            if (lane == 31) v += xp[(q + 1) * 3];
            v += xp[(q + 1) * 3 + 1 + lane - 30];
          }
          const int r = w.row0 + n * a.Wp + j;
          int line, x, dpl, y;
          fdivmod(r, a.fWp, line, x);
          fdivmod(line, a.fHp, dpl, y);
          if (j < 126 && x >= 1 && x <= a.W && y >= 1 && y <= a.H && dpl >= 1 && dpl <= a.D) {
            const long long o = (((long long)w.b * a.D + (dpl - 1)) * a.H + (y - 1)) * a.W + (x - 1);
            a.out_f32[o] = v + (a.skip ? __ldg(a.skip + o) : 0.f);
          }
          continue;
        }
        float m0[8], m1[8], m2[8], k0[8], k1[8], k2[8];
        c8_ld8(taddr, m0);
        c8_ld8(taddr + 8, m1);
        c8_ld8(taddr + 16, m2);
        c8_ld8(taddr + 24, k0);
        c8_ld8(taddr + 32, k1);
        c8_ld8(taddr + 40, k2);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(t_empty + g);
        float out[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float e0 = fmaf(k0[c], c1, m0[c] * c0);
          m1[c] = fmaf(k1[c], c1, m1[c] * c0);
          m2[c] = fmaf(k2[c], c1, m2[c] * c0);
          const float s1 = __shfl_down_sync(0xffffffffu, m1[c], 1);
          const float s2 = __shfl_down_sync(0xffffffffu, m2[c], 2);
          out[c] = e0 + (lane < 31 ? s1 : 0.f) + (lane < 30 ? s2 : 0.f);
        }
        if (lane < 2) {
          float4* d2 = reinterpret_cast<float4*>(xq + (1 + lane) * 8);
          d2[0] = make_float4(m2[0], m2[1], m2[2], m2[3]), d2[1] = make_float4(m2[4], m2[5], m2[6], m2[7]);
          if (lane == 0) {
            float4* d1 = reinterpret_cast<float4*>(xq);
            d1[0] = make_float4(m1[0], m1[1], m1[2], m1[3]), d1[1] = make_float4(m1[4], m1[5], m1[6], m1[7]);
          }
        }
        // the bulk stores of this group's previous tile must have read the staging buffer
        if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        named_bar_sync(bar_id, 128);
        if (q < 3 && lane >= 30) {
          if (lane == 31) {
            const float4* p1 = reinterpret_cast<const float4*>(xn);
            const float4 u = p1[0], v = p1[1];
            out[0] += u.x, out[1] += u.y, out[2] += u.z, out[3] += u.w, out[4] += v.x, out[5] += v.y, out[6] += v.z, out[7] += v.w;
          }
          const float4* p2 = reinterpret_cast<const float4*>(xn + (1 + lane - 30) * 8);
          const float4 u = p2[0], v = p2[1];
          out[0] += u.x, out[1] += u.y, out[2] += u.z, out[3] += u.w, out[4] += v.x, out[5] += v.y, out[6] += v.z, out[7] += v.w;
        }
        // voxel coordinates -> border
        const int r = w.row0 + n * a.Wp + j;
        int line, x, dpl, y;
        fdivmod(r, a.fWp, line, x);
        fdivmod(line, a.fHp, dpl, y);
        const bool border = x < 1 || x > a.W || y < 1 || y > a.H || dpl < 1 || dpl > a.D;  // tiles may spill into the pad plane
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          float v0 = fmaxf(out[2 * p] + bias[2 * p], 0.f), v1 = fmaxf(out[2 * p + 1] + bias[2 * p + 1], 0.f);
          v0 = border ? 0.f : v0, v1 = border ? 0.f : v1;
          const __half2 h = f2h2_sat(v0, v1);
          const float2 f = __half22float2(h);
          const __half2 l = f2h2_sat((v0 - f.x) * 2048.f, (v1 - f.y) * 2048.f);
          hi[p] = *reinterpret_cast<const uint32_t*>(&h), lo[p] = *reinterpret_cast<const uint32_t*>(&l);
        }
        *reinterpret_cast<uint4*>(stage + j * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(stage + 2048 + j * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        fence_proxy_async_smem();
        named_bar_sync(bar_id, 128);
        if (issuer) {
          const long long v = (long long)w.b * a.vox_b + w.row0 + (long long)n * a.Wp;
          bulk_store(a.out_hi + v * 16, stage, 126 * 16);
          bulk_store(a.out_lo + v * 16, stage + 2048, 126 * 16);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    }
    if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(C8_NGRP * 64));
}

// ================================================================================================================================
// Plane-group version: one step = the tiles (dp0 + i, y, ct), i < L <= 3, of THREE neighbouring d planes at once.
// The kernel above reads every input line three times (as kd = 0, 1, 2 of three different output planes: 12 KB of shared-memory
// fills and 9 MMAs per 4 KB of output, ~6 TB/s of L2 -> SM traffic) and is bound by that.  Here a line group is the same line of
// the L + 2 planes dp0 - 1 .. dp0 + L (10 bulk copies per step = 1.67 per output tile instead of 6), and the kd taps are folded
// into N next to kw: with the operand table of a kh tap stored as [W(kd=2); W(kd=1); W(kd=0)] (3 x 48 rows), input plane
// t = -1 .. L feeds the accumulators i = max(0, t-1) .. min(L-1, t+1), which are neighbouring column blocks of TMEM, through ONE
// MMA whose B operand is a row sub-range of that table:  N = 48, 96, 144, 96, 48 for L = 3  ->  15 MMAs / 1248 issue cycles per
// three tiles instead of 27 / 1728.  The plane that feeds all L accumulators goes first with accumulate = 0.
// Tiles are numbered k = CP_L * step + i and handled round robin by the four epilogue groups; TMEM holds two steps (2 x 192 columns).
constexpr int CP_L = 3;                       // d planes per group.  The kernel is generic up to 5 (CP_L = 5, CP_NG = 5, CP_TCOLS = 256,
                                              // CP_NT = 2: D = 9 as groups of 5 + 4, 13 input planes per 9 outputs instead of 15) --
                                              // measured equal within noise (stage-3 stack 434 vs 442 us, stage-2 165 vs 159 per 4 pairs);
                                              // re-measured after the zero planes stopped coming from DRAM: 350 instead of 416 MB of
                                              // DRAM reads per stage-3 layer but 125 instead of 118 us (two step buffers in TMEM)
constexpr int CP_NG = 6;                      // ring of line groups
constexpr int CP_GBYTES = (CP_L + 2) * 4096;  // (L + 2) planes x (hi, lo) x 128 voxels x 16 B
constexpr int CP_OFF_RING = 14336;
constexpr int CP_OFF_STAGE = CP_OFF_RING + CP_NG * CP_GBYTES;
constexpr int CP_OFF_XCH = CP_OFF_STAGE + C8_NGRP * 4096;
constexpr int CP_OFF_EX = CP_OFF_XCH + C8_NGRP * 4 * 3 * 8 * 4;   // [NGRP groups][4 planes: E1 lo/hi, E2 lo/hi channels][128 rows] float4
constexpr int CP_OFF_BAR = CP_OFF_EX + C8_NGRP * 4 * 128 * 16;
constexpr int CP_SMEM = CP_OFF_BAR + 256 + 128;
constexpr int CP_TCOLS = 160;                 // TMEM columns per step buffer (CP_L x 48 = 144 used)
constexpr int CP_NT = 3;                      // step buffers in TMEM (3 x 160 <= 512 columns)

struct CPArgs {
  C8Args c;       // tensors, sizes, fast divisors (line0 / strip_len / total_tiles / ct_per_line unused)
  int ct, ndg;    // column tiles per line, d groups
  int ch;         // lines per chunk of the work order
  int total_steps;
  FastDiv fcolsteps, fct;  // steps per (b, column tile) = ndg * Hp
};
struct CPItem {
  int b, L, row0, nsteps;  // row0: first output voxel (inside the batch element) of accumulator 0 of step 0
  int dp0, y0, x0;         // its plane, line and column
};
// Work order: (b, column tile) -> chunks of `ch` lines -> d group -> line; every CTA takes one contiguous range of it (a range
// start or a new (chunk, d group) costs two warm-up line groups).  Neighbouring d groups share two of their five input planes.
// Small chunks turn that re-read into an L2 hit (measured on 8 KITTI pairs: 482 MB of DRAM reads with ch = Hp, 378 MB with
// ch = 4) but the kernel is bound by its epilogue's instruction issue, not by DRAM (176 us with ch = Hp, 190 us with ch = 4),
// so the default is ch = Hp: plain strips.
struct CPSched {
  int g, g1;
  __device__ __forceinline__ CPSched(const CPArgs& a) {
    g = (int)((long long)a.total_steps * blockIdx.x / gridDim.x);
    g1 = (int)((long long)a.total_steps * (blockIdx.x + 1) / gridDim.x);
  }
  __device__ __forceinline__ bool next(const CPArgs& a, CPItem& it) {
    if (g >= g1) return false;
    int col, s, cti;
    fdivmod(g, a.fcolsteps, col, s);          // col = b * ct + cti;  s = step inside the column: chunk-major
    fdivmod(col, a.fct, it.b, cti);
    const int yc = s / (a.ndg * a.ch);         // only the last chunk is short, so full-chunk strides locate every chunk
    const int rem = s - yc * a.ndg * a.ch;
    const int chl = min(a.ch, a.c.Hp - yc * a.ch);
    const int dg = rem / chl, yy = rem - dg * chl;
    const int y0 = yc * a.ch + yy;
    it.nsteps = min(chl - yy, g1 - g);
    it.L = min(CP_L, a.c.D - CP_L * dg);
    it.dp0 = 1 + CP_L * dg, it.y0 = y0, it.x0 = cti * 126;
    it.row0 = (it.dp0 * a.c.Hp + y0) * a.c.Wp + it.x0;
    g += it.nsteps;
    return true;
  }
};

// MMAs of one line group (one kh) for L accumulators.  Input plane t = -1 .. L feeds accumulators max(0, t-1) .. min(L-1, t+1).
// INIT (first kh of a step): the planes t = 1, 4, ... cover disjoint accumulator sets that together are all of them; they go first
// and overwrite, everything else accumulates.
template <uint32_t NB>
__device__ __forceinline__ void cp_mma(uint32_t tmem_buf, uint32_t g_lo, uint32_t b_kh, uint64_t a_hi, uint64_t b_hi, int t, int L,
                                       uint32_t acc) {
  const int i_min = t - 1 > 0 ? t - 1 : 0, i_max = t + 1 < L - 1 ? t + 1 : L - 1;
  const uint32_t N = NB * (uint32_t)(i_max - i_min + 1);
  const uint32_t idesc = (1u << 4) | ((N >> 3) << 17) | ((128u >> 4) << 24);
  const uint64_t da = a_hi | (uint64_t)(g_lo + (uint32_t)(t + 1) * (4096 >> 4));
  const uint64_t db = b_hi | (uint64_t)(b_kh + (uint32_t)(1 - t + i_min) * NB);  // row start (16-byte units)
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_buf + (uint32_t)i_min * NB),
      "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
template <int L, uint32_t NB>
__device__ __forceinline__ void cp_issue_group(uint32_t tmem_buf, uint32_t g_lo, uint32_t b_kh, uint64_t a_hi, uint64_t b_hi, bool init) {
#pragma unroll
  for (int t = 1; t <= L; t += 3) cp_mma<NB>(tmem_buf, g_lo, b_kh, a_hi, b_hi, t, L, init ? 0u : 1u);
#pragma unroll
  for (int t = -1; t <= L; ++t)
    if (!(t >= 1 && (t - 1) % 3 == 0)) cp_mma<NB>(tmem_buf, g_lo, b_kh, a_hi, b_hi, t, L, 1u);
}

template <bool LAST>
__global__ void __launch_bounds__(C8_THREADS, 1) conv3d_c8p_kernel(const CPArgs pa) {
  const C8Args& a = pa.c;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sB = smem;
  uint8_t* sRing = smem + CP_OFF_RING;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + CP_OFF_BAR);
  uint64_t* g_full = bars;              // [NG]
  uint64_t* g_empty = g_full + CP_NG;   // [NG]
  uint64_t* t_full = g_empty + CP_NG;   // [NT]
  uint64_t* t_empty = t_full + CP_NT;   // [NT]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + CP_NT);
  constexpr uint32_t NB = LAST ? 16 : 48;  // accumulator columns per output tile = B rows per kd

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < CP_NG; ++i) mbar_init(g_full + i, 1), mbar_init(g_empty + i, 1);
    for (int i = 0; i < CP_NT; ++i) mbar_init(t_full + i, 1), mbar_init(t_empty + i, CP_L * 4);
    mbar_fence_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // operand table: global block (kd*3 + kh) = [chunk 0: NB rows x 16 B][chunk 1]  ->  per kh: [chunk][(2 - kd) * NB + row]
  for (int u = tid; u < 9 * (int)NB * 2; u += C8_THREADS) {
    const int blk = u / ((int)NB * 2), rem = u - blk * (int)NB * 2;
    const int c = rem / (int)NB, r = rem - c * (int)NB;
    const int kd = blk / 3, kh = blk - kd * 3;
    reinterpret_cast<uint4*>(sB)[kh * 3 * (int)NB * 2 + c * 3 * (int)NB + (2 - kd) * (int)NB + r] =
        __ldg(reinterpret_cast<const uint4*>(a.wtab) + u);
  }
  fence_proxy_async_smem();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  const int plane = a.Hp * a.Wp;  // voxels per d plane

  if (warp == 0) {
    // ================================ producer: lanes 0 .. 2(L+2)-1 issue one 2 KB bulk copy each ================================
    const int pl = lane >> 1, part = lane & 1;
    const uint8_t* src_plane = part ? a.in_lo : a.in_hi;
    uint32_t slot = 0, ph = 0;
    CPSched sched(pa);
    CPItem w;
    while (sched.next(pa, w)) {
      const long long base = (long long)w.b * a.vox_b + w.row0 - 1 + (long long)(pl - 1) * plane;  // GEMM row 0 (Toeplitz shift 1)
      // the two zero planes that pad d (2 of the 15 plane reads of a D = 9 volume in groups of 3) are not streamed from DRAM: every
      // copy of them reads the same 2 KB of the zeroed slack in front of the tensor, an L2 hit (the kernel runs at the DRAM roofline)
      const int dpa = w.dp0 + pl - 1;
      const bool zero_plane = dpa <= 0 || dpa > a.D;
      for (int n = 0; n < w.nsteps; ++n) {
        for (int kh = n == 0 ? 0 : 2; kh < 3; ++kh) {  // later steps of a strip only need the line below
          mbar_wait(g_empty + slot, ph ^ 1);
          if (lane == 0) mbar_expect_tx(g_full + slot, (uint32_t)(w.L + 2) * 4096);
          __syncwarp();
          if (lane < 2 * (w.L + 2))
            bulk_load(sRing + slot * CP_GBYTES + pl * 4096 + part * 2048,
                      zero_plane ? src_plane - 4096 : src_plane + (base + (long long)(n + kh - 1) * a.Wp) * 16, 2048, g_full + slot);
          if (++slot == CP_NG) slot = 0, ph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer (whole schedule inside one elected lane) ================================
    const uint64_t a_hi = ((uint64_t)(2048 >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46);
    const uint64_t b_hi = ((uint64_t)((3 * NB * 16) >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46);
    const uint32_t ring_lo = (smem_u32(sRing) & 0x3FFFF) >> 4, b_lo = (smem_u32(sB) & 0x3FFFF) >> 4;
    if (elect_one_sync()) {
      uint32_t bslot = 0, bph = 0, st = 0;
      CPSched sched(pa);
      CPItem w;
      while (sched.next(pa, w)) {
        for (int n = 0; n < w.nsteps; ++n, ++st) {
          const bool last = n == w.nsteps - 1;
          const uint32_t tb = st % CP_NT;
          mbar_wait(t_empty + tb, ((st / CP_NT) & 1) ^ 1);
          uint32_t slot = bslot, ph = bph;
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
            if (n == 0 || kh == 2) mbar_wait(g_full + slot, ph);  // the other groups were waited for by the previous step
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t g_lo = ring_lo + slot * (CP_GBYTES >> 4);
            const uint32_t b_kh = b_lo + (uint32_t)kh * ((3 * NB * 32) >> 4);
            const uint32_t tbuf = tmem + tb * CP_TCOLS;
            if (w.L == 5) cp_issue_group<5, NB>(tbuf, g_lo, b_kh, a_hi, b_hi, kh == 0);
            else if (w.L == 4) cp_issue_group<4, NB>(tbuf, g_lo, b_kh, a_hi, b_hi, kh == 0);
            else if (w.L == 3) cp_issue_group<3, NB>(tbuf, g_lo, b_kh, a_hi, b_hi, kh == 0);
            else if (w.L == 2) cp_issue_group<2, NB>(tbuf, g_lo, b_kh, a_hi, b_hi, kh == 0);
            else cp_issue_group<1, NB>(tbuf, g_lo, b_kh, a_hi, b_hi, kh == 0);
            if (kh == 0 || last)
              asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(g_empty + slot))
                           : "memory");
            if (kh == 2)
              asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(t_full + tb))
                           : "memory");
            if (++slot == CP_NG) slot = 0, ph ^= 1;
          }
          if (++bslot == CP_NG) bslot = 0, bph ^= 1;
        }
        bslot += 2;  // the strip's last step consumed its remaining two groups
        if (bslot >= CP_NG) bslot -= CP_NG, bph ^= 1;
      }
    }
    __syncwarp();
  } else {
    // ================================ epilogue: group g handles the tiles k = 3 * step + i with k % 4 == g ================================
    const int g = (warp - 2) >> 2;
    const int q = warp & 3;            // TMEM lane quarter
    const int j = q * 32 + lane;       // GEMM row; this thread produces output voxel j (staging row j), valid for j < 126
    const bool issuer = ((warp - 2) & 3) == 0 && lane == 0;
    const float* scl = reinterpret_cast<const float*>(a.wtab + (LAST ? 9 * 512 : C8_BBYTES));
    const float c0 = __ldg(scl) * (LAST ? 1.f / kDwsepActScale : 1.f), c1 = __ldg(scl + 1) * (LAST ? 1.f / kDwsepActScale : 1.f);
    float bias[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) bias[c] = __ldg(a.bias + c) * kDwsepActScale;
    uint8_t* stage = smem + CP_OFF_STAGE + g * 4096;
    float* xch = reinterpret_cast<float*>(smem + CP_OFF_XCH) + g * (4 * 3 * 8);
    float* xq = xch + q * 24;                    // this quarter publishes e1 of lane 0, e2 of lanes 0 and 1
    const float* xn = xch + ((q + 1) & 3) * 24;  // next quarter's
    const int bar_id = 1 + g;
    uint32_t st0 = 0, mine = 0;  // first step of the current item; tiles this group has handled (double-buffers the LAST exchange rows)
    CPSched sched(pa);
    CPItem w;
    while (sched.next(pa, w)) {
      // this thread's voxel column and line offset are the same for every step of the item (a tile may wrap over line ends)
      int xj, carry;
      fdivmod(w.x0 + j, a.fWp, carry, xj);
      const bool xborder = xj < 1 || xj > a.W || j >= 126;
      const int yj = w.y0 + carry;  // line of this thread's voxel at step 0; >= Hp: next plane (never stored from here)
      // this group's tiles of the item: k = 3 * step + i with k % 4 == g, visited directly (no skipped iterations)
      const uint32_t k_begin = st0 * CP_L, k_end = (st0 + (uint32_t)w.nsteps) * CP_L;
#pragma unroll 1
      for (uint32_t k = k_begin + (((uint32_t)g - k_begin) & 3u); k < k_end; k += 4) {
        {
          const uint32_t st = k / CP_L;
          const int i = (int)(k - st * CP_L), n = (int)(st - st0);
          const uint32_t tb = st % CP_NT;
          // LAST: the output index and the skip value do not depend on the accumulator: fetch the skip before waiting for it
          long long o_last = -1;
          float sk = 0.f;
          if (LAST && i < w.L) {
            const int y = yj + n;
            if (!xborder && y >= 1 && y <= a.H) {
              o_last = (((long long)w.b * a.D + (w.dp0 + i - 1)) * a.H + (y - 1)) * a.W + (xj - 1);
              if (a.skip) sk = __ldg(a.skip + o_last);
            }
          }
          mbar_wait(t_full + tb, (st / CP_NT) & 1);
          if (i >= w.L) {  // no such plane in this d group: only release the buffer
            __syncwarp();
            if (lane == 0) mbar_arrive(t_empty + tb);
            continue;
          }
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + tb * CP_TCOLS + (uint32_t)i * NB;
          const int r = w.row0 + n * a.Wp + i * plane + j;  // output voxel (inside the batch element) of this thread
          if (LAST) {
            float m[8], k[8];
            c8_ld8(taddr, m);
            c8_ld8(taddr + 8, k);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(t_empty + tb);
            const float e0 = fmaf(k[0], c1, m[0] * c0), e1 = fmaf(k[1], c1, m[1] * c0), e2 = fmaf(k[2], c1, m[2] * c0);
            const float s1 = __shfl_down_sync(0xffffffffu, e1, 1), s2 = __shfl_down_sync(0xffffffffu, e2, 2);
            float v = e0 + (lane < 31 ? s1 : 0.f) + (lane < 30 ? s2 : 0.f);
            float* xp = xch + (mine & 1) * 12;  // this group's exchange rows, double-buffered by its tile parity
            ++mine;
            if (lane == 0) xp[q * 3] = e1;
            if (lane < 2) xp[q * 3 + 1 + lane] = e2;
            named_bar_sync(bar_id, 128);
            if (q < 3 && lane >= 30) {
              if (lane == 31) v += xp[(q + 1) * 3];
              v += xp[(q + 1) * 3 + 1 + lane - 30];
            }
            // a tile that runs past the end of its plane (y > H) would produce voxels of the next plane, which that plane's own
            // tile accumulates in a different kd order (other accumulator index): leave them to their owner so results are unique
            if (o_last >= 0) a.out_f32[o_last] = v + sk;
            continue;
          }
          float m0[8], m1[8], m2[8], k0[8], k1[8], k2[8];
          c8_ld8(taddr, m0);
          c8_ld8(taddr + 8, m1);
          c8_ld8(taddr + 16, m2);
          c8_ld8(taddr + 24, k0);
          c8_ld8(taddr + 32, k1);
          c8_ld8(taddr + 40, k2);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(t_empty + tb);
          // E = c0 * (main + corr * 2^-11) with c0 a power of two: the scale commutes with every rounding below, so it is applied
          // once, together with the bias, after the shifted sum (bit-identical to scaling each E)
          // The Toeplitz shift out[r] = E0[r] + E1[r+1] + E2[r+2] goes through shared memory (4 conflict-free float4 planes per
          // group): 4 stores + 4 loads per thread instead of 16 shuffles + 16 selects and a separate cross-quarter exchange.
          float out[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            out[c] = fmaf(k0[c], 0x1p-11f, m0[c]);
            m1[c] = fmaf(k1[c], 0x1p-11f, m1[c]);
            m2[c] = fmaf(k2[c], 0x1p-11f, m2[c]);
          }
          float4* ex = reinterpret_cast<float4*>(smem + CP_OFF_EX) + g * (4 * 128);
          ex[j] = make_float4(m1[0], m1[1], m1[2], m1[3]), ex[128 + j] = make_float4(m1[4], m1[5], m1[6], m1[7]);
          ex[256 + j] = make_float4(m2[0], m2[1], m2[2], m2[3]), ex[384 + j] = make_float4(m2[4], m2[5], m2[6], m2[7]);
          // the bulk stores of this group's previous tile must have read the staging buffer
          if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          named_bar_sync(bar_id, 128);
          if (j < 126) {  // rows 126, 127 are the halo of the next tile
            const float4 a0 = ex[j + 1], a1 = ex[128 + j + 1], b0 = ex[256 + j + 2], b1 = ex[384 + j + 2];
            out[0] = out[0] + a0.x + b0.x, out[1] = out[1] + a0.y + b0.y, out[2] = out[2] + a0.z + b0.z, out[3] = out[3] + a0.w + b0.w;
            out[4] = out[4] + a1.x + b1.x, out[5] = out[5] + a1.y + b1.y, out[6] = out[6] + a1.z + b1.z, out[7] = out[7] + a1.w + b1.w;
          }
          const bool border = xborder || yj + n < 1 || yj + n > a.H;  // beyond the plane end: not stored from this tile anyway
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            float v0 = fmaxf(fmaf(out[2 * p], c0, bias[2 * p]), 0.f), v1 = fmaxf(fmaf(out[2 * p + 1], c0, bias[2 * p + 1]), 0.f);
            v0 = border ? 0.f : v0, v1 = border ? 0.f : v1;
            const __half2 h = f2h2_sat(v0, v1);
            const float2 f = __half22float2(h);
            // (v - hi) * 2^11 on channel pairs (FADD2 + FMUL2)
            const float2 d = split_lo2(v0, v1, f);
            const __half2 l = f2h2_sat(d.x, d.y);
            hi[p] = *reinterpret_cast<const uint32_t*>(&h), lo[p] = *reinterpret_cast<const uint32_t*>(&l);
          }
          *reinterpret_cast<uint4*>(stage + j * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(stage + 2048 + j * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          fence_proxy_async_smem();
          named_bar_sync(bar_id, 128);
          if (issuer) {
            const int r0 = r - j;  // the tile's first voxel
            const long long v = (long long)w.b * a.vox_b + r0;
            // stop at the end of the tile's own plane: the voxels beyond belong to the next plane's tile, which accumulates its
            // kd taps in a different order (other accumulator index) -- one writer per voxel keeps the result unique
            const uint32_t nv = (uint32_t)min(126, (w.dp0 + i + 1) * plane - r0);
            bulk_store(a.out_hi + v * 16, stage, nv * 16);
            bulk_store(a.out_lo + v * 16, stage + 2048, nv * 16);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
      }
      st0 += (uint32_t)w.nsteps;
    }
    if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// ---- first conv 1 -> 8 on the raw cost (BN_0 affine + ReLU on the taps), writes every voxel of the padded hi/lo planes ----
// One thread = NV voxels that are neighbours in y (same padded x): the 27-tap window of the NV voxels is (NV + 2) rows x 3 x 3, so
// the cost loads (coalesced along x) and the broadcast weight loads from shared memory are shared NV ways (NV = 8: 90 + 54 per 1728
// FFMAs) and the 16-byte voxel stores of a warp are contiguous.  Grid = (pair x padded plane, NV-row group, 128-voxel x segment).
// 8 resident blocks (64 registers, ~50 bytes of spills): the kernel is bound by the latency of its predicated global tap loads, and
// 32 instead of 20 resident warps take it from 223 to 191 us per 8 pairs (6 blocks: 217; 10 blocks / 48 registers spill heavily: 345)
template <int NV>
__global__ void __launch_bounds__(128, NV == 4 ? 8 : 3)
    conv3d_first_c8_kernel(const float* __restrict__ cost, const float* __restrict__ w /*[27][8]*/, const float* __restrict__ bias,
                           const float* __restrict__ affine, uint4* __restrict__ out_hi, uint4* __restrict__ out_lo, int D, int H,
                           int W) {
  __shared__ __align__(16) float sW[27 * 8];
  for (int i = threadIdx.x; i < 27 * 8; i += blockDim.x) sW[i] = __ldg(w + i);
  __syncthreads();
  const int Wp = W + 2, Hp = H + 2, Dp = D + 2;
  const int xp = blockIdx.z * 128 + threadIdx.x;
  if (xp >= Wp) return;
  const int b = blockIdx.x / Dp, dp = blockIdx.x - b * Dp;
  const int yp0 = blockIdx.y * NV;
  const int x = xp - 1, d = dp - 1;
  const long long hw = (long long)H * W;
  const long long vox0 = (((long long)b * Dp + dp) * Hp + yp0) * Wp + xp;
  const bool zero_all = d < 0 || d >= D || x < 0 || x >= W;
  // accumulators as float2 pairs: FFMA2 (fma.rn.f32x2, sm_100) does two output channels per issued instruction -- the kernel is
  // bound by instruction issue (45 % of its instructions were scalar FFMAs), not by the FP32 pipe
  float2 acc[NV][4];
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[v][j] = make_float2(0.f, 0.f);
  if (!zero_all) {
    const float s0 = __ldg(affine), t0 = __ldg(affine + 1);
    const float* cb = cost + (long long)b * D * hw;  // 32-bit offsets inside one pair's volume (host checks D*H*W < 2^31)
    const int ihw = H * W;
    int offr[NV + 2];
    bool okr[NV + 2];
#pragma unroll
    for (int r = 0; r < NV + 2; ++r) {
      const int yy = yp0 + r - 2;  // input row of (voxel vy, tap kh) with vy + kh = r
      okr[r] = (unsigned)yy < (unsigned)H;
      offr[r] = d * ihw + yy * W + x;
    }
#pragma unroll
    for (int kd = 0; kd < 3; ++kd) {
      const bool okd = (unsigned)(d + kd - 1) < (unsigned)D;  // block-uniform
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const bool okx = okd && (unsigned)(x + kw - 1) < (unsigned)W;
        const int tap = (kd - 1) * ihw + (kw - 1);
        float v[NV + 2];
#pragma unroll
        for (int r = 0; r < NV + 2; ++r) {
          const bool ok = okx && okr[r];
          const float c = ok ? __ldg(cb + (offr[r] + tap)) : 0.f;
          v[r] = ok ? fmaxf(fmaf(c, s0, t0), 0.f) : 0.f;
        }
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
          const float4 wa = *reinterpret_cast<const float4*>(sW + (kd * 9 + kh * 3 + kw) * 8);
          const float4 wb = *reinterpret_cast<const float4*>(sW + (kd * 9 + kh * 3 + kw) * 8 + 4);
#pragma unroll
          const float2 w01 = make_float2(wa.x, wa.y), w23 = make_float2(wa.z, wa.w);
          const float2 w45 = make_float2(wb.x, wb.y), w67 = make_float2(wb.z, wb.w);
#pragma unroll
          for (int vy = 0; vy < NV; ++vy) {
            const float2 t = make_float2(v[vy + kh], v[vy + kh]);
            acc[vy][0] = __ffma2_rn(t, w01, acc[vy][0]), acc[vy][1] = __ffma2_rn(t, w23, acc[vy][1]);
            acc[vy][2] = __ffma2_rn(t, w45, acc[vy][2]), acc[vy][3] = __ffma2_rn(t, w67, acc[vy][3]);
          }
        }
      }
    }
  }
  float bv[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) bv[j] = __ldg(bias + j);
#pragma unroll
  for (int vy = 0; vy < NV; ++vy) {
    const int yp = yp0 + vy;
    if (yp >= Hp) break;
    const bool border = zero_all || yp == 0 || yp == Hp - 1;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const float a0 = border ? 0.f : fmaxf(acc[vy][p].x + bv[2 * p], 0.f) * kDwsepActScale;
      const float a1 = border ? 0.f : fmaxf(acc[vy][p].y + bv[2 * p + 1], 0.f) * kDwsepActScale;
      const __half2 h = f2h2_sat(a0, a1);
      const float2 f = __half22float2(h);
      const float2 dl = split_lo2(a0, a1, f);
      const __half2 l = f2h2_sat(dl.x, dl.y);
      hi[p] = *reinterpret_cast<const uint32_t*>(&h), lo[p] = *reinterpret_cast<const uint32_t*>(&l);
    }
    out_hi[vox0 + (long long)vy * Wp] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    out_lo[vox0 + (long long)vy * Wp] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

// ---- first conv 1 -> 8, version 2: the activated tap window staged once in shared memory --------------------------------------
// Version 1 above fetches its 54 taps per thread straight from global memory: 3 address instructions + 1 predicate per load and
// the BN_0 affine + ReLU recomputed for every use (ncu: 28 % of its instructions are the FFMA2s that do the work).  Here a block
// (one padded output plane dp, NV rows, 128 columns) first stages ReLU(BN_0(cost)) of the 3 x (NV + 2) x 130 window, zero outside
// the volume (the conv pads the ACTIVATED tensor), so the taps are immediate-offset LDS and the activation is applied once per
// element.  Thread = one column x NV rows x 8 channels (channel pairs through FFMA2).
template <int NV>
__global__ void __launch_bounds__(128, NV == 8 ? 4 : 6)
    conv3d_first_c8_v2_kernel(const float* __restrict__ cost, const float* __restrict__ w /*[27][8]*/, const float* __restrict__ bias,
                              const float* __restrict__ affine, uint4* __restrict__ out_hi, uint4* __restrict__ out_lo, int D, int H,
                              int W) {
  constexpr int TR = NV + 2, TC = 130;
  __shared__ __align__(16) float sW[27 * 8];
  __shared__ float sIn[3 * TR * TC];
  const int Wp = W + 2, Hp = H + 2, Dp = D + 2;
  const int b = blockIdx.x / Dp, dp = blockIdx.x - b * Dp;
  const int yp0 = blockIdx.y * NV;
  const int xp0 = blockIdx.z * 128;
  const int tid = threadIdx.x;
  const int xp = xp0 + tid;
  const int d = dp - 1;
  const long long vox0 = (((long long)b * Dp + dp) * Hp + yp0) * Wp + xp;
  if (d < 0 || d >= D) {  // padding plane: zeros only (block-uniform)
    if (xp < Wp)
      for (int vy = 0; vy < NV && yp0 + vy < Hp; ++vy)
        out_hi[vox0 + (long long)vy * Wp] = make_uint4(0, 0, 0, 0), out_lo[vox0 + (long long)vy * Wp] = make_uint4(0, 0, 0, 0);
    return;
  }
  for (int i = tid; i < 27 * 8; i += 128) sW[i] = __ldg(w + i);
  {
    const float s0 = __ldg(affine), t0 = __ldg(affine + 1);
    const long long hw = (long long)H * W;
    const float* cb = cost + (long long)b * D * hw;
    // tile element (kd, r, c) = activated cost at plane d + kd - 1, row yp0 + r - 2, column xp0 + c - 2; a thread stages column
    // c = tid of every (kd, r) row (threads 0 and 1 also the two halo columns 128, 129): no index arithmetic per element
    const int xx0 = xp0 + tid - 2, xx1 = xp0 + 128 + tid - 2;
    const bool okx0 = (unsigned)xx0 < (unsigned)W, okx1 = tid < 2 && (unsigned)xx1 < (unsigned)W;
#pragma unroll
    for (int kd = 0; kd < 3; ++kd) {
      const int dd = d + kd - 1;
      const bool okd = (unsigned)dd < (unsigned)D;
      const float* cd = cb + (long long)dd * hw;
#pragma unroll
      for (int r = 0; r < TR; ++r) {
        const int yy = yp0 + r - 2;
        const bool ok = okd && (unsigned)yy < (unsigned)H;  // block-uniform
        const float* row = cd + (long long)yy * W;
        sIn[(kd * TR + r) * TC + tid] = (ok && okx0) ? fmaxf(fmaf(__ldg(row + xx0), s0, t0), 0.f) : 0.f;
        if (tid < 2) sIn[(kd * TR + r) * TC + 128 + tid] = (ok && okx1) ? fmaxf(fmaf(__ldg(row + xx1), s0, t0), 0.f) : 0.f;
      }
    }
  }
  __syncthreads();
  if (xp >= Wp) return;
  float2 acc[NV][4];
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[v][j] = make_float2(0.f, 0.f);
  const int x = xp - 1;
  const bool zero_all = x < 0 || x >= W;
  if (!zero_all) {
#pragma unroll
    for (int kd = 0; kd < 3; ++kd) {
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        float v[TR];
#pragma unroll
        for (int r = 0; r < TR; ++r) v[r] = sIn[(kd * TR + r) * TC + tid + kw];  // column xp0 + tid + kw - 2 = x + kw - 1
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
          const float4 wa = *reinterpret_cast<const float4*>(sW + (kd * 9 + kh * 3 + kw) * 8);
          const float4 wb = *reinterpret_cast<const float4*>(sW + (kd * 9 + kh * 3 + kw) * 8 + 4);
          const float2 w01 = make_float2(wa.x, wa.y), w23 = make_float2(wa.z, wa.w);
          const float2 w45 = make_float2(wb.x, wb.y), w67 = make_float2(wb.z, wb.w);
#pragma unroll
          for (int vy = 0; vy < NV; ++vy) {
            const float2 t = make_float2(v[vy + kh], v[vy + kh]);
            acc[vy][0] = __ffma2_rn(t, w01, acc[vy][0]), acc[vy][1] = __ffma2_rn(t, w23, acc[vy][1]);
            acc[vy][2] = __ffma2_rn(t, w45, acc[vy][2]), acc[vy][3] = __ffma2_rn(t, w67, acc[vy][3]);
          }
        }
      }
    }
  }
  float bv[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) bv[j] = __ldg(bias + j);
#pragma unroll
  for (int vy = 0; vy < NV; ++vy) {
    const int yp = yp0 + vy;
    if (yp >= Hp) break;
    const bool border = zero_all || yp == 0 || yp == Hp - 1;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const float a0 = border ? 0.f : fmaxf(acc[vy][p].x + bv[2 * p], 0.f) * kDwsepActScale;
      const float a1 = border ? 0.f : fmaxf(acc[vy][p].y + bv[2 * p + 1], 0.f) * kDwsepActScale;
      const __half2 h = f2h2_sat(a0, a1);
      const float2 f = __half22float2(h);
      const float2 dl = split_lo2(a0, a1, f);
      const __half2 l = f2h2_sat(dl.x, dl.y);
      hi[p] = *reinterpret_cast<const uint32_t*>(&h), lo[p] = *reinterpret_cast<const uint32_t*>(&l);
    }
    out_hi[vox0 + (long long)vy * Wp] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    out_lo[vox0 + (long long)vy * Wp] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

// ---- host -------------------------------------------------------------------------------------------------------------------
static long long c8_plane_bytes(int B, int D, int H, int W) {
  const long long vox = (long long)B * (D + 2) * (H + 2) * (W + 2);
  const long long slack = (long long)(H + 2) * (W + 2) + (W + 2) + 256;  // one plane + one line + one tile before and after
  return (vox + 2 * slack) * 16;
}
size_t conv3d_c8_workspace_bytes(int B, int D, int H, int W) { return (size_t)4 * ((c8_plane_bytes(B, D, H, W) + 255) / 256 * 256); }

// wtab[l]: 9 x 1536 B operand table + scales[2]; bias_mid[l]: [8]; w_last_tab: 9 x 512 B table of the closing conv + scales[2]
int conv3d_stack_c8(const float* cost, const float* affine, const float* w_first, const float* b_first, const float* const* wtab,
                    const float* const* bias_mid, int layers, const float* w_last_tab, float* out, void* ws, int B, int D, int H,
                    int W, int add_skip, cudaStream_t st) {
  const int Wp = W + 2, Hp = H + 2, Dp = D + 2;
  const long long vox_b = (long long)Dp * Hp * Wp;
  if (vox_b * B >= (1ll << 31) - (1 << 20) || (long long)D * H * W >= (1ll << 31)) return LWS_ERR_BAD_SHAPE;
  const long long pb = (c8_plane_bytes(B, D, H, W) + 255) / 256 * 256;
  const long long slack = ((long long)Hp * Wp + Wp + 256) * 16;
  uint8_t* base = (uint8_t*)ws;
  uint8_t* plane[4];  // A.hi, A.lo, B.hi, B.lo (voxel 0)
  for (int i = 0; i < 4; ++i) plane[i] = base + i * pb + slack;
  cudaError_t e;
  // the slack in front of / behind every plane and the d-padding planes of the second buffer pair are only ever read
  for (int i = 0; i < 4; ++i) {
    if ((e = cudaMemsetAsync(base + i * pb, 0, slack, st)) != cudaSuccess) return (int)e;
    if ((e = cudaMemsetAsync(plane[i] + vox_b * B * 16, 0, slack, st)) != cudaSuccess) return (int)e;
  }
  for (int i = 2; i < 4; ++i) {
    const size_t pl = (size_t)Hp * Wp * 16;
    if ((e = cudaMemset2DAsync(plane[i], vox_b * 16, 0, pl, B, st)) != cudaSuccess) return (int)e;
    if ((e = cudaMemset2DAsync(plane[i] + (size_t)(Dp - 1) * pl, vox_b * 16, 0, pl, B, st)) != cudaSuccess) return (int)e;
  }
  {
    // C = 8: the global-memory version stays the default (measured r02i, 8 pairs: 52 + 178 us against 61 + 201 us (NV = 4) and
    // 74 + 238 us (NV = 8) for the staged version: with 4-6 resident blocks the staging phase is not overlapped); option
    // "first_conv" = 4 / 8 selects the staged version for A/B.  (C = 32 uses its staged version by default: 152 -> 125 us.)
    const int ver = opt(OPT_FIRST_CONV);
    if (ver != 4 && ver != 8) {
      constexpr int NV = 4;  // NV = 8 (168 registers, 2-3 blocks / SM) measured slower
      if ((Hp + NV - 1) / NV > 65535 || (Wp + 127) / 128 > 65535) return LWS_ERR_BAD_SHAPE;
      dim3 grid(B * Dp, (Hp + NV - 1) / NV, (Wp + 127) / 128);
      conv3d_first_c8_kernel<NV><<<grid, 128, 0, st>>>(cost, w_first, b_first, affine, (uint4*)plane[0], (uint4*)plane[1], D, H, W);
    } else if (ver == 4) {
      if ((Hp + 3) / 4 > 65535 || (Wp + 127) / 128 > 65535) return LWS_ERR_BAD_SHAPE;
      dim3 grid(B * Dp, (Hp + 3) / 4, (Wp + 127) / 128);
      conv3d_first_c8_v2_kernel<4><<<grid, 128, 0, st>>>(cost, w_first, b_first, affine, (uint4*)plane[0], (uint4*)plane[1], D, H, W);
    } else {
      if ((Hp + 7) / 8 > 65535 || (Wp + 127) / 128 > 65535) return LWS_ERR_BAD_SHAPE;
      dim3 grid(B * Dp, (Hp + 7) / 8, (Wp + 127) / 128);
      conv3d_first_c8_v2_kernel<8><<<grid, 128, 0, st>>>(cost, w_first, b_first, affine, (uint4*)plane[0], (uint4*)plane[1], D, H, W);
    }
    if ((e = cudaPeekAtLastError()) != cudaSuccess) return (int)e;
  }
  // lws_set_option("c8_v1", 1): the one-plane-per-tile kernel
  const bool plane_groups = opt(OPT_C8_V1) == 0 && (long long)B * ((Wp + 125) / 126) * ((D + CP_L - 1) / CP_L) * Hp < (1ll << 31);
  LWS_SET_SMEM_ONCE(conv3d_c8p_kernel<false>, CP_SMEM);
  LWS_SET_SMEM_ONCE(conv3d_c8p_kernel<true>, CP_SMEM);
  LWS_SET_SMEM_ONCE(conv3d_c8_kernel<false>, C8_SMEM);
  LWS_SET_SMEM_ONCE(conv3d_c8_kernel<true>, C8_SMEM);
  e = cudaSuccess;
  if (e != cudaSuccess) return (int)e;
  int cur = 0;
  for (int l = 0; l <= layers; ++l) {  // l == layers: the closing 8 -> 1 conv
    C8Args a;
    memset(&a, 0, sizeof(a));
    const bool last = l == layers;
    a.in_hi = plane[cur], a.in_lo = plane[cur + 1], a.out_hi = plane[2 - cur], a.out_lo = plane[3 - cur];
    a.wtab = (const uint8_t*)(last ? w_last_tab : wtab[l]), a.bias = last ? bias_mid[0] : bias_mid[l];
    a.skip = add_skip ? cost : nullptr, a.out_f32 = out;
    a.vox_b = vox_b, a.Hp = Hp, a.Wp = Wp, a.H = H, a.W = W, a.D = D, a.fWp = make_fastdiv(Wp), a.fHp = make_fastdiv(Hp);
    a.line0 = Hp, a.strip_len = D * Hp;
    const int ct = (Wp + 125) / 126;
    a.ct_per_line = make_fastdiv(ct);
    a.total_tiles = B * ct * a.strip_len;
    if (plane_groups) {
      CPArgs pa;
      memset(&pa, 0, sizeof(pa));
      pa.c = a, pa.ct = ct, pa.ndg = (D + CP_L - 1) / CP_L;
      pa.total_steps = B * ct * pa.ndg * Hp;
      pa.fcolsteps = make_fastdiv(pa.ndg * Hp), pa.fct = make_fastdiv(ct);
      const int chs = opt(OPT_C8_CHUNK);  // lws_set_option("c8_chunk", n)
      pa.ch = chs > 0 ? chs : Hp;
      const int grid = pa.total_steps < kNumSMs ? pa.total_steps : kNumSMs;
      if (last) conv3d_c8p_kernel<true><<<grid, C8_THREADS, CP_SMEM, st>>>(pa);
      else conv3d_c8p_kernel<false><<<grid, C8_THREADS, CP_SMEM, st>>>(pa);
    } else {
      const int grid = a.total_tiles < kNumSMs ? a.total_tiles : kNumSMs;
      if (last) conv3d_c8_kernel<true><<<grid, C8_THREADS, C8_SMEM, st>>>(a);
      else conv3d_c8_kernel<false><<<grid, C8_THREADS, C8_SMEM, st>>>(a);
    }
    if ((e = cudaPeekAtLastError()) != cudaSuccess) return (int)e;
    cur = 2 - cur;
  }
  return LWS_OK;
}

}  // namespace lws
