// K6: the BN-ReLU-DW3x3(dil)-PW1x1 block of the colour-guidance refinement (reference models/submodules.py:236-261, used by
// refinement1 / refinement2, :282-327) on channels-last bordered ("CLP") tensors  act[b][y][x][32] fp32, border RP = 16.
//
// HBM-bound by design (reads and writes each 32-channel pixel once: 256 B / pixel); everything else is arranged so that it
// stays under that:
//   * TMA producer warp: one input line segment ((128 + 2*dil) pixels x 128 B) per step into a 5-slot shared-memory ring.
//   * 8 depthwise warps: a thread owns one channel quad of 4 pixels `dil` apart and walks down image lines of one row phase
//     (y = py + i*dil), so the 3x3 dilated window lives in registers: per output line it reads 6 float4 from the ring
//     (1.5 loads per output instead of 9) and issues 144 FFMA.  The result is split x = hi + lo*2^-11 into two fp16 values
//     and written as the A operand (row = pixel: [32 hi | 32 lo] halves = 128 B, SWIZZLE_128B).
//   * MMA warp: the 32x32 pointwise product as 4 tcgen05.mma kind::f16 (M=128): hi x [wh | wl] (N=64) and lo x wh (N=32,
//     accumulated onto the hi*wl columns, both carry the 2^-11 weight); products of 11-bit significands are exact in the
//     fp32 accumulator, so the result has the accuracy of an fp32 FMA chain (the dropped lo*lo term is 2^-22 relative).
//     At these small N an SS MMA is bound by the shared-memory port (max(N/2, (4096 + 32 N) / 128) cycles, tools/umma_ts_bench.cu),
//     hence few, wide MMAs; one elected lane runs the warp's whole schedule.
//   * 4 epilogue warps: tcgen05.ld, rescale + bias + ReLU, zero the x border; every warp stages its own 32 x 128 B quarter in
//     shared memory and TMA-stores it (clipped at the line end by the tensor map), so the tile's critical path has no
//     cross-warp barrier.
// The y-border lines of the output are zero-filled by all threads at kernel start.
#include "dwsep_common.cuh"

namespace lws {

// Schedule: a strip = (b, x tile, row phase py) = the image lines yi = py + i*dil, i in [0, rows_phase), of one 128-pixel
// column tile.  All B * nxt * dil * rows_phase tile-rows are numbered strip-major and every CTA takes one contiguous range of
// them, so the load is balanced to within one tile-row and a CTA pays the two warm-up line loads only once per (partial) strip.
// All warp roles walk the same static schedule through DsSched.
struct DsItem {
  int b, x0, yi0, nrows;
};
struct DsSched {
  long long g, g1;
  __device__ __forceinline__ DsSched(const DsArgs& a) {
    g = a.total_rows * blockIdx.x / gridDim.x;
    g1 = a.total_rows * (blockIdx.x + 1) / gridDim.x;
  }
  // next (partial) strip of this CTA; false when the range is exhausted
  __device__ __forceinline__ bool next(const DsArgs& a, DsItem& it) {
    while (g < g1) {
      const long long strip = g / a.rows_phase;
      const int i0 = (int)(g - strip * a.rows_phase);
      const int n = (int)min((long long)(a.rows_phase - i0), g1 - g);
      g += n;
      int t = (int)strip;
      const int py = t % a.dil;
      t /= a.dil;
      const int xt = t % a.nxt;
      it.b = t / a.nxt;
      it.x0 = xt * 128;
      const int rows_here = py < a.H ? (a.H - py + a.dil - 1) / a.dil : 0;  // this phase may be one line shorter
      it.nrows = min(n, rows_here - i0);
      it.yi0 = py + i0 * a.dil;
      if (it.nrows > 0) return true;
    }
    return false;
  }
};

// CIN = 0: the depthwise-separable block described above.  CIN = 3 / 1: the dense 3x3 first conv of a refinement branch
// (reference models/submodules.py:284-300: conv CIN -> 32 on the NCHW image / disparity), same back end with an im2col front end:
// the eight front-end warps gather the CIN*9 taps of a pixel straight from the NCHW input (coalesced along x), split them into
// hi/lo halves and write the A row [K = ci*9 + tap, zero padded to 32]; the 32 x 32 "pointwise" operand is the folded conv weight.
template <int CIN>
__global__ void __launch_bounds__(DS_THREADS, 1)
    dwsep_f16_kernel(const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_out, const DsArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int OFF_W = CIN > 0 ? DS_OFF_IN : DS_OFF_W, OFF_BAR = OFF_W + 9 * 32 * 4;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* in_full = bars;                 // [NIN]
  uint64_t* in_empty = in_full + DS_NIN;    // [NIN]
  uint64_t* a_full = in_empty + DS_NIN;     // [NA]
  uint64_t* a_empty = a_full + DS_NA;       // [NA]
  uint64_t* t_full = a_empty + DS_NA;       // [NT]
  uint64_t* t_empty = t_full + DS_NT;       // [NT]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + DS_NT);
  float* sW = reinterpret_cast<float*>(smem + OFF_W);
  uint64_t* img_full = reinterpret_cast<uint64_t*>(smem + DS_OFF_IMGBAR);  // CIN > 0 only: [NIMG] + [NIMG]
  uint64_t* img_empty = img_full + DS_NIMG;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int dil = a.dil;

  if (tid == 0) {
    for (int i = 0; i < DS_NIN; ++i) mbar_init(in_full + i, 1), mbar_init(in_empty + i, DS_DW_WARPS);
    for (int i = 0; i < DS_NA; ++i) mbar_init(a_full + i, DS_DW_WARPS), mbar_init(a_empty + i, 1);
    for (int i = 0; i < DS_NT; ++i) mbar_init(t_full + i, 1), mbar_init(t_empty + i, 4);
    if (CIN > 0)
      for (int i = 0; i < DS_NIMG; ++i) mbar_init(img_full + i, 1), mbar_init(img_empty + i, DS_DW_WARPS);
    mbar_fence_init();
    tma_prefetch_desc(&map_in);
    tma_prefetch_desc(&map_out);
  }
  if (warp == DS_MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(DS_NT * 64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // pointwise operand tile: 64 rows x 64 B used of a 128-byte SWIZZLE_128B row
  for (int idx = tid; idx < 64 * 4; idx += DS_THREADS) {
    const int n = idx >> 2, c = idx & 3;
    *reinterpret_cast<uint4*>(smem + DS_OFF_B + n * 128 + ((c ^ (n & 7)) << 4)) =
        __ldg(reinterpret_cast<const uint4*>(a.pwh + n * 32 + c * 8));
  }
  if (CIN == 0)
    for (int idx = tid; idx < 9 * 32; idx += DS_THREADS) sW[idx] = __ldg(a.dw + (idx & 31) * 9 + (idx >> 5)) * DS_ACT_SCALE;
  // zero the y-border lines of the output
  {
    const long long line4 = (long long)a.Wp * 8;  // float4 per line
    const long long total = (long long)a.B * 2 * DS_RP * line4;
    float4* o = reinterpret_cast<float4*>(a.out);
    for (long long i = (long long)blockIdx.x * DS_THREADS + tid; i < total; i += (long long)gridDim.x * DS_THREADS) {
      const long long ln = i / line4, r = i - ln * line4;
      const int b = (int)(ln / (2 * DS_RP)), k = (int)(ln % (2 * DS_RP));
      const int y = k < DS_RP ? k : a.Hp - 2 * DS_RP + k;
      o[((long long)b * a.Hp + y) * line4 + r] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  fence_proxy_async_smem();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  if (warp == DS_PROD_WARP) {
    // ================================ TMA producer ================================
    if (CIN > 0 && a.img_tma && !(a.dbg & 4) && elect_one_sync()) {
      // image lines yi0 - 1 .. yi0 + nrows of every (partial) strip, CIN channel lines of DS_IMG_ROW floats per ring slot; columns left of
      // 0 / right of W - 1 and lines above 0 / below H - 1 are the TMA's zero fill = the conv's zero padding
      uint32_t it = 0;
      DsSched sched(a);
      DsItem w;
      while (sched.next(a, w)) {
        for (int k = 0; k < w.nrows + 2; ++k, ++it) {
          const uint32_t slot = it % DS_NIMG;
          mbar_wait(img_empty + slot, ((it / DS_NIMG) & 1) ^ 1);
          mbar_expect_tx(img_full + slot, (uint32_t)(CIN * DS_IMG_ROW * 4));
          asm volatile(
              "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                  smem_u32(smem + DS_OFF_IMG + slot * DS_IMG_SLOT)),
              "l"(reinterpret_cast<uint64_t>(&map_in)), "r"(smem_u32(img_full + slot)), "r"(w.x0 - DS_RP - DS_IMG_X0), "r"(w.yi0 - 1 + k),
              "r"(w.b * CIN)
              : "memory");
        }
      }
    }
    if (CIN == 0 && elect_one_sync()) {
      uint32_t it = 0;
      const uint32_t bytes = (uint32_t)(128 + 2 * dil) * 128;
      DsSched sched(a);
        DsItem w;
        while (sched.next(a, w)) {
        const int line0 = w.b * a.Hp + DS_RP + w.yi0 - dil;  // first line needed: y - dil
        for (int k = 0; k < w.nrows + 2; ++k, ++it) {
          const uint32_t slot = it % DS_NIN;
          mbar_wait(in_empty + slot, ((it / DS_NIN) & 1) ^ 1);
          mbar_expect_tx(in_full + slot, bytes);
          asm volatile(
              "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                  smem_u32(smem + DS_OFF_IN + slot * DS_INBYTES)),
              "l"(reinterpret_cast<uint64_t>(&map_in)), "r"(smem_u32(in_full + slot)), "r"(0), "r"(w.x0 - dil),
              "r"(line0 + k * dil)
              : "memory");
        }
      }
    }
  } else if (warp == DS_MMA_WARP) {
    // ================================ MMA issuer ================================
    // kind::f16, fp32 accumulate, K-major A and B, M = 128, N = 64 / 32
    const uint32_t idesc64 = (1u << 4) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t idesc32 = (1u << 4) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t desc_hi = ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);  // SBO, version, SW128
    const uint32_t b_lo = ((smem_u32(smem + DS_OFF_B) & 0x3FFFF) >> 4) | (1u << 16);
    uint32_t t = 0;
    if (elect_one_sync()) {  // one lane runs the whole schedule (no per-tile elect + reconvergence on the tile's critical path)
    DsSched sched(a);
      DsItem w;
      while (sched.next(a, w)) {
      for (int i = 0; i < w.nrows; ++i, ++t) {
        const uint32_t ab = t % DS_NA, tb = t % DS_NT;
        mbar_wait(t_empty + tb, ((t / DS_NT) & 1) ^ 1);
        mbar_wait(a_full + ab, (t / DS_NA) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        {
          const uint32_t a_lo = ((smem_u32(smem + DS_OFF_A + ab * DS_TILE) & 0x3FFFF) >> 4) | (1u << 16);
          const uint32_t d = tmem + tb * 64;
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const uint32_t acc = k > 0;
            asm volatile(
                "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d),
                "l"(desc_hi | (uint64_t)(a_lo + k * 2)), "l"(desc_hi | (uint64_t)(b_lo + k * 2)), "r"(idesc64), "r"(acc)
                : "memory");
          }
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            asm volatile(
                "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d + 32),
                "l"(desc_hi | (uint64_t)(a_lo + 4 + k * 2)), "l"(desc_hi | (uint64_t)(b_lo + k * 2)), "r"(idesc32), "r"(1u)
                : "memory");
          }
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(a_empty + ab))
                       : "memory");
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(t_full + tb))
                       : "memory");
        }
      }
    }
    }
    __syncwarp();
  } else if (warp >= DS_EPI_WARP0) {
    // ================================ epilogue (warps 8..11) ================================
    const int quarter = warp & 3;
    const int p = quarter * 32 + lane;  // pixel of the tile = TMEM lane
    float bias[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) bias[c] = __ldg(a.bias + c);
    const float c0 = __ldg(a.scales), c1 = __ldg(a.scales + 1);
    const float lo = a.relu ? 0.f : -INFINITY;
    uint32_t t = 0;
    DsSched sched(a);
      DsItem w;
      while (sched.next(a, w)) {
      const int xpix = w.x0 + p;
      const bool border = xpix < DS_RP || xpix >= a.Wp - DS_RP;
      for (int i = 0; i < w.nrows; ++i, ++t) {
        const uint32_t tb = t % DS_NT, ob = t % DS_NOUT;
        mbar_wait(t_full + tb, (t / DS_NT) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t taddr = tmem + ((uint32_t)(quarter * 32) << 16) + tb * 64;
        float m[32], c[32];
        ds_ld32(taddr, m);
        ds_ld32(taddr + 32, c);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(t_empty + tb);
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const float2 u = __ffma2_rn(make_float2(c[j], c[j + 1]), make_float2(c1, c1),
                                      __ffma2_rn(make_float2(m[j], m[j + 1]), make_float2(c0, c0), make_float2(bias[j], bias[j + 1])));
          m[j] = border ? 0.f : fmaxf(u.x, lo);
          m[j + 1] = border ? 0.f : fmaxf(u.y, lo);
        }
        if (a.dbg & 8) {  // timing experiment: no staging, no store
          if (m[0] == 123.456f) a.out[0] = m[1];
          continue;
        }
        // every epilogue warp stages and stores its own 32 pixels (no cross-warp barrier on the tile's critical path);
        // staging quarter `ob` was last read by the TMA store this warp issued DS_NOUT tiles ago
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(DS_NOUT - 1) : "memory");
        __syncwarp();
        const uint32_t so = smem_u32(smem + DS_OFF_OUT + ob * DS_TILE) + p * 128;
        if (a.out_split) {
#pragma unroll
          for (int c8 = 0; c8 < 4; ++c8) {
            uint32_t hi[4], lw[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float v0 = m[c8 * 8 + 2 * k] * DS_ACT_SCALE, v1 = m[c8 * 8 + 2 * k + 1] * DS_ACT_SCALE;
              const __half2 h = f2h2_sat(v0, v1);
              const float2 f = __half22float2(h);
              const float2 d = split_lo2(v0, v1, f);
              hi[k] = h2_bits(h), lw[k] = h2_bits(f2h2_sat(d.x, d.y));
            }
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(so + ((c8 ^ (p & 7)) << 4)), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]),
                         "r"(hi[3])
                         : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(so + (((c8 + 4) ^ (p & 7)) << 4)), "r"(lw[0]), "r"(lw[1]),
                         "r"(lw[2]), "r"(lw[3])
                         : "memory");
          }
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) sts128(so + ((j ^ (p & 7)) << 4), make_float4(m[4 * j], m[4 * j + 1], m[4 * j + 2], m[4 * j + 3]));
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                           reinterpret_cast<uint64_t>(&map_out)),
                       "r"(smem_u32(smem + DS_OFF_OUT + ob * DS_TILE + quarter * 4096)), "r"(0), "r"(w.x0 + quarter * 32),
                       "r"(w.b * a.Hp + DS_RP + w.yi0 + i * dil)
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  } else if (CIN > 0) {
    // ================================ im2col warps (0..7) ================================
    const int h = warp & 1;                  // which 16 of the 32 K slots (warp-uniform)
    const int p = (warp >> 1) * 32 + lane;   // pixel of the tile
    const uint32_t a_base = smem_u32(smem + DS_OFF_A);
    const uint32_t row = (uint32_t)p * 128;
    const long long hw = (long long)a.H * a.W;
    uint32_t t_ = 0, it_img = 0;
    DsSched sched(a);
      DsItem w;
      while (sched.next(a, w)) {
      const int x = w.x0 + p - DS_RP;  // image column of this pixel
      const float* src0 = a.img + (long long)w.b * CIN * hw + x;
      // K slots are ordered by tap column c = ci*3 + kx with the three ky taps of a column adjacent (slot = 3c + ky for c < 5,
      // 16 + 3(c-5) + ky otherwise; slot 15 is padding), so each half warp owns whole columns and a step down the image is a
      // register shift: only the new bottom row (5 / 4 values per thread instead of 16) is read per step -- from the TMA ring of
      // image lines when the rows are 16-byte aligned (a.img_tma), otherwise from global memory PF steps ahead through a rotating
      // register FIFO (a new image line comes from DRAM, ~1.5 us under load, longer than one step).
      constexpr int PF = 4, NCOL = CIN * 3;
      if (a.img_tma) {
        // ---- image lines from the TMA ring: the line that becomes the window's bottom row is read from shared memory when it is needed
        const uint32_t img_base = smem_u32(smem + DS_OFF_IMG) + (uint32_t)(p + DS_IMG_X0 - 1) * 4;  // tap kx = 0 of pixel p
        auto take_row = [&](float (&r)[5]) {
          const uint32_t slot = it_img % DS_NIMG;
          if (!(a.dbg & 4)) mbar_wait(img_full + slot, (it_img / DS_NIMG) & 1);
          const uint32_t sb = img_base + slot * DS_IMG_SLOT;
#pragma unroll
          for (int lc = 0; lc < 5; ++lc) {
            const int c0 = lc, c1 = 5 + lc;  // both halves are unrolled and the warp-uniform h selects one
            float v0 = 0.f, v1 = 0.f;
            if (h == 0) {
              if (c0 < NCOL) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v0) : "r"(sb + (uint32_t)((c0 / 3) * DS_IMG_ROW + c0 % 3) * 4));
            } else {
              if (c1 < NCOL) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v1) : "r"(sb + (uint32_t)((c1 / 3) * DS_IMG_ROW + c1 % 3) * 4));
            }
            r[lc] = h == 0 ? v0 : v1;
          }
          __syncwarp();
          if (lane == 0 && !(a.dbg & 4)) mbar_arrive(img_empty + slot);
          ++it_img;
        };
        float t[16];
        {
          float r0[5], r1[5];
          take_row(r0);
          take_row(r1);
#pragma unroll
          for (int lc = 0; lc < 5; ++lc) t[3 * lc] = 0.f, t[3 * lc + 1] = r0[lc], t[3 * lc + 2] = r1[lc];
          t[15] = 0.f;
        }
        for (int i = 0; i < w.nrows; ++i, ++t_) {
          float nr[5];
          take_row(nr);
#pragma unroll
          for (int lc = 0; lc < 5; ++lc) t[3 * lc] = t[3 * lc + 1], t[3 * lc + 1] = t[3 * lc + 2], t[3 * lc + 2] = nr[lc];
          float v[16];
#pragma unroll
          for (int kk = 0; kk < 16; ++kk) v[kk] = t[kk] * DS_ACT_SCALE;
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const __half2 hh = f2h2_sat(v[2 * k], v[2 * k + 1]);
            const float2 f = __half22float2(hh);
            const float2 d = split_lo2(v[2 * k], v[2 * k + 1], f);
            hi[k] = h2_bits(hh), lo[k] = h2_bits(f2h2_sat(d.x, d.y));
          }
          const uint32_t ab = t_ % DS_NA;
          mbar_wait(a_empty + ab, ((t_ / DS_NA) & 1) ^ 1);
          const uint32_t dst = a_base + ab * DS_TILE + row;
#pragma unroll
          for (int c = 0; c < 2; ++c) {  // hi chunks 2h, 2h+1; lo chunks 4+2h, 4+2h+1
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + (((2 * h + c) ^ (p & 7)) << 4)), "r"(hi[4 * c]),
                         "r"(hi[4 * c + 1]), "r"(hi[4 * c + 2]), "r"(hi[4 * c + 3])
                         : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + (((4 + 2 * h + c) ^ (p & 7)) << 4)), "r"(lo[4 * c]),
                         "r"(lo[4 * c + 1]), "r"(lo[4 * c + 2]), "r"(lo[4 * c + 3])
                         : "memory");
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(a_full + ab);
        }
        continue;
      }
      auto load_row = [&](int yrow, float (&r)[5]) {
        const bool oky = (unsigned)yrow < (unsigned)a.H && !(a.dbg & 4);
        const float* src = src0 + (long long)yrow * a.W;
#pragma unroll
        for (int lc = 0; lc < 5; ++lc) {
          // both halves are unrolled and the warp-uniform h selects one
          const int c0 = lc, c1 = 5 + lc;
          float v0 = 0.f, v1 = 0.f;
          if (h == 0) {
            if (c0 < NCOL) {
              const int ci = c0 / 3, kx = c0 % 3;
              if (oky && (unsigned)(x + kx - 1) < (unsigned)a.W) v0 = __ldg(src + ci * hw + (kx - 1));
            }
          } else {
            if (c1 < NCOL) {
              const int ci = c1 / 3, kx = c1 % 3;
              if (oky && (unsigned)(x + kx - 1) < (unsigned)a.W) v1 = __ldg(src + ci * hw + (kx - 1));
            }
          }
          r[lc] = h == 0 ? v0 : v1;
        }
      };
      float t[16], fifo[PF][5];
      {
        float r0[5], r1[5];
        load_row(w.yi0 - 1, r0);
        load_row(w.yi0, r1);
#pragma unroll
        for (int lc = 0; lc < 5; ++lc) t[3 * lc] = 0.f, t[3 * lc + 1] = r0[lc], t[3 * lc + 2] = r1[lc];
        t[15] = 0.f;
#pragma unroll
        for (int f = 0; f < PF; ++f) load_row(w.yi0 + 1 + f, fifo[f]);
      }
      // One step down the image.  `slot` holds the line that becomes the window's bottom row; it is consumed and then refilled with
      // the line PF steps further down.  The FIFO is a ROTATION of PF statically named register sets (the row loop is unrolled by
      // PF): shifting the values through registers instead would make every step wait for the load issued ONE step earlier (a move
      // out of a load's destination register waits for that load), i.e. a prefetch distance of 1 whatever PF is -- that move was
      // 23 % of the kernel's stall samples in profiles/r02_ncu_full_r02_dwsep_f16_3.txt.
      auto step = [&](int i, float (&slot)[5]) {
        const int y = w.yi0 + i;
#pragma unroll
        for (int lc = 0; lc < 5; ++lc) t[3 * lc] = t[3 * lc + 1], t[3 * lc + 1] = t[3 * lc + 2], t[3 * lc + 2] = slot[lc];
        load_row(y + 1 + PF, slot);
        float v[16];
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) v[kk] = t[kk] * DS_ACT_SCALE;
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const __half2 hh = f2h2_sat(v[2 * k], v[2 * k + 1]);
          const float2 f = __half22float2(hh);
          const float2 d = split_lo2(v[2 * k], v[2 * k + 1], f);
          hi[k] = h2_bits(hh), lo[k] = h2_bits(f2h2_sat(d.x, d.y));
        }
        const uint32_t ab = t_ % DS_NA;
        mbar_wait(a_empty + ab, ((t_ / DS_NA) & 1) ^ 1);
        const uint32_t dst = a_base + ab * DS_TILE + row;
#pragma unroll
        for (int c = 0; c < 2; ++c) {  // hi chunks 2h, 2h+1; lo chunks 4+2h, 4+2h+1
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + (((2 * h + c) ^ (p & 7)) << 4)), "r"(hi[4 * c]), "r"(hi[4 * c + 1]),
                       "r"(hi[4 * c + 2]), "r"(hi[4 * c + 3])
                       : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + (((4 + 2 * h + c) ^ (p & 7)) << 4)), "r"(lo[4 * c]),
                       "r"(lo[4 * c + 1]), "r"(lo[4 * c + 2]), "r"(lo[4 * c + 3])
                       : "memory");
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(a_full + ab);
        ++t_;
      };
      for (int i = 0; i < w.nrows; i += PF) {
        step(i, fifo[0]);
        if (i + 1 < w.nrows) step(i + 1, fifo[1]);
        if (i + 2 < w.nrows) step(i + 2, fifo[2]);
        if (PF > 3 && i + 3 < w.nrows) step(i + 3, fifo[PF - 1]);
      }
    }
  } else {
    // ================================ depthwise warps (0..7) ================================
    const int q = tid & 7;    // channel quad
    const int pg = tid >> 3;  // 0..31: pixel group = (block of 4*dil pixels, offset g inside the stride)
    const int blk = pg / dil, g = pg - blk * dil;
    const int pbase = blk * 4 * dil + g;                      // first of this thread's 4 pixels (stride dil)
    const uint32_t in_off = (uint32_t)pbase * 128 + q * 16;   // ring column of tap kx = 0 of pixel 0 is pbase (box starts at x0 - dil)
    const uint32_t jstep = (uint32_t)dil * 128;
    const uint32_t in_base = smem_u32(smem + DS_OFF_IN), a_base = smem_u32(smem + DS_OFF_A);
    const uint32_t w_base = smem_u32(sW) + q * 16;
    uint32_t a_off[4];  // byte offsets of the hi halves of this thread's 4 pixels inside an A tile (lo = chunk + 4)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int p = pbase + j * dil;
      a_off[j] = (uint32_t)p * 128 + (q & 1) * 8 + ((uint32_t)((q >> 1) ^ (p & 7)) << 4);
    }
    uint32_t it = 0, t = 0;
    float4 win[3][6];
    DsSched sched(a);
      DsItem w;
      while (sched.next(a, w)) {
      int npend = 0;  // ring slots read but not yet released (released once their values have been consumed)
      auto load_line = [&](float4(&dst)[6]) {
        const uint32_t slot = it % DS_NIN;
        mbar_wait(in_full + slot, (it / DS_NIN) & 1);
        const uint32_t s = in_base + slot * DS_INBYTES + in_off;
#pragma unroll
        for (int j = 0; j < 6; ++j) dst[j] = lds128(s + j * jstep);
        ++it, ++npend;
      };
      auto emit = [&](const float4(&top)[6], const float4(&mid)[6], const float4(&bot)[6]) {
        // channel pairs through FFMA2 (fma.rn.f32x2, sm_100): 72 issued FMAs per output line instead of 144, same roundings
        float2 alo[4], ahi[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) alo[j] = make_float2(0.f, 0.f), ahi[j] = make_float2(0.f, 0.f);
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const float4(&row)[6] = ky == 0 ? top : (ky == 1 ? mid : bot);
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const float4 k = lds128(w_base + (ky * 3 + kx) * 128);
            const float2 k01 = make_float2(k.x, k.y), k23 = make_float2(k.z, k.w);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              alo[j] = __ffma2_rn(make_float2(row[j + kx].x, row[j + kx].y), k01, alo[j]);
              ahi[j] = __ffma2_rn(make_float2(row[j + kx].z, row[j + kx].w), k23, ahi[j]);
            }
          }
        }
        float4 acc[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] = make_float4(alo[j].x, alo[j].y, ahi[j].x, ahi[j].y);
        __syncwarp();
        if (lane == 0)
          for (; npend > 0; --npend) mbar_arrive(in_empty + (it - npend) % DS_NIN);
        npend = 0;
        const uint32_t ab = t % DS_NA;
        mbar_wait(a_empty + ab, ((t / DS_NA) & 1) ^ 1);
        const uint32_t dst = a_base + ab * DS_TILE;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const __half2 h01 = f2h2_sat(acc[j].x, acc[j].y), h23 = f2h2_sat(acc[j].z, acc[j].w);
          const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
          const float2 d01 = split_lo2(acc[j].x, acc[j].y, f01), d23 = split_lo2(acc[j].z, acc[j].w, f23);
          const __half2 l01 = f2h2_sat(d01.x, d01.y);
          const __half2 l23 = f2h2_sat(d23.x, d23.y);
          sts64(dst + a_off[j], h2_bits(h01), h2_bits(h23));
          sts64(dst + (a_off[j] ^ 64u), h2_bits(l01), h2_bits(l23));  // chunk + 4 (bit 6 of the swizzled offset)
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(a_full + ab);
        ++t;
      };
      load_line(win[0]);
      load_line(win[1]);
      for (int i = 0; i < w.nrows; i += 3) {
        load_line(win[2]);
        emit(win[0], win[1], win[2]);
        if (i + 1 >= w.nrows) break;
        load_line(win[0]);
        emit(win[1], win[2], win[0]);
        if (i + 2 >= w.nrows) break;
        load_line(win[1]);
        emit(win[2], win[0], win[1]);
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == DS_MMA_WARP) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(DS_NT * 64));
}

// in/out: CLP [B][H + 2*RP][W + 2*RP][32]; dw [32][9]; pwh/scales: see DsArgs; dil in {1,2,4,8,16}
int launch_dwsep_f16(const float* in, float* out, const float* dw, const void* pwh, const float* scales, const float* bias, int B,
                     int H, int W, int dil, int relu, int out_split, cudaStream_t st) {
  if (dil < 1 || dil > DS_RP || (32 % dil) != 0) return LWS_ERR_UNSUPPORTED;
  LWS_SET_SMEM_ONCE(dwsep_f16_kernel<0>, DS_SMEM);
  cudaError_t e;
  DsArgs a;
  memset(&a, 0, sizeof(a));
  a.dw = dw, a.pwh = (const __half*)pwh, a.scales = scales, a.bias = bias, a.out = out;
  a.B = B, a.Hp = H + 2 * DS_RP, a.Wp = W + 2 * DS_RP, a.H = H, a.dil = dil, a.relu = relu, a.out_split = out_split;
  a.nxt = (a.Wp + 127) / 128;
  a.rows_phase = (H + dil - 1) / dil;
  a.total_rows = (long long)B * a.nxt * dil * a.rows_phase;
  a.dbg = opt(OPT_CHAIN_DEBUG);
  const int grid = a.total_rows < kNumSMs ? (int)a.total_rows : kNumSMs;
  CUtensorMap map_in, map_out;
  const uint64_t dims[3] = {32, (uint64_t)a.Wp, (uint64_t)B * a.Hp}, strides[2] = {128, (uint64_t)a.Wp * 128};
  const uint32_t box_in[3] = {32, (uint32_t)(128 + 2 * dil), 1}, box_out[3] = {32, 32, 1};  // one TMA store per epilogue warp
  int rc = make_tensor_map_f32(&map_in, in, 3, dims, strides, box_in, false);
  if (rc) return rc;
  rc = make_tensor_map_f32(&map_out, out, 3, dims, strides, box_out, true);
  if (rc) return rc;
  dwsep_f16_kernel<0><<<grid, DS_THREADS, DS_SMEM, st>>>(map_in, map_out, a);
  e = cudaPeekAtLastError();
  return e == cudaSuccess ? LWS_OK : (int)e;
}

// first conv of a refinement branch: img NCHW [B][CIN][H][W] fp32 -> CLP [B][H+32][W+32][32] fp32 = ReLU(conv3x3 + bias)
// (bias = the folded BatchNorm of the following block); wtab: [64][32] halves (row co / 32 + co = hi / lo of w[k][co] * sw, K slot
// k = ci*9 + tap, zero padded), scales as for the pointwise tables
int launch_conv0_f16(const float* img, float* out, const void* wtab, const float* scales, const float* bias, int B, int CIN, int H,
                     int W, cudaStream_t st) {
  if (CIN != 1 && CIN != 3) return LWS_ERR_UNSUPPORTED;
  if (CIN == 3) LWS_SET_SMEM_ONCE(dwsep_f16_kernel<3>, DS_SMEM_IM2COL);
  else LWS_SET_SMEM_ONCE(dwsep_f16_kernel<1>, DS_SMEM_IM2COL);
  cudaError_t e;
  DsArgs a;
  memset(&a, 0, sizeof(a));
  a.dw = bias /*unused*/, a.pwh = (const __half*)wtab, a.scales = scales, a.bias = bias, a.out = out, a.img = img, a.W = W;
  a.B = B, a.Hp = H + 2 * DS_RP, a.Wp = W + 2 * DS_RP, a.H = H, a.dil = 1, a.relu = 1, a.out_split = 0;
  a.nxt = (a.Wp + 127) / 128;
  a.rows_phase = H;
  a.total_rows = (long long)B * a.nxt * H;
  a.dbg = opt(OPT_CHAIN_DEBUG);
  const int grid = a.total_rows < kNumSMs ? (int)a.total_rows : kNumSMs;
  CUtensorMap map_out, map_img;
  const uint64_t dims[3] = {32, (uint64_t)a.Wp, (uint64_t)B * a.Hp}, strides[2] = {128, (uint64_t)a.Wp * 128};
  const uint32_t box_out[3] = {32, 32, 1};
  int rc = make_tensor_map_f32(&map_out, out, 3, dims, strides, box_out, true);
  if (rc) return rc;
  // image lines through TMA when the NCHW rows are 16-byte aligned (W % 4 == 0); otherwise the front end loads its taps itself
  a.img_tma = (W % 4 == 0) && ((uintptr_t)img % 16 == 0) && W >= 4;
  map_img = map_out;
  if (a.img_tma) {
    const uint64_t dimg[3] = {(uint64_t)W, (uint64_t)H, (uint64_t)B * CIN}, simg[2] = {(uint64_t)W * 4, (uint64_t)H * W * 4};
    const uint32_t bimg[3] = {(uint32_t)DS_IMG_ROW, 1, (uint32_t)CIN};
    if (make_tensor_map_f32(&map_img, img, 3, dimg, simg, bimg, false) != LWS_OK) a.img_tma = 0, map_img = map_out;
  }
  if (CIN == 3) dwsep_f16_kernel<3><<<grid, DS_THREADS, DS_SMEM_IM2COL, st>>>(map_img, map_out, a);
  else dwsep_f16_kernel<1><<<grid, DS_THREADS, DS_SMEM_IM2COL, st>>>(map_img, map_out, a);
  e = cudaPeekAtLastError();
  return e == cudaSuccess ? LWS_OK : (int)e;
}

}  // namespace lws
