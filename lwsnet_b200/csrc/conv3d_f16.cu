// K3 (C = 32) and the dense refinement conv on the 5th-generation tensor cores with split-fp16 operands.
//
// Replaces the 32 -> 32 layers of post_3dconvs(4, 32) (reference models/submodules.py:190-221) and, with TZ = 8, the dense
// 64 -> 32 dilation-8 conv of refinement2 (models/submodules.py:302-312).
//
// Numerics: every activation x and folded weight w is carried as two fp16 values, x*sa = xh + xl * 2^-11 (sa = 2^-6),
// w*sw = wh + wl * 2^-11 (sw = per-layer power of two).  acc_main = sum xh*wh, acc_corr = sum (xh*wl + xl*wh) in fp32 (TMEM);
// products of two 11-bit significands are exact in fp32, the dropped xl*wl term is 2^-22 relative: fp32-grade results at the
// fp16 tensor rate (plain TF32/BF16 operands move stage-1 disparities by whole pixels).
//
// Cost model (tools/umma_ts_bench.cu, r02): one tcgen05.mma with both operands in shared memory takes max(N/2, (4096 + 32 N) / 128)
// cycles at M = 128, K = 16 -- the tensor pipe's N/2 once N >= 128, the shared-memory port (4 KB A slice + 32 N bytes of B) below --
// so the work is arranged as few, wide MMAs ("Toeplitz-N"):
//   * a GEMM row is one voxel = 128 bytes [32 hi | 32 lo] halves; voxels are stored with the Toeplitz axis fastest
//     (3D stack: act[b][y][x][d], d padded by one zero voxel each side, x padded, y padding = TMA out-of-bounds zero fill);
//   * the three taps along the fastest axis are folded into N: D[r', t*32+co] = sum_ci x[r', ci] * w[t][ci][co], t = 0..2, and
//     the epilogue adds the three column blocks of three neighbouring rows: out[r] = D[r-TZ, 0] + D[r, 1] + D[r+TZ, 2].
//     With the hi/lo weight halves side by side, N = 192 for the xh operand and N = 96 (accumulated onto the wl columns) for
//     the xl operand: 4 MMAs per (stage, window) instead of 24;
//   * the taps along the middle axis are row-shifted windows (UMMA descriptor offsets) of ONE TMA box per stage, the taps
//     along the slowest axis are the stages; tiles walk down the slowest axis in strips, so two of a tile's three boxes are already
//     in the shared-memory ring: 1 TMA load of 23 KB per 126 output voxels for the 3D stack;
//   * all weights of the layer (27 x 32 x 32 hi + lo = 108 KB) stay resident in shared memory.
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = MMA issuer (ONE elected lane runs the whole schedule) + TMEM owner (2
// accumulators x 192 columns),
// warps 2-9 = epilogue, two per TMEM lane quarter, 16 output channels each (tcgen05.ld, Toeplitz row shifts by warp shuffles + a small shared-memory exchange at the warp
// boundaries, bias + ReLU, border zeroing, re-split into hi/lo, staged through shared memory into one TMA store).
#include <cuda_fp16.h>
#include <math.h>
#include <string.h>

#include "lws_common.cuh"
#include "tma_utils.cuh"
#include "conv3d_f16.cuh"

namespace lws {

constexpr int TZ_THREADS = 320;  // warp 0 producer, warp 1 MMA, warps 2..9 epilogue (two per TMEM lane quarter)
constexpr int TZ_XS = 36;  // floats per row of the epilogue's boundary-row exchange: 32 + 4 so that the rows of a quarter-warp's 128-bit
                           // accesses fall into different banks (a 32-float pitch put all of them into the same four: 8-way conflicts)
constexpr int TZ_BTILE = 192 * 128;  // one weight tile: 192 rows x [block 2i | block 2i+1] x 32 halves

struct TzArgs {
  const float* bias;    // [32]
  const float* skip;    // LAST: raw cost [B,D,H,W] or null
  float* out_f32;       // LAST: [B,D,H,W]
  int out_mode;         // LAST: 0 = rows ordered (y, x, d) -> NCDHW; 1 = rows ordered (y, x) -> NCHW
  const float* scales;  // [2] device: 1/sw, 1/(sw * 2^11)
  float out_mul;        // epilogue multiplier on top of scales: 1 for split-fp16 output (values stay scaled by sa), 1/sa for fp32
  float bias_mul;       // sa for split-fp16 output, 1 for fp32
  int out_split;        // 1: rows [32 hi | 32 lo] halves; 0: rows of 32 fp32
  int relu;
  int R;                // rows per batch element
  FastDiv n0, n1;       // padded lengths of the fastest and the middle axis
  int p0, i0, p1, i1;   // pad and interior length of each: rows outside the interior are written as 0
  // strip schedule: the row space of one batch element is cut into super lines of `srow` rows (the distance between
  // consecutive taps of the slowest axis); a strip = (b, column tile ct of the super line) walked down the super lines; tile n of
  // a strip outputs rows ct*OUTR + n*srow + [0, OUTR) and shares all but its last `G` stage boxes with tile n-1, so they stay
  // in the shared-memory ring.  srow = R gives plain linear tiling (strips of one tile).
  int srow, strip_len, total_tiles;  // total_tiles = B * ct_per_sl * strip_len, split evenly (contiguously) over the CTAs
  int seg_len, segs_per_strip, nseg;  // segment schedule (TzSched): nseg = B * segs_per_strip * ct_per_sl
  int dbg;  // option tz_debug, TIMING EXPERIMENTS ONLY (results are wrong): 1 = no A loads after the first ring fill, 2 = epilogue only frees
            // the accumulator, 4 = epilogue without staging / TMA store
  FastDiv ct_per_sl;
  int nstages, G, nshift, nbtiles, nslot;
  int slot_bytes, box_bytes;
  int st_off[TZ_MAXST];  // row offset of the stage's box relative to the tile's first GEMM row
  int st_src[TZ_MAXST];
  int shift_rows[3];     // row offset of window k inside the box
};

struct TzItem {
  int b, orow0, ntiles;
};
// A strip = (b, column tile ct) walked down all strip_len super lines.  Two ways of handing strips to the CTAs:
//  * seg_len == 0: tiles are numbered strip-major and every CTA takes one contiguous range (balanced to one tile; a CTA reloads the
//    shared stage boxes only at the start of a (partial) strip);
//  * seg_len > 0: a strip is cut into segments of seg_len super lines; segments are numbered (b, segment of the strip, ct) with ct
//    fastest and dealt round-robin, so CTAs k and k+1 walk NEIGHBOURING column tiles down the SAME super lines at the same time:
//    the rows their boxes share (the halo of the middle-axis taps) are fetched from DRAM once and hit L2 for the neighbour.
// NCTA = 2: a CTA pair walks the schedule together (OUTR = rows of the pair's double tile; CTA r of the pair takes its r-th half).
template <int OUTR, int NCTA = 1>
struct TzSched {
  int g, g1, step;
  __device__ __forceinline__ TzSched(const TzArgs& a) {
    const int bid = blockIdx.x / NCTA, nb = gridDim.x / NCTA;
    if (a.strip_len == 1) {  // linear tiling: interleave the tiles over the CTAs (neighbouring tiles run concurrently: L2 reuse)
      g = bid, g1 = a.total_tiles, step = nb;
    } else if (a.seg_len > 0) {
      g = bid, g1 = a.nseg, step = nb;
    } else {
      g = (int)((long long)a.total_tiles * bid / nb);
      g1 = (int)((long long)a.total_tiles * (bid + 1) / nb);
      step = 0;
    }
  }
  __device__ __forceinline__ bool next(const TzArgs& a, TzItem& it) {
    if (g >= g1) return false;
    if (a.seg_len > 0) {
      int bs, ct;
      fdivmod(g, a.ct_per_sl, bs, ct);
      it.b = bs / a.segs_per_strip;
      const int n0 = (bs - it.b * a.segs_per_strip) * a.seg_len;
      it.ntiles = min(a.seg_len, a.strip_len - n0);
      it.orow0 = ct * OUTR + n0 * a.srow;
      g += step;
      return true;
    }
    const int strip = g / a.strip_len, n0 = g - strip * a.strip_len;
    int ct;
    fdivmod(strip, a.ct_per_sl, it.b, ct);
    it.ntiles = step ? 1 : min(a.strip_len - n0, g1 - g);
    it.orow0 = ct * OUTR + n0 * a.srow;
    g += step ? step : it.ntiles;
    return true;
  }
};

__device__ __forceinline__ void tz_mma(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d),
      "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tz_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// ---- CTA-pair (cta_group::2) forms: one MMA over M = 256 = the two CTAs' 128-row A tiles; each CTA supplies half of B's N rows from
// its own shared memory at the descriptor's address, so the B operand bytes per CTA (the shared-memory port load) are halved.
__device__ __forceinline__ void tz_mma2(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d),
      "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
// arrives on the barrier at this shared-memory offset in BOTH CTAs of the pair once the MMAs issued so far have completed
__device__ __forceinline__ void tz_commit2(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // shared::cluster address of the same offset in the pair's even (leader) CTA
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load into this CTA's shared memory whose bytes are counted on the LEADER CTA's mbarrier
__device__ __forceinline__ void tma2_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma2_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                   smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
               : "memory");
}
// arrive on the barrier at this offset in the pair's leader CTA
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerBitMask) : "memory");
}
constexpr int TZ_BTILE2 = 144 * 128;  // pair: a CTA's share of one weight tile = 96 rows (its half of N = 192) + 48 rows (its half of N = 96)

__device__ __forceinline__ void tz_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[j]);
}

// TZ = Toeplitz row shift (1: 3D stack along d; 8: dilated 2D conv along x).  Output rows per tile: 128 - 2*TZ.
// NST / NSH: stages per tile and row-shifted windows per stage (compile-time so the MMA issue loop is fully unrolled: the
// single issuing thread must spend only a few uniform-datapath instructions per MMA or it, not the tensor pipe, is the limit).
// LAST: the C -> 1 convolution that closes the stack.  Same pipeline with N = 16 per MMA: the weight block of a (stage,
// window) is 16 rows x [B1 | B2] with B1 = rows {t: wh_t, 8+t: wl_t} for the xh operand and B2 = rows {8+t: wh_t} for the xl
// operand, so acc[:, t] = main and acc[:, 8+t] = corr of Toeplitz tap t; the epilogue writes fp32 NCDHW (+ skip).
// PAIR: launched as clusters of two CTAs that walk the schedule together on double tiles (CTA r = the r-th 128-row half); the leader's
// single thread issues every MMA for both (cta_group::2), each CTA's producer fills its own A ring and its half of the weights, each
// CTA's epilogue drains its own TMEM.
template <int TZ, int NST, int NSH, bool LAST, bool PAIR = false>
__global__ void __launch_bounds__(TZ_THREADS, 1)
    tz_gemm_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                   const __grid_constant__ CUtensorMap mapB, const __grid_constant__ CUtensorMap mapOut, const TzArgs a) {
  constexpr int OUTR = 128 - 2 * TZ;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sB = smem;                                   // [nbtiles][24576]
  uint8_t* sOut = sB + a.nbtiles * (PAIR ? TZ_BTILE2 : TZ_BTILE);  // [16384] staging tile
  uint8_t* sA = sOut + 16384;                            // [nslot][slot_bytes]
  uint8_t* tail = sA + a.nslot * a.slot_bytes;
  float* xch = reinterpret_cast<float*>(tail);           // [4 quarters][3*TZ rows][TZ_XS] boundary rows for the Toeplitz shifts
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail + 4 * 3 * TZ * TZ_XS * 4);
  uint64_t* a_full = bars;                 // [8]
  uint64_t* a_empty = a_full + 8;          // [8]
  uint64_t* b_full = a_empty + 8;          // [1]
  uint64_t* t_full = b_full + 1;           // [2]
  uint64_t* t_empty = t_full + 2;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  constexpr int SOUTR = PAIR ? 2 * OUTR : OUTR;  // rows of a schedule tile
  constexpr int NCTA = PAIR ? 2 : 1;
  if (tid == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(a_full + i, 1), mbar_init(a_empty + i, 1);
    mbar_init(b_full, 1);
    for (int i = 0; i < 2; ++i) mbar_init(t_full + i, 1), mbar_init(t_empty + i, LAST ? 4 : 8 * NCTA);
    mbar_fence_init();
    tma_prefetch_desc(&mapA0);
    tma_prefetch_desc(&mapA1);
    tma_prefetch_desc(&mapB);
    tma_prefetch_desc(&mapOut);
  }
  if (PAIR) {  // the peer's barriers must exist before anything of this CTA signals them
    __syncthreads();
    cluster_sync_all();
  }
  if (warp == 1) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  constexpr int nst = NST;
  const uint32_t nslot = (uint32_t)a.nslot;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (elect_one_sync()) {
      if (PAIR) {  // mapB's box is 48 rows here: rows [96 r, 96 r + 96) and [48 r, 48 r + 48) of every 192-row weight tile
        if (rank == 0) mbar_expect_tx(b_full, (uint32_t)a.nbtiles * TZ_BTILE2 * 2);
        for (int i = 0; i < a.nbtiles; ++i) {
          tma2_load_2d(sB + i * TZ_BTILE2, &mapB, b_full, 0, i * 192 + (int)rank * 96);
          tma2_load_2d(sB + i * TZ_BTILE2 + 48 * 128, &mapB, b_full, 0, i * 192 + (int)rank * 96 + 48);
          tma2_load_2d(sB + i * TZ_BTILE2 + 96 * 128, &mapB, b_full, 0, i * 192 + (int)rank * 48);
        }
      } else if (LAST) {  // one box: NST*NSH blocks x 16 rows
        mbar_expect_tx(b_full, (uint32_t)(NST * NSH) * 2048);
        tma_load_2d(sB, &mapB, b_full, 0, 0);
      } else {
        mbar_expect_tx(b_full, (uint32_t)a.nbtiles * TZ_BTILE);
        for (int i = 0; i < a.nbtiles; ++i) tma_load_2d(sB + i * TZ_BTILE, &mapB, b_full, 0, i * 192);
      }
      uint32_t slot = 0, ph = 0;  // ring position and phase of the next entry
      TzSched<SOUTR, NCTA> sched(a);
        TzItem w;
        while (sched.next(a, w)) {
        for (int n = 0; n < w.ntiles; ++n) {
          const int row0 = w.orow0 + (int)rank * OUTR + n * a.srow - TZ;  // GEMM row 0 of the tile
          for (int s = n == 0 ? 0 : nst - a.G; s < nst; ++s) {  // later tiles of a strip only load their newest stage group
            mbar_wait(a_empty + slot, ph ^ 1);
            const CUtensorMap* src = a.st_src[s] ? &mapA1 : &mapA0;
            if (PAIR) {  // both CTAs' boxes are counted on the leader's barrier, armed by the leader for both
              if (rank == 0) mbar_expect_tx(a_full + slot, (uint32_t)a.box_bytes * 2);
              tma2_load_3d(sA + slot * a.slot_bytes, src, a_full + slot, 0, row0 + a.st_off[s], w.b);
              if (++slot == nslot) slot = 0, ph ^= 1;
              continue;
            }
            if ((a.dbg & 1) && ph) {  // timing experiment: the ring keeps its first contents
              mbar_arrive(a_full + slot);
              if (++slot == nslot) slot = 0, ph ^= 1;
              continue;
            }
            mbar_expect_tx(a_full + slot, (uint32_t)a.box_bytes);
            // 3D map {32 words, R, B}: rows outside [0, R) come back as zeros (the padding of the slowest axis)
            asm volatile(
                "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], "
                "[%2];" ::"r"(smem_u32(sA + slot * a.slot_bytes)),
                "l"(reinterpret_cast<uint64_t>(src)), "r"(smem_u32(a_full + slot)), "r"(0), "r"(row0 + a.st_off[s]), "r"(w.b)
                : "memory");
            if (++slot == nslot) slot = 0, ph ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (!PAIR || rank == 0) {
    // ================================ MMA issuer ================================
    constexpr uint32_t MM = PAIR ? 256u : 128u;
    const uint32_t idesc192 = (1u << 4) | ((192u >> 3) << 17) | ((MM >> 4) << 24);  // f16 x f16 -> f32, K-major, M = 128 (256 per pair)
    const uint32_t idesc96 = (1u << 4) | ((96u >> 3) << 17) | ((MM >> 4) << 24);
    const uint32_t idesc16 = (1u << 4) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t desc_hi = ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);  // SBO, version, SW128
    const uint32_t b_lo = ((smem_u32(sB) & 0x3FFFF) >> 4) | (1u << 16);
    const uint32_t shq[3] = {(uint32_t)a.shift_rows[0] * 8, (uint32_t)a.shift_rows[1] * 8, (uint32_t)a.shift_rows[2] * 8};
    const uint32_t a_lo0 = ((smem_u32(sA) & 0x3FFFF) >> 4) | (1u << 16), slot16 = (uint32_t)a.slot_bytes >> 4;
    const int G = a.G;
    // ONE elected lane runs the whole schedule: a per-stage elect + reconvergence costs ~100 cycles, which is nothing next to twelve
    // 96-cycle MMAs but was half of a tile's time in the closing C -> 1 layers (12 small MMAs per tile)
    if (elect_one_sync()) {
    mbar_wait(b_full, 0);
    uint32_t bslot = 0, bph = 0, ti = 0;  // ring position / phase of stage 0 of the current tile
    auto advance = [&](uint32_t k) {
      bslot += k;
      if (bslot >= nslot) bslot -= nslot, bph ^= 1;
    };
    TzSched<SOUTR, NCTA> sched(a);
      TzItem w;
      while (sched.next(a, w)) {
      for (int n = 0; n < w.ntiles; ++n, ++ti, advance(G)) {
        const bool last = n == w.ntiles - 1;
        const uint32_t tb = ti & 1;
        mbar_wait(t_empty + tb, ((ti >> 1) & 1) ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_main = tmem + tb * (LAST ? 32 : 256);
        uint32_t slot = bslot, ph = bph;
#pragma unroll
        for (int s = 0; s < NST; ++s) {
          mbar_wait(a_full + slot, ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          {
            const uint32_t x_lo = a_lo0 + slot * slot16;
#pragma unroll
            for (int k = 0; k < NSH; ++k) {
              const int blk = s * NSH + k;
              const uint32_t wa = x_lo + shq[k];                                                  // window base (16-byte units)
              if (LAST) {
                const uint32_t wb = b_lo + (uint32_t)blk * (2048 >> 4);
                if (a.dbg & 8) continue;  // timing experiment: no MMAs, only the commits
#pragma unroll
                for (int u = 0; u < 4; ++u)  // xh k-steps against B1, xl k-steps against B2, all into the same 16 columns
                  tz_mma(d_main, desc_hi | (uint64_t)(wa + 2 * u), desc_hi | (uint64_t)(wb + 2 * u), idesc16, (s | k | u) == 0 ? 0u : 1u);
              } else if (PAIR) {
                const uint32_t wb = b_lo + (uint32_t)(blk >> 1) * (TZ_BTILE2 >> 4) + (blk & 1) * 4;  // this CTA's half of [wh | wl]
                const uint32_t wc = wb + ((96 * 128) >> 4);                                          // this CTA's half of wh
                tz_mma2(d_main, desc_hi | (uint64_t)(wa + 0), desc_hi | (uint64_t)(wb + 0), idesc192, (s | k) == 0 ? 0u : 1u);
                tz_mma2(d_main, desc_hi | (uint64_t)(wa + 2), desc_hi | (uint64_t)(wb + 2), idesc192, 1u);
                tz_mma2(d_main + 96, desc_hi | (uint64_t)(wa + 4), desc_hi | (uint64_t)(wc + 0), idesc96, 1u);
                tz_mma2(d_main + 96, desc_hi | (uint64_t)(wa + 6), desc_hi | (uint64_t)(wc + 2), idesc96, 1u);
              } else {
                const uint32_t wb = b_lo + (uint32_t)(blk >> 1) * (TZ_BTILE >> 4) + (blk & 1) * 4;  // weight block
                tz_mma(d_main, desc_hi | (uint64_t)(wa + 0), desc_hi | (uint64_t)(wb + 0), idesc192, (s | k) == 0 ? 0u : 1u);
                tz_mma(d_main, desc_hi | (uint64_t)(wa + 2), desc_hi | (uint64_t)(wb + 2), idesc192, 1u);
                tz_mma(d_main + 96, desc_hi | (uint64_t)(wa + 4), desc_hi | (uint64_t)(wb + 0), idesc96, 1u);
                tz_mma(d_main + 96, desc_hi | (uint64_t)(wa + 6), desc_hi | (uint64_t)(wb + 2), idesc96, 1u);
              }
            }
            if (PAIR) {
              if (s < G || last) tz_commit2(a_empty + slot);
              if (s == NST - 1) tz_commit2(t_full + tb);
            } else {
              if (s < G || last) tz_commit(a_empty + slot);  // the other boxes are stages s - G of the next tile of the strip
              if (s == NST - 1) tz_commit(t_full + tb);
            }
          }
          if (++slot == nslot) slot = 0, ph ^= 1;
        }
      }
      advance(nst - G);  // the strip's last tile consumed all of its entries
    }
    }
    __syncwarp();
    }
  } else if (LAST) {
    // ================================ epilogue of the C -> 1 layer ================================
    // two groups of four warps (2..5, 6..9) take alternate tiles (group = accumulator = ti & 1): a tile's epilogue is a long
    // dependent chain (TMEM load, shuffles, skip load, scattered store) while its MMAs take only ~600 cycles
    {
      const int grp = (warp - 2) >> 2;
      const int q = warp & 3;
      const int j = q * 32 + lane;  // GEMM row; produces output row j (relative to the tile's first output row), valid for j < OUTR
      const float c0 = __ldg(a.scales) * a.out_mul, c1 = __ldg(a.scales + 1) * a.out_mul;
      const bool has1 = lane + TZ < 32, has2 = lane + 2 * TZ < 32;
      const int Hdim = a.R / (a.n0.d * a.n1.d);  // interior length of the slowest axis
      float* xg = xch + grp * (4 * 3 * TZ);
      uint32_t ti = 0;
      TzSched<OUTR> sched(a);
      TzItem w;
      while (sched.next(a, w)) {
        for (int n = 0; n < w.ntiles; ++n, ++ti) {
          if ((int)(ti & 1) != grp) continue;
          const int orow0 = w.orow0 + n * a.srow;
          // output coordinates and the skip value do not depend on the accumulator: fetch them before waiting for it
          const int r = orow0 + j;
          int rq, c0i, c1i, c2i;
          fdivmod(r, a.n0, rq, c0i);
          fdivmod(rq, a.n1, c2i, c1i);
          const bool ok = j < OUTR && c0i >= a.p0 && c0i < a.p0 + a.i0 && c1i >= a.p1 && c1i < a.p1 + a.i1 && c2i < Hdim;
          // NCDHW index for rows ordered (slowest = y, middle = x, fastest = d); NCHW for rows ordered (y, x)
          const long long o = a.out_mode ? ((long long)w.b * a.i1 + (c1i - a.p1)) * a.i0 + (c0i - a.p0)
                                         : (((long long)w.b * a.i0 + (c0i - a.p0)) * Hdim + c2i) * a.i1 + (c1i - a.p1);
          const float sk = (ok && a.skip) ? __ldg(a.skip + o) : 0.f;
          mbar_wait(t_full + grp, (ti >> 1) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (a.dbg & 2) {  // timing experiment: free the accumulator at once
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(t_empty + grp);
            continue;
          }
          float m[8], k[8];
          tz_ld8(tmem + ((uint32_t)(q * 32) << 16) + grp * 32, m);
          tz_ld8(tmem + ((uint32_t)(q * 32) << 16) + grp * 32 + 8, k);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(t_empty + grp);
          const float e0 = fmaf(k[0], c1, m[0] * c0), e1 = fmaf(k[1], c1, m[1] * c0), e2 = fmaf(k[2], c1, m[2] * c0);
          const float s1 = __shfl_down_sync(0xffffffffu, e1, TZ), s2 = __shfl_down_sync(0xffffffffu, e2, (2 * TZ) & 31);
          float v = e0 + (has1 ? s1 : 0.f) + (has2 ? s2 : 0.f);
          float* xq = xg + q * (3 * TZ);
          named_bar_sync(1 + grp, 128);  // the group's previous tile has been read by everyone
          if (lane < TZ) xq[lane] = e1;
          if (lane < 2 * TZ) xq[TZ + lane] = e2;
          named_bar_sync(1 + grp, 128);
          if (q < 3) {
            const float* xn = xg + (q + 1) * (3 * TZ);
            if (!has1) v += xn[lane + TZ - 32];
            if (!has2) v += xn[TZ + lane + 2 * TZ - 32];
          }
          if (ok) a.out_f32[o] = v + sk;
        }
      }
    }
  } else {
    // ================================ epilogue (warps 2..9) ================================
    const int q = warp & 3;           // TMEM lane quarter
    const int hf = (warp - 2) >> 2;   // which 16 of the 32 output channels
    const int j = q * 32 + lane;      // GEMM row of the tile; this thread produces output row j + TZ -> staging row j
    const bool issuer = warp == 2 && lane == 0;
    const float c0 = __ldg(a.scales) * a.out_mul, c1 = __ldg(a.scales + 1) * a.out_mul;
    const float relu_lo = a.relu ? 0.f : -INFINITY;
    const bool has1 = lane + TZ < 32, has2 = lane + 2 * TZ < 32;
    float bias[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) bias[c] = __ldg(a.bias + hf * 16 + c) * a.bias_mul;
    // quarter q publishes for quarter q-1: rows [0,TZ) = e1 of lanes 0..TZ-1, rows [TZ,3TZ) = e2 of lanes 0..2TZ-1
    float* xq = xch + q * (3 * TZ * TZ_XS) + hf * 16;
    const float* xn = xch + ((q + 1) & 3) * (3 * TZ * TZ_XS) + hf * 16;
    uint32_t ti = 0;
    TzSched<SOUTR, NCTA> sched(a);
      TzItem w;
      while (sched.next(a, w)) {
      for (int n = 0; n < w.ntiles; ++n, ++ti) {
        const uint32_t tb = ti & 1;
        const int orow0 = w.orow0 + (int)rank * OUTR + n * a.srow;  // first output row of the tile
        mbar_wait(t_full + tb, (ti >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + tb * 256 + hf * 16;
        if (a.dbg & 2) {  // timing experiment: free the accumulator at once
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            if (PAIR) mbar_arrive_leader(t_empty + tb);
            else mbar_arrive(t_empty + tb);
          }
          continue;
        }
        float out[16];
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {  // 8 output channels at a time
          float m0[8], m1[8], m2[8], k0[8], k1[8], k2[8];
          tz_ld8(taddr + cc * 8, m0);
          tz_ld8(taddr + 32 + cc * 8, m1);
          tz_ld8(taddr + 64 + cc * 8, m2);
          tz_ld8(taddr + 96 + cc * 8, k0);
          tz_ld8(taddr + 128 + cc * 8, k1);
          tz_ld8(taddr + 160 + cc * 8, k2);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          // tap t of GEMM row j contributes to staging row j - t*TZ:
          // staging row j = e0 of GEMM row j + e1 of GEMM row j + TZ + e2 of GEMM row j + 2 TZ
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float e0 = fmaf(k0[c], c1, m0[c] * c0);
            m1[c] = fmaf(k1[c], c1, m1[c] * c0);
            m2[c] = fmaf(k2[c], c1, m2[c] * c0);
            const float s1 = __shfl_down_sync(0xffffffffu, m1[c], TZ);
            const float s2 = __shfl_down_sync(0xffffffffu, m2[c], (2 * TZ) & 31);
            out[cc * 8 + c] = e0 + (has1 ? s1 : 0.f) + (has2 ? s2 : 0.f);
          }
          if (lane < 2 * TZ) {  // rows the previous quarter needs
            float4* d2 = reinterpret_cast<float4*>(xq + (TZ + lane) * TZ_XS + cc * 8);
            d2[0] = make_float4(m2[0], m2[1], m2[2], m2[3]), d2[1] = make_float4(m2[4], m2[5], m2[6], m2[7]);
            if (lane < TZ) {
              float4* d1 = reinterpret_cast<float4*>(xq + lane * TZ_XS + cc * 8);
              d1[0] = make_float4(m1[0], m1[1], m1[2], m1[3]), d1[1] = make_float4(m1[4], m1[5], m1[6], m1[7]);
            }
          }
          __syncwarp();
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          if (PAIR) mbar_arrive_leader(t_empty + tb);
          else mbar_arrive(t_empty + tb);
        }
        if (a.dbg & 4) {  // timing experiment: no staging, no store
          if (out[0] == 123.456f) a.out_f32[0] = out[1];
          continue;
        }
        // previous tile's TMA store must have finished reading the staging tile before anyone overwrites it
        if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        named_bar_sync(1, 256);
        if (q < 3) {  // rows of the next quarter (the last quarter's missing rows belong to the next tile)
          if (!has1) {
            const float4* p1 = reinterpret_cast<const float4*>(xn + (lane + TZ - 32) * TZ_XS);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const float4 v = p1[c];
              out[4 * c] += v.x, out[4 * c + 1] += v.y, out[4 * c + 2] += v.z, out[4 * c + 3] += v.w;
            }
          }
          if (!has2) {
            const float4* p2 = reinterpret_cast<const float4*>(xn + (TZ + lane + 2 * TZ - 32) * TZ_XS);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const float4 v = p2[c];
              out[4 * c] += v.x, out[4 * c + 1] += v.y, out[4 * c + 2] += v.z, out[4 * c + 3] += v.w;
            }
          }
        }
        // bias, ReLU, border; output row index inside the batch element
        const int r = orow0 + j;
        int rq, c0i, c1i, rq2;
        fdivmod(r, a.n0, rq, c0i);
        fdivmod(rq, a.n1, rq2, c1i);
        const bool border = c0i < a.p0 || c0i >= a.p0 + a.i0 || c1i < a.p1 || c1i >= a.p1 + a.i1;
        const uint32_t so = smem_u32(sOut) + j * 128;
        if (a.out_split) {
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            const int c8 = hf * 2 + cc;
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int p = 0; p < 4; ++p) {
              const int c = cc * 8 + 2 * p;
              float v0 = fmaxf(out[c] + bias[c], relu_lo);
              float v1 = fmaxf(out[c + 1] + bias[c + 1], relu_lo);
              v0 = border ? 0.f : v0, v1 = border ? 0.f : v1;
              const __half2 h = f2h2_sat(v0, v1);
              const float2 f = __half22float2(h);
              const float2 dl = split_lo2(v0, v1, f);
              const __half2 l = f2h2_sat(dl.x, dl.y);
              hi[p] = *reinterpret_cast<const uint32_t*>(&h), lo[p] = *reinterpret_cast<const uint32_t*>(&l);
            }
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(so + ((c8 ^ (j & 7)) << 4)), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]),
                         "r"(hi[3])
                         : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(so + (((c8 + 4) ^ (j & 7)) << 4)), "r"(lo[0]), "r"(lo[1]),
                         "r"(lo[2]), "r"(lo[3])
                         : "memory");
          }
        } else {
#pragma unroll
          for (int cq = 0; cq < 4; ++cq) {
            const int c4 = hf * 4 + cq;
            float v[4];
#pragma unroll
            for (int p = 0; p < 4; ++p) {
              const int c = cq * 4 + p;
              const float t = fmaxf(out[c] + bias[c], relu_lo);
              v[p] = border ? 0.f : t;
            }
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(so + ((c4 ^ (j & 7)) << 4)), "f"(v[0]), "f"(v[1]), "f"(v[2]),
                         "f"(v[3])
                         : "memory");
          }
        }
        fence_proxy_async_smem();
        named_bar_sync(1, 256);
        if (issuer) {
          asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                           reinterpret_cast<uint64_t>(&mapOut)),
                       "r"(smem_u32(sOut)), "r"(0), "r"(orow0), "r"(w.b)
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    }
    if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (PAIR) {
    cluster_sync_all();  // the leader's MMAs read this CTA's shared memory and write its TMEM until the pair's last tile
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
  } else if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
  }
}

// ---- host: one Toeplitz-N GEMM layer (TzLayer: conv3d_f16.cuh) ------------------------------------------------------------
static size_t tz_smem_bytes(int nbtiles, int nslot, int slot_bytes, int tz, bool pair = false) {
  return (size_t)nbtiles * (pair ? TZ_BTILE2 : TZ_BTILE) + 16384 + (size_t)nslot * slot_bytes + 4 * 3 * tz * TZ_XS * 4 + 256 + 1024;
}

int launch_tz_gemm(const TzLayer& L, cudaStream_t st) {
  if ((L.tz != 1 && L.tz != 8) || L.nstages < 1 || L.nstages > TZ_MAXST || L.nshift < 1 || L.nshift > 3 || L.box_rows > 256)
    return LWS_ERR_UNSUPPORTED;
  TzArgs a;
  memset(&a, 0, sizeof(a));
  const int nblk = L.nstages * L.nshift;
  a.nbtiles = L.last ? 1 : (nblk + 1) / 2;
  if (L.last && nblk * 2048 > TZ_BTILE) return LWS_ERR_UNSUPPORTED;
  a.box_bytes = L.box_rows * 128;
  a.slot_bytes = (a.box_bytes + 1023) / 1024 * 1024;
  const int outr = 128 - 2 * L.tz;
  const bool is3d = L.tz == 1 && L.nstages == 3 && L.nshift == 3, is2d = L.tz == 8 && L.nstages == 6 && L.nshift == 1 && !L.last;
  const bool is2dlast = L.tz == 1 && L.nstages == 3 && L.nshift == 1 && L.last;
  if (!is3d && !is2d && !is2dlast) return LWS_ERR_UNSUPPORTED;
  // CTA pairs (cta_group::2): the 32 -> 32 3D layers on the strip schedule; a schedule tile is the pair's 2 * outr rows
  const bool pair = L.pair && is3d && !L.last && L.srow > 0;
  const int soutr = pair ? 2 * outr : outr;
  a.nstages = L.nstages, a.nshift = L.nshift;
  if (L.srow > 0) {  // strips along the slowest tap axis
    a.G = L.G, a.srow = L.srow;
    a.ct_per_sl = make_fastdiv((L.srow + soutr - 1) / soutr);
    a.strip_len = (L.R + L.srow - 1) / L.srow;
  } else {
    a.G = L.nstages, a.srow = L.R;
    a.ct_per_sl = make_fastdiv((L.R + outr - 1) / outr);
    a.strip_len = 1;
  }
  a.total_tiles = L.B * a.ct_per_sl.d * a.strip_len;
  a.dbg = opt(OPT_TZ_DEBUG);
  const int grid = pair ? (2 * a.total_tiles < kNumSMs ? 2 * a.total_tiles : kNumSMs / 2 * 2) : (a.total_tiles < kNumSMs ? a.total_tiles : kNumSMs);
  if (a.strip_len > 1 && L.seg_mode && !pair) {
    // segments per strip: the count that minimises the busiest CTA's tiles + box reloads (a segment start loads nstages - G boxes
    // more than a tile inside a strip; counted as (nstages - G) / nstages of a tile)
    double best = 1e30;
    for (int sp = 1; sp <= a.strip_len && sp <= 16; ++sp) {
      const int sl = (a.strip_len + sp - 1) / sp, spe = (a.strip_len + sl - 1) / sl;
      const long long nseg = (long long)L.B * spe * a.ct_per_sl.d;
      const double cost = (double)((nseg + grid - 1) / grid) * (sl + (double)(L.nstages - a.G) / L.nstages);
      if (cost < best) best = cost, a.seg_len = sl, a.segs_per_strip = spe, a.nseg = (int)nseg;
    }
  }
  // ring: one tile's stages plus as many more as fit (at most 8)
  a.nslot = 8;
  while (a.nslot > L.nstages && tz_smem_bytes(a.nbtiles, a.nslot, a.slot_bytes, L.tz, pair) > 232448) --a.nslot;
  const size_t smem = tz_smem_bytes(a.nbtiles, a.nslot, a.slot_bytes, L.tz, pair);
  if (smem > 232448 || a.nslot < L.nstages) return LWS_ERR_UNSUPPORTED;
  // every variant may use up to the 227 KB cap (the per-call size is `smem`): set once per variant and device
  if (is2dlast) LWS_SET_SMEM_ONCE((tz_gemm_kernel<1, 3, 1, true>), 232448);
  else if (is2d) LWS_SET_SMEM_ONCE((tz_gemm_kernel<8, 6, 1, false>), 232448);
  else if (L.last) LWS_SET_SMEM_ONCE((tz_gemm_kernel<1, 3, 3, true>), 232448);
  else if (pair) LWS_SET_SMEM_ONCE((tz_gemm_kernel<1, 3, 3, false, true>), 232448);
  else LWS_SET_SMEM_ONCE((tz_gemm_kernel<1, 3, 3, false>), 232448);
  cudaError_t e;
  a.bias = L.bias, a.scales = L.wtab + (L.last ? (size_t)nblk * 16 * 32 : (size_t)a.nbtiles * 192 * 32);
  a.skip = L.skip, a.out_f32 = L.out_f32, a.out_mode = L.out_mode;
  a.out_split = L.out_split, a.relu = L.relu;
  a.out_mul = L.out_split ? 1.f : 1.f / kDwsepActScale;
  a.bias_mul = L.out_split ? kDwsepActScale : 1.f;
  a.R = L.R, a.n0 = make_fastdiv(L.n0), a.p0 = L.p0, a.i0 = L.i0, a.n1 = make_fastdiv(L.n1), a.p1 = L.p1, a.i1 = L.i1;
  for (int s = 0; s < L.nstages; ++s) a.st_off[s] = L.st_off[s], a.st_src[s] = L.st_src[s];
  for (int k = 0; k < 3; ++k) a.shift_rows[k] = k < L.nshift ? L.shift_rows[k] : 0;
  CUtensorMap mapA0, mapA1, mapB, mapOut;
  const uint64_t dimsA[3] = {32, (uint64_t)L.R, (uint64_t)L.B}, strA[2] = {128, (uint64_t)L.R * 128};
  const uint32_t boxA[3] = {32, (uint32_t)L.box_rows, 1}, boxO[3] = {32, (uint32_t)outr, 1};
  int rc = make_tensor_map_f32(&mapA0, L.src0, 3, dimsA, strA, boxA, true);
  if (rc) return rc;
  rc = make_tensor_map_f32(&mapA1, L.src1, 3, dimsA, strA, boxA, true);
  if (rc) return rc;
  rc = make_tensor_map_f32(&mapOut, L.last ? L.src0 : L.out, 3, dimsA, strA, boxO, true);  // unused by the C -> 1 layer
  if (rc) return rc;
  const uint64_t dimsB[2] = {32, L.last ? (uint64_t)nblk * 16 : (uint64_t)a.nbtiles * 192}, strB[1] = {128};
  const uint32_t boxB[2] = {32, L.last ? (uint32_t)nblk * 16 : (pair ? 48u : 192u)};
  rc = make_tensor_map_f32(&mapB, L.wtab, 2, dimsB, strB, boxB, true);
  if (rc) return rc;
  if (pair) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid), cfg.blockDim = dim3(TZ_THREADS), cfg.dynamicSmemBytes = smem, cfg.stream = st;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2, attr.val.clusterDim.y = 1, attr.val.clusterDim.z = 1;
    cfg.attrs = &attr, cfg.numAttrs = 1;
    // a persistent grid must be co-resident: GPCs with an odd number of usable SMs leave one SM without a partner
    static int max_clusters[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && max_clusters[dev] == 0) {
      int n = 0;
      cfg.gridDim = dim3(kNumSMs / 2 * 2);
      if (cudaOccupancyMaxActiveClusters(&n, tz_gemm_kernel<1, 3, 3, false, true>, &cfg) != cudaSuccess || n <= 0) n = kNumSMs / 2;
      max_clusters[dev] = n;
    }
    const int maxc = dev >= 0 && dev < 64 ? max_clusters[dev] : kNumSMs / 2;
    cfg.gridDim = dim3(grid < 2 * maxc ? grid : 2 * maxc);
    e = cudaLaunchKernelEx(&cfg, tz_gemm_kernel<1, 3, 3, false, true>, mapA0, mapA1, mapB, mapOut, a);
    return e == cudaSuccess ? LWS_OK : (int)e;
  }
  if (is2dlast) tz_gemm_kernel<1, 3, 1, true><<<grid, TZ_THREADS, smem, st>>>(mapA0, mapA1, mapB, mapOut, a);
  else if (is2d) tz_gemm_kernel<8, 6, 1, false><<<grid, TZ_THREADS, smem, st>>>(mapA0, mapA1, mapB, mapOut, a);
  else if (L.last) tz_gemm_kernel<1, 3, 3, true><<<grid, TZ_THREADS, smem, st>>>(mapA0, mapA1, mapB, mapOut, a);
  else tz_gemm_kernel<1, 3, 3, false><<<grid, TZ_THREADS, smem, st>>>(mapA0, mapA1, mapB, mapOut, a);
  e = cudaPeekAtLastError();
  return e == cudaSuccess ? LWS_OK : (int)e;
}

// ---- 3D stack, C = 32: rows ordered (y, x, d), every d column and every x line preceded by one zero voxel ---------------------
// The zero voxel that FOLLOWS a column (line) is the one that precedes the next column (line) -- kPadT = 0 -- and the rows past the end
// of the row space are the TMA's out-of-bounds zero fill: (W + 1)(D + 1) rows per image line instead of (W + 2)(D + 2), 4.5 % fewer rows
// (and 31 instead of 33 column tiles per line at KITTI's 154 x 24) for kernels that run at the tensor roofline.
constexpr int kPadT = 0;
// first conv 1 -> 32 on the raw cost (BN_0 affine + ReLU applied to the taps), output rows split-fp16 (scaled by sa).
// 4 lanes per voxel group (8 output channels each); a group owns 4 voxels that are neighbours in d (consecutive output rows), so
// its 27-tap window is 6 planes x 3 x 3: cost loads and weight loads are shared 4 ways (54 + 54 per 864 FFMAs per lane).  Groups
// run over (d-group, x) with x fastest so the cost reads of a warp coalesce; grid = (group chunk, y, pair).
__global__ void __launch_bounds__(256)
    conv3d_first_ydx_kernel(const float* __restrict__ cost, const float* __restrict__ w /*[27][32]*/, const float* __restrict__ bias,
                            const float* __restrict__ affine, uint4* __restrict__ out, int D, int H, int W, FastDiv fWp) {
  __shared__ __align__(16) float sW[27 * 32];
  for (int i = threadIdx.x; i < 27 * 32; i += blockDim.x) sW[i] = __ldg(w + i);
  __syncthreads();
  const int sub = threadIdx.x & 3;
  const int Wp = W + 1 + kPadT, Dp = D + 1 + kPadT;
  const int ngd = (Dp + 3) >> 2;
  const int gid = blockIdx.x * 64 + (threadIdx.x >> 2);
  if (gid >= Wp * ngd) return;
  int dpg, xp;
  fdivmod(gid, fWp, dpg, xp);
  const int y = blockIdx.y, b = blockIdx.z;
  const int x = xp - 1, dp0 = dpg * 4;
  const long long hw = (long long)H * W;
  float2 acc[4][4];  // float2 pairs: FFMA2 (fma.rn.f32x2) does two output channels per issued instruction
#pragma unroll
  for (int v = 0; v < 4; ++v)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[v][j] = make_float2(0.f, 0.f);
  const bool xborder = x < 0 || x >= W;
  if (!xborder) {
    const float s0 = __ldg(affine), t0 = __ldg(affine + 1);
    const float* cb = cost + (long long)b * D * hw;  // 32-bit offsets inside one pair's volume (host checks D*H*W < 2^31)
    const int ihw = H * W;
    const float* pr[6];
    bool okr[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      const int dd = dp0 + r - 2;  // input plane of (voxel vd, tap kd) with vd + kd = r  (d = dp - 1)
      okr[r] = (unsigned)dd < (unsigned)D;
      pr[r] = cb + (dd * ihw + y * W + x);
    }
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const bool oky = (unsigned)(y + kh - 1) < (unsigned)H;  // block-uniform
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const bool okx = oky && (unsigned)(x + kw - 1) < (unsigned)W;
        const int tap = (kh - 1) * W + (kw - 1);
        float v[6];
#pragma unroll
        for (int r = 0; r < 6; ++r) {
          const bool ok = okx && okr[r];
          const float c = ok ? __ldg(pr[r] + tap) : 0.f;
          v[r] = ok ? fmaxf(fmaf(c, s0, t0), 0.f) : 0.f;
        }
#pragma unroll
        for (int kd = 0; kd < 3; ++kd) {
          const float4 wa = *reinterpret_cast<const float4*>(sW + (kd * 9 + kh * 3 + kw) * 32 + sub * 8);
          const float4 wb = *reinterpret_cast<const float4*>(sW + (kd * 9 + kh * 3 + kw) * 32 + sub * 8 + 4);
#pragma unroll
          const float2 w01 = make_float2(wa.x, wa.y), w23 = make_float2(wa.z, wa.w);
          const float2 w45 = make_float2(wb.x, wb.y), w67 = make_float2(wb.z, wb.w);
#pragma unroll
          for (int vd = 0; vd < 4; ++vd) {
            const float2 t = make_float2(v[vd + kd], v[vd + kd]);
            acc[vd][0] = __ffma2_rn(t, w01, acc[vd][0]), acc[vd][1] = __ffma2_rn(t, w23, acc[vd][1]);
            acc[vd][2] = __ffma2_rn(t, w45, acc[vd][2]), acc[vd][3] = __ffma2_rn(t, w67, acc[vd][3]);
          }
        }
      }
    }
  }
  float bv[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) bv[j] = __ldg(bias + sub * 8 + j);
  const long long row0 = (((long long)b * H + y) * Wp + xp) * Dp + dp0;
#pragma unroll
  for (int vd = 0; vd < 4; ++vd) {
    const int dp = dp0 + vd;
    if (dp >= Dp) break;
    const bool border = xborder || dp == 0 || dp > D;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const float a0 = border ? 0.f : fmaxf(acc[vd][p].x + bv[2 * p], 0.f) * kDwsepActScale;
      const float a1 = border ? 0.f : fmaxf(acc[vd][p].y + bv[2 * p + 1], 0.f) * kDwsepActScale;
      const __half2 h = f2h2_sat(a0, a1);
      const float2 f = __half22float2(h);
      const float2 dl = split_lo2(a0, a1, f);
      const __half2 l = f2h2_sat(dl.x, dl.y);
      hi[p] = *reinterpret_cast<const uint32_t*>(&h), lo[p] = *reinterpret_cast<const uint32_t*>(&l);
    }
    out[(row0 + vd) * 8 + sub] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    out[(row0 + vd) * 8 + 4 + sub] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

// ---- first conv 1 -> 32, version 2: the activated tap window of a (row, 64-column chunk) staged once in shared memory ----------------
// Version 1 above loads its 54 taps per lane from global memory (address + predicate instructions per load, the BN_0 affine + ReLU
// recomputed per use).  Here a block = (y, 64 padded columns, pair) stages ReLU(BN_0(cost)) of rows y-1..y+1, all planes (two zero
// planes either side, so no tap needs a predicate) and 66 columns; the 64 voxel groups of the block (4 lanes x 8 output channels
// each) then walk the (4-plane group, column) items with immediate-offset LDS taps.
constexpr int F32V2_XT = 64;
__global__ void __launch_bounds__(256, 2)
    conv3d_first_ydx_v2_kernel(const float* __restrict__ cost, const float* __restrict__ w /*[27][32]*/, const float* __restrict__ bias,
                               const float* __restrict__ affine, uint4* __restrict__ out, int D, int H, int W) {
  extern __shared__ __align__(16) float smem_f[];
  float* sW = smem_f;             // [27][32]
  float* sIn = smem_f + 27 * 32;  // [3 kh][NP planes][66 cols]
  constexpr int TC = F32V2_XT + 2;
  const int Wp = W + 1 + kPadT, Dp = D + 1 + kPadT;
  const int ngd = (Dp + 3) >> 2;
  const int NP = 4 * ngd + 2;     // tile plane pi holds input plane pi - 2; the last plane group reads planes up to 4*ngd + 1
  const int y = blockIdx.y, b = blockIdx.z;
  const int xc0 = blockIdx.x * F32V2_XT;
  const int tid = threadIdx.x;
  for (int i = tid; i < 27 * 32; i += 256) sW[i] = __ldg(w + i);
  {
    const float s0 = __ldg(affine), t0 = __ldg(affine + 1);
    const long long hw = (long long)H * W;
    const float* cb = cost + (long long)b * D * hw;
    const int rr = tid / TC, cc = tid - rr * TC;  // 3 tile rows of 66 columns per pass (198 of the 256 threads)
    const int xx = xc0 + cc - 2;
    const bool okx = rr < 3 && (unsigned)xx < (unsigned)W;
    for (int row = rr; row < 3 * NP && rr < 3; row += 3) {
      const int kh = row / NP, pi = row - kh * NP;
      const int dd = pi - 2, yy = y + kh - 1;
      float v = 0.f;
      if (okx && (unsigned)dd < (unsigned)D && (unsigned)yy < (unsigned)H)
        v = fmaxf(fmaf(__ldg(cb + (long long)dd * hw + (long long)yy * W + xx), s0, t0), 0.f);
      sIn[row * TC + cc] = v;
    }
  }
  __syncthreads();
  const int sub = tid & 3, grp = tid >> 2;
  float bv[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) bv[j] = __ldg(bias + sub * 8 + j);
  const int ncols = min(F32V2_XT, Wp - xc0);
  const int nitems = ngd * ncols;
  for (int item = grp; item < nitems; item += 64) {
    const int dpg = item / ncols, i = item - dpg * ncols;  // x fastest: the 8 groups of a warp read 8 neighbouring columns
    const int xp = xc0 + i, dp0 = dpg * 4;
    float2 acc[4][4];
#pragma unroll
    for (int v = 0; v < 4; ++v)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[v][j] = make_float2(0.f, 0.f);
    const bool xborder = xp == 0 || xp > W;
    if (!xborder) {
      // voxel vd (padded plane dp0 + vd, d = dp - 1), tap kd: input plane d + kd - 1 = dp0 + (vd + kd) - 2 -> tile plane dp0 + vd + kd
      const float* base = sIn + dp0 * TC + i;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          float v[6];
#pragma unroll
          for (int r = 0; r < 6; ++r) v[r] = base[(kh * NP + r) * TC + kw];
#pragma unroll
          for (int kd = 0; kd < 3; ++kd) {
            const float4 wa = *reinterpret_cast<const float4*>(sW + (kd * 9 + kh * 3 + kw) * 32 + sub * 8);
            const float4 wb = *reinterpret_cast<const float4*>(sW + (kd * 9 + kh * 3 + kw) * 32 + sub * 8 + 4);
            const float2 w01 = make_float2(wa.x, wa.y), w23 = make_float2(wa.z, wa.w);
            const float2 w45 = make_float2(wb.x, wb.y), w67 = make_float2(wb.z, wb.w);
#pragma unroll
            for (int vd = 0; vd < 4; ++vd) {
              const float2 t = make_float2(v[vd + kd], v[vd + kd]);
              acc[vd][0] = __ffma2_rn(t, w01, acc[vd][0]), acc[vd][1] = __ffma2_rn(t, w23, acc[vd][1]);
              acc[vd][2] = __ffma2_rn(t, w45, acc[vd][2]), acc[vd][3] = __ffma2_rn(t, w67, acc[vd][3]);
            }
          }
        }
      }
    }
    const long long row0 = (((long long)b * H + y) * Wp + xp) * Dp + dp0;
#pragma unroll
    for (int vd = 0; vd < 4; ++vd) {
      const int dp = dp0 + vd;
      if (dp >= Dp) break;
      const bool border = xborder || dp == 0 || dp > D;
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const float a0 = border ? 0.f : fmaxf(acc[vd][p].x + bv[2 * p], 0.f) * kDwsepActScale;
        const float a1 = border ? 0.f : fmaxf(acc[vd][p].y + bv[2 * p + 1], 0.f) * kDwsepActScale;
        const __half2 h = f2h2_sat(a0, a1);
        const float2 f = __half22float2(h);
        const float2 dl = split_lo2(a0, a1, f);
        const __half2 l = f2h2_sat(dl.x, dl.y);
        hi[p] = *reinterpret_cast<const uint32_t*>(&h), lo[p] = *reinterpret_cast<const uint32_t*>(&l);
      }
      out[(row0 + vd) * 8 + sub] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      out[(row0 + vd) * 8 + 4 + sub] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
  }
}

// ---- north_star item 2: the stage-1 volume build FUSED into the first conv ------------------------------------------------------
// cost[b,d,y,x] = sum_c |L[b,c,y,x] - R[b,c,y,x-d]| (LWSNet._build_volume_2d, models/models.py:58-76) is computed straight into the
// shared-memory tap window of the first 1 -> 32 conv (post_3dconvs, models/submodules.py:216-218): a block = (TY = 4 output rows,
// 64 padded columns, pair) builds rows y0-1 .. y0+TY of the volume with the register-window scheme of cost_volume.cu (a thread-tile
// = 1 row x 4 pixels x 8 disparities, 64-bit loads of L and of the aligned R window, 2*C FADD-class operations per value), writes
// the rows it owns to HBM once (the skip connection of models/models.py:137 needs the raw volume) and keeps ReLU(BN_0(cost)) for the
// conv.  The volume is rebuilt (TY+2)/TY = 1.5x (halo rows), which costs ~8 % more instructions than the conv alone; in exchange the
// stand-alone volume launch and its re-read disappear.  Same FMA order as the unfused kernels: bit-identical.
constexpr int FUS_TY = 4;
template <int DT>
__global__ void __launch_bounds__(256, 2)
    cost_first_conv_fused_kernel(const float* __restrict__ L, const float* __restrict__ R, float* __restrict__ cost,
                                 const float* __restrict__ w /*[27][32]*/, const float* __restrict__ bias, const float* __restrict__ affine,
                                 uint4* __restrict__ out, int Cf, int D, int H, int W) {
  extern __shared__ __align__(16) float smem_f[];
  float* sW = smem_f;             // [27][32]
  float* sIn = smem_f + 27 * 32;  // [TY + 2 rows][NP planes][66 cols]: ReLU(BN_0(cost)), zero outside the volume
  constexpr int TC = F32V2_XT + 2, TR = FUS_TY + 2;
  const int Wp = W + 1 + kPadT, Dp = D + 1 + kPadT;
  const int ngd = (Dp + 3) >> 2;
  const int NP = 4 * ngd + 2;  // tile plane pi holds input plane d = pi - 2
  const int y0 = blockIdx.y * FUS_TY, b = blockIdx.z;
  const int xc0 = blockIdx.x * F32V2_XT;
  const int tid = threadIdx.x;
  const long long hw = (long long)H * W;
  for (int i = tid; i < 27 * 32; i += 256) sW[i] = __ldg(w + i);
  for (int i = tid; i < TR * NP * TC; i += 256) sIn[i] = 0.f;
  __syncthreads();
  {
    // ---- phase 1: the cost window.  thread-tile = (row r, pixel quad xq, disparity tile dt of DT); window column c <-> x = xc0 - 2 + c
    const float s0 = __ldg(affine), t0 = __ldg(affine + 1);
    constexpr int NQ = (TC + 3) / 4;  // 17 pixel quads (the last one covers 2 columns beyond the window)
    const int ndt = (D + DT - 1) / DT;
    const float* Lb = L + (long long)b * Cf * hw;
    const float* Rb = R + (long long)b * Cf * hw;
    float* cb = cost + (long long)b * D * hw;
    for (int tile = tid; tile < TR * NQ * ndt; tile += 256) {
      const int dt = tile % ndt, rq = tile / ndt, xq = rq % NQ, r = rq / NQ;
      const int yy = y0 + r - 1, x0 = xc0 - 2 + xq * 4, d0 = dt * DT;
      if ((unsigned)yy >= (unsigned)H || x0 >= W || x0 + 3 < 0) continue;  // rows / quads outside the image stay zero
      const float* Lp = Lb + (long long)yy * W + x0;
      const int xs = x0 - d0 - DT;  // wv[i] = R[xs + i]; R[x0 + p - (d0 + j)] = wv[p - j + DT]
      const float* Rp = Rb + (long long)yy * W + xs;
      constexpr int NW = (DT + 4) / 2;
      bool okw[NW];
#pragma unroll
      for (int i = 0; i < NW; ++i) okw[i] = xs + 2 * i >= 0 && xs + 2 * i < W;  // xs and W are even: a pair is wholly in or out
      const bool okl0 = x0 >= 0, okl1 = x0 + 2 < W;  // x0 is even
      float acc[4][DT];
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int j = 0; j < DT; ++j) acc[p][j] = 0.f;
      for (int c = 0; c < Cf; ++c) {
        const float2 a = okl0 ? __ldg(reinterpret_cast<const float2*>(Lp + c * hw)) : make_float2(0.f, 0.f);
        const float2 bq = okl1 ? __ldg(reinterpret_cast<const float2*>(Lp + c * hw + 2)) : make_float2(0.f, 0.f);
        const float l[4] = {a.x, a.y, bq.x, bq.y};
        float wv[DT + 4];
#pragma unroll
        for (int i = 0; i < NW; ++i) {
          const float2 v = okw[i] ? __ldg(reinterpret_cast<const float2*>(Rp + c * hw) + i) : make_float2(0.f, 0.f);
          wv[2 * i] = v.x, wv[2 * i + 1] = v.y;
        }
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
          for (int j = 0; j < DT; ++j) acc[p][j] += fabsf(l[p] - wv[p - j + DT]);
      }
      const bool own_row = r >= 1 && r <= FUS_TY;  // rows y0 .. y0+TY-1 belong to this block (the halo rows to its neighbours)
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const int col = xq * 4 + p, xx = x0 + p;
        if (col >= TC || (unsigned)xx >= (unsigned)W) continue;
        const bool own = own_row && col >= 1 && col <= F32V2_XT;  // x = xp - 1 for this chunk's padded columns xp
#pragma unroll
        for (int j = 0; j < DT; ++j) {
          const int d = d0 + j;
          if (d >= D) break;
          if (own) cb[(long long)d * hw + (long long)yy * W + xx] = acc[p][j];
          sIn[(r * NP + d + 2) * TC + col] = fmaxf(fmaf(acc[p][j], s0, t0), 0.f);
        }
      }
    }
  }
  __syncthreads();
  // ---- phase 2: the 1 -> 32 conv from the window (as conv3d_first_ydx_v2_kernel, over TY rows) ---------------------------------------
  const int sub = tid & 3, grp = tid >> 2;
  float bv[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) bv[j] = __ldg(bias + sub * 8 + j);
  const int ncols = min(F32V2_XT, Wp - xc0);
  const int per_row = ngd * ncols;
  const int nrows = min(FUS_TY, H - y0);
  for (int item = grp; item < nrows * per_row; item += 64) {
    const int ty = item / per_row, rem = item - ty * per_row;
    const int dpg = rem / ncols, i = rem - dpg * ncols;
    const int xp = xc0 + i, dp0 = dpg * 4, y = y0 + ty;
    float2 acc[4][4];
#pragma unroll
    for (int v = 0; v < 4; ++v)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[v][j] = make_float2(0.f, 0.f);
    const bool xborder = xp == 0 || xp > W;
    if (!xborder) {
      const float* base = sIn + (ty * NP + dp0) * TC + i;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          float v[6];
#pragma unroll
          for (int r = 0; r < 6; ++r) v[r] = base[(kh * NP + r) * TC + kw];
#pragma unroll
          for (int kd = 0; kd < 3; ++kd) {
            const float4 wa = *reinterpret_cast<const float4*>(sW + (kd * 9 + kh * 3 + kw) * 32 + sub * 8);
            const float4 wb = *reinterpret_cast<const float4*>(sW + (kd * 9 + kh * 3 + kw) * 32 + sub * 8 + 4);
            const float2 w01 = make_float2(wa.x, wa.y), w23 = make_float2(wa.z, wa.w);
            const float2 w45 = make_float2(wb.x, wb.y), w67 = make_float2(wb.z, wb.w);
#pragma unroll
            for (int vd = 0; vd < 4; ++vd) {
              const float2 t = make_float2(v[vd + kd], v[vd + kd]);
              acc[vd][0] = __ffma2_rn(t, w01, acc[vd][0]), acc[vd][1] = __ffma2_rn(t, w23, acc[vd][1]);
              acc[vd][2] = __ffma2_rn(t, w45, acc[vd][2]), acc[vd][3] = __ffma2_rn(t, w67, acc[vd][3]);
            }
          }
        }
      }
    }
    const long long row0 = (((long long)b * H + y) * Wp + xp) * Dp + dp0;
#pragma unroll
    for (int vd = 0; vd < 4; ++vd) {
      const int dp = dp0 + vd;
      if (dp >= Dp) break;
      const bool border = xborder || dp == 0 || dp > D;
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const float a0 = border ? 0.f : fmaxf(acc[vd][p].x + bv[2 * p], 0.f) * kDwsepActScale;
        const float a1 = border ? 0.f : fmaxf(acc[vd][p].y + bv[2 * p + 1], 0.f) * kDwsepActScale;
        const __half2 h = f2h2_sat(a0, a1);
        const float2 f = __half22float2(h);
        const float2 dl = split_lo2(a0, a1, f);
        const __half2 l = f2h2_sat(dl.x, dl.y);
        hi[p] = *reinterpret_cast<const uint32_t*>(&h), lo[p] = *reinterpret_cast<const uint32_t*>(&l);
      }
      out[(row0 + vd) * 8 + sub] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      out[(row0 + vd) * 8 + 4 + sub] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
  }
}

// 0 when the fused volume + first-conv kernel applies: stride 1, even W and even feature-channel count (64-bit loads), window fits
int cost_first_conv_fused_supported(int Cf, int D, int H, int W) {
  const int Dp = D + 1 + kPadT;
  const size_t smem = (size_t)(27 * 32 + (FUS_TY + 2) * (4 * ((Dp + 3) / 4) + 2) * (F32V2_XT + 2)) * sizeof(float);
  if ((W & 1) || Cf <= 0 || D <= 0 || smem > 100 * 1024 || H > 65535 * FUS_TY) return LWS_ERR_UNSUPPORTED;
  return LWS_OK;
}

size_t conv3d_f16_workspace_bytes(int B, int D, int H, int W) {
  const size_t rows = (size_t)B * H * (W + 1 + kPadT) * (D + 1 + kPadT);
  return 2 * (rows * 128 + 1024);
}

// wtab[l]: per mid layer the Toeplitz operand table (5 tiles x 192 rows x 128 B, then scales[2]); bias_mid[l]: [32];
// w_last_tab: operand table of the closing 32 -> 1 conv (9 blocks x 16 rows x 128 B, then scales[2])
// featL / featR non-null: `cost` is an OUTPUT -- the stage-1 volume is built inside the first conv kernel (cost_first_conv_fused_kernel)
int conv3d_stack_f16(const float* cost, const float* affine, const float* w_first, const float* b_first, const float* const* wtab,
                     const float* const* bias_mid, int layers, const float* w_last_tab, float* out, void* ws, int B, int D, int H,
                     int W, int add_skip, cudaStream_t st, const float* featL, const float* featR, int Cf) {
  const int Wp = W + 1 + kPadT, Dp = D + 1 + kPadT;
  const long long R = (long long)H * Wp * Dp;
  if (R >= (1ll << 31) - 65536 || 128 + 2 * Dp > 256 || (long long)D * H * W >= (1ll << 31)) return LWS_ERR_UNSUPPORTED;
  const size_t half_ws = conv3d_f16_workspace_bytes(B, D, H, W) / 2;
  float* bufA = (float*)ws;
  float* bufB = (float*)((char*)ws + half_ws);
  const long long nrows = (long long)B * R;
  cudaError_t e;
  {
    if (H > 65535 || B > 65535) return LWS_ERR_BAD_SHAPE;
    const size_t smem2 = (size_t)(27 * 32 + 3 * (4 * ((Dp + 3) / 4) + 2) * (F32V2_XT + 2)) * sizeof(float);
    if (featL) {
      if (cost_first_conv_fused_supported(Cf, D, H, W) != LWS_OK) return LWS_ERR_UNSUPPORTED;
      const size_t smemf = (size_t)(27 * 32 + (FUS_TY + 2) * (4 * ((Dp + 3) / 4) + 2) * (F32V2_XT + 2)) * sizeof(float);
      dim3 grid((Wp + F32V2_XT - 1) / F32V2_XT, (H + FUS_TY - 1) / FUS_TY, B);
      if (opt(OPT_FUSE_VOLUME) != 2 && D % 12 == 0) {  // 12-disparity tiles: the 6 x 17 x D/12 tiles of a block fit one pass of its 256 threads at D = 24
        LWS_SET_SMEM_ONCE(cost_first_conv_fused_kernel<12>, 100 * 1024);
        cost_first_conv_fused_kernel<12><<<grid, 256, smemf, st>>>(featL, featR, const_cast<float*>(cost), w_first, b_first, affine,
                                                                   (uint4*)bufA, Cf, D, H, W);
      } else {
        LWS_SET_SMEM_ONCE(cost_first_conv_fused_kernel<8>, 100 * 1024);
        cost_first_conv_fused_kernel<8><<<grid, 256, smemf, st>>>(featL, featR, const_cast<float*>(cost), w_first, b_first, affine,
                                                                  (uint4*)bufA, Cf, D, H, W);
      }
    } else if (opt(OPT_FIRST_CONV) != 0 && smem2 <= 48 * 1024) {  // shared-memory staged tap window
      dim3 grid((Wp + F32V2_XT - 1) / F32V2_XT, H, B);
      conv3d_first_ydx_v2_kernel<<<grid, 256, smem2, st>>>(cost, w_first, b_first, affine, (uint4*)bufA, D, H, W);
    } else {
      const int groups = Wp * ((Dp + 3) / 4);
      dim3 grid((groups + 63) / 64, H, B);
      conv3d_first_ydx_kernel<<<grid, 256, 0, st>>>(cost, w_first, b_first, affine, (uint4*)bufA, D, H, W, make_fastdiv(Wp));
    }
    if ((e = cudaPeekAtLastError()) != cudaSuccess) return (int)e;
  }
  float* cur = bufA;
  float* nxt = bufB;
  for (int l = 0; l < layers; ++l) {
    TzLayer L;
    memset(&L, 0, sizeof(L));
    L.src0 = L.src1 = cur, L.wtab = wtab[l], L.bias = bias_mid[l], L.out = nxt, L.B = B, L.R = (int)R;
    L.n0 = Dp, L.p0 = 1, L.i0 = D, L.n1 = Wp, L.p1 = 1, L.i1 = W;
    L.tz = 1, L.nstages = 3, L.nshift = 3, L.box_rows = 128 + 2 * Dp;
    for (int kh = 0; kh < 3; ++kh) L.st_off[kh] = (kh - 1) * Wp * Dp - Dp, L.st_src[kh] = 0;  // box starts one x column early
    for (int kw = 0; kw < 3; ++kw) L.shift_rows[kw] = kw * Dp;
    // strips down y: a tile shares its ky = 0, 1 boxes with the tile above it (2: strips cut into segments dealt round-robin)
    if (opt(OPT_TZ_STRIPS) & 3) L.srow = Wp * Dp, L.G = 1, L.seg_mode = (opt(OPT_TZ_STRIPS) & 3) == 2, L.pair = (opt(OPT_TZ_STRIPS) & 3) == 3;
    L.out_split = 1, L.relu = 1;
    int rc = launch_tz_gemm(L, st);
    if (rc) return rc;
    float* t = cur;
    cur = nxt, nxt = t;
  }
  {
    TzLayer L;
    memset(&L, 0, sizeof(L));
    L.src0 = L.src1 = cur, L.wtab = w_last_tab, L.B = B, L.R = (int)R;
    L.n0 = Dp, L.p0 = 1, L.i0 = D, L.n1 = Wp, L.p1 = 1, L.i1 = W;
    L.tz = 1, L.nstages = 3, L.nshift = 3, L.box_rows = 128 + 2 * Dp;
    for (int kh = 0; kh < 3; ++kh) L.st_off[kh] = (kh - 1) * Wp * Dp - Dp, L.st_src[kh] = 0;
    for (int kw = 0; kw < 3; ++kw) L.shift_rows[kw] = kw * Dp;
    if (opt(OPT_TZ_STRIPS) & 4) L.srow = Wp * Dp, L.G = 1;
    L.last = 1, L.skip = add_skip ? cost : nullptr, L.out_f32 = out;
    int rc = launch_tz_gemm(L, st);
    if (rc) return rc;
  }
  return LWS_OK;
}

}  // namespace lws
