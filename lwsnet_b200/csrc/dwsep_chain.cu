// K6 chain: 2..4 CONSECUTIVE BN-ReLU-DW3x3(dil)-PW1x1 blocks of the colour-guidance refinement (reference
// models/submodules.py:236-261, chained at :282-327) in ONE persistent launch, with the tensors between the blocks kept in
// small row RINGS that stay resident in the 126 MB L2 instead of making a round trip through HBM.
//
// One block per launch (dwsep_tc.cu) is HBM-bound: it reads and writes every 32-channel pixel once (256 B / pixel) and twelve
// such launches are a third of the 4-stage step.  Here block k+1 consumes the rows of block k a few hundred microseconds after
// they were produced:
//   * the launch is a dataflow over ITEMS = (block k, pair b, band of MB image rows, 128-pixel column tile, 16 tile-rows of the
//     band's row phases).  Items are numbered in wavefront order -- macro step s holds band s of block 0, band s - lag of
//     block 1, band s - 2*lag of block 2 ... -- and the 148 persistent CTAs pop them from one global counter;
//   * every (block, band) has a completion counter.  Before an item is published to the CTA's warp roles the scheduler thread
//     waits (ld.acquire.gpu) until the bands of block k-1 it reads are complete and until block k+1 has finished with the
//     ring rows it is about to overwrite.  Every dependency points to an item that was popped EARLIER, and a CTA finishes
//     the items it has published without the scheduler thread, so the wait always ends (no co-scheduling assumption beyond
//     "a popped item runs to completion"); a watchdog turns a would-be hang into an error flag in the control words;
//   * the ring between two blocks holds RB = lag + 3 bands: image row y of pair b lives in ring row (b*nmb*MB + y) mod (RB*MB);
//     rows above / below the image come from one all-zero line behind the ring.  `lag` is chosen so that a consumer band is
//     popped >= ~150 items after the last item of the producer band it needs (nobody spins in steady state).
// Inside an item the machinery is that of dwsep_tc.cu: TMA line ring -> depthwise 3x3 in registers along a row phase ->
// split-fp16 A operand -> 4 tcgen05.mma -> epilogue warps -> per-warp TMA stores; the per-block operands (pointwise tile,
// depthwise taps) of all blocks of the chain stay in shared memory.
#include "dwsep_common.cuh"

namespace lws {

constexpr int CH_MAXBLK = 4;
constexpr int CH_NS = 4;  // item slots between the scheduler thread and the other warp roles
constexpr int CH_OFF_A = 0;
constexpr int CH_OFF_OUT = CH_OFF_A + DS_NA * DS_TILE;
constexpr int CH_OFF_B = CH_OFF_OUT + DS_NOUT * DS_TILE;        // [CH_MAXBLK][8192] pointwise operand tiles
constexpr int CH_OFF_IN = CH_OFF_B + CH_MAXBLK * 8192;
constexpr int CH_OFF_W = CH_OFF_IN + DS_NIN * DS_INBYTES;       // [CH_MAXBLK][9][32] depthwise taps
constexpr int CH_OFF_BAR = CH_OFF_W + CH_MAXBLK * 9 * 32 * 4;
constexpr int CH_SMEM = CH_OFF_BAR + 512 + 1024 /*align slack*/;
static_assert(CH_SMEM <= 232448, "chain kernel shared memory");
constexpr int CH_CTRL_WORDS = 16;  // [0] item queue, [1] watchdog flag, then the band counters

struct ChainBlk {
  const float* dw;      // [32][9]
  const __half* pwh;    // [64][32] split-fp16 pointwise table (see DsArgs)
  const float* scales;  // [2]
  const float* bias;    // [32]
  int dil, relu, out_split, pad_;
};
struct ChainArgs {
  ChainBlk blk[CH_MAXBLK];
  float* out_full;       // the last block's output (CLP, for the y-border zero fill)
  unsigned* ctrl;        // control words: zeroed before the launch
  int nblk, B, H, Hp, Wp, nxt;
  int MB, nmb, RB, lag;  // band rows, bands per pair, ring bands, queue lag between consecutive blocks (macro steps)
  int ipb;               // items per (band, column tile) = MB / 16
  int nbands;            // B * nmb
  int per_blk;           // nxt * ipb items per (block, band)
  int items_per_step;    // nblk * per_blk
  int total_items;
  unsigned done_target;  // per_blk * 4 epilogue-warp arrivals complete a band
  int debug;             // option "chain_debug" (timing experiments)
};
struct ChainMaps {
  CUtensorMap in[CH_MAXBLK];   // block k input: k = 0 the CLP input tensor, k > 0 ring k-1 (box = 128 + 2*dil_k pixels)
  CUtensorMap out[CH_MAXBLK];  // block k output: ring k, or the CLP output tensor for the last block (box = 32 pixels)
};

struct ChainItem {
  int k, b, g, x0, dil;
  int y0, y1;      // band rows [y0, y1)
  int p0, np;      // row phases [p0, p0 + np)
  int r0, nr;      // phase-rows [r0, r0 + nr) of each phase inside the band
};
// false: the macro step has no band for this block (pipeline fill / drain)
__device__ __forceinline__ bool chain_decode(const ChainArgs& a, int n, ChainItem& it) {
  const int s = n / a.items_per_step;
  int r = n - s * a.items_per_step;
  it.k = r / a.per_blk;
  r -= it.k * a.per_blk;
  const int q = r / a.nxt, xt = r - q * a.nxt;
  it.g = s - it.k * a.lag;
  if (it.g < 0 || it.g >= a.nbands) return false;
  it.b = it.g / a.nmb;
  const int m = it.g - it.b * a.nmb;
  it.x0 = xt * 128;
  it.dil = a.blk[it.k].dil;
  it.y0 = m * a.MB;
  it.y1 = min(it.y0 + a.MB, a.H);
  const int rp = a.MB / it.dil;  // rows of one phase in a band
  if (rp >= 16) {
    it.np = 1, it.p0 = q % it.dil, it.r0 = (q / it.dil) * 16, it.nr = 16;
  } else {
    it.np = 16 / rp, it.p0 = q * it.np, it.r0 = 0, it.nr = rp;
  }
  return true;
}
// the item's run of phase p: first image row and number of rows (0: nothing of this phase inside the image)
__device__ __forceinline__ int chain_run(const ChainItem& it, int p, int& yi0) {
  yi0 = it.y0 + p + it.dil * it.r0;
  if (yi0 >= it.y1) return 0;
  return min(it.nr, (it.y1 - yi0 + it.dil - 1) / it.dil);
}
// TMA row coordinate of image row y of pair b: in the CLP tensor (full) or in a ring (rows outside the image -> the zero line)
__device__ __forceinline__ int chain_row(const ChainArgs& a, bool ring, int b, int y) {
  if (!ring) return b * a.Hp + DS_RP + y;
  if (y < 0 || y >= a.H) return a.RB * a.MB;
  return (b * a.nmb * a.MB + y) % (a.RB * a.MB);
}

__device__ __forceinline__ void chain_wait(const unsigned* cnt, unsigned target, unsigned* flag) {
  unsigned spins = 0;
  while (true) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(cnt) : "memory");
    if (v >= target) return;
    __nanosleep(64);
    if ((++spins & 1023u) == 0) {  // watchdog: never hang the device; the flag is checked by the host-side tests
      unsigned f;
      asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(f) : "l"(flag) : "memory");
      if (f || spins > (1u << 21)) {
        atomicExch(flag, 1u);
        return;
      }
    }
  }
}

__global__ void __launch_bounds__(DS_THREADS, 1)
    dwsep_chain_kernel(const __grid_constant__ ChainMaps maps, const __grid_constant__ ChainArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + CH_OFF_BAR);
  uint64_t* in_full = bars;                 // [NIN]
  uint64_t* in_empty = in_full + DS_NIN;    // [NIN]
  uint64_t* a_full = in_empty + DS_NIN;     // [NA]
  uint64_t* a_empty = a_full + DS_NA;       // [NA]
  uint64_t* t_full = a_empty + DS_NA;       // [NT]
  uint64_t* t_empty = t_full + DS_NT;       // [NT]
  uint64_t* s_full = t_empty + DS_NT;       // [NS]
  uint64_t* s_empty = s_full + CH_NS;       // [NS]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_empty + CH_NS);
  volatile int* sched = reinterpret_cast<volatile int*>(tmem_slot + 4);  // [NS] published item numbers (-1: queue exhausted)
  float* sW = reinterpret_cast<float*>(smem + CH_OFF_W);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int i = 0; i < DS_NIN; ++i) mbar_init(in_full + i, 1), mbar_init(in_empty + i, DS_DW_WARPS);
    for (int i = 0; i < DS_NA; ++i) mbar_init(a_full + i, DS_DW_WARPS), mbar_init(a_empty + i, 1);
    for (int i = 0; i < DS_NT; ++i) mbar_init(t_full + i, 1), mbar_init(t_empty + i, 4);
    for (int i = 0; i < CH_NS; ++i) mbar_init(s_full + i, 1), mbar_init(s_empty + i, DS_DW_WARPS + 4 + 1);
    mbar_fence_init();
    for (int k = 0; k < a.nblk; ++k) tma_prefetch_desc(&maps.in[k]), tma_prefetch_desc(&maps.out[k]);
  }
  if (warp == DS_MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(DS_NT * 64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // per-block operands: pointwise tiles (64 rows x 64 B used of a 128-byte SWIZZLE_128B row) and depthwise taps [tap][channel]
  for (int idx = tid; idx < a.nblk * 64 * 4; idx += DS_THREADS) {
    const int k = idx >> 8, n = (idx >> 2) & 63, c = idx & 3;
    *reinterpret_cast<uint4*>(smem + CH_OFF_B + k * 8192 + n * 128 + ((c ^ (n & 7)) << 4)) =
        __ldg(reinterpret_cast<const uint4*>(a.blk[k].pwh + n * 32 + c * 8));
  }
  for (int idx = tid; idx < a.nblk * 9 * 32; idx += DS_THREADS) {
    const int k = idx / 288, r = idx - k * 288;
    sW[idx] = __ldg(a.blk[k].dw + (r & 31) * 9 + (r >> 5)) * DS_ACT_SCALE;
  }
  // zero the y-border lines of the chain's output tensor
  {
    const long long line4 = (long long)a.Wp * 8;  // float4 per line
    const long long total = (long long)a.B * 2 * DS_RP * line4;
    float4* o = reinterpret_cast<float4*>(a.out_full);
    for (long long i = (long long)blockIdx.x * DS_THREADS + tid; i < total; i += (long long)gridDim.x * DS_THREADS) {
      const long long ln = i / line4, r = i - ln * line4;
      const int b = (int)(ln / (2 * DS_RP)), kk = (int)(ln % (2 * DS_RP));
      const int y = kk < DS_RP ? kk : a.Hp - 2 * DS_RP + kk;
      o[((long long)b * a.Hp + y) * line4 + r] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  fence_proxy_async_smem();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  // every consumer role takes the next published item: false when the queue is exhausted
  uint32_t si = 0;
  auto item_published = [&]() -> bool {  // non-blocking: has the scheduler thread published the next item already?
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(smem_u32(s_full + si % CH_NS)), "r"((si / CH_NS) & 1)
        : "memory");
    return ok != 0;
  };
  auto next_item = [&](ChainItem& it) -> bool {
    const uint32_t slot = si % CH_NS;
    mbar_wait(s_full + slot, (si / CH_NS) & 1);
    const int n = sched[slot];
    __syncwarp();
    if (lane == 0) mbar_arrive(s_empty + slot);
    ++si;
    if (n < 0) return false;
    chain_decode(a, n, it);
    return true;
  };

  if (warp == DS_PROD_WARP) {
    // ================================ scheduler + TMA producer (one thread) ================================
    if (elect_one_sync()) {
      unsigned* queue = a.ctrl;
      unsigned* flag = a.ctrl + 1;
      const unsigned* done = a.ctrl + CH_CTRL_WORDS;
      uint32_t it = 0, pub = 0;
      while (true) {
        const int n = (int)atomicAdd(queue, 1u);
        ChainItem w;
        const bool end = n >= a.total_items;
        if (!end) {
          if (!chain_decode(a, n, w)) continue;
          const int m = w.g - w.b * a.nmb;
          if (!(a.debug & 1) && w.k > 0) {  // the bands of block k-1 this item reads (one band of halo either side, inside the pair)
            const unsigned* d = done + (w.k - 1) * a.nbands + w.b * a.nmb;
            for (int mm = max(m - 1, 0); mm <= min(m + 1, a.nmb - 1); ++mm) chain_wait(d + mm, a.done_target, flag);
          }
          if (!(a.debug & 1) && w.k + 1 < a.nblk) {  // block k+1 is done with the ring rows this band overwrites (global band g - RB and its halo readers)
            const unsigned* d = done + (w.k + 1) * a.nbands;
            for (int gg = max(w.g - a.RB - 1, 0); gg <= w.g - a.RB + 1; ++gg) chain_wait(d + gg, a.done_target, flag);
          }
          asm volatile("fence.proxy.async;" ::: "memory");  // acquired generic-proxy view -> the TMA (async proxy) reads below
        }
        const uint32_t slot = pub % CH_NS;
        mbar_wait(s_empty + slot, ((pub / CH_NS) & 1) ^ 1);
        sched[slot] = end ? -1 : n;
        mbar_arrive(s_full + slot);
        ++pub;
        if (end) break;
        const int dil = w.dil;
        const uint32_t bytes = (uint32_t)(128 + 2 * dil) * 128;
        const CUtensorMap* map = &maps.in[w.k];
        const bool ring = w.k > 0;
        for (int p = w.p0; p < w.p0 + w.np; ++p) {
          int yi0;
          const int nrows = chain_run(w, p, yi0);
          if (nrows == 0) continue;
          for (int kk = 0; kk < nrows + 2; ++kk, ++it) {
            const uint32_t slot_in = it % DS_NIN;
            mbar_wait(in_empty + slot_in, ((it / DS_NIN) & 1) ^ 1);
            mbar_expect_tx(in_full + slot_in, bytes);
            asm volatile(
                "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                    smem_u32(smem + CH_OFF_IN + slot_in * DS_INBYTES)),
                "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(in_full + slot_in)), "r"(0), "r"(w.x0 - dil),
                "r"(chain_row(a, ring, w.b, yi0 + (kk - 1) * dil))
                : "memory");
          }
        }
      }
    }
  } else if (warp == DS_MMA_WARP) {
    // ================================ MMA issuer ================================
    const uint32_t idesc64 = (1u << 4) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t idesc32 = (1u << 4) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t desc_hi = ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);  // SBO, version, SW128
    uint32_t t = 0;
    ChainItem w;
    while (next_item(w)) {
      const uint32_t b_lo = ((smem_u32(smem + CH_OFF_B + w.k * 8192) & 0x3FFFF) >> 4) | (1u << 16);
      for (int p = w.p0; p < w.p0 + w.np; ++p) {
        int yi0;
        const int nrows = chain_run(w, p, yi0);
        for (int i = 0; i < nrows; ++i, ++t) {
          const uint32_t ab = t % DS_NA, tb = t % DS_NT;
          mbar_wait(t_empty + tb, ((t / DS_NT) & 1) ^ 1);
          mbar_wait(a_full + ab, (t / DS_NA) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (elect_one_sync()) {
            const uint32_t a_lo = ((smem_u32(smem + CH_OFF_A + ab * DS_TILE) & 0x3FFFF) >> 4) | (1u << 16);
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
          __syncwarp();
        }
      }
    }
  } else if (warp >= DS_EPI_WARP0) {
    // ================================ epilogue (warps 8..11) ================================
    const int quarter = warp & 3;
    const int p = quarter * 32 + lane;  // pixel of the tile = TMEM lane
    unsigned* done = a.ctrl + CH_CTRL_WORDS;
    uint32_t t = 0;              // tiles stored so far by this warp (= bulk groups committed)
    unsigned* pend = nullptr;    // band counter of the previous item, signalled once its stores have completed
    uint32_t pend_t = 0;         // value of t after the previous item's last tile
    auto signal = [&](unsigned* cnt) {
      // this warp's TMA stores of the item have completed (wait_group without .read): publish them device-wide
      asm volatile("fence.proxy.async;" ::: "memory");
      __threadfence();
      atomicAdd(cnt, 1u);
    };
    ChainItem w;
    while (true) {
      // the deferred signal must not wait for an item that is not there yet: this CTA's own scheduler thread may be spinning on
      // exactly that band counter before it publishes the next item
      if (lane == 0 && pend && !item_published()) {
        if (!(a.debug & 2)) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        signal(pend);
        pend = nullptr;
      }
      __syncwarp();
      if (!next_item(w)) break;
      const ChainBlk& blk = a.blk[w.k];
      float bias[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) bias[c] = __ldg(blk.bias + c);
      const float c0 = __ldg(blk.scales), c1 = __ldg(blk.scales + 1);
      const float lo = blk.relu ? 0.f : -INFINITY;
      const int out_split = blk.out_split;
      const bool ring_out = w.k + 1 < a.nblk;
      const CUtensorMap* map = &maps.out[w.k];
      const int xpix = w.x0 + p;
      const bool border = xpix < DS_RP || xpix >= a.Wp - DS_RP;
      for (int ph = w.p0; ph < w.p0 + w.np; ++ph) {
        int yi0;
        const int nrows = chain_run(w, ph, yi0);
        for (int i = 0; i < nrows; ++i, ++t) {
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
          if (lane == 0) {
            asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(DS_NOUT - 1) : "memory");
            if (pend && t >= pend_t + 3) {  // all but the 3 newest groups have completed: the previous item is in L2
              asm volatile("cp.async.bulk.wait_group 3;" ::: "memory");
              signal(pend);
              pend = nullptr;
            }
          }
          __syncwarp();
          const uint32_t so = smem_u32(smem + CH_OFF_OUT + ob * DS_TILE) + p * 128;
          if (out_split) {
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
                             reinterpret_cast<uint64_t>(map)),
                         "r"(smem_u32(smem + CH_OFF_OUT + ob * DS_TILE + quarter * 4096)), "r"(0), "r"(w.x0 + quarter * 32),
                         "r"(chain_row(a, ring_out, w.b, yi0 + i * w.dil))
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
      }
      // end of the item: its band counter is signalled once its stores are known to have completed
      if (lane == 0) {
        if (pend) {  // the item before was too short for the deferred path
          if (!(a.debug & 2)) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
          signal(pend);
          pend = nullptr;
          signal(done + w.k * a.nbands + w.g);
        } else if (t == pend_t) {  // nothing of this item lies inside the image: no stores to wait for
          signal(done + w.k * a.nbands + w.g);
        } else {
          pend = done + w.k * a.nbands + w.g;
        }
        pend_t = t;
      }
    }
    if (lane == 0) {
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      if (pend) signal(pend);
    }
  } else {
    // ================================ depthwise warps (0..7) ================================
    const int q = tid & 7;    // channel quad
    const int pg = tid >> 3;  // 0..31: pixel group = (block of 4*dil pixels, offset g inside the stride)
    const uint32_t in_base = smem_u32(smem + CH_OFF_IN), a_base = smem_u32(smem + CH_OFF_A);
    uint32_t it = 0, t = 0;
    float4 win[3][6];
    ChainItem w;
    while (next_item(w)) {
      const int dil = w.dil;
      const int blk = pg / dil, g = pg - blk * dil;
      const int pbase = blk * 4 * dil + g;                      // first of this thread's 4 pixels (stride dil)
      const uint32_t in_off = (uint32_t)pbase * 128 + q * 16;   // ring column of tap kx = 0 of pixel 0 is pbase (box starts at x0 - dil)
      const uint32_t jstep = (uint32_t)dil * 128;
      const uint32_t w_base = smem_u32(sW + w.k * 288) + q * 16;
      uint32_t a_off[4];  // byte offsets of the hi halves of this thread's 4 pixels inside an A tile (lo = chunk + 4)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int px = pbase + j * dil;
        a_off[j] = (uint32_t)px * 128 + (q & 1) * 8 + ((uint32_t)((q >> 1) ^ (px & 7)) << 4);
      }
      for (int ph = w.p0; ph < w.p0 + w.np; ++ph) {
        int yi0;
        const int nrows = chain_run(w, ph, yi0);
        if (nrows == 0) continue;
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
        for (int i = 0; i < nrows; i += 3) {
          load_line(win[2]);
          emit(win[0], win[1], win[2]);
          if (i + 1 >= nrows) break;
          load_line(win[0]);
          emit(win[1], win[2], win[0]);
          if (i + 2 >= nrows) break;
          load_line(win[1]);
          emit(win[2], win[0], win[1]);
        }
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == DS_MMA_WARP) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(DS_NT * 64));
}

// ---- host ------------------------------------------------------------------------------------------------------------------
// Geometry of a chain: band height (>= 4 rows per phase for the largest dilation, multiple of 16), queue lag and ring size.
struct ChainGeom {
  int MB, nmb, lag, RB, nxt, ipb;
  size_t ring_bytes;   // one ring incl. its zero line
  size_t ctrl_bytes;
};
static ChainGeom chain_geom(const int* dil, int nblk, int B, int H, int W, int sep_items) {
  ChainGeom g;
  int dmax = 1;
  for (int k = 0; k < nblk; ++k) dmax = dil[k] > dmax ? dil[k] : dmax;
  g.MB = 4 * dmax < 16 ? 16 : 4 * dmax;
  g.nmb = (H + g.MB - 1) / g.MB;
  g.nxt = (W + 2 * DS_RP + 127) / 128;
  g.ipb = g.MB / 16;
  const int per_step = nblk * g.nxt * g.ipb;
  g.lag = (sep_items + per_step - 1) / per_step + 1;
  g.RB = g.lag + 3;
  g.ring_bytes = ((size_t)g.RB * g.MB + 1) * (size_t)(W + 2 * DS_RP) * 128;
  g.ctrl_bytes = ((size_t)CH_CTRL_WORDS + (size_t)nblk * B * g.nmb) * 4;
  g.ctrl_bytes = (g.ctrl_bytes + 255) / 256 * 256;
  return g;
}

// workspace for the worst case over the option "chain_sep_items" values a caller may switch to between sizing and launching is
// not attempted: size and launch must see the same option value (the launch re-checks against `ws_bytes`)
size_t dwsep_chain_workspace_bytes(const int* dil, int nblk, int B, int H, int W) {
  const ChainGeom g = chain_geom(dil, nblk, B, H, W, opt(OPT_CHAIN_SEP_ITEMS));
  return g.ctrl_bytes + (size_t)(nblk - 1) * ((g.ring_bytes + 255) / 256 * 256);
}

// in / out: CLP [B][H + 2*RP][W + 2*RP][32]; ws: dwsep_chain_workspace_bytes (control words, then the rings)
int launch_dwsep_chain(const float* in, float* out, const ChainBlockDesc* blocks, int nblk, void* ws, size_t ws_bytes, int B,
                       int H, int W, cudaStream_t st) {
  if (nblk < 2 || nblk > CH_MAXBLK) return LWS_ERR_UNSUPPORTED;
  int dil[CH_MAXBLK];
  for (int k = 0; k < nblk; ++k) {
    dil[k] = blocks[k].dil;
    if (dil[k] < 1 || dil[k] > DS_RP || (32 % dil[k]) != 0) return LWS_ERR_UNSUPPORTED;
    if (k + 1 < nblk && blocks[k].out_split) return LWS_ERR_UNSUPPORTED;  // the rings carry fp32 rows
  }
  LWS_SET_SMEM_ONCE(dwsep_chain_kernel, CH_SMEM);
  const ChainGeom g = chain_geom(dil, nblk, B, H, W, opt(OPT_CHAIN_SEP_ITEMS));
  if (ws_bytes < g.ctrl_bytes + (size_t)(nblk - 1) * ((g.ring_bytes + 255) / 256 * 256)) return LWS_ERR_WORKSPACE_TOO_SMALL;
  if ((((uintptr_t)ws) & 255) != 0) return LWS_ERR_BAD_ALIGN;
  ChainArgs a;
  memset(&a, 0, sizeof(a));
  for (int k = 0; k < nblk; ++k) {
    a.blk[k].dw = blocks[k].dw, a.blk[k].pwh = (const __half*)blocks[k].pwh, a.blk[k].scales = blocks[k].scales;
    a.blk[k].bias = blocks[k].bias, a.blk[k].dil = dil[k], a.blk[k].relu = blocks[k].relu, a.blk[k].out_split = blocks[k].out_split;
  }
  a.out_full = out;
  a.ctrl = (unsigned*)ws;
  a.nblk = nblk, a.B = B, a.H = H, a.Hp = H + 2 * DS_RP, a.Wp = W + 2 * DS_RP, a.nxt = g.nxt;
  a.MB = g.MB, a.nmb = g.nmb, a.RB = g.RB, a.lag = g.lag, a.ipb = g.ipb;
  a.nbands = B * g.nmb;
  a.per_blk = g.nxt * g.ipb;
  a.items_per_step = nblk * a.per_blk;
  const long long steps = (long long)a.nbands + (long long)(nblk - 1) * g.lag;
  if (steps * a.items_per_step >= (1ll << 30)) return LWS_ERR_BAD_SHAPE;
  a.total_items = (int)(steps * a.items_per_step);
  a.done_target = (unsigned)a.per_blk * 4u;
  a.debug = opt(OPT_CHAIN_DEBUG);

  cudaError_t e = cudaMemsetAsync(ws, 0, g.ctrl_bytes, st);
  if (e != cudaSuccess) return (int)e;
  const size_t ring_stride = (g.ring_bytes + 255) / 256 * 256;
  char* rings = (char*)ws + g.ctrl_bytes;
  const size_t line_bytes = (size_t)a.Wp * 128;
  for (int k = 0; k + 1 < nblk; ++k) {  // the zero line behind every ring
    e = cudaMemsetAsync(rings + k * ring_stride + (size_t)g.RB * g.MB * line_bytes, 0, line_bytes, st);
    if (e != cudaSuccess) return (int)e;
  }
  ChainMaps maps;
  const uint64_t dims_full[3] = {32, (uint64_t)a.Wp, (uint64_t)B * a.Hp}, strides[2] = {128, (uint64_t)a.Wp * 128};
  const uint64_t dims_ring[3] = {32, (uint64_t)a.Wp, (uint64_t)g.RB * g.MB + 1};
  const uint32_t box_out[3] = {32, 32, 1};
  for (int k = 0; k < nblk; ++k) {
    const uint32_t box_in[3] = {32, (uint32_t)(128 + 2 * dil[k]), 1};
    const float* src = k == 0 ? in : (const float*)(rings + (k - 1) * ring_stride);
    int rc = make_tensor_map_f32(&maps.in[k], src, 3, k == 0 ? dims_full : dims_ring, strides, box_in, false);
    if (rc) return rc;
    const float* dst = k + 1 == nblk ? out : (const float*)(rings + k * ring_stride);
    rc = make_tensor_map_f32(&maps.out[k], dst, 3, k + 1 == nblk ? dims_full : dims_ring, strides, box_out, true);
    if (rc) return rc;
  }
  const int grid = a.total_items < kNumSMs ? a.total_items : kNumSMs;
  dwsep_chain_kernel<<<grid, DS_THREADS, CH_SMEM, st>>>(maps, a);
  e = cudaPeekAtLastError();
  return e == cudaSuccess ? LWS_OK : (int)e;
}

}  // namespace lws
