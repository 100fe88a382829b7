// Shared constants and device helpers of the depthwise-separable refinement kernels (dwsep_tc.cu: one block per launch;
// dwsep_chain.cu: several consecutive blocks per launch with the tensors between them kept in L2-resident row rings).
#pragma once
#include <cuda_fp16.h>
#include <math.h>
#include <string.h>

#include "lws_common.cuh"
#include "tma_utils.cuh"

namespace lws {

constexpr int DS_RP = 16;          // border of the refinement CLP tensors
constexpr int DS_DW_WARPS = 8;     // warps 0..7
constexpr int DS_EPI_WARP0 = 8;    // warps 8..11 (warp % 4 = TMEM lane quarter)
constexpr int DS_PROD_WARP = 12;
constexpr int DS_MMA_WARP = 13;
constexpr int DS_THREADS = 14 * 32;
constexpr int DS_NIN = 5;                   // input ring slots
constexpr int DS_INBYTES = 160 * 128;       // (128 + 2*16) pixels x 128 B
constexpr int DS_NA = 3;                    // A-operand tiles
constexpr int DS_NT = 2;                    // TMEM accumulators (64 columns each)
constexpr int DS_NOUT = 2;                  // output staging tiles
constexpr int DS_TILE = 128 * 128;          // 16 KB
constexpr int DS_OFF_A = 0;
constexpr int DS_OFF_OUT = DS_OFF_A + DS_NA * DS_TILE;
constexpr int DS_OFF_B = DS_OFF_OUT + DS_NOUT * DS_TILE;
constexpr int DS_OFF_IN = DS_OFF_B + 8192;
constexpr int DS_OFF_W = DS_OFF_IN + DS_NIN * DS_INBYTES;   // depthwise weights [9][32] fp32
constexpr int DS_OFF_BAR = DS_OFF_W + 9 * 32 * 4;
constexpr int DS_SMEM = DS_OFF_BAR + 256 + 1024 /*align slack*/;
// the im2col modes (CIN > 0) have no activation ring; instead the image lines arrive through a small TMA ring (DS_NIMG lines of
// CIN x 136 floats: the 128 pixels of the tile + 4 columns each side -- the innermost start coordinate of a tiled TMA load has to be
// 16-byte aligned (tools/tma_probe.cu: an unaligned one is an illegal instruction), so the box starts at x0 - 4, not x0 - 1) whose
// out-of-bounds zero fill is the conv's zero padding
constexpr int DS_NIMG = 16;
constexpr int DS_IMG_ROW = 136;                     // floats per channel line in the ring
constexpr int DS_IMG_X0 = 4;                        // image column of ring column 0, relative to the tile's first pixel: x - 4
constexpr int DS_IMG_SLOT = 13 * 128;               // >= 3 * 136 * 4 bytes, 128-byte aligned
constexpr int DS_OFF_IMGBAR = DS_OFF_IN + 9 * 32 * 4 + 256;  // after the CIN > 0 layout's weights + barriers (dwsep_f16_kernel)
constexpr int DS_OFF_IMG = DS_OFF_IMGBAR + 256;
constexpr int DS_SMEM_IM2COL = DS_OFF_IMG + DS_NIMG * DS_IMG_SLOT + 1024 /*align slack*/;
static_assert(DS_OFF_IMG % 128 == 0, "TMA destination alignment");
constexpr float DS_ACT_SCALE = kDwsepActScale;

struct DsArgs {
  const float* dw;      // [32][9] depthwise weights
  const __half* pwh;    // [64][32]: rows 0..31 = fp16(w*sw) (row = cout, col = cin), rows 32..63 = fp16((w*sw - hi) * 2^11)
  const float* scales;  // [2]: c0 = 1 / (DS_ACT_SCALE * sw), c1 = c0 * 2^-11
  const float* bias;    // [32]
  float* out;           // CLP [B][Hp][Wp][32] (for the y-border zero fill; the interior goes through the TMA store map)
  const float* img;     // CIN > 0 (first conv of a refinement branch): NCHW fp32 input [B][CIN][H][W]
  int W;
  int B, Hp, Wp, H, dil, relu;
  int out_split;        // 1: write rows as [32 hi | 32 lo] halves of act * 2^-6 (operand format of conv3d_f16.cu) instead of fp32
  int nxt, rows_phase;     // x tiles per line; lines per row phase (max over phases)
  long long total_rows;     // B * nxt * dil * rows_phase tile-rows, split evenly (contiguously) over the CTAs
  int img_tma;              // CIN > 0: image lines through the TMA ring (map_in = {W, H, B*CIN} fp32) instead of per-thread loads
  int dbg;                  // option chain_debug, TIMING EXPERIMENTS ONLY: bit 2 (4) = the im2col front end does not read the image
};

__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts64(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ uint32_t h2_bits(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }
__device__ __forceinline__ void ds_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,"
      "%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
}


// ---- host entry points ---------------------------------------------------------------------------------------------------------
// one BN-ReLU-DW(dil)-PW block per launch (dwsep_tc.cu); in / out: CLP [B][H + 2*RP][W + 2*RP][32]
int launch_dwsep_f16(const float* in, float* out, const float* dw, const void* pwh, const float* scales, const float* bias, int B,
                     int H, int W, int dil, int relu, int out_split, cudaStream_t st);
int launch_conv0_f16(const float* img, float* out, const void* wtab, const float* scales, const float* bias, int B, int CIN, int H,
                     int W, cudaStream_t st);
// 2..4 consecutive blocks per launch through L2-resident rings (dwsep_chain.cu)
struct ChainBlockDesc {
  const float* dw;      // [32][9]
  const void* pwh;      // [64][32] split-fp16 pointwise table
  const float* scales;  // [2]
  const float* bias;    // [32]
  int dil, relu, out_split;
};
size_t dwsep_chain_workspace_bytes(const int* dil, int nblk, int B, int H, int W);
int launch_dwsep_chain(const float* in, float* out, const ChainBlockDesc* blocks, int nblk, void* ws, size_t ws_bytes, int B,
                       int H, int W, cudaStream_t st);

}  // namespace lws
