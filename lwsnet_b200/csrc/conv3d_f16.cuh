// Host interface of the split-fp16 Toeplitz-N implicit-GEMM kernel (conv3d_f16.cu).
#pragma once
#include <cuda_runtime.h>

namespace lws {

constexpr int TZ_MAXST = 6;

struct TzLayer {
  const float* src0;    // rows [B][R][128 B]
  const float* src1;    // second source (dense refinement conv) or src0
  const float* wtab;    // device: [nbtiles][192][32 words] operand table, then scales[2]
  const float* bias;
  float* out;           // rows [B][R][128 B]
  int B, R;
  int n0, p0, i0, n1, p1, i1;
  int tz;               // 1 or 8
  int nstages, nshift;
  int st_off[TZ_MAXST], st_src[TZ_MAXST], shift_rows[3];
  int box_rows;         // rows per TMA box (<= 256)
  int out_split, relu;
  int srow;             // rows between consecutive taps of the slowest axis if the stages are ordered [tap][G sources] and
                        // st_off[tap*G + g] = (tap - 1) * srow + const (strip schedule with ring reuse); 0 = linear tiling
  int G;
  int pair;             // strip schedule, 3D 32 -> 32 layers only: 1 = CTA pairs (tcgen05 cta_group::2, M = 256, weights split over the pair)
  int seg_mode;         // strip schedule only: 1 = strips cut into segments dealt round-robin to the CTAs (TzSched), 0 = contiguous ranges
  int last;             // 1: the closing C -> 1 layer: wtab = [nstages*nshift blocks][16 rows][32 words] + scales[2], fp32 NCDHW
                        //    output out_f32 (+ skip) instead of rows
  const float* skip;
  float* out_f32;
  int out_mode;         // last: 0 = rows (y, x, d) -> NCDHW, 1 = rows (y, x) -> NCHW
};

int launch_tz_gemm(const TzLayer& L, cudaStream_t st);

}  // namespace lws
