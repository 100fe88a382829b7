/*
 * lws.h — C ABI of the B200-native LWSNet stereo hot path (liblws_b200.so).
 *
 * The reference (PrinceVictor/LWSNet) has no FFI / operator-plugin layer: its hot path is a chain of
 * PaddlePaddle operator calls made from models/models.py and models/submodules.py.  Each entry point
 * below replaces the operator chain of one reference function (file:line given per function); the
 * Python host in lwsnet_b200/ keeps the reference's class / method names and calls these through ctypes,
 * and INTEGRATION.md shows the Paddle custom-op binding a maintainer would add on top of the same symbols.
 *
 * Contract (all functions unless noted):
 *   - every tensor pointer is a DEVICE pointer to contiguous fp32, NCHW (NCDHW for volumes);
 *   - the caller owns every buffer including workspaces; nothing is allocated or freed here;
 *   - work is enqueued asynchronously on `stream` (a cudaStream_t passed as void*); no synchronisation;
 *   - re-entrant; the only process-wide state is the option table below, changed only by lws_set_option (the library
 *     never reads the environment), so concurrent calls see whatever options were set before them;
 *   - return 0 on success, <0 = LWS_ERR_*, >0 = the cudaError_t of a failed launch.
 *   - lws_pack_* functions are pure HOST functions (host pointers in, host blob out).
 */
#ifndef LWS_H_
#define LWS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* lws_stream_t; /* cudaStream_t */

#if defined(__GNUC__)
#define LWS_API __attribute__((visibility("default")))
#else
#define LWS_API
#endif

enum {
  LWS_OK = 0,
  LWS_ERR_BAD_SHAPE = -1,
  LWS_ERR_BAD_ALIGN = -2,
  LWS_ERR_NULL_PTR = -3,
  LWS_ERR_WORKSPACE_TOO_SMALL = -4,
  LWS_ERR_UNSUPPORTED = -5
};

LWS_API const char* lws_status_string(int status);
/* "lws_b200 <semver> sm_100a" */
LWS_API const char* lws_version(void);

/* ---- options ---------------------------------------------------------------------------------------------
 * Explicit, process-wide switches (int values).  Unknown key: LWS_ERR_UNSUPPORTED; value out of range: LWS_ERR_BAD_SHAPE.
 *   "conv3d_tc"       1   3D stacks on tcgen05 split-fp16 (1) or on the exact-fp32 FFMA kernels (0)
 *   "refine_tc"       1   refinement on tcgen05 split-fp16 (1) or on the exact-fp32 FFMA kernels (0)
 *   "refine_chain"    0   consecutive BN-ReLU-DW-PW blocks per L2-resident chain launch (0 / 1: one block per launch; 2; 4).
 *                         Off by default: it cuts the blocks' DRAM traffic 1.7x but is slower (profiles/r02_chain_ab.txt)
 *   "chain_min_bands" 24  chains are used from this many (pair, 64-row band) units per launch on
 *   "chain_sep_items" 160 chain kernel: queue distance between a producer band and its consumers (sizes the L2 rings;
 *                         workspace sizes depend on it: set it before lws_refinement_workspace_bytes)
 *   "warp_div_mode"   0   lws_warp_* coordinate normalisation x / (size-1): 0 = x * fl32(1/(size-1)) (Paddle 2.0 scalar
 *                         division = scale op), 1 = IEEE division
 *   "first_conv"      1   first 1 -> C conv of a 3D stack: 0 = taps read from global memory, 1 = C = 32 with the tap window staged
 *                         in shared memory (default), 4 / 8 = C = 8 staged as well (measured slower; A/B)
 *   "fused_tail"      0   hint for hosts: run the stage tail through lws_regression_tail_f32 (one kernel) instead of the three
 *                         stand-alone entries.  Bit-identical; measured slower at 24 pairs per launch (profiles/r02_tail_ab.txt)
 *   "fuse_volume"     1   hint for hosts: stage 1 through lws_cost_volume_conv3d_stack_f32 (volume built inside the first conv kernel);
 *                         0 = two calls, 2 = fused with 8-disparity tiles.  Bit-identical (profiles/r02_fuse_volume_ab.txt)
 *   "tz_strips"       5   C = 32 stack schedule.  Bits 0-1, the 32 -> 32 layers: 0 = every tile loads its three ky boxes, 1 = tiles walk
 *                         down y in strips (default), 2 = strips in round-robin segments, 3 = CTA pairs (tcgen05 cta_group::2);
 *                         bit 2: strips for the closing conv too.  All bit-identical (profiles/r02_tz_strips_ab.txt)
 *   "k5_int"          1   lws_scale_upsample_add_f32: closed-form periodic taps for the integer scales 2 / 4 / 8 (0 = generic kernel)
 *   "fe_tma"          1   feature pyramid: input tiles through TMA where the maps have 16-byte aligned rows (0 = per-thread loads)
 *   "c8_group"        0   C = 8 stacks depth-first over groups of this many pairs (0 = whole batch layer by layer; measured slower)
 *   "c8_v1" 0, "c8_chunk" 0, "k1_dt" 8   developer A/B switches of the C = 8 stack and the stage-1 volume kernel
 *   "chain_debug" 0, "tz_debug" 0   TIMING EXPERIMENTS ONLY (wrong results): kernels with waits / loads / epilogues / stores switched off */
LWS_API int lws_set_option(const char* key, int value);
LWS_API int lws_get_option(const char* key, int* value);

/* ---- a1: LWSNet._build_volume_2d  (models/models.py:58-76) ------------------------------------------
 * cost[b,d/stride,y,x] = sum_c | L[b,c,y,x] - (x-d >= 0 ? R[b,c,y,x-d] : 0) |,  d = 0,stride,..,maxdisp-stride.
 * L,R [B,C,H,W] -> cost [B,maxdisp/stride,H,W].  LWS_ERR_BAD_SHAPE if maxdisp % stride != 0 (the reference asserts). */
LWS_API int lws_cost_volume_l1_f32(const float* L, const float* R, float* cost, int B, int C, int H, int W, int maxdisp,
                           int stride, lws_stream_t stream);

/* ---- a2: wflow, LWSNet.forward (models/models.py:119-121) -------------------------------------------
 * wflow = (bilinear_resize_halfpixel(pred_full, h, w) * float(h)) * fl32(1/H).  pred_full [B,1,H,W] -> wflow [B,1,h,w]. */
LWS_API int lws_disp_to_scale_f32(const float* pred_full, float* wflow, int B, int H, int W, int h, int w, lws_stream_t stream);

/* ---- a3: LWSNet.warp  (models/models.py:28-55) -------------------------------------------------------
 * out[n,c,y,x] = bilinear sample of x[n,c] at (x - disp[n,0,y,x], y), zero padding, grid_sample(align_corners=True)
 * semantics with the reference's fp32 normalise/un-normalise round trip replayed op by op (no FMA contraction). */
LWS_API int lws_warp_bilinear_f32(const float* x, const float* disp, float* out, int N, int C, int H, int W, lws_stream_t stream);

/* Test hook for "warp sampling indices must be bit-exact": runs the same device function the warp kernels use and
 * exports, for warp argument (disp - shift): x0 [N,H,W] int32 (clamped to [-2, W+1]), y0 [H] int32,
 * wx [N,H,W,2] = (x1-ix, ix-x0), wy [H,2] = (y1-iy, iy-y0). */
LWS_API int lws_warp_taps_f32(const float* disp, float shift, int32_t* x0, int32_t* y0, float* wx, float* wy, int N, int H,
                      int W, lws_stream_t stream);

/* ---- a4: LWSNet._build_volume_2d3  (models/models.py:78-104) ---------------------------------------
 * cost[b,k,y,x] = sum_c | L[b,c,y,x] - warp(R[b], disp[b] - (k-(m-1))*stride)[c,y,x] |,  k = 0..2m-2.
 * L,R [B,C,H,W], disp [B,1,H,W] -> cost [B,2m-1,H,W].  The 9x replicated batch of the reference is never materialised. */
LWS_API int lws_warp_residual_volume_l1_f32(const float* L, const float* R, const float* disp, float* cost, int B, int C,
                                    int H, int W, int m, int stride, lws_stream_t stream);

/* ---- a5: post_3dconvs + skip  (models/submodules.py:190-221, models/models.py:136-138) ---------------
 * out = cost + Conv_{n-1}(ReLU(BN_{n-1}( ... Conv_0(ReLU(BN_0(cost))) ... ))),  n = layers + 2 convs,
 * channels 1 -> C -> ... -> C -> 1, every conv 3x3x3 / stride 1 / zero pad 1 / no bias, BN in inference mode.
 * cost,out [B,D,H,W] (the singleton channel is implicit).  Any C >= 1 (the reference takes any channels_3d * growth_rate,
 * models/models.py:19-22): C = 8 and 32 (the reference's configuration) have tcgen05 kernels, every other width runs on the
 * exact-fp32 FFMA kernel (16 natively, the rest in groups of 8 output channels, zero-padded to a multiple of 8).
 * add_skip != 0 fuses the `+ cost` of models/models.py:137 into the last conv; add_skip == 0 returns the bare
 * post_3dconvs(cost) (what calling the reference's nn.Sequential alone returns). */
LWS_API size_t lws_conv3d_stack_packed_floats(int C, int layers);
/* HOST: fold BN into the conv weights and lay them out for the kernels.
 * conv_w[i]  : [Cout_i, Cin_i, 3,3,3] fp32 (Paddle Conv3D layout), i = 0..layers+1
 * bn_*[i]    : [Cin_i] weight / bias / _mean / _variance of the BatchNorm3D in front of conv i
 * packed     : host buffer of lws_conv3d_stack_packed_floats(C, layers) floats */
LWS_API int lws_pack_conv3d_stack_weights(const float* const* conv_w, const float* const* bn_weight,
                                  const float* const* bn_bias, const float* const* bn_mean,
                                  const float* const* bn_var, float eps, int C, int layers, float* packed);
LWS_API size_t lws_conv3d_stack_workspace_bytes(int B, int D, int H, int W, int C, int layers);
/* kernel launches lws_conv3d_stack_f32 enqueues (layers + 2 for C = 8, 16, 32) */
LWS_API int lws_conv3d_stack_launches(int C, int layers);
LWS_API int lws_conv3d_stack_f32(const float* cost, const float* packed_weights, float* out, void* ws, size_t ws_bytes, int B,
                         int D, int H, int W, int C, int layers, int add_skip, lws_stream_t stream);

/* a2 + a5 in one call (north_star item 2: "volume build fused into the first conv"): stage 1 of LWSNet.forward,
 * models/models.py:126-138 with _build_volume_2d (models/models.py:58-76) computed inside the first conv kernel's shared-memory
 * tap window.  L,R [B,Cf,H,W] features, maxdisp = the stage's disparity count D (stride 1); cost_out [B,D,H,W] receives the raw
 * volume (the skip input), out [B,D,H,W] the stack's result; bit-identical to lws_cost_volume_l1_f32 + lws_conv3d_stack_f32.
 * _supported returns LWS_OK when the fused kernel applies (C = 32 tensor-core path, even W, window fits in shared memory),
 * LWS_ERR_UNSUPPORTED otherwise -- the caller then issues the two calls. */
LWS_API int lws_cost_volume_conv3d_stack_supported(int B, int Cf, int H, int W, int maxdisp, int C, int layers);
LWS_API int lws_cost_volume_conv3d_stack_f32(const float* L, const float* R, const float* packed_weights, float* cost_out,
                                     float* out, void* ws, size_t ws_bytes, int B, int Cf, int H, int W, int maxdisp, int C,
                                     int layers, int add_skip, lws_stream_t stream);

/* One BN-folded C -> C layer of the stack (out = ReLU(conv3x3x3(in, w_folded) + bias)), in/out [B,C,D,H,W] post-activation,
 * w_folded [Cin][27][Cout].  The stack's dominant kernel on its own: for per-layer tests and for timing it in isolation. */
LWS_API int lws_conv3d_bnrelu_layer_f32(const float* in, const float* w_folded, const float* bias, float* out, int B, int C,
                                        int D, int H, int W, lws_stream_t stream);

/* ---- a6: F.softmax(-cost, axis=1) + disparity_regression  (models/models.py:142,151-152,167-179) ----
 * low[b,0,y,x] = sum_j softmax_j(-cost[b,:,y,x]) * (start + j*step).  One pass over the volume. */
LWS_API int lws_softmax_regression_f32(const float* cost, float* low, int B, int D, int H, int W, float start, float step,
                               lws_stream_t stream);

/* The stand-alone class disparity_regression(start, end, stride).forward(input) (models/models.py:167-179) on an ALREADY
 * soft-maxed (or any other) volume: out[b,0,y,x] = sum_j input[b,j,y,x] * (start + j*step), no renormalisation, fp32
 * left-to-right sum like paddle.sum over axis 1. */
LWS_API int lws_disparity_regression_f32(const float* prob, float* out, int B, int D, int H, int W, float start, float step,
                                         lws_stream_t stream);

/* ---- a7: rescale + upsample + skip  (models/models.py:145-148,153-156) ------------------------------
 * pred = bilinear_resize_halfpixel((low * float(H)) * fl32(1/h), H, W) (+ prev if prev != NULL).
 * low [B,1,h,w], prev/pred [B,1,H,W]. */
LWS_API int lws_scale_upsample_add_f32(const float* low, const float* prev_or_null, float* pred, int B, int h, int w, int H,
                               int W, lws_stream_t stream);

/* ---- a6 + a7 (+ a2 of the next stage) fused: the tail of one iteration of the stage loop (models/models.py:142-156, :119-121) ----
 * pred = bilinear_resize_halfpixel((softmax_regression(cost) * float(H)) * fl32(1/h), H, W) (+ prev);
 * wflow_next = (bilinear_resize_halfpixel(pred, hn, wn) * float(hn)) * fl32(1/H)   (only if wflow_next_or_null != NULL).
 * One pass: the filtered volume is streamed once, the low-resolution disparity never reaches HBM and pred is not read back.
 * Bit-identical to lws_softmax_regression_f32 + lws_scale_upsample_add_f32 + lws_disp_to_scale_f32.  Applies when H/h == W/w is
 * an integer <= 8 dividing 32 and hn, wn decimate H, W by 1 or an even factor dividing the 32 x 256 tile
 * (lws_regression_tail_supported == 0); otherwise LWS_ERR_UNSUPPORTED and the caller runs the three stand-alone entries. */
LWS_API int lws_regression_tail_supported(int h, int w, int H, int W, int hn, int wn);
LWS_API int lws_regression_tail_f32(const float* cost, const float* prev_or_null, float* pred, float* wflow_next_or_null, int B,
                                    int D, int h, int w, int H, int W, int hn, int wn, float start, float step, lws_stream_t stream);

/* ---- a8+a9: refinement1_left / refinement1_disp / refinement2 + skip
 *             (models/submodules.py:223-327, models/models.py:158-162) --------------------------------
 * pred4 = pred3 + R2(concat[R1_left(left), R1_disp(pred3)]).  left [B,3,H,W], pred3/pred4 [B,1,H,W]. */
LWS_API size_t lws_refinement_packed_floats(void);
/* HOST.  Tensor order in `tensors` (each a host fp32 pointer), with BN = (weight, bias, _mean, _variance):
 *   R1_left : conv0 [32,3,3,3];  then for blocks 1..4: BN(32) x4 tensors, dw [32,1,3,3], pw [32,32,1,1]   -> 25 tensors
 *   R1_disp : conv0 [32,1,3,3];  blocks 1..4 as above                                                      -> 25 tensors
 *   R2      : BN(64) x4 tensors, conv [32,64,3,3]; blocks 1..4 as above; conv_last [1,32,3,3]               -> 30 tensors
 * n_tensors must be 80. */
LWS_API int lws_pack_refinement_weights(const float* const* tensors, int n_tensors, float eps, float* packed);
LWS_API size_t lws_refinement_workspace_bytes(int B, int H, int W);
/* kernel launches lws_refinement_f32 enqueues for this shape under the current options (16 with one block per launch) */
LWS_API int lws_refinement_launches(int B, int H, int W);
LWS_API int lws_refinement_f32(const float* left, const float* pred3, const float* packed_weights, float* pred4, void* ws,
                       size_t ws_bytes, int B, int H, int W, lws_stream_t stream);

/* ---- a8 / a9 as stand-alone layers: the reference calls refinement1_left(x), refinement1_disp(x), refinement2(x) as layers
 *      (models/models.py:158-160).  Module-local BN folding, NCHW in / out, exact-fp32 FFMA kernels; the model's forward uses the
 *      fused lws_refinement_f32 above.
 * lws_refinement1_f32: x [B,in_channels,H,W] (in_channels 1 or 3) -> out [B,32,H,W]  (models/submodules.py:282-300)
 *   HOST pack, 25 tensors: conv0 [32,in,3,3]; blocks 1..4: BN(32) weight, bias, _mean, _variance, dw [32,1,3,3], pw [32,32,1,1]
 * lws_refinement2_f32: x [B,64,H,W] -> out [B,1,H,W] (no skip add; models/submodules.py:302-327)
 *   HOST pack, 30 tensors: BN(64) x4; conv [32,64,3,3]; blocks 1..4 as above; conv_last [1,32,3,3] */
LWS_API size_t lws_refinement1_packed_floats(int in_channels);
LWS_API int lws_pack_refinement1_weights(const float* const* tensors, int n_tensors, int in_channels, float eps, float* packed);
LWS_API size_t lws_refinement1_workspace_bytes(int B, int H, int W);
LWS_API int lws_refinement1_f32(const float* x, const float* packed_weights, float* out, void* ws, size_t ws_bytes, int B,
                                int in_channels, int H, int W, lws_stream_t stream);
LWS_API size_t lws_refinement2_packed_floats(void);
LWS_API int lws_pack_refinement2_weights(const float* const* tensors, int n_tensors, float eps, float* packed);
LWS_API size_t lws_refinement2_workspace_bytes(int B, int H, int W);
LWS_API int lws_refinement2_f32(const float* x, const float* packed_weights, float* out, void* ws, size_t ws_bytes, int B, int H,
                                int W, lws_stream_t stream);

/* One BN-ReLU-DW3x3(dil)-PW1x1 block of the refinement (models/submodules.py:236-261) on its own, on the channels-last bordered
 * tensors the refinement uses internally: act[b][H+32][W+32][32] fp32, 16-pixel zero border (lws_refinement_clp_floats elements).
 * branch 0/1/2 = refinement1_left / refinement1_disp / refinement2, block 0..3 selects the block's weights and dilation from
 * the packed blob.  The refinement's dominant kernel: for per-block tests and for timing it against the HBM roofline. */
LWS_API size_t lws_refinement_clp_floats(int B, int H, int W);
LWS_API int lws_refinement_block_clp_f32(const float* in_clp, float* out_clp, const float* packed_weights, int branch, int block,
                                         int B, int H, int W, lws_stream_t stream);

/* blocks [block0, block0 + nblk) (nblk = 2..4) of a branch in ONE launch: the tensors between the blocks live in row rings that
 * stay resident in L2 (dwsep_chain.cu).  Same layout as lws_refinement_block_clp_f32; the result is bit-identical to nblk
 * single-block calls.  ws: lws_refinement_chain_workspace_bytes (256-byte aligned). */
LWS_API size_t lws_refinement_chain_workspace_bytes(int branch, int block0, int nblk, int B, int H, int W);
LWS_API int lws_refinement_chain_clp_f32(const float* in_clp, float* out_clp, const float* packed_weights, int branch, int block0,
                                         int nblk, void* ws, size_t ws_bytes, int B, int H, int W, lws_stream_t stream);

/* ---- n1 (SURVEY.md 8(f) "next"): feature_extraction  (models/submodules.py:5-188) ----------------------------
 * img [B,3,H,W] -> f8 [B,16,H/8,W/8], f4 [B,16,H/4,W/4], f2 [B,8,H/2,W/2]; H, W multiples of 8.  12 launches, fp32.
 * HOST pack: tensors = for each of the 12 convs in execution order (dres0.0, dres0.2, dres1.0, dres1.2, dres2.conv1..4,
 * dres2.conv5, dres2.conv6 (Conv2DTranspose, weight [Cin,Cout,3,3]), classif1.0, classif1.2): conv weight, then — except
 * for the last conv, which has none — its BatchNorm (weight, bias, _mean, _variance).  n_tensors must be 56. */
LWS_API size_t lws_feature_extraction_packed_floats(void);
LWS_API int lws_pack_feature_extraction_weights(const float* const* tensors, int n_tensors, float eps, float* packed);
LWS_API size_t lws_feature_extraction_workspace_bytes(int B, int H, int W);
LWS_API int lws_feature_extraction_f32(const float* img, const float* packed_weights, float* f8, float* f4, float* f2,
                                       void* ws, size_t ws_bytes, int B, int H, int W, lws_stream_t stream);

/* ---- n2 (SURVEY.md 8(f) "next"): the steps either side of the model in the reference's inference loop ----------------
 * lws_preprocess_bgr_u8  (inference.py:93-103): img [B,h,w,3] uint8 HWC BGR (cv2.imread) -> bottom-right crop th x tw, BGR->RGB,
 *   ToTensor + Normalize -> out [B,3,th,tw] fp32.  lut [3][256] (device, RGB order) holds ((v/255) - mean[c]) / std[c] evaluated by
 *   the host in fp32 exactly as the reference does, so the output is bit-identical to the CPU preprocessing.
 * lws_disparity_to_u8    (inference.py:114-115): gray[i] = (uint8)disp[i] (numpy astype: truncate, wrap modulo 256) and/or
 *   bgr[i] = COLORMAP_JET[gray[i]] (cv2.applyColorMap; convertScaleAbs(alpha=1, beta=0) is the identity on uint8). */
LWS_API int lws_preprocess_bgr_u8(const uint8_t* img, const float* lut, float* out, int B, int h, int w, int th, int tw,
                                  lws_stream_t stream);
LWS_API int lws_disparity_to_u8(const float* disp, uint8_t* gray_or_null, uint8_t* bgr_or_null, long long n, lws_stream_t stream);

/* ---- n4 (SURVEY.md 8(f) "next"): training path of the hot-path kernels (reference train.py:127-166) ---------------------------
 * Backward of a1 / a4 / a6 and the multi-stage smooth-L1 loss.  Gradients are those autograd gives the reference's formulation:
 * d|x| = sign(x) (0 at 0); bilinear sampling with zero padding differentiated w.r.t. the sampled features and the disparity.
 * The convolution stacks' backward is not part of this tier.
 * lws_cost_volume_l1_bwd_f32: gcost [B,maxdisp/stride,H,W] -> gL, gR [B,C,H,W]
 * lws_warp_residual_volume_l1_bwd_f32: gcost [B,2m-1,H,W] -> gL, gR [B,C,H,W] (gR is zero-filled here, then accumulated with
 *   atomics: the summation order, hence the last bits, may vary from run to run), gdisp [B,1,H,W]
 * lws_softmax_regression_bwd_f32: cost [B,D,H,W], glow [B,1,H,W] -> gcost [B,D,H,W]
 * lws_smooth_l1_multistage_loss_f32: out[s] = weights[s] * mean_{gt < maxdisp} smooth_l1(preds[s] - gt), s < n_stages <= 4, and
 *   out[4] = number of masked pixels (all losses 0 when it is 0: the reference skips such batches, train.py:139-140);
 *   grads[s] (optional) = d out[s] / d preds[s].  preds / grads: HOST arrays of n_stages device pointers to n floats each;
 *   weights: HOST array; out: DEVICE float[5]; deterministic (fixed reduction order). */
LWS_API int lws_cost_volume_l1_bwd_f32(const float* L, const float* R, const float* gcost, float* gL, float* gR, int B, int C, int H,
                                       int W, int maxdisp, int stride, lws_stream_t stream);
LWS_API int lws_warp_residual_volume_l1_bwd_f32(const float* L, const float* R, const float* disp, const float* gcost, float* gL,
                                                float* gR, float* gdisp, int B, int C, int H, int W, int m, int stride,
                                                lws_stream_t stream);
LWS_API int lws_softmax_regression_bwd_f32(const float* cost, const float* glow, float* gcost, int B, int D, int H, int W, float start,
                                           float step, lws_stream_t stream);
LWS_API size_t lws_smooth_l1_loss_workspace_bytes(long long n);
LWS_API int lws_smooth_l1_multistage_loss_f32(const float* const* preds, const float* gt, const float* weights, int n_stages,
                                              long long n, float maxdisp, float* out, float* const* grads_or_null, void* ws,
                                              size_t ws_bytes, lws_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* LWS_H_ */
