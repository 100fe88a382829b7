// paddle_binding/lws_paddle_ops.cc -- PaddlePaddle (>= 2.1, custom-operator API) binding of liblws_b200 (include/lws.h).
//
// This is the file a maintainer of the reference (PrinceVictor/LWSNet) adds to call the B200 kernels from models/models.py and
// models/submodules.py: every op below forwards to ONE C-ABI entry point, takes / returns paddle::Tensor (contiguous fp32 NCHW
// on the GPU), allocates outputs and scratch with Paddle's allocator and enqueues on the tensor's stream.
//
// Build (on a machine with Paddle):   python paddle_binding/setup_paddle.py install     (CUDAExtension, links -llws_b200)
// Here Paddle is not installable (no network): the file is type-checked against paddle_binding/stub/paddle/extension.h by
// tests/test_abi.py::test_paddle_binding_compiles; packed weight blobs come from the lws_pack_* host functions (run once per
// checkpoint over the layer's state dict, see models_patch.py).
#include <cstdint>
#include <vector>

#include "paddle/extension.h"

#include "lws.h"

namespace {

void lws_check(int rc, const char* what) { PD_CHECK(rc == 0, what, ": ", lws_status_string(rc)); }
inline int i32(int64_t v) { return static_cast<int>(v); }
paddle::Tensor scratch(size_t bytes, const paddle::Tensor& like) {
  return paddle::empty({static_cast<int64_t>(bytes < 256 ? 256 : bytes)}, paddle::DataType::UINT8, like.place());
}

// ---- a1: LWSNet._build_volume_2d (models/models.py:58-76) ----------------------------------------------------------------------
std::vector<paddle::Tensor> CostVolumeL1(const paddle::Tensor& L, const paddle::Tensor& R, int maxdisp, int stride) {
  PD_CHECK(maxdisp % stride == 0, "maxdisp % stride != 0");  // the reference's assert (models/models.py:63)
  const auto s = L.shape();                                   // [B,C,H,W]
  auto cost = paddle::empty({s[0], maxdisp / stride, s[2], s[3]}, paddle::DataType::FLOAT32, L.place());
  lws_check(lws_cost_volume_l1_f32(L.data<float>(), R.data<float>(), cost.data<float>(), i32(s[0]), i32(s[1]), i32(s[2]), i32(s[3]),
                                   maxdisp, stride, L.stream()),
            "lws_cost_volume_l1_f32");
  return {cost};
}

// ---- a2: wflow (models/models.py:119-121) --------------------------------------------------------------------------------------
std::vector<paddle::Tensor> DispToScale(const paddle::Tensor& pred_full, int h, int w) {
  const auto s = pred_full.shape();  // [B,1,H,W]
  auto wflow = paddle::empty({s[0], 1, h, w}, paddle::DataType::FLOAT32, pred_full.place());
  lws_check(lws_disp_to_scale_f32(pred_full.data<float>(), wflow.data<float>(), i32(s[0]), i32(s[2]), i32(s[3]), h, w,
                                  pred_full.stream()),
            "lws_disp_to_scale_f32");
  return {wflow};
}

// ---- a3: LWSNet.warp (models/models.py:28-55) ----------------------------------------------------------------------------------
std::vector<paddle::Tensor> WarpBilinear(const paddle::Tensor& x, const paddle::Tensor& disp) {
  const auto s = x.shape();  // [N,C,H,W]
  auto out = paddle::empty_like(x);
  lws_check(lws_warp_bilinear_f32(x.data<float>(), disp.data<float>(), out.data<float>(), i32(s[0]), i32(s[1]), i32(s[2]), i32(s[3]),
                                  x.stream()),
            "lws_warp_bilinear_f32");
  return {out};
}

// ---- a4: LWSNet._build_volume_2d3 (models/models.py:78-104) --------------------------------------------------------------------
std::vector<paddle::Tensor> WarpResidualVolumeL1(const paddle::Tensor& L, const paddle::Tensor& R, const paddle::Tensor& disp,
                                                 int maxdisp, int stride) {
  const auto s = L.shape();
  auto cost = paddle::empty({s[0], 2 * maxdisp - 1, s[2], s[3]}, paddle::DataType::FLOAT32, L.place());
  lws_check(lws_warp_residual_volume_l1_f32(L.data<float>(), R.data<float>(), disp.data<float>(), cost.data<float>(), i32(s[0]),
                                            i32(s[1]), i32(s[2]), i32(s[3]), maxdisp, stride, L.stream()),
            "lws_warp_residual_volume_l1_f32");
  return {cost};
}

// ---- a5: volume_postprocess[scale](cost) + cost (models/submodules.py:190-221, models/models.py:136-138) -----------------------
// `packed`: lws_pack_conv3d_stack_weights(...) over the stack's state dict, copied to the device once per checkpoint.
std::vector<paddle::Tensor> Conv3dStack(const paddle::Tensor& cost, const paddle::Tensor& packed, int C, int layers, int add_skip) {
  const auto s = cost.shape();  // [B,D,H,W]
  auto out = paddle::empty_like(cost);
  const size_t ws_bytes = lws_conv3d_stack_workspace_bytes(i32(s[0]), i32(s[1]), i32(s[2]), i32(s[3]), C, layers);
  auto ws = scratch(ws_bytes, cost);
  lws_check(lws_conv3d_stack_f32(cost.data<float>(), packed.data<float>(), out.data<float>(), ws.data<uint8_t>(), ws_bytes, i32(s[0]),
                                 i32(s[1]), i32(s[2]), i32(s[3]), C, layers, add_skip, cost.stream()),
            "lws_conv3d_stack_f32");
  return {out};
}

// ---- a2 + a5 fused: stage 1 of LWSNet.forward (models/models.py:131-138), the volume built inside the first conv kernel ----------
std::vector<paddle::Tensor> CostVolumeConv3dStack(const paddle::Tensor& L, const paddle::Tensor& R, const paddle::Tensor& packed,
                                                  int maxdisp, int C, int layers, int add_skip) {
  const auto s = L.shape();  // [B,Cf,H,W]
  auto cost = paddle::empty({s[0], maxdisp, s[2], s[3]}, paddle::DataType::FLOAT32, L.place());
  auto out = paddle::empty_like(cost);
  const size_t ws_bytes = lws_conv3d_stack_workspace_bytes(i32(s[0]), maxdisp, i32(s[2]), i32(s[3]), C, layers);
  auto ws = scratch(ws_bytes, L);
  lws_check(lws_cost_volume_conv3d_stack_f32(L.data<float>(), R.data<float>(), packed.data<float>(), cost.data<float>(),
                                             out.data<float>(), ws.data<uint8_t>(), ws_bytes, i32(s[0]), i32(s[1]), i32(s[2]),
                                             i32(s[3]), maxdisp, C, layers, add_skip, L.stream()),
            "lws_cost_volume_conv3d_stack_f32");
  return {cost, out};
}

// ---- a6: F.softmax(-cost, axis=1) + disparity_regression (models/models.py:142,151-152,167-179) --------------------------------
std::vector<paddle::Tensor> SoftmaxRegression(const paddle::Tensor& cost, float start, float step) {
  const auto s = cost.shape();  // [B,D,H,W]
  auto low = paddle::empty({s[0], 1, s[2], s[3]}, paddle::DataType::FLOAT32, cost.place());
  lws_check(lws_softmax_regression_f32(cost.data<float>(), low.data<float>(), i32(s[0]), i32(s[1]), i32(s[2]), i32(s[3]), start, step,
                                       cost.stream()),
            "lws_softmax_regression_f32");
  return {low};
}
// the stand-alone class disparity_regression(start, end, stride).forward(prob) (models/models.py:167-179)
std::vector<paddle::Tensor> DisparityRegression(const paddle::Tensor& prob, float start, float step) {
  const auto s = prob.shape();
  auto out = paddle::empty({s[0], 1, s[2], s[3]}, paddle::DataType::FLOAT32, prob.place());
  lws_check(lws_disparity_regression_f32(prob.data<float>(), out.data<float>(), i32(s[0]), i32(s[1]), i32(s[2]), i32(s[3]), start,
                                         step, prob.stream()),
            "lws_disparity_regression_f32");
  return {out};
}

// ---- a7: rescale + F.interpolate + skip (models/models.py:145-148,153-156) -----------------------------------------------------
std::vector<paddle::Tensor> ScaleUpsampleAdd(const paddle::Tensor& low, const paddle::Tensor& prev, int H, int W, int has_prev) {
  const auto s = low.shape();  // [B,1,h,w]
  auto pred = paddle::empty({s[0], 1, H, W}, paddle::DataType::FLOAT32, low.place());
  lws_check(lws_scale_upsample_add_f32(low.data<float>(), has_prev ? prev.data<float>() : nullptr, pred.data<float>(), i32(s[0]),
                                       i32(s[2]), i32(s[3]), H, W, low.stream()),
            "lws_scale_upsample_add_f32");
  return {pred};
}

// ---- a6 + a7 (+ a2 of the next stage) fused: the tail of one stage-loop iteration (models/models.py:142-156, :119-121) ----------
// has_prev = 0 for the first stage; hn = wn = 0 when there is no next stage (Wflow is then an empty [B,1,0,0] tensor)
std::vector<paddle::Tensor> RegressionTail(const paddle::Tensor& cost, const paddle::Tensor& prev, int H, int W, int hn, int wn,
                                           float start, float step, int has_prev) {
  const auto s = cost.shape();  // [B,D,h,w]
  auto pred = paddle::empty({s[0], 1, H, W}, paddle::DataType::FLOAT32, cost.place());
  auto wflow = paddle::empty({s[0], 1, hn, wn}, paddle::DataType::FLOAT32, cost.place());
  lws_check(lws_regression_tail_f32(cost.data<float>(), has_prev ? prev.data<float>() : nullptr, pred.data<float>(),
                                    hn > 0 ? wflow.data<float>() : nullptr, i32(s[0]), i32(s[1]), i32(s[2]), i32(s[3]), H, W, hn, wn,
                                    start, step, cost.stream()),
            "lws_regression_tail_f32");
  return {pred, wflow};
}

// ---- a8 + a9 fused: pred3 + refinement2(concat[refinement1_left(left), refinement1_disp(pred3)]) (models/models.py:158-162) ----
std::vector<paddle::Tensor> Refinement(const paddle::Tensor& left, const paddle::Tensor& pred3, const paddle::Tensor& packed) {
  const auto s = left.shape();  // [B,3,H,W]
  auto pred4 = paddle::empty_like(pred3);
  const size_t ws_bytes = lws_refinement_workspace_bytes(i32(s[0]), i32(s[2]), i32(s[3]));
  auto ws = scratch(ws_bytes, left);
  lws_check(lws_refinement_f32(left.data<float>(), pred3.data<float>(), packed.data<float>(), pred4.data<float>(), ws.data<uint8_t>(),
                               ws_bytes, i32(s[0]), i32(s[2]), i32(s[3]), left.stream()),
            "lws_refinement_f32");
  return {pred4};
}
// ---- a8 / a9 as layers: refinement1(in, 32)(x), refinement2(64, 32)(x) (models/submodules.py:282-327) --------------------------
std::vector<paddle::Tensor> Refinement1(const paddle::Tensor& x, const paddle::Tensor& packed) {
  const auto s = x.shape();  // [B,in,H,W]
  auto out = paddle::empty({s[0], 32, s[2], s[3]}, paddle::DataType::FLOAT32, x.place());
  const size_t ws_bytes = lws_refinement1_workspace_bytes(i32(s[0]), i32(s[2]), i32(s[3]));
  auto ws = scratch(ws_bytes, x);
  lws_check(lws_refinement1_f32(x.data<float>(), packed.data<float>(), out.data<float>(), ws.data<uint8_t>(), ws_bytes, i32(s[0]),
                                i32(s[1]), i32(s[2]), i32(s[3]), x.stream()),
            "lws_refinement1_f32");
  return {out};
}
std::vector<paddle::Tensor> Refinement2(const paddle::Tensor& x, const paddle::Tensor& packed) {
  const auto s = x.shape();  // [B,64,H,W]
  auto out = paddle::empty({s[0], 1, s[2], s[3]}, paddle::DataType::FLOAT32, x.place());
  const size_t ws_bytes = lws_refinement2_workspace_bytes(i32(s[0]), i32(s[2]), i32(s[3]));
  auto ws = scratch(ws_bytes, x);
  lws_check(lws_refinement2_f32(x.data<float>(), packed.data<float>(), out.data<float>(), ws.data<uint8_t>(), ws_bytes, i32(s[0]),
                                i32(s[2]), i32(s[3]), x.stream()),
            "lws_refinement2_f32");
  return {out};
}

// ---- n1: feature_extraction (models/submodules.py:113-188) ---------------------------------------------------------------------
std::vector<paddle::Tensor> FeatureExtraction(const paddle::Tensor& img, const paddle::Tensor& packed) {
  const auto s = img.shape();  // [B,3,H,W], H and W multiples of 8
  auto f8 = paddle::empty({s[0], 16, s[2] / 8, s[3] / 8}, paddle::DataType::FLOAT32, img.place());
  auto f4 = paddle::empty({s[0], 16, s[2] / 4, s[3] / 4}, paddle::DataType::FLOAT32, img.place());
  auto f2 = paddle::empty({s[0], 8, s[2] / 2, s[3] / 2}, paddle::DataType::FLOAT32, img.place());
  const size_t ws_bytes = lws_feature_extraction_workspace_bytes(i32(s[0]), i32(s[2]), i32(s[3]));
  auto ws = scratch(ws_bytes, img);
  lws_check(lws_feature_extraction_f32(img.data<float>(), packed.data<float>(), f8.data<float>(), f4.data<float>(), f2.data<float>(),
                                       ws.data<uint8_t>(), ws_bytes, i32(s[0]), i32(s[2]), i32(s[3]), img.stream()),
            "lws_feature_extraction_f32");
  return {f8, f4, f2};
}

// ---- n2: the steps either side of the model in inference.py:93-115 -------------------------------------------------------------
std::vector<paddle::Tensor> PreprocessBgrU8(const paddle::Tensor& img, const paddle::Tensor& lut, int th, int tw) {
  const auto s = img.shape();  // [B,h,w,3] uint8
  auto out = paddle::empty({s[0], 3, th, tw}, paddle::DataType::FLOAT32, img.place());
  lws_check(lws_preprocess_bgr_u8(img.data<uint8_t>(), lut.data<float>(), out.data<float>(), i32(s[0]), i32(s[1]), i32(s[2]), th, tw,
                                  img.stream()),
            "lws_preprocess_bgr_u8");
  return {out};
}
std::vector<paddle::Tensor> DisparityToU8(const paddle::Tensor& disp) {
  auto shape = disp.shape();
  int64_t n = 1;
  for (auto d : shape) n *= d;
  auto gray = paddle::empty(shape, paddle::DataType::UINT8, disp.place());
  shape.push_back(3);
  auto bgr = paddle::empty(shape, paddle::DataType::UINT8, disp.place());
  lws_check(lws_disparity_to_u8(disp.data<float>(), gray.data<uint8_t>(), bgr.data<uint8_t>(), n, disp.stream()), "lws_disparity_to_u8");
  return {gray, bgr};
}

// ---- n4: backward of a1 / a4 / a6 (registered as the grad ops of the forward ops) and the loss of train.py:145-155 --------------
std::vector<paddle::Tensor> CostVolumeL1Grad(const paddle::Tensor& L, const paddle::Tensor& R, const paddle::Tensor& gcost, int maxdisp,
                                             int stride) {
  const auto s = L.shape();
  auto gL = paddle::empty_like(L), gR = paddle::empty_like(R);
  lws_check(lws_cost_volume_l1_bwd_f32(L.data<float>(), R.data<float>(), gcost.data<float>(), gL.data<float>(), gR.data<float>(),
                                       i32(s[0]), i32(s[1]), i32(s[2]), i32(s[3]), maxdisp, stride, L.stream()),
            "lws_cost_volume_l1_bwd_f32");
  return {gL, gR};
}
std::vector<paddle::Tensor> WarpResidualVolumeL1Grad(const paddle::Tensor& L, const paddle::Tensor& R, const paddle::Tensor& disp,
                                                     const paddle::Tensor& gcost, int maxdisp, int stride) {
  const auto s = L.shape();
  auto gL = paddle::empty_like(L), gR = paddle::empty_like(R), gdisp = paddle::empty_like(disp);
  lws_check(lws_warp_residual_volume_l1_bwd_f32(L.data<float>(), R.data<float>(), disp.data<float>(), gcost.data<float>(),
                                                gL.data<float>(), gR.data<float>(), gdisp.data<float>(), i32(s[0]), i32(s[1]),
                                                i32(s[2]), i32(s[3]), maxdisp, stride, L.stream()),
            "lws_warp_residual_volume_l1_bwd_f32");
  return {gL, gR, gdisp};
}
std::vector<paddle::Tensor> SoftmaxRegressionGrad(const paddle::Tensor& cost, const paddle::Tensor& glow, float start, float step) {
  const auto s = cost.shape();
  auto gcost = paddle::empty_like(cost);
  lws_check(lws_softmax_regression_bwd_f32(cost.data<float>(), glow.data<float>(), gcost.data<float>(), i32(s[0]), i32(s[1]),
                                           i32(s[2]), i32(s[3]), start, step, cost.stream()),
            "lws_softmax_regression_bwd_f32");
  return {gcost};
}
// loss[s] = w_s * smooth_l1(pred_s[gt < maxdisp], gt[gt < maxdisp], mean) for the four stage outputs, with their gradients
std::vector<paddle::Tensor> SmoothL1MultistageLoss(const paddle::Tensor& p0, const paddle::Tensor& p1, const paddle::Tensor& p2,
                                                   const paddle::Tensor& p3, const paddle::Tensor& gt, float w0, float w1, float w2,
                                                   float w3, float maxdisp) {
  int64_t n = 1;
  for (auto d : gt.shape()) n *= d;
  auto out = paddle::empty({5}, paddle::DataType::FLOAT32, gt.place());
  auto g0 = paddle::empty_like(p0), g1 = paddle::empty_like(p1), g2 = paddle::empty_like(p2), g3 = paddle::empty_like(p3);
  const float* preds[4] = {p0.data<float>(), p1.data<float>(), p2.data<float>(), p3.data<float>()};
  float* grads[4] = {g0.data<float>(), g1.data<float>(), g2.data<float>(), g3.data<float>()};
  const float w[4] = {w0, w1, w2, w3};
  const size_t ws_bytes = lws_smooth_l1_loss_workspace_bytes(n);
  auto ws = scratch(ws_bytes, gt);
  lws_check(lws_smooth_l1_multistage_loss_f32(preds, gt.data<float>(), w, 4, n, maxdisp, out.data<float>(), grads, ws.data<uint8_t>(),
                                              ws_bytes, gt.stream()),
            "lws_smooth_l1_multistage_loss_f32");
  return {out, g0, g1, g2, g3};
}

}  // namespace

PD_BUILD_GRAD_OP(lws_cost_volume_l1).Inputs({"L", "R", "Cost@GRAD"}).Outputs({"L@GRAD", "R@GRAD"}).Attrs({"maxdisp: int", "stride: int"}).SetKernelFn(PD_KERNEL(CostVolumeL1Grad));
PD_BUILD_GRAD_OP(lws_warp_residual_volume_l1).Inputs({"L", "R", "Disp", "Cost@GRAD"}).Outputs({"L@GRAD", "R@GRAD", "Disp@GRAD"}).Attrs({"maxdisp: int", "stride: int"}).SetKernelFn(PD_KERNEL(WarpResidualVolumeL1Grad));
PD_BUILD_GRAD_OP(lws_softmax_regression).Inputs({"Cost", "Low@GRAD"}).Outputs({"Cost@GRAD"}).Attrs({"start: float", "step: float"}).SetKernelFn(PD_KERNEL(SoftmaxRegressionGrad));
PD_BUILD_OP(lws_smooth_l1_multistage_loss).Inputs({"P0", "P1", "P2", "P3", "Gt"}).Outputs({"Loss", "G0", "G1", "G2", "G3"}).Attrs({"w0: float", "w1: float", "w2: float", "w3: float", "maxdisp: float"}).SetKernelFn(PD_KERNEL(SmoothL1MultistageLoss));
PD_BUILD_OP(lws_cost_volume_l1).Inputs({"L", "R"}).Outputs({"Cost"}).Attrs({"maxdisp: int", "stride: int"}).SetKernelFn(PD_KERNEL(CostVolumeL1));
PD_BUILD_OP(lws_disp_to_scale).Inputs({"PredFull"}).Outputs({"Wflow"}).Attrs({"h: int", "w: int"}).SetKernelFn(PD_KERNEL(DispToScale));
PD_BUILD_OP(lws_warp_bilinear).Inputs({"X", "Disp"}).Outputs({"Out"}).SetKernelFn(PD_KERNEL(WarpBilinear));
PD_BUILD_OP(lws_warp_residual_volume_l1).Inputs({"L", "R", "Disp"}).Outputs({"Cost"}).Attrs({"maxdisp: int", "stride: int"}).SetKernelFn(PD_KERNEL(WarpResidualVolumeL1));
PD_BUILD_OP(lws_conv3d_stack).Inputs({"Cost", "Packed"}).Outputs({"Out"}).Attrs({"C: int", "layers: int", "add_skip: int"}).SetKernelFn(PD_KERNEL(Conv3dStack));
PD_BUILD_OP(lws_cost_volume_conv3d_stack).Inputs({"L", "R", "Packed"}).Outputs({"Cost", "Out"}).Attrs({"maxdisp: int", "C: int", "layers: int", "add_skip: int"}).SetKernelFn(PD_KERNEL(CostVolumeConv3dStack));
PD_BUILD_OP(lws_softmax_regression).Inputs({"Cost"}).Outputs({"Low"}).Attrs({"start: float", "step: float"}).SetKernelFn(PD_KERNEL(SoftmaxRegression));
PD_BUILD_OP(lws_disparity_regression).Inputs({"Prob"}).Outputs({"Out"}).Attrs({"start: float", "step: float"}).SetKernelFn(PD_KERNEL(DisparityRegression));
PD_BUILD_OP(lws_scale_upsample_add).Inputs({"Low", "Prev"}).Outputs({"Pred"}).Attrs({"H: int", "W: int", "has_prev: int"}).SetKernelFn(PD_KERNEL(ScaleUpsampleAdd));
PD_BUILD_OP(lws_regression_tail).Inputs({"Cost", "Prev"}).Outputs({"Pred", "WflowNext"}).Attrs({"H: int", "W: int", "hn: int", "wn: int", "start: float", "step: float", "has_prev: int"}).SetKernelFn(PD_KERNEL(RegressionTail));
PD_BUILD_OP(lws_refinement).Inputs({"Left", "Pred3", "Packed"}).Outputs({"Pred4"}).SetKernelFn(PD_KERNEL(Refinement));
PD_BUILD_OP(lws_refinement1).Inputs({"X", "Packed"}).Outputs({"Out"}).SetKernelFn(PD_KERNEL(Refinement1));
PD_BUILD_OP(lws_refinement2).Inputs({"X", "Packed"}).Outputs({"Out"}).SetKernelFn(PD_KERNEL(Refinement2));
PD_BUILD_OP(lws_feature_extraction).Inputs({"Img", "Packed"}).Outputs({"F8", "F4", "F2"}).SetKernelFn(PD_KERNEL(FeatureExtraction));
PD_BUILD_OP(lws_preprocess_bgr_u8).Inputs({"Img", "Lut"}).Outputs({"Out"}).Attrs({"th: int", "tw: int"}).SetKernelFn(PD_KERNEL(PreprocessBgrU8));
PD_BUILD_OP(lws_disparity_to_u8).Inputs({"Disp"}).Outputs({"Gray", "Bgr"}).SetKernelFn(PD_KERNEL(DisparityToU8));
