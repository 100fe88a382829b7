"""What changes in the REFERENCE's models/models.py once `lws_paddle_ops` (paddle_binding/lws_paddle_ops.cc) is built: the bodies of
the hot-path methods become single custom-op calls; names, signatures and the state dict stay as they are.  This file documents the
edit (it imports paddle, so it does not run in this image); the same mapping, executed through ctypes + torch, is lwsnet_b200/.

    reference models/models.py                              ->  custom op (C-ABI entry)
    LWSNet.warp(x, disp)                          :28-55    ->  lws_warp_bilinear          (lws_warp_bilinear_f32)
    LWSNet._build_volume_2d(...)                  :58-76    ->  lws_cost_volume_l1         (lws_cost_volume_l1_f32)
    LWSNet._build_volume_2d3(...)                 :78-104   ->  lws_warp_residual_volume_l1
    forward: wflow                                :119-121  ->  lws_disp_to_scale
    forward: volume_postprocess[s](cost) + cost   :136-138  ->  lws_conv3d_stack (packed = lws_pack_conv3d_stack_weights)
    forward: softmax + disparity_regression       :142-152  ->  lws_softmax_regression
    forward: * H / h, interpolate, + pred[s-1]    :145-156  ->  lws_scale_upsample_add
    forward: refinement1_left/_disp, refinement2  :158-162  ->  lws_refinement (fused) | lws_refinement1 / lws_refinement2 (layers)
    disparity_regression.forward                  :167-179  ->  lws_disparity_regression
    feature_extraction.forward (submodules.py:176-188)      ->  lws_feature_extraction
"""
import ctypes

import numpy as np
import paddle
import lws_paddle_ops as lws  # built by paddle_binding/setup_paddle.py

_c = ctypes.CDLL("liblws_b200.so")


def _pack_stack(seq, C, layers):
    """lws_pack_conv3d_stack_weights over a post_3dconvs Sequential (children: [BatchNorm3D, ReLU, Conv3D])."""
    n = layers + 2
    arr = lambda ts: (ctypes.c_void_p * n)(*[t.ctypes.data for t in ts])
    keep = [[np.ascontiguousarray(getattr(blk[0], k).numpy(), np.float32) for blk in seq] for k in ("weight", "bias", "_mean", "_variance")]
    conv = [np.ascontiguousarray(blk[2].weight.numpy(), np.float32) for blk in seq]
    _c.lws_conv3d_stack_packed_floats.restype = ctypes.c_size_t
    packed = np.zeros(_c.lws_conv3d_stack_packed_floats(C, layers), np.float32)
    rc = _c.lws_pack_conv3d_stack_weights(arr(conv), arr(keep[0]), arr(keep[1]), arr(keep[2]), arr(keep[3]), ctypes.c_float(1e-5), C,
                                          layers, packed.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return paddle.to_tensor(packed)


class LWSNetPatch:
    """Method bodies to paste over the reference's LWSNet (models/models.py)."""

    def warp(self, x, disp):
        return lws.lws_warp_bilinear(x, disp)

    def _build_volume_2d(self, feat_l, feat_r, maxdisp, stride=1):
        assert maxdisp % stride == 0
        return lws.lws_cost_volume_l1(feat_l, feat_r, maxdisp, stride)

    def _build_volume_2d3(self, feat_l, feat_r, maxdisp, disp, stride=1):
        return lws.lws_warp_residual_volume_l1(feat_l, feat_r, disp, maxdisp, stride)

    def forward(self, left_input, right_input):
        img_h, img_w = left_input.shape[2], left_input.shape[3]
        feats_l = self.feature_extraction(left_input)
        feats_r = self.feature_extraction(right_input)
        pred = []
        for scale in range(len(feats_l)):
            if scale > 0:
                wflow = lws.lws_disp_to_scale(pred[scale - 1], feats_l[scale].shape[2], feats_l[scale].shape[3])
                cost = self._build_volume_2d3(feats_l[scale], feats_r[scale], self.maxdisplist[scale], wflow, stride=1)
                start = float(-self.maxdisplist[scale] + 1)
            else:
                cost = self._build_volume_2d(feats_l[scale], feats_r[scale], self.maxdisplist[scale], stride=1)
                start = 0.0
            C = self.channels_3d * self.growth_rate[scale]
            cost = lws.lws_conv3d_stack(cost, self._packed_stack[scale], C, self.layers_3d, 1)
            low = lws.lws_softmax_regression(cost, start, 1.0)
            prev = pred[scale - 1] if scale > 0 else low
            pred.append(lws.lws_scale_upsample_add(low, prev, img_h, img_w, int(scale > 0)))
        pred.append(lws.lws_refinement(left_input, pred[2], self._packed_refinement))
        return pred
