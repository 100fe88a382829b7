"""Mirror of the reference's models/models.py: same class / method names and signatures, B200 kernels underneath.

``LWSNet(args)`` reads the same four attributes (``maxdisplist, layers_3d, channels_3d, growth_rate``; reference
models/models.py:11-14), owns sub-layers under the same attribute names (``feature_extraction``,
``volume_postprocess``, ``refinement1_left``, ``refinement1_disp``, ``refinement2``; models.py:16-26) so state-dict keys
match the Paddle checkpoint grammar, and ``forward(left, right)`` returns the same list of four ``[B,1,H,W]``
disparity maps.  Every hot-path function runs through the C ABI (``ops``); there is no CPU path.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from ._lib import LwsError, get_option
from .submodules import BN_EPS, _pack_key, feature_extraction, post_3dconvs, refinement1, refinement2, refinement_tensor_list


class LWSNet(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.maxdisplist = args.maxdisplist
        self.layers_3d = args.layers_3d
        self.channels_3d = args.channels_3d
        self.growth_rate = args.growth_rate

        self.feature_extraction = feature_extraction()
        self.volume_postprocess = nn.ModuleList(
            [post_3dconvs(self.layers_3d, self.channels_3d * self.growth_rate[i]) for i in range(3)])
        self.refinement1_left = refinement1(in_channels=3, out_channels=32)
        self.refinement1_disp = refinement1(in_channels=1, out_channels=32)
        self.refinement2 = refinement2(in_channels=64, out_channels=32)
        self._ref_packed = None
        self._ref_key = None
        self.eval()  # inference only: BatchNorm always uses its running statistics

    # -- Paddle Layer API used by the reference's scripts (inference.py:45, train.py:84-85) ------------------------
    def set_state_dict(self, state_dict, use_structured_name=True):
        """``model.set_state_dict(paddle.load(path))``: accepts {key: ndarray | tensor | (name, ndarray)} with the Paddle key
        grammar; unlike Paddle, a key or shape mismatch raises instead of warning."""
        from .checkpoint import convert_state
        res = self.load_state_dict(convert_state(state_dict, self.state_dict()), strict=True)
        self.repack()
        return res

    set_dict = set_state_dict  # Paddle 2.0 alias

    def load_pdparams(self, path):
        """Restore a reference ``.pdparams`` checkpoint (restricted unpickler, no Paddle needed)."""
        from .checkpoint import load_pdparams
        return self.set_state_dict(load_pdparams(path))

    # -- reference models/models.py:28-55 ------------------------------------------------------------------------
    def warp(self, x, disp):
        """x [B,C,H,W] (right features), disp [B,1,H,W] -> x sampled at (x - disp, y), bilinear, zero padding."""
        return ops.warp_bilinear(x, disp)

    # -- reference models/models.py:58-76 ------------------------------------------------------------------------
    def _build_volume_2d(self, feat_l, feat_r, maxdisp, stride=1):
        assert maxdisp % stride == 0
        return ops.cost_volume_l1(feat_l, feat_r, maxdisp, stride)

    # -- reference models/models.py:78-104 -----------------------------------------------------------------------
    def _build_volume_2d3(self, feat_l, feat_r, maxdisp, disp, stride=1):
        return ops.warp_residual_volume_l1(feat_l, feat_r, disp, maxdisp, stride)

    def _refinement_packed(self, device):
        mods = (self.refinement1_left, self.refinement1_disp, self.refinement2)
        key = tuple(_pack_key(m, device) for m in mods)
        if self._ref_packed is None or self._ref_key != key:
            self._ref_packed = ops.pack_refinement(refinement_tensor_list(*mods), BN_EPS).to(device)
            self._ref_key = key
        return self._ref_packed

    def repack(self):
        """Drop every cached BN-folded weight blob.  The caches notice ordinary weight updates by themselves (load_state_dict,
        set_state_dict, in-place tensor ops); edits through ``p.data`` bypass torch's version counter and need this call."""
        self._ref_packed = None
        for m in self.modules():
            if m is not self and hasattr(m, "repack"):
                m.repack()

    def pack_state(self, device):
        """Identity of the packed weight blobs the next forward on `device` will read (StereoEngine compares it with what a
        captured CUDA graph baked in)."""
        blobs = [self.feature_extraction.packed(device), self._refinement_packed(device)]
        blobs += [vp.packed(device) for vp in self.volume_postprocess]
        return tuple((b.data_ptr(), b._version) for b in blobs)

    # one iteration of the stage loop, reference models/models.py:115-156
    def _stage(self, scale, feat_l, feat_r, prev_pred, img_h, img_w, out=None, wflow=None, next_hw=None):
        """One iteration of the stage loop.  ``wflow``: this stage's wflow if the previous stage's fused tail already produced it;
        ``next_hw``: feature size of the next stage -- then the tail also emits the next wflow and (pred, wflow_next) is returned."""
        if scale > 0:
            if wflow is None:
                wflow = ops.disp_to_scale(prev_pred, feat_l.shape[2], feat_l.shape[3])             # models.py:119-121
            cost = self._build_volume_2d3(feat_l, feat_r, self.maxdisplist[scale], wflow, stride=1)  # :123-127
            start = float(-self.maxdisplist[scale] + 1)
            cost = self.volume_postprocess[scale].run(cost, add_skip=True)                          # :136-138
        else:
            start = 0.0
            vp, D = self.volume_postprocess[scale], self.maxdisplist[scale]
            B, Cf, h, w = feat_l.shape
            if get_option("fuse_volume") and ops.cost_volume_conv3d_stack_supported(B, Cf, h, w, D, vp.channels, vp.layers):
                # :131-138 in one call: the volume is built inside the first conv kernel's shared-memory tap window
                _, cost = ops.cost_volume_conv3d_stack(feat_l, feat_r, D, vp.packed(feat_l.device), vp.channels, vp.layers)
            else:
                cost = self._build_volume_2d(feat_l, feat_r, D, stride=1)                            # :131-134
                cost = vp.run(cost, add_skip=True)                                                   # :136-138
        # :142 / :151-152 softmax + regression, :145-148 / :153-156 rescale + upsample + skip, and the next stage's :119-121
        pred, wflow_next = ops.regression_tail(cost, prev_pred if scale > 0 else None, img_h, img_w, start, 1.0, next_hw=next_hw,
                                               out=out)
        return pred if next_hw is None else (pred, wflow_next)

    def _refine(self, left_input, pred3, out=None):
        return ops.refinement(left_input, pred3, self._refinement_packed(left_input.device), out=out)  # :158-162

    # -- reference models/models.py:106-164 ----------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, left_input, right_input, out=None):
        """Returns the list of 4 stage disparities [B,1,H,W] like the reference.  `out` (optional, an extension used by
        StereoEngine): a preallocated [4,B,1,H,W] fp32 tensor the four stages are written into (the returned list are its views)."""
        if not (left_input.is_cuda and right_input.is_cuda):
            raise LwsError("LWSNet.forward: lwsnet_b200 has no CPU path; inputs must be CUDA tensors")
        if left_input.shape != right_input.shape or left_input.dim() != 4 or left_input.shape[1] != 3:
            raise ValueError("left_input / right_input must be equal-shape [B,3,H,W]")
        img_h, img_w = left_input.shape[2], left_input.shape[3]
        if img_h % 8 or img_w % 8:
            raise ValueError("H and W must be multiples of 8 (hourglass skip adds, reference models/submodules.py:103)")
        left_input = left_input.contiguous()
        # the reference runs the shared-weight extractor twice (models.py:110-111); one launch set over the stacked pair is
        # the same arithmetic per image (every kernel is batch-independent) and keeps the small 1/8-resolution layers fuller
        n = left_input.shape[0]
        right_input = right_input.contiguous()
        if (left_input.untyped_storage().data_ptr() == right_input.untyped_storage().data_ptr()
                and right_input.storage_offset() == left_input.storage_offset() + left_input.numel()):
            # the pair already sits back to back in one allocation (StereoEngine's buffers): view it as the stacked batch
            stacked = left_input.new_empty(0).set_(left_input.untyped_storage(), left_input.storage_offset(),
                                                   (2 * n,) + tuple(left_input.shape[1:]))
        else:
            stacked = torch.cat([left_input, right_input])
        feats = self.feature_extraction(stacked)
        feats_l = [f[:n] for f in feats]
        feats_r = [f[n:] for f in feats]
        if out is not None and (tuple(out.shape) != (4, n, 1, img_h, img_w) or out.dtype != torch.float32 or not out.is_contiguous()
                                or out.device != left_input.device):
            raise ValueError("out must be a contiguous fp32 [4,B,1,H,W] tensor on the inputs' device")
        pred, wflow = [], None
        for scale in range(len(feats_l)):
            nxt = tuple(feats_l[scale + 1].shape[2:]) if scale + 1 < len(feats_l) else None
            res = self._stage(scale, feats_l[scale].contiguous(), feats_r[scale].contiguous(),
                              pred[scale - 1] if scale > 0 else None, img_h, img_w, out=None if out is None else out[scale],
                              wflow=wflow, next_hw=nxt)
            p, wflow = res if nxt is not None else (res, None)
            pred.append(p)
        pred.append(self._refine(left_input, pred[2], out=None if out is None else out[3]))
        return pred


class disparity_regression(nn.Module):
    """reference models/models.py:167-179: expectation of arange(start*stride, end*stride, stride) under `input`.

    `input` is the already soft-maxed volume (as in the reference).  Inside LWSNet the softmax and this expectation run
    as one fused kernel (ops.softmax_regression); this stand-alone class evaluates exactly what the reference's class does,
    sum_j input[:, j] * disp_j, with no renormalisation of `input` (ops.disparity_regression).
    """

    def __init__(self, start, end, stride=1):
        super().__init__()
        self.start, self.end, self.stride = start, end, stride
        self.my_steplength = len(range(start * stride, end * stride, stride))

    def forward(self, input):
        if not input.is_cuda:
            raise LwsError("disparity_regression: no CPU path")
        if input.shape[1] != self.my_steplength:
            raise ValueError(f"expected {self.my_steplength} disparity planes, got {input.shape[1]}")
        return ops.disparity_regression(input, float(self.start * self.stride), float(self.stride))
