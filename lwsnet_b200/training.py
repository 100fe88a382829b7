"""Training path of the hot-path kernels (reference train.py:127-166).

``torch.autograd.Function`` wrappers around the forward kernels and their hand-written backward kernels (csrc/training.cu):

* ``cost_volume_l1(feat_l, feat_r, maxdisp, stride)``                      -- LWSNet._build_volume_2d   (models/models.py:58-76)
* ``warp_residual_volume_l1(feat_l, feat_r, disp, maxdisp, stride)``       -- LWSNet._build_volume_2d3  (models/models.py:78-104)
* ``softmax_regression(cost, start, step)``                                -- softmax + disparity_regression (models.py:142-179)
* ``multistage_smooth_l1_loss(preds, gt, maxdisp, weights)``               -- the loss of train.py:145-155

Gradients flow into the feature maps, the disparity the warp is driven by, the cost volume and the stage predictions.  The
convolution stacks between them have no backward in this tier (SURVEY.md 8(f) ranks the full training loop last), so this module
is the differentiable skeleton of the stage loop, not a trainer.  There is no CPU path.
"""
from __future__ import annotations

import ctypes
from typing import Sequence

import torch

from . import ops
from ._lib import check, lib


def _c(t):
    return t if t.is_contiguous() else t.contiguous()


class _CostVolumeL1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat_l, feat_r, maxdisp, stride):
        ctx.save_for_backward(feat_l, feat_r)
        ctx.maxdisp, ctx.stride = maxdisp, stride
        return ops.cost_volume_l1(feat_l, feat_r, maxdisp, stride)

    @staticmethod
    def backward(ctx, gcost):
        L, R = ctx.saved_tensors
        L, R, gcost = ops._f32c(L), ops._f32c(R), ops._f32c(gcost)
        B, C, H, W = L.shape
        gL, gR = torch.empty_like(L), torch.empty_like(R)
        with torch.cuda.device(L.device):
            check(lib.lws_cost_volume_l1_bwd_f32(ops._ptr(L, "L"), ops._ptr(R, "R"), ops._ptr(gcost, "gcost"), ops._ptr(gL, "gL"),
                                                 ops._ptr(gR, "gR"), B, C, H, W, ctx.maxdisp, ctx.stride, ops._stream(L)),
                  "lws_cost_volume_l1_bwd_f32")
        ops.LAUNCHES[0] += 1
        return gL, gR, None, None


class _WarpResidualVolumeL1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat_l, feat_r, disp, maxdisp, stride):
        ctx.save_for_backward(feat_l, feat_r, disp)
        ctx.maxdisp, ctx.stride = maxdisp, stride
        return ops.warp_residual_volume_l1(feat_l, feat_r, disp, maxdisp, stride)

    @staticmethod
    def backward(ctx, gcost):
        L, R, disp = ctx.saved_tensors
        L, R, disp, gcost = ops._f32c(L), ops._f32c(R), ops._f32c(disp), ops._f32c(gcost)
        B, C, H, W = L.shape
        gL, gR, gd = torch.empty_like(L), torch.empty_like(R), torch.empty_like(disp)
        with torch.cuda.device(L.device):
            check(lib.lws_warp_residual_volume_l1_bwd_f32(ops._ptr(L, "L"), ops._ptr(R, "R"), ops._ptr(disp, "disp"),
                                                          ops._ptr(gcost, "gcost"), ops._ptr(gL, "gL"), ops._ptr(gR, "gR"),
                                                          ops._ptr(gd, "gdisp"), B, C, H, W, ctx.maxdisp, ctx.stride, ops._stream(L)),
                  "lws_warp_residual_volume_l1_bwd_f32")
        ops.LAUNCHES[0] += 1
        return gL, gR, gd, None, None


class _SoftmaxRegression(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cost, start, step):
        ctx.save_for_backward(cost)
        ctx.start, ctx.step = start, step
        return ops.softmax_regression(cost, start, step)

    @staticmethod
    def backward(ctx, glow):
        (cost,) = ctx.saved_tensors
        cost, glow = ops._f32c(cost), ops._f32c(glow)
        B, D, H, W = cost.shape
        gcost = torch.empty_like(cost)
        with torch.cuda.device(cost.device):
            check(lib.lws_softmax_regression_bwd_f32(ops._ptr(cost, "cost"), ops._ptr(glow, "glow"), ops._ptr(gcost, "gcost"), B, D, H, W,
                                                     float(ctx.start), float(ctx.step), ops._stream(cost)),
                  "lws_softmax_regression_bwd_f32")
        ops.LAUNCHES[0] += 1
        return gcost, None, None


def cost_volume_l1(feat_l, feat_r, maxdisp, stride=1):
    assert maxdisp % stride == 0  # the reference's own assert (models/models.py:63)
    return _CostVolumeL1.apply(feat_l, feat_r, maxdisp, stride)


def warp_residual_volume_l1(feat_l, feat_r, disp, maxdisp, stride=1):
    return _WarpResidualVolumeL1.apply(feat_l, feat_r, disp, maxdisp, stride)


def softmax_regression(cost, start, step=1.0):
    return _SoftmaxRegression.apply(cost, float(start), float(step))


class _MultistageLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gt, maxdisp, weights, *preds):
        S = len(preds)
        gt = ops._f32c(gt)
        ps = [ops._f32c(p) for p in preds]
        n = gt.numel()
        for p in ps:
            if p.numel() != n:
                raise ValueError("every stage prediction must have as many elements as gt")
        dev = gt.device
        out = torch.empty(5, dtype=torch.float32, device=dev)
        grads = [torch.empty_like(p) for p in ps]
        nbytes = int(lib.lws_smooth_l1_loss_workspace_bytes(n))
        ws = ops.workspace(dev, "loss", nbytes)
        arr = lambda ts: (ctypes.c_void_p * S)(*[t.data_ptr() for t in ts])
        w = (ctypes.c_float * S)(*[float(x) for x in weights])
        with torch.cuda.device(dev):
            check(lib.lws_smooth_l1_multistage_loss_f32(arr(ps), ops._ptr(gt, "gt"), w, S, n, float(maxdisp), ops._ptr(out, "out"),
                                                        arr(grads), ctypes.c_void_p(ws.data_ptr()), nbytes, ops._stream(gt)),
                  "lws_smooth_l1_multistage_loss_f32")
        ops.LAUNCHES[0] += 3
        ctx.save_for_backward(*grads)
        ctx.shapes = [p.shape for p in preds]
        return out[:S].clone(), out[4].clone()

    @staticmethod
    def backward(ctx, gloss, _gcount):
        grads = ctx.saved_tensors
        return (None, None, None) + tuple((g * gloss[s]).reshape(ctx.shapes[s]) for s, g in enumerate(grads))


def multistage_smooth_l1_loss(preds: Sequence[torch.Tensor], gt: torch.Tensor, maxdisp: float, weights: Sequence[float]):
    """train.py:137-155: ``mask = gt < maxdisp``; stage loss s = ``weights[s] * smooth_l1(preds[s][mask], gt[mask], mean)``.
    Returns (stage_losses [S] -- sum them for the reference's ``paddle.add_n(stage_loss)`` --, number of masked pixels).  With an
    empty mask every loss is 0 (the reference skips such batches, train.py:139-140)."""
    if len(preds) != len(weights) or not 1 <= len(preds) <= 4:
        raise ValueError("1..4 stage predictions with one weight each")
    return _MultistageLoss.apply(gt, float(maxdisp), tuple(weights), *preds)
