"""Mirror of the reference's models/submodules.py factories (same names, same arguments, same state-dict keys).

The reference builds ``paddle.nn.Sequential`` stacks of cuDNN-backed layers.  Here the same factories build
parameter containers whose ``forward`` enqueues the hand-written sm_100a kernels of liblws_b200 (via ``ops``):

* ``post_3dconvs``                      -> lws_conv3d_stack_f32         (reference models/submodules.py:190-221)
* ``refinement1`` / ``refinement2``      -> lws_refinement_f32 (fused, driven from models.LWSNet.forward)
                                           (reference models/submodules.py:223-327)
* ``feature_extraction`` / ``hourglass`` -> lws_feature_extraction_f32   (reference models/submodules.py:35-188;
  SURVEY.md section 8(f) row n1, the step in front of the hot path)

State-dict keys follow the Paddle key grammar: Sequential children by index, BatchNorm
tensors ``weight, bias, _mean, _variance``.  There is no CPU path: CPU tensors raise.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from ._lib import LwsError

BN_EPS = 1e-5  # paddle.nn.BatchNorm2D/3D default epsilon


def _pack_key(module, device):
    """Cache key of a folded-weight blob: storage address and in-place version of every tensor.  In-place edits through
    ``p.data`` do not bump the version: call ``repack()`` (or ``model.repack()``) after such weight surgery."""
    return (str(device),) + tuple((p.data_ptr(), p._version) for p in list(module.parameters()) + list(module.buffers()))


def _require_cuda(x: torch.Tensor, who: str) -> None:
    if not x.is_cuda:
        raise LwsError(f"{who}: lwsnet_b200 has no CPU path; move the model and its inputs to a CUDA device")


# The per-layer classes below are parameter holders: their parents (feature_extraction, Post3DConvs, the refinement parts) execute
# them fused through the C ABI.  A per-layer torch / cuDNN forward exists only as a GPU-side cross-check for the tests and is
# refused outside `torch_crosscheck()`, so the product can never run a library kernel by accident.
_CROSSCHECK = [False]


class torch_crosscheck:
    def __enter__(self):
        self._old = _CROSSCHECK[0]
        _CROSSCHECK[0] = True

    def __exit__(self, *exc):
        _CROSSCHECK[0] = self._old


def _require_crosscheck(who: str) -> None:
    if not _CROSSCHECK[0]:
        raise LwsError(f"{who} is a parameter holder executed fused by its parent layer through liblws_b200; "
                       "a stand-alone per-layer forward (torch / cuDNN) is not part of the product path")


class BatchNorm(nn.Module):
    """Inference-mode BatchNorm2D/3D parameter holder with Paddle's state keys."""

    def __init__(self, num_features: int, eps: float = BN_EPS):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(num_features))
        self.bias = nn.Parameter(torch.zeros(num_features))
        self.register_buffer("_mean", torch.zeros(num_features))
        self.register_buffer("_variance", torch.ones(num_features))

    def tensors(self):
        return (self.weight, self.bias, self._mean, self._variance)

    def scale_shift(self):
        s = self.weight.double() / torch.sqrt(self._variance.double() + self.eps)
        t = self.bias.double() - self._mean.double() * s
        return s, t

    def forward(self, x):
        _require_cuda(x, "BatchNorm")
        _require_crosscheck("BatchNorm")
        s, t = self.scale_shift()
        shape = [1, -1] + [1] * (x.dim() - 2)
        return x * s.float().view(shape) + t.float().view(shape)


class _Conv(nn.Module):
    """Bias-free conv parameter holder (KaimingNormal fan-in init like the reference's weight_attr)."""

    def __init__(self, shape, stride=1, padding=0, dilation=1, groups=1, transposed=False, output_padding=0):
        super().__init__()
        self.stride, self.padding, self.dilation, self.groups = stride, padding, dilation, groups
        self.transposed, self.output_padding = transposed, output_padding
        w = torch.empty(shape)
        fan_in = shape[1] * int(math.prod(shape[2:]))
        nn.init.normal_(w, 0.0, math.sqrt(2.0 / fan_in))
        self.weight = nn.Parameter(w)


class Conv2D(_Conv):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1):
        super().__init__((out_channels, in_channels // groups, kernel_size, kernel_size), stride, padding, dilation, groups)

    def forward(self, x):  # feature extractor only (off the hot path, see module docstring)
        _require_cuda(x, "Conv2D")
        _require_crosscheck("Conv2D")
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):  # fp32 operands: TF32 moves disparities by px
            return F.conv2d(x, self.weight, None, self.stride, self.padding, self.dilation, self.groups)


class Conv2DTranspose(_Conv):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, output_padding=0):
        super().__init__((in_channels, out_channels, kernel_size, kernel_size), stride, padding, 1, 1, True, output_padding)

    def forward(self, x):
        _require_cuda(x, "Conv2DTranspose")
        _require_crosscheck("Conv2DTranspose")
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            return F.conv_transpose2d(x, self.weight, None, self.stride, self.padding, self.output_padding)


class Conv3D(_Conv):
    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=1):
        super().__init__((out_channels, in_channels, kernel_size, kernel_size, kernel_size), stride, padding)


class ReLU(nn.Module):
    def forward(self, x):
        _require_cuda(x, "ReLU")
        _require_crosscheck("ReLU")
        return torch.relu(x)


# ---------------------------------------------------------------------------- feature extractor (off hot path)
def convbn(in_channels, out_channels, kernel_size, stride, padding, dilation=1,
           conv_param_attr=None, conv_bias_attr=None, bn_param_attr=None, bn_bias_attr=None):
    """reference models/submodules.py:5-18."""
    return nn.Sequential(Conv2D(in_channels, out_channels, kernel_size, stride,
                                dilation if dilation > 1 else padding, dilation),
                         BatchNorm(out_channels))


def deconvbn(in_channels, out_channels, kernel_size, stride, padding, output_padding=1, dilation=1,
             conv_param_attr=None, conv_bias_attr=None, bn_param_attr=None, bn_bias_attr=None):
    """reference models/submodules.py:20-33."""
    return nn.Sequential(Conv2DTranspose(in_channels, out_channels, kernel_size, stride, padding, output_padding),
                         BatchNorm(out_channels))


class hourglass(nn.Module):
    """reference models/submodules.py:35-109."""

    def __init__(self, init_channel=8):
        super().__init__()
        c = self.init_channel = init_channel
        self.conv1 = nn.Sequential(convbn(c, c * 2, 3, 2, 1), ReLU())
        self.conv2 = nn.Sequential(convbn(c * 2, c * 2, 3, 1, 1), ReLU())
        self.conv3 = nn.Sequential(convbn(c * 2, c * 2, 3, 2, 1), ReLU())
        self.conv4 = nn.Sequential(convbn(c * 2, c * 2, 3, 1, 1), ReLU())
        self.conv5 = deconvbn(c * 2, c * 2, 3, 2, 1, 1)
        self.conv6 = deconvbn(c * 2, c, 3, 2, 1, 1)

    def forward(self, input):
        res = []
        output = self.conv1(input)
        pre = self.conv2(output)
        output = self.conv3(pre)
        output = self.conv4(output)
        res.append(output)
        post = torch.relu(self.conv5(output) + pre)
        res.append(post)
        output = self.conv6(post)
        res.append(output)
        return res


class feature_extraction(nn.Module):
    """reference models/submodules.py:113-188, executed by lws_feature_extraction_f32 (12 fused conv+BN(+ReLU/+skip) launches)."""

    def __init__(self):
        super().__init__()
        self.dres0 = nn.Sequential(convbn(3, 4, 3, 2, 1, 2), ReLU(), convbn(4, 8, 3, 1, 1, 4), ReLU())
        self.dres1 = nn.Sequential(convbn(8, 4, 3, 1, 1, 2), ReLU(), convbn(4, 8, 3, 1, 1, 2))
        self.dres2 = hourglass(8)
        self.classif1 = nn.Sequential(convbn(8, 8, 3, 1, 1, 1), ReLU(), Conv2D(8, 8, 3, 1, 1))
        self._packed = None
        self._packed_key = None

    def tensor_list(self):
        """The 56 tensors of include/lws.h:lws_pack_feature_extraction_weights, in execution order."""
        h = self.dres2
        cbs = [self.dres0[0], self.dres0[2], self.dres1[0], self.dres1[2], h.conv1[0], h.conv2[0], h.conv3[0], h.conv4[0],
               h.conv5, h.conv6, self.classif1[0]]
        out = []
        for cb in cbs:
            out += [cb[0].weight] + list(cb[1].tensors())
        out.append(self.classif1[2].weight)
        return out

    def repack(self):
        self._packed = None

    def packed(self, device):
        key = _pack_key(self, device)
        if self._packed is None or self._packed_key != key:
            self._packed = ops.pack_feature_extraction(self.tensor_list(), BN_EPS).to(device)
            self._packed_key = key
        return self._packed

    def forward(self, input):
        _require_cuda(input, "feature_extraction")
        return ops.feature_extraction(input, self.packed(input.device))

    def forward_torch(self, input):
        """The same graph through torch's conv ops (cuDNN, fp32).  Not used by the product path; kept as a GPU-side
        cross-check for the tests of the feature-pyramid kernel."""
        _require_cuda(input, "feature_extraction")
        with torch_crosscheck():
            output = self.dres0(input)
            output = self.dres1(output) + output
            res = self.dres2(output)
            output = res[-1] + output
            output = self.classif1(output)
        res.pop(-1)
        res.append(output)
        return res


# ---------------------------------------------------------------------------- 3D stack (hot path, a5)
def batch_relu_conv3d(in_channels, out_channels, kernel_size=3, stride=1, padding=1, bn3d=True,
                      conv_param_attr=None, conv_bias_attr=False, bn_param_attr=None, bn_bias_attr=None):
    """reference models/submodules.py:190-214 (parameter holder; executed by Post3DConvs)."""
    if kernel_size != 3 or stride != 1 or padding != 1 or not bn3d:
        raise LwsError("batch_relu_conv3d: only the reference's 3x3x3 / stride 1 / pad 1 / BN configuration is built")
    return nn.Sequential(BatchNorm(in_channels), ReLU(), Conv3D(in_channels, out_channels, kernel_size, stride, padding))


class Post3DConvs(nn.Sequential):
    """post_3dconvs (reference models/submodules.py:216-221) executed by lws_conv3d_stack_f32."""

    def __init__(self, layers, channels):
        net = [batch_relu_conv3d(1, channels)]
        net += [batch_relu_conv3d(channels, channels) for _ in range(layers)]
        net += [batch_relu_conv3d(channels, 1)]
        super().__init__(*net)
        self.layers, self.channels = layers, channels
        self._packed = None
        self._packed_key = None

    def repack(self):
        self._packed = None

    def packed(self, device):
        key = _pack_key(self, device)
        if self._packed is None or self._packed_key != key:
            convs = [blk[2].weight for blk in self]
            bns = [blk[0].tensors() for blk in self]
            self._packed = ops.pack_conv3d_stack(convs, bns, BN_EPS, self.channels, self.layers).to(device)
            self._packed_key = key
        return self._packed

    def run(self, cost4d, add_skip):
        _require_cuda(cost4d, "post_3dconvs")
        return ops.conv3d_stack(cost4d, self.packed(cost4d.device), self.channels, self.layers, add_skip)

    def forward(self, cost):
        """cost [B,1,D,H,W] -> [B,1,D,H,W] (no skip), like calling the reference's Sequential."""
        if cost.dim() != 5 or cost.shape[1] != 1:
            raise ValueError("post_3dconvs expects [B,1,D,H,W]")
        return self.run(cost[:, 0], add_skip=False).unsqueeze(1)


def post_3dconvs(layers, channels):
    return Post3DConvs(layers, channels)


# ---------------------------------------------------------------------------- refinement (hot path, a8 + a9)
def preconv2d(in_channels, out_channels, kernel_size, stride, pad, dilation=1, bn=True):
    """reference models/submodules.py:223-235."""
    if bn:
        return nn.Sequential(BatchNorm(in_channels), ReLU(),
                             Conv2D(in_channels, out_channels, kernel_size, stride,
                                    dilation if dilation > 1 else pad, dilation))
    return None  # the reference returns None here too; never called that way


def preconv2d_depthseperated(in_channels, out_channels, kernel_size, stride, pad, dilation=1, bn=True):
    """reference models/submodules.py:238-280."""
    p = dilation if dilation > 1 else pad
    dw = Conv2D(in_channels, in_channels, kernel_size, stride, p, dilation, groups=in_channels)
    pw = Conv2D(in_channels, out_channels, 1, 1, 0)
    if bn:
        return nn.Sequential(BatchNorm(in_channels), ReLU(), dw, pw)
    return nn.Sequential(ReLU(), dw, pw)


class _RefinementPart(nn.Sequential):
    """refinement1 / refinement2 stacks.  Inside LWSNet.forward they execute fused (lws_refinement_f32, BN folded across the
    module boundaries); called as a layer, like the reference does at models/models.py:158-160, they run their own C-ABI entry
    (lws_refinement1_f32 / lws_refinement2_f32, module-local BN folding, NCHW in / out)."""

    def __init__(self, *layers):
        super().__init__(*layers)
        self._packed = None
        self._packed_key = None

    def repack(self):
        self._packed = None

    def _cached(self, device, make):
        key = _pack_key(self, device)
        if self._packed is None or self._packed_key != key:
            self._packed = make().to(device)
            self._packed_key = key
        return self._packed


class Refinement1(_RefinementPart):
    def tensor_list(self):
        out = [self[0].weight]
        for j in range(1, 5):
            out += list(self[j][0].tensors()) + [self[j][2].weight, self[j][3].weight]
        return out

    def forward(self, input):
        _require_cuda(input, "refinement1")
        cin = self[0].weight.shape[1]
        if input.dim() != 4 or input.shape[1] != cin:
            raise ValueError(f"refinement1 expects [B,{cin},H,W], got {tuple(input.shape)}")
        return ops.refinement1(input, self._cached(input.device, lambda: ops.pack_refinement1(self.tensor_list(), cin, BN_EPS)))


class Refinement2(_RefinementPart):
    def tensor_list(self):
        out = list(self[0][0].tensors()) + [self[0][2].weight]
        for j in range(1, 5):
            out += list(self[j][0].tensors()) + [self[j][2].weight, self[j][3].weight]
        out.append(self[5].weight)
        return out

    def forward(self, input):
        _require_cuda(input, "refinement2")
        return ops.refinement2(input, self._cached(input.device, lambda: ops.pack_refinement2(self.tensor_list(), BN_EPS)))


def refinement1(in_channels, out_channels):
    """reference models/submodules.py:282-300."""
    if in_channels not in (1, 3) or out_channels != 32:
        raise LwsError("refinement1: built for the reference's two uses, refinement1(3, 32) and refinement1(1, 32)")
    net = [Conv2D(in_channels, out_channels, 3, 1, 1)]
    net += [preconv2d_depthseperated(out_channels, out_channels, 3, 1, 1, dilation=2 ** (k + 1)) for k in range(4)]
    return Refinement1(*net)


def refinement2(in_channels, out_channels):
    """reference models/submodules.py:302-327."""
    if in_channels != 64 or out_channels != 32:
        raise LwsError("refinement2: built for the reference's use, refinement2(64, 32)")
    net = [preconv2d(in_channels, out_channels, 3, 1, 1, dilation=8)]
    net += [preconv2d_depthseperated(out_channels, out_channels, 3, 1, 1, dilation=2 ** k) for k in reversed(range(4))]
    net += [Conv2D(out_channels, 1, 3, 1, 1)]
    return Refinement2(*net)


def refinement_tensor_list(r1_left, r1_disp, r2):
    """The 80 tensors of include/lws.h:lws_pack_refinement_weights, in order."""
    out = []
    for r1 in (r1_left, r1_disp):
        out.append(r1[0].weight)
        for j in range(1, 5):
            out += list(r1[j][0].tensors()) + [r1[j][2].weight, r1[j][3].weight]
    out += list(r2[0][0].tensors()) + [r2[0][2].weight]
    for j in range(1, 5):
        out += list(r2[j][0].tensors()) + [r2[j][2].weight, r2[j][3].weight]
    out.append(r2[5].weight)
    assert len(out) == 80
    return out
