"""PyTorch-CPU restatement of the reference model (TEST INFRASTRUCTURE — see oracle/__init__.py).

PARITY UNPINNED (no Paddle in this image, no reference tests/golden vectors).

Follows /root/reference/models/models.py:8-179 and
/root/reference/models/submodules.py:5-327 line by line, with the Paddle 2.0.0rc0
operator semantics of SURVEY.md Appendix C:

* ``F.grid_sample`` defaults bilinear / zeros / align_corners=True (models.py:53)
* ``F.interpolate(mode="bilinear")`` defaults align_corners=False, align_mode=0
  (half-pixel; models.py:119,146,154,161)
* tensor (*|/) python-scalar is a separately rounded ``scale`` op, division is a
  reciprocal multiply (models.py:47-48,121,145,153)
* BatchNorm eval, eps=1e-5, state keys weight/bias/_mean/_variance
* KaimingNormal fan-in init, std = sqrt(2 / (shape[1]*prod(shape[2:])))

State-dict keys equal the Paddle key grammar so weights
move between this oracle and the CUDA product with ``load_state_dict``.
The model works in fp32 or fp64 (``.double()``): fp64 is the noise-floor reference.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


def default_args(maxdisplist=(24, 5, 5), layers_3d=4, channels_3d=8, growth_rate=(4, 1, 1)):
    """The four attributes LWSNet reads from ``args`` (inference.py:23-26, models.py:11-14)."""
    return SimpleNamespace(maxdisplist=list(maxdisplist), layers_3d=layers_3d,
                           channels_3d=channels_3d, growth_rate=list(growth_rate))


class BatchNorm(nn.Module):
    """Eval-mode BatchNorm{2D,3D} with Paddle's state keys (Appendix C.6)."""

    def __init__(self, num_features, eps=1e-5):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(num_features))
        self.bias = nn.Parameter(torch.zeros(num_features))
        self.register_buffer("_mean", torch.zeros(num_features))
        self.register_buffer("_variance", torch.ones(num_features))

    def forward(self, x):
        shape = [1, -1] + [1] * (x.dim() - 2)
        inv = 1.0 / torch.sqrt(self._variance + self.eps)
        return (x - self._mean.view(shape)) * inv.view(shape) * self.weight.view(shape) + self.bias.view(shape)


# ---------------------------------------------------------------- submodules.py:5-33
def convbn(in_channels, out_channels, kernel_size, stride, padding, dilation=1):
    return nn.Sequential(
        nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride,
                  padding=dilation if dilation > 1 else padding, dilation=dilation, bias=False),
        BatchNorm(out_channels))


def deconvbn(in_channels, out_channels, kernel_size, stride, padding, output_padding=1):
    return nn.Sequential(
        nn.ConvTranspose2d(in_channels, out_channels, kernel_size, stride=stride, padding=padding,
                           output_padding=output_padding, bias=False),
        BatchNorm(out_channels))


# ---------------------------------------------------------------- submodules.py:35-109
class hourglass(nn.Module):
    def __init__(self, init_channel=8):
        super().__init__()
        c = init_channel
        self.conv1 = nn.Sequential(convbn(c, c * 2, 3, 2, 1), nn.ReLU())
        self.conv2 = nn.Sequential(convbn(c * 2, c * 2, 3, 1, 1), nn.ReLU())
        self.conv3 = nn.Sequential(convbn(c * 2, c * 2, 3, 2, 1), nn.ReLU())
        self.conv4 = nn.Sequential(convbn(c * 2, c * 2, 3, 1, 1), nn.ReLU())
        self.conv5 = deconvbn(c * 2, c * 2, 3, 2, 1, 1)
        self.conv6 = deconvbn(c * 2, c, 3, 2, 1, 1)

    def forward(self, x):
        res = []
        out = self.conv1(x)
        pre = self.conv2(out)
        out = self.conv3(pre)
        out = self.conv4(out)
        res.append(out)
        post = F.relu(self.conv5(out) + pre)
        res.append(post)
        out = self.conv6(post)
        res.append(out)
        return res


# ---------------------------------------------------------------- submodules.py:113-188
class feature_extraction(nn.Module):
    def __init__(self):
        super().__init__()
        self.dres0 = nn.Sequential(convbn(3, 4, 3, 2, 1, 2), nn.ReLU(), convbn(4, 8, 3, 1, 1, 4), nn.ReLU())
        self.dres1 = nn.Sequential(convbn(8, 4, 3, 1, 1, 2), nn.ReLU(), convbn(4, 8, 3, 1, 1, 2))
        self.dres2 = hourglass(8)
        self.classif1 = nn.Sequential(convbn(8, 8, 3, 1, 1, 1), nn.ReLU(),
                                      nn.Conv2d(8, 8, 3, padding=1, stride=1, bias=False))

    def forward(self, x):
        out = self.dres0(x)
        out = self.dres1(out) + out
        res = self.dres2(out)
        out = res[-1] + out
        out = self.classif1(out)
        res.pop(-1)
        res.append(out)
        return res


# ---------------------------------------------------------------- submodules.py:190-221
def batch_relu_conv3d(in_channels, out_channels, kernel_size=3, stride=1, padding=1):
    return nn.Sequential(BatchNorm(in_channels), nn.ReLU(),
                         nn.Conv3d(in_channels, out_channels, kernel_size, padding=padding, stride=stride,
                                   bias=False))


def post_3dconvs(layers, channels):
    net = [batch_relu_conv3d(1, channels)]
    net += [batch_relu_conv3d(channels, channels) for _ in range(layers)]
    net += [batch_relu_conv3d(channels, 1)]
    return nn.Sequential(*net)


# ---------------------------------------------------------------- submodules.py:223-327
def preconv2d(in_channels, out_channels, kernel_size, stride, pad, dilation=1):
    return nn.Sequential(BatchNorm(in_channels), nn.ReLU(),
                         nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride,
                                   padding=dilation if dilation > 1 else pad, dilation=dilation, bias=False))


def preconv2d_depthseperated(in_channels, out_channels, kernel_size, stride, pad, dilation=1):
    return nn.Sequential(BatchNorm(in_channels), nn.ReLU(),
                         nn.Conv2d(in_channels, in_channels, kernel_size, stride=stride,
                                   padding=dilation if dilation > 1 else pad, dilation=dilation, bias=False,
                                   groups=in_channels),
                         nn.Conv2d(in_channels, out_channels, 1, stride=1, padding=0, bias=False))


def refinement1(in_channels, out_channels):
    net = [nn.Conv2d(in_channels, out_channels, 3, stride=1, padding=1, bias=False)]
    net += [preconv2d_depthseperated(out_channels, out_channels, 3, 1, 1, dilation=2 ** (k + 1)) for k in range(4)]
    return nn.Sequential(*net)


def refinement2(in_channels, out_channels):
    net = [preconv2d(in_channels, out_channels, 3, 1, 1, dilation=8)]
    net += [preconv2d_depthseperated(out_channels, out_channels, 3, 1, 1, dilation=2 ** k) for k in reversed(range(4))]
    net += [nn.Conv2d(out_channels, 1, 3, stride=1, padding=1, bias=False)]
    return nn.Sequential(*net)


# ---------------------------------------------------------------- Paddle op semantics
def _recip(c, dtype):
    """x / c in Paddle 2.0 dygraph == x * fl(1/c) (Appendix C.2)."""
    if dtype == torch.float32:
        return float(np.float32(1.0 / c))
    return 1.0 / c


def interpolate_bilinear(x, size):
    """paddle F.interpolate(mode='bilinear'): align_corners=False, align_mode=0 (Appendix C.4)."""
    return F.interpolate(x, size=list(size), mode="bilinear", align_corners=False)


def grid_sample_zeros_ac(x, ix, iy):
    """Bilinear gather at un-normalised coords (ix, iy) [N,H,W], zero padding (Paddle CPU grid_sampler)."""
    N, C, H, W = x.shape
    x0 = torch.floor(ix)
    y0 = torch.floor(iy)
    x1 = x0 + 1
    y1 = y0 + 1
    d_w = ix - x0
    d_e = x1 - ix
    d_n = iy - y0
    d_s = y1 - iy
    flat = x.reshape(N, C, H * W)

    def tap(yy, xx):
        inb = (xx >= 0) & (xx <= W - 1) & (yy >= 0) & (yy <= H - 1)
        xi = xx.clamp(0, W - 1).long()
        yi = yy.clamp(0, H - 1).long()
        idx = (yi * W + xi).reshape(N, 1, -1).expand(N, C, -1)
        v = torch.gather(flat, 2, idx).reshape(N, C, *ix.shape[1:])
        return v * inb.unsqueeze(1).to(x.dtype)

    return (tap(y0, x0) * (d_e * d_s).unsqueeze(1) + tap(y0, x1) * (d_w * d_s).unsqueeze(1)
            + tap(y1, x0) * (d_e * d_n).unsqueeze(1) + tap(y1, x1) * (d_w * d_n).unsqueeze(1))


def warp_coords(disp, H, W):
    """models.py:36-48 + grid_sampler un-normalise (A.3): returns (ix, iy) [N,H,W], every op rounded on its own."""
    dt = disp.dtype
    N = disp.shape[0]
    xx = torch.arange(0, W, dtype=dt).view(1, 1, 1, W).expand(N, 1, H, W)
    yy = torch.arange(0, H, dtype=dt).view(1, 1, H, 1).expand(N, 1, H, W)
    vx = xx - disp
    vy = yy
    gx = (2.0 * vx) * _recip(max(W - 1, 1), dt) - 1.0
    gy = (2.0 * vy) * _recip(max(H - 1, 1), dt) - 1.0
    ix = (gx + 1.0) * ((W - 1) * 0.5)
    iy = (gy + 1.0) * ((H - 1) * 0.5)
    return ix[:, 0], iy[:, 0]


def warp(x, disp):
    """LWSNet.warp (models.py:28-55)."""
    N, C, H, W = x.shape
    ix, iy = warp_coords(disp, H, W)
    return grid_sample_zeros_ac(x, ix, iy)


def build_volume_2d(feat_l, feat_r, maxdisp, stride=1):
    """LWSNet._build_volume_2d (models.py:58-76), literal loop."""
    assert maxdisp % stride == 0
    B, C, H, W = feat_l.shape
    cost = torch.zeros((B, maxdisp // stride, H, W), dtype=feat_l.dtype)
    for i in range(0, maxdisp, stride):
        if i > 0:
            cost[:, i // stride, :, :i] = feat_l[:, :, :, :i].abs().sum(dim=1)
            cost[:, i // stride, :, i:] = (feat_l[:, :, :, i:] - feat_r[:, :, :, :-i]).abs().sum(dim=1)
        else:
            cost[:, i // stride, :, i:] = (feat_l - feat_r).abs().sum(dim=1)
    return cost


def build_volume_2d3(feat_l, feat_r, maxdisp, disp, stride=1):
    """LWSNet._build_volume_2d3 (models.py:78-104), literal 9x replicated batch."""
    B, C, H, W = feat_l.shape
    K = maxdisp * 2 - 1
    batch_disp = disp.unsqueeze(1).expand(B, K, 1, H, W).reshape(-1, 1, H, W)
    batch_shift = torch.arange(-maxdisp + 1, maxdisp, dtype=disp.dtype)
    batch_shift = batch_shift.expand(B, K).reshape(-1).view(-1, 1, 1, 1) * stride
    batch_disp = batch_disp - batch_shift
    batch_feat_l = feat_l.unsqueeze(1).expand(B, K, C, H, W).reshape(-1, C, H, W)
    batch_feat_r = feat_r.unsqueeze(1).expand(B, K, C, H, W).reshape(-1, C, H, W)
    cost = (batch_feat_l - warp(batch_feat_r, batch_disp)).abs().sum(dim=1)
    return cost.reshape(B, -1, H, W)


class disparity_regression(nn.Module):
    """models.py:167-179."""

    def __init__(self, start, end, stride=1):
        super().__init__()
        self.disp = torch.arange(start * stride, end * stride, stride, dtype=torch.float32).view(1, -1, 1, 1)

    def forward(self, input):
        disp = self.disp.to(input.dtype).expand(input.shape[0], self.disp.shape[1], input.shape[2], input.shape[3])
        return torch.sum(input * disp, dim=1, keepdim=True)


class LWSNet(nn.Module):
    """models.py:8-164.  forward returns the 4 predictions; ``forward_trace`` also returns every intermediate."""

    def __init__(self, args=None):
        super().__init__()
        args = args or default_args()
        self.maxdisplist = args.maxdisplist
        self.layers_3d = args.layers_3d
        self.channels_3d = args.channels_3d
        self.growth_rate = args.growth_rate
        self.feature_extraction = feature_extraction()
        self.volume_postprocess = nn.ModuleList(
            [post_3dconvs(self.layers_3d, self.channels_3d * self.growth_rate[i]) for i in range(3)])
        self.refinement1_left = refinement1(3, 32)
        self.refinement1_disp = refinement1(1, 32)
        self.refinement2 = refinement2(64, 32)
        self.eval()

    warp = staticmethod(warp)
    _build_volume_2d = staticmethod(build_volume_2d)
    _build_volume_2d3 = staticmethod(build_volume_2d3)

    def forward(self, left_input, right_input):
        return self.forward_trace(left_input, right_input)[0]

    @torch.no_grad()
    def stage(self, scale, feat_l, feat_r, prev_pred, img_hw, trace=None):
        """One iteration of the stage loop (models.py:115-156).  prev_pred is pred[scale-1] (None at scale 0)."""
        dt = feat_l.dtype
        H_img, W_img = img_hw
        h, w = feat_l.shape[2], feat_l.shape[3]
        t = {} if trace is None else trace
        if scale > 0:
            wflow = interpolate_bilinear(prev_pred, (h, w)) * float(h) * _recip(H_img, dt)
            t[f"wflow{scale}"] = wflow
            cost = build_volume_2d3(feat_l, feat_r, self.maxdisplist[scale], wflow, stride=1)
        else:
            cost = build_volume_2d(feat_l, feat_r, self.maxdisplist[scale], stride=1)
        t[f"cost_raw{scale}"] = cost
        cost = cost.unsqueeze(1)
        cost = self.volume_postprocess[scale](cost) + cost
        cost = cost.squeeze(1)
        t[f"cost_post{scale}"] = cost
        if scale == 0:
            reg = disparity_regression(0, self.maxdisplist[0])
        else:
            reg = disparity_regression(-self.maxdisplist[scale] + 1, self.maxdisplist[scale])
        low = reg(F.softmax(-cost, dim=1))
        t[f"low{scale}"] = low
        low = low * float(H_img) * _recip(low.shape[2], dt)
        up = interpolate_bilinear(low, (H_img, W_img))
        return up if scale == 0 else up + prev_pred

    @torch.no_grad()
    def refine(self, left_input, pred3, trace=None):
        """models.py:158-162."""
        t = {} if trace is None else trace
        refined_left = self.refinement1_left(left_input)
        refined_disp = self.refinement1_disp(pred3)
        t["refined_left"] = refined_left
        t["refined_disp"] = refined_disp
        disp = self.refinement2(torch.cat([refined_left, refined_disp], 1))
        t["refine_res"] = disp
        disp_up = interpolate_bilinear(disp, left_input.shape[2:])
        return pred3 + disp_up

    @torch.no_grad()
    def forward_trace(self, left_input, right_input):
        img_hw = (left_input.shape[2], left_input.shape[3])
        trace = {}
        feats_l = self.feature_extraction(left_input)
        feats_r = self.feature_extraction(right_input)
        for s in range(3):
            trace[f"feat_l{s}"] = feats_l[s]
            trace[f"feat_r{s}"] = feats_r[s]
        pred = []
        for scale in range(len(feats_l)):
            pred.append(self.stage(scale, feats_l[scale], feats_r[scale],
                                   pred[scale - 1] if scale > 0 else None, img_hw, trace))
        pred.append(self.refine(left_input, pred[2], trace))
        for i, p in enumerate(pred):
            trace[f"pred{i}"] = p
        return pred, trace


# ---------------------------------------------------------------- init / inputs
def kaiming_normal_init_(model, seed=0):
    """nn.initializer.KaimingNormal on every conv (Appendix C.7); BN gamma=1 beta=0 mean=0 var=1."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in sorted(model.named_parameters()):
            if p.dim() >= 4:
                fan_in = p.shape[1] * int(np.prod(p.shape[2:]))
                p.copy_(torch.randn(p.shape, generator=g, dtype=torch.float32) * math.sqrt(2.0 / fan_in))
        for m in model.modules():
            if isinstance(m, BatchNorm):
                m.weight.fill_(1.0)
                m.bias.zero_()
                m._mean.zero_()
                m._variance.fill_(1.0)
    return model


def randomize_bn_(model, seed=1):
    """Non-trivial BN statistics (as a trained checkpoint has) so BN folding is actually exercised."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, m in sorted(model.named_modules()):
            if isinstance(m, BatchNorm):
                n = m.weight.numel()
                m.weight.copy_(0.75 + 0.5 * torch.rand(n, generator=g))
                m.bias.copy_(0.2 * torch.randn(n, generator=g))
                m._mean.copy_(0.2 * torch.randn(n, generator=g))
                m._variance.copy_(0.6 + 0.8 * torch.rand(n, generator=g))
    return model


IMAGENET_MEAN = (0.485, 0.456, 0.406)  # dataloader/dataloader.py:10-11
IMAGENET_STD = (0.229, 0.224, 0.225)


def preprocess_bgr_uint8(img_bgr, th=368, tw=1232):
    """inference.py:93-103: bottom-right crop, BGR->RGB, ToTensor (/255), Normalize."""
    h, w, _ = img_bgr.shape
    crop = img_bgr[h - th:h, w - tw:w, ::-1].astype(np.float32) / 255.0
    crop = (crop - np.asarray(IMAGENET_MEAN, np.float32)) / np.asarray(IMAGENET_STD, np.float32)
    return torch.from_numpy(np.ascontiguousarray(crop.transpose(2, 0, 1))).unsqueeze(0)


def synthetic_pair(B, H, W, seed=1234, max_disp=150.0):
    """Synthetic stereo pairs (SURVEY.md 8(d) configs[2]): smooth texture, right = left warped by a smooth disparity."""
    outs_l, outs_r = [], []
    for b in range(B):
        g = torch.Generator().manual_seed(seed + b)
        tex = torch.randn(1, 3, H // 4 + 2, W // 4 + 2, generator=g)
        left = F.interpolate(tex, size=(H, W), mode="bicubic", align_corners=False)
        left = left + 0.25 * torch.randn(1, 3, H, W, generator=g)
        d = torch.rand(1, 1, 4, 8, generator=g) * max_disp
        d = F.interpolate(d, size=(H, W), mode="bicubic", align_corners=False).clamp_(0, max_disp)
        xs = torch.arange(W, dtype=torch.float32).view(1, 1, 1, W) + d  # right(x) = left(x + d)
        gx = 2.0 * xs / (W - 1) - 1.0
        gy = (2.0 * torch.arange(H, dtype=torch.float32) / (H - 1) - 1.0).view(1, 1, H, 1).expand(1, 1, H, W)
        grid = torch.cat([gx, gy], 1).permute(0, 2, 3, 1)
        right = F.grid_sample(left, grid, mode="bilinear", padding_mode="border", align_corners=True)
        right = right + 0.05 * torch.randn(1, 3, H, W, generator=g)
        outs_l.append(left)
        outs_r.append(right)
    return torch.cat(outs_l), torch.cat(outs_r)


def build_oracle(seed=0, args=None, dtype=torch.float32, random_bn=False):
    m = kaiming_normal_init_(LWSNet(args), seed)
    if random_bn:
        randomize_bn_(m, seed + 1)
    return m.to(dtype).eval()
