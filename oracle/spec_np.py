"""Closed-form NumPy specification of the hot-path functions (TEST INFRASTRUCTURE — see oracle/__init__.py).

PARITY UNPINNED.  Independent of oracle/lwsnet_torch.py (which restates the reference's literal
loop / replicated-batch formulation); tests/test_oracle.py checks the two against each other.
Formulas: SURVEY.md Appendix A.  All fp32 arithmetic is done op by op on np.float32 arrays, so every
operation is rounded on its own exactly like the reference's separate Paddle operator launches
(no FMA contraction).
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


def cost_volume_l1(L, R, maxdisp, stride=1):
    """A.1 / models/models.py:58-76.  L,R [B,C,H,W] -> [B,maxdisp//stride,H,W]."""
    assert maxdisp % stride == 0
    B, C, H, W = L.shape
    out = np.zeros((B, maxdisp // stride, H, W), L.dtype)
    for d in range(0, maxdisp, stride):
        Rs = np.zeros_like(R)
        if d < W:
            Rs[..., d:] = R[..., :W - d]
        out[:, d // stride] = np.abs(L - Rs).sum(axis=1, dtype=L.dtype)
    return out


def resize_src_index(out_size, in_size):
    """Half-pixel source index (Appendix C.4): src=(dst+0.5)*(in/out)-0.5 clamped at 0; i0, i1, lambda1."""
    scale = f32(in_size) / f32(out_size)
    dst = np.arange(out_size, dtype=f32)
    src = scale * (dst + f32(0.5)) - f32(0.5)
    src = np.maximum(src, f32(0))
    i0 = np.minimum(src.astype(np.int32), in_size - 1)
    i1 = np.minimum(i0 + 1, in_size - 1)
    l1 = (src - i0.astype(f32)).astype(f32)
    return i0, i1, l1


def bilinear_resize_halfpixel(x, oh, ow):
    """paddle F.interpolate(mode='bilinear') defaults; x [B,C,h,w] fp32."""
    y0, y1, ly = resize_src_index(oh, x.shape[2])
    x0, x1, lx = resize_src_index(ow, x.shape[3])
    ly1 = ly[None, None, :, None]
    ly0 = (f32(1) - ly)[None, None, :, None]
    lx1 = lx[None, None, None, :]
    lx0 = (f32(1) - lx)[None, None, None, :]
    top = x[:, :, y0][:, :, :, x0] * lx0 + x[:, :, y0][:, :, :, x1] * lx1
    bot = x[:, :, y1][:, :, :, x0] * lx0 + x[:, :, y1][:, :, :, x1] * lx1
    return (ly0 * top + ly1 * bot).astype(x.dtype)


def disp_to_scale(pred_full, h, w):
    """A.2 / models/models.py:119-121: (resize(pred) * float(h)) * fl32(1/H_img)."""
    H_img = pred_full.shape[2]
    r = bilinear_resize_halfpixel(pred_full, h, w)
    return (r * f32(h)) * f32(1.0 / H_img)


def warp_taps(dsp, H, W, div_mode=0):
    """A.3: dsp [N,H,W] fp32 (the *argument* of warp, i.e. disp - shift).

    div_mode 0: `2*v / max(size-1,1)` as the reciprocal multiply of Paddle 2.0's scale op (the oracle's choice);
    div_mode 1: as a true IEEE division (later Paddle versions).

    Returns x0 [N,H,W] int32, y0 [H] int32, (wx0, wx1) [N,H,W] fp32 = (x1-ix, ix-x0), (wy0, wy1) [H] fp32.
    """
    xs = np.arange(W, dtype=f32)[None, None, :]
    ys = np.arange(H, dtype=f32)
    if div_mode:
        gx = ((f32(2.0) * (xs - dsp)) / f32(max(W - 1, 1))).astype(f32) - f32(1.0)
        gy = ((f32(2.0) * ys) / f32(max(H - 1, 1))).astype(f32) - f32(1.0)
    else:
        rW = f32(1.0 / max(W - 1, 1))
        rH = f32(1.0 / max(H - 1, 1))
        gx = (f32(2.0) * (xs - dsp)) * rW - f32(1.0)
        gy = (f32(2.0) * ys) * rH - f32(1.0)
    ix = (gx + f32(1.0)) * f32((W - 1) * 0.5)
    iy = (gy + f32(1.0)) * f32((H - 1) * 0.5)
    x0f = np.floor(ix)
    y0f = np.floor(iy)
    wx1 = ix - x0f
    wx0 = (x0f + f32(1)) - ix
    wy1 = iy - y0f
    wy0 = (y0f + f32(1)) - iy
    # clip before the int cast: wildly out-of-range coordinates are out of bounds either way
    x0 = np.clip(x0f, -2.0, W + 1.0).astype(np.int32)
    y0 = np.clip(y0f, -2.0, H + 1.0).astype(np.int32)
    return x0, y0, (wx0.astype(f32), wx1.astype(f32)), (wy0.astype(f32), wy1.astype(f32))


def warp_bilinear(x, disp):
    """A.3 / models/models.py:28-55.  x [N,C,H,W], disp [N,1,H,W]."""
    N, C, H, W = x.shape
    x0, y0, (wx0, wx1), (wy0, wy1) = warp_taps(disp[:, 0], H, W)
    xp = np.zeros((N, C, H + 6, W + 6), x.dtype)
    xp[:, :, 3:H + 3, 3:W + 3] = x
    X0 = x0 + 3
    Y0 = np.broadcast_to((y0 + 3)[None, :, None], x0.shape)
    n = np.arange(N)[:, None, None, None]
    c = np.arange(C)[None, :, None, None]

    def tap(yy, xx):
        return xp[n, c, yy[:, None], xx[:, None]]

    wy0b = wy0[None, None, :, None]
    wy1b = wy1[None, None, :, None]
    return (tap(Y0, X0) * (wx0[:, None] * wy0b) + tap(Y0, X0 + 1) * (wx1[:, None] * wy0b)
            + tap(Y0 + 1, X0) * (wx0[:, None] * wy1b) + tap(Y0 + 1, X0 + 1) * (wx1[:, None] * wy1b)).astype(x.dtype)


def warp_residual_volume_l1(L, R, disp, m, stride=1):
    """A.4 / models/models.py:78-104.  disp [B,1,H,W] -> cost [B,2m-1,H,W]; plane k uses warp arg disp - (k-(m-1))*stride."""
    B, C, H, W = L.shape
    K = 2 * m - 1
    out = np.zeros((B, K, H, W), L.dtype)
    for k in range(K):
        shift = f32((k - (m - 1)) * stride)
        wr = warp_bilinear(R, disp - shift)
        out[:, k] = np.abs(L - wr).sum(axis=1, dtype=L.dtype)
    return out


def softmax_regression(cost, start, step=1.0):
    """A.6 / models/models.py:142,151-152,167-179: sum_j softmax_j(-cost) * (start + j*step)."""
    z = -cost
    z = z - z.max(axis=1, keepdims=True)
    e = np.exp(z)
    p = e / e.sum(axis=1, keepdims=True)
    v = (start + step * np.arange(cost.shape[1])).astype(cost.dtype)[None, :, None, None]
    return (p * v).sum(axis=1, keepdims=True).astype(cost.dtype)


def scale_upsample_add(low, prev, H, W):
    """A.7 / models/models.py:145-148,153-156: resize((low*float(H))*fl32(1/h), (H,W)) (+ prev)."""
    h = low.shape[2]
    up = bilinear_resize_halfpixel((low * f32(H)) * f32(1.0 / h), H, W)
    return up if prev is None else (up + prev).astype(low.dtype)
