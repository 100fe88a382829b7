"""Generate tests/golden/*.npz from the oracle (TEST INFRASTRUCTURE — see oracle/__init__.py).

PARITY UNPINNED: Paddle cannot be imported here and the reference ships no fixtures, so these vectors are outputs of
the oracle itself (oracle/lwsnet_torch.py, cross-checked against oracle/spec_np.py by tests/test_oracle.py).  They
pin the oracle against silent drift and travel to the GPU box, where the CUDA kernels are compared with them.

Run from the repo root:  python -m oracle.make_golden
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import lwsnet_torch as O
from . import spec_np as S

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def rng_t(seed, *shape, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


def stack_state(mod):
    return {k: v.numpy() for k, v in mod.state_dict().items()}


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)  # deterministic summation order

    # ---- a1 cost volume -----------------------------------------------------------------------------------
    cases = {}
    for name, (B, C, H, W, D) in {"odd": (2, 16, 6, 37, 24), "d12": (1, 8, 5, 40, 12), "dgtw": (1, 4, 3, 10, 16)}.items():
        L, R = rng_t(1, B, C, H, W, scale=2.0), rng_t(2, B, C, H, W, scale=2.0)
        cases[f"{name}_L"], cases[f"{name}_R"] = L.numpy(), R.numpy()
        cases[f"{name}_cost"] = O.build_volume_2d(L, R, D).numpy()
    np.savez_compressed(os.path.join(OUT, "cost_volume.npz"), **cases)

    # ---- a2/a3/a4 warp + residual volume --------------------------------------------------------------------
    B, C, H, W, m = 2, 16, 7, 38, 5
    L, R = rng_t(3, B, C, H, W, scale=2.0), rng_t(4, B, C, H, W, scale=2.0)
    disp = rng_t(5, B, 1, H, W, scale=12.0) + 6.0  # pushes samples out of bounds on both sides
    x0, y0, (wx0, wx1), (wy0, wy1) = S.warp_taps(disp[:, 0].numpy() - np.float32(2.0), H, W)
    pred_full = rng_t(6, B, 1, 4 * H, 4 * W, scale=20.0)
    np.savez_compressed(
        os.path.join(OUT, "warp_volume.npz"), L=L.numpy(), R=R.numpy(), disp=disp.numpy(),
        warped=O.warp(R, disp).numpy(), cost=O.build_volume_2d3(L, R, m, disp).numpy(),
        taps_shift=np.float32(2.0), x0=x0, y0=y0, wx=np.stack([wx0, wx1], -1), wy=np.stack([wy0, wy1], -1),
        pred_full=pred_full.numpy(),
        wflow=(O.interpolate_bilinear(pred_full, (H, W)) * float(H) * O._recip(4 * H, torch.float32)).numpy())

    # ---- a6/a7 regression + upsample --------------------------------------------------------------------------
    c24, c9 = rng_t(7, 2, 24, 6, 20, scale=8.0), rng_t(8, 2, 9, 5, 13, scale=30.0)
    low = rng_t(9, 2, 1, 6, 20, scale=5.0)
    prev = rng_t(10, 2, 1, 48, 160, scale=30.0)
    up = O.interpolate_bilinear(low * 48.0 * O._recip(6, torch.float32), (48, 160))
    np.savez_compressed(
        os.path.join(OUT, "regression.npz"), c24=c24.numpy(), c9=c9.numpy(),
        low24=O.disparity_regression(0, 24)(torch.softmax(-c24, 1)).numpy(),
        low9=O.disparity_regression(-4, 5)(torch.softmax(-c9, 1)).numpy(),
        low=low.numpy(), prev=prev.numpy(), up=up.numpy(), up_prev=(up + prev).numpy())

    # ---- a5 3D stacks (random BN statistics so the folding is exercised); outputs from the fp64 oracle ----------
    out = {}
    for name, (C, B, D, H, W) in {"c8": (8, 1, 9, 10, 36), "c32": (32, 1, 6, 9, 34)}.items():
        torch.manual_seed(11)
        net = O.post_3dconvs(4, C)
        holder = torch.nn.Module()
        holder.net = net
        O.kaiming_normal_init_(holder, 12)
        O.randomize_bn_(holder, 13)
        cost = rng_t(14, B, D, H, W, scale=6.0).abs()
        net64 = net.double()
        with torch.no_grad():
            y = net64(cost.double().unsqueeze(1)) + cost.double().unsqueeze(1)
        out[f"{name}_cost"] = cost.numpy()
        out[f"{name}_out64"] = y[:, 0].numpy()
        for k, v in net.float().state_dict().items():
            out[f"{name}_w_{k}"] = v.numpy()
    np.savez_compressed(os.path.join(OUT, "conv3d_stack.npz"), **out)

    # ---- a8+a9 refinement (random BN) and end-to-end, small image ------------------------------------------------
    m32 = O.build_oracle(seed=0, random_bn=True)
    m64 = O.build_oracle(seed=0, random_bn=True, dtype=torch.float64)
    left, right = O.synthetic_pair(1, 64, 128, seed=77, max_disp=20.0)
    pred3 = rng_t(15, 1, 1, 64, 128, scale=10.0) + 20.0
    with torch.no_grad():
        r64 = m64.refine(left.double(), pred3.double())
        p32, _ = m32.forward_trace(left, right)
        p64, _ = m64.forward_trace(left.double(), right.double())
    np.savez_compressed(
        os.path.join(OUT, "model_small.npz"), left=left.numpy(), right=right.numpy(), pred3=pred3.numpy(),
        refine64=r64.numpy(), **{f"pred32_{i}": p.numpy() for i, p in enumerate(p32)},
        **{f"pred64_{i}": p.numpy() for i, p in enumerate(p64)})
    total = sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT))
    print(f"wrote {sorted(os.listdir(OUT))} ({total / 1e6:.2f} MB) to {OUT}")


def make_jet_lut():
    """cv2.COLORMAP_JET as a 256 x 3 BGR table (inference.py:115 applyColorMap): the fixture the n2 output kernel is checked against."""
    import cv2
    lut = cv2.applyColorMap(np.arange(256, dtype=np.uint8).reshape(256, 1), cv2.COLORMAP_JET).reshape(256, 3)
    np.save(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "jet_lut_bgr.npy"), lut)


if __name__ == "__main__":
    main()
    make_jet_lut()
