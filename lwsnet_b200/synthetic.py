"""Synthetic workloads owned by the package (BASELINE.json configs): KITTI / SceneFlow-shaped stereo pairs and the
reference's default hyper-parameters, so that `import lwsnet_b200` alone reproduces the benchmark (no dataset, no checkpoint:
both are unavailable offline).  Input generation runs on the host with torch CPU ops; it is not part of the hot path.
"""
from __future__ import annotations

from types import SimpleNamespace

import torch
import torch.nn.functional as F

# the BASELINE.json configs: (name, H, W, maxdisplist, pairs per step on one GPU)
CONFIGS = {
    0: dict(name="configs[0]: reference/left_test.png + right_test.png (KITTI 1232x368 crop), batch 1", H=368, W=1232,
            maxdisplist=(24, 5, 5), batch=1),
    1: dict(name="configs[1]: stage 1 only (cost volume + C=32 3D stack + regression), KITTI 1232x368, batch 8", H=368, W=1232,
            maxdisplist=(24, 5, 5), batch=8),
    2: dict(name="configs[2]: full 4-stage inference, KITTI 1232x368, batch 64 per GPU", H=368, W=1232, maxdisplist=(24, 5, 5),
            batch=64),
    3: dict(name="configs[3]: SceneFlow-shaped 960x544 4-stage inference, batch 256 sharded over the GPUs", H=544, W=960,
            maxdisplist=(24, 5, 5), batch=256),
    4: dict(name="configs[4]: 1920x1088, maxdisp 384 (D = 48 at 1/8), 4-stage inference, batch 32 sharded over the GPUs", H=1088,
            W=1920, maxdisplist=(48, 5, 5), batch=32),
}


def default_args(maxdisplist=(24, 5, 5), layers_3d=4, channels_3d=8, growth_rate=(4, 1, 1)):
    """The reference's argparse defaults (inference.py:23-26, train.py) as the namespace LWSNet(args) reads."""
    return SimpleNamespace(maxdisplist=list(maxdisplist), layers_3d=layers_3d, channels_3d=channels_3d,
                           growth_rate=list(growth_rate))


def random_init_model(seed=0, args=None, device=None):
    """LWSNet with the reference's random initialisation (KaimingNormal fan-in convs, identity BatchNorm), seeded."""
    from .models import LWSNet
    gen_state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    try:
        model = LWSNet(args or default_args())
    finally:
        torch.random.set_rng_state(gen_state)
    return model.to(device) if device is not None else model


def synthetic_pair(B, H, W, seed=1234, max_disp=150.0):
    """B stereo pairs [B,3,H,W] fp32 in normalised-image range: a band-limited random texture plus fine noise as the left view, the
    right view = the left view displaced by a smooth random disparity field in [0, max_disp] px plus sensor noise."""
    lefts, rights = [], []
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
        lefts.append(left)
        rights.append(right)
    return torch.cat(lefts), torch.cat(rights)


def synthetic_images_u8(B, h=375, w=1242, seed=99):
    """B pairs of uint8 HWC BGR images as cv2.imread returns them (KITTI frame size by default)."""
    g = torch.Generator().manual_seed(seed)
    left = torch.randint(0, 256, (B, h, w, 3), dtype=torch.uint8, generator=g)
    right = torch.randint(0, 256, (B, h, w, 3), dtype=torch.uint8, generator=g)
    return left, right
