"""CPU: the oracle against its independent closed-form spec, torch's own kernels, and the committed golden vectors.

PARITY UNPINNED (see oracle/__init__.py): the reference ships no tests or fixtures and Paddle cannot be installed here,
so these tests pin the oracle's two formulations against each other and against oracle-generated golden files.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import lwsnet_torch as O
from oracle import spec_np as S
from util import golden, rnd


def test_cost_volume_literal_vs_closed_form():
    L, R = rnd(1, 2, 16, 9, 41, scale=2.0), rnd(2, 2, 16, 9, 41, scale=2.0)
    a, b = O.build_volume_2d(L, R, 24).numpy(), S.cost_volume_l1(L.numpy(), R.numpy(), 24)
    assert np.abs(a - b).max() <= 1e-5 * max(1.0, np.abs(b).max())
    # D > W: planes with d >= W are the pure occlusion branch
    a = O.build_volume_2d(L[..., :10], R[..., :10], 16).numpy()
    assert np.allclose(a[:, 12], L[..., :10].abs().sum(1).numpy(), atol=1e-5)


def test_residual_volume_literal_vs_closed_form():
    L, R = rnd(3, 2, 8, 7, 38, scale=2.0), rnd(4, 2, 8, 7, 38, scale=2.0)
    disp = rnd(5, 2, 1, 7, 38, scale=12.0) + 6.0
    a, b = O.build_volume_2d3(L, R, 5, disp).numpy(), S.warp_residual_volume_l1(L.numpy(), R.numpy(), disp.numpy(), 5)
    assert np.abs(a - b).max() <= 1e-5 * max(1.0, np.abs(b).max())


def test_warp_matches_torch_grid_sample():
    """Paddle grid_sample defaults == torch grid_sample(bilinear, zeros, align_corners=True) (Appendix C.1)."""
    x = rnd(6, 2, 5, 11, 29)
    disp = rnd(7, 2, 1, 11, 29, scale=8.0)
    N, C, H, W = x.shape
    xx = torch.arange(W, dtype=torch.float32).view(1, 1, 1, W).expand(N, 1, H, W) - disp
    yy = torch.arange(H, dtype=torch.float32).view(1, 1, H, 1).expand(N, 1, H, W)
    gx = (2.0 * xx) * float(np.float32(1.0 / (W - 1))) - 1.0
    gy = (2.0 * yy) * float(np.float32(1.0 / (H - 1))) - 1.0
    ref = F.grid_sample(x, torch.cat([gx, gy], 1).permute(0, 2, 3, 1), mode="bilinear", padding_mode="zeros",
                        align_corners=True)
    assert (O.warp(x, disp) - ref).abs().max().item() < 5e-6
    assert np.abs(S.warp_bilinear(x.numpy(), disp.numpy()) - ref.numpy()).max() < 5e-6


def test_fp32_coordinate_round_trip_is_not_identity():
    """for H=92, 22 integer rows come back as j - eps (floor = j-1)."""
    for n, expected in ((46, 11), (92, 22), (184, 58), (308, 126)):
        _, y0, _, _ = S.warp_taps(np.zeros((1, n, 1), np.float32), n, 1)
        assert int((y0 != np.arange(n)).sum()) == expected, n


def test_regression_and_resize_vs_closed_form():
    c = rnd(8, 2, 24, 6, 20, scale=8.0)
    a = O.disparity_regression(0, 24)(torch.softmax(-c, 1)).numpy()
    assert np.abs(a - S.softmax_regression(c.numpy(), 0.0)).max() <= 2e-5
    p = rnd(9, 2, 1, 48, 160, scale=30.0)
    a = (O.interpolate_bilinear(p, (12, 40)) * 12.0 * O._recip(48, torch.float32)).numpy()
    assert np.abs(a - S.disp_to_scale(p.numpy(), 12, 40)).max() <= 1e-5
    low = rnd(10, 2, 1, 12, 40)
    a = O.interpolate_bilinear(low * 48.0 * O._recip(12, torch.float32), (48, 160)).numpy()
    assert np.abs(a - S.scale_upsample_add(low.numpy(), None, 48, 160)).max() <= 1e-5


def test_parameter_inventory():
    """226 state tensors, 177,890 parameters (conv weights + BN affine)."""
    m = O.build_oracle(0)
    assert len(m.state_dict()) == 226
    assert sum(p.numel() for p in m.parameters()) == 177890
    assert m.volume_postprocess[0][1][2].weight.shape == (32, 32, 3, 3, 3)
    assert m.refinement2[0][2].weight.shape == (32, 64, 3, 3)


def test_oracle_against_golden():
    torch.set_num_threads(1)
    g = golden("cost_volume")
    assert np.array_equal(O.build_volume_2d(torch.from_numpy(g["odd_L"]), torch.from_numpy(g["odd_R"]), 24).numpy(), g["odd_cost"])
    g = golden("warp_volume")
    L, R, disp = (torch.from_numpy(g[k]) for k in ("L", "R", "disp"))
    assert np.allclose(O.build_volume_2d3(L, R, 5, disp).numpy(), g["cost"], atol=1e-6)
    x0, y0, (wx0, wx1), _ = S.warp_taps(g["disp"][:, 0] - g["taps_shift"], 7, 38)
    assert np.array_equal(x0, g["x0"]) and np.array_equal(y0, g["y0"])
    assert np.array_equal(np.stack([wx0, wx1], -1).view(np.uint32), g["wx"].view(np.uint32))
    g = golden("regression")
    assert np.allclose(S.softmax_regression(g["c24"], 0.0), g["low24"], atol=2e-5)
    assert np.allclose(S.scale_upsample_add(g["low"], g["prev"], 48, 160), g["up_prev"], atol=1e-4)


def test_oracle_model_against_golden():
    g = golden("model_small")
    m = O.build_oracle(seed=0, random_bn=True)
    pred = m(torch.from_numpy(g["left"]), torch.from_numpy(g["right"]))
    for s in range(4):
        ref = torch.from_numpy(g[f"pred32_{s}"])
        assert (pred[s] - ref).abs().max().item() <= 2e-3 * (1 + ref.abs().max().item()), s  # thread-count dependent sums


def test_preprocess_matches_reference_crop():
    """inference.py:93-103: bottom-right 368x1232 crop, BGR->RGB, /255, ImageNet normalise."""
    img = (np.arange(375 * 1242 * 3) % 251).astype(np.uint8).reshape(375, 1242, 3)
    t = O.preprocess_bgr_uint8(img)
    assert tuple(t.shape) == (1, 3, 368, 1232)
    r = img[374, 1241, 2] / 255.0
    assert abs(t[0, 0, -1, -1].item() - (r - 0.485) / 0.229) < 1e-6


def test_jet_fixture_matches_opencv_definition():
    """tests/golden/jet_lut_bgr.npy (made by oracle/make_golden.py:make_jet_lut with cv2.applyColorMap, inference.py:115):
    the known anchor colours of COLORMAP_JET, and agreement with cv2 itself when it is importable."""
    import os
    lut = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "jet_lut_bgr.npy"))
    assert lut.shape == (256, 3) and lut.dtype == np.uint8
    assert tuple(lut[0]) == (128, 0, 0) and tuple(lut[255]) == (0, 0, 128)      # dark blue -> dark red (BGR)
    assert tuple(lut[96]) == (255, 255, 0) or lut[96][1] == 255                    # cyan band
    try:
        import cv2
    except ImportError:
        return
    ref = cv2.applyColorMap(np.arange(256, dtype=np.uint8).reshape(256, 1), cv2.COLORMAP_JET).reshape(256, 3)
    assert np.array_equal(lut, ref)
