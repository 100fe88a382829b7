"""GPU parity tests: every hot-path kernel, called through the C ABI, against the CPU oracle (SURVEY.md 8(c) protocol 1).

Tolerances (fp32 kernels, stated per test):
  K1/K2 volumes     |d| <= 1e-5 * max(1, |cost|)      (summation order differs from the oracle's)
  warp taps         indices and both weights BIT-EXACT
  K4 regression     |d| <= 2e-5 low-res px
  K5 / wflow        |d| <= 1e-5 * max(1, |x|)
  K3 / K6 convs     |d| <= 1e-4 * (1 + |y|) against the fp64 oracle
"""
import numpy as np
import pytest
import torch

from util import cu, golden, rnd

pytestmark = pytest.mark.gpu


def ops():
    from lwsnet_b200 import ops as o
    return o


def close_rel(a, b, tol, floor=1.0):
    a, b = a.double().cpu(), torch.as_tensor(b).double()
    bound = tol * torch.clamp(b.abs(), min=floor)
    bad = (a - b).abs() > bound
    assert not bad.any(), f"{int(bad.sum())} / {bad.numel()} beyond tol; max err {(a - b).abs().max().item():.3e}"


# ------------------------------------------------------------------------------------------------ a1
@pytest.mark.parametrize("case,D", [("odd", 24), ("d12", 12), ("dgtw", 16)])
def test_cost_volume_golden(case, D):
    g = golden("cost_volume")
    out = ops().cost_volume_l1(cu(g[f"{case}_L"]), cu(g[f"{case}_R"]), D)
    close_rel(out, g[f"{case}_cost"], 1e-5)


@pytest.mark.parametrize("B,C,H,W,D", [(2, 16, 46, 154, 24), (1, 16, 23, 77, 12), (1, 16, 20, 64, 48), (1, 8, 9, 33, 20),
                                        (1, 16, 4, 1300, 24), (3, 6, 5, 31, 8), (1, 3, 4, 9, 7)])
def test_cost_volume_vs_spec(B, C, H, W, D):
    from oracle import spec_np as S
    L, R = rnd(1, B, C, H, W, scale=2.0), rnd(2, B, C, H, W, scale=2.0)
    out = ops().cost_volume_l1(L.cuda(), R.cuda(), D)
    close_rel(out, S.cost_volume_l1(L.numpy(), R.numpy(), D), 1e-5)


def test_cost_volume_stride2_and_assert():
    from oracle import lwsnet_torch as O
    L, R = rnd(3, 1, 8, 6, 40), rnd(4, 1, 8, 6, 40)
    out = ops().cost_volume_l1(L.cuda(), R.cuda(), 12, stride=2)
    close_rel(out, O.build_volume_2d(L, R, 12, stride=2), 1e-5)
    with pytest.raises(AssertionError):  # reference: assert maxdisp % stride == 0 (models.py:63)
        ops().cost_volume_l1(L.cuda(), R.cuda(), 13, stride=2)


def test_cost_volume_linearity_fullsize():
    """Size-independent property at the BASELINE configs[1] size: cost(aL, aR) == a * cost(L, R) exactly for a = 2."""
    L, R = rnd(5, 8, 16, 46, 154, scale=2.0).cuda(), rnd(6, 8, 16, 46, 154, scale=2.0).cuda()
    c1 = ops().cost_volume_l1(L, R, 24)
    c2 = ops().cost_volume_l1(2 * L, 2 * R, 24)
    assert torch.equal(c2, 2 * c1)
    # plane 0 is the plain L1 distance; plane d at x < d is |L|_1 (occlusion branch)
    close_rel(c1[:, 0], (L - R).abs().sum(1).cpu(), 1e-5)
    close_rel(c1[:, 23, :, :23], L[..., :23].abs().sum(1).cpu(), 1e-5)


# ------------------------------------------------------------------------------------------------ a2 / a3 / a4
def test_warp_taps_bit_exact_golden():
    g = golden("warp_volume")
    x0, y0, wx, wy = ops().warp_taps(cu(g["disp"]), float(g["taps_shift"]))
    assert torch.equal(x0.cpu(), torch.from_numpy(g["x0"]))
    assert torch.equal(y0.cpu(), torch.from_numpy(g["y0"]))
    assert np.array_equal(wx.cpu().numpy().view(np.uint32), g["wx"].view(np.uint32))
    assert np.array_equal(wy.cpu().numpy().view(np.uint32), g["wy"].view(np.uint32))


@pytest.mark.parametrize("H,W", [(92, 308), (184, 616), (46, 154), (136, 240), (272, 480), (7, 2), (1, 1)])
def test_warp_taps_bit_exact(H, W):
    from oracle import spec_np as S
    disp = rnd(7, 2, 1, H, W, scale=40.0) + 30.0
    disp[0, 0, 0, 0] = 0.0
    disp[1, 0, -1, -1] = 1e9   # wildly out of range both ways
    disp[1, 0, 0, -1] = -1e9
    # both readings of the reference's `tensor / python scalar`: reciprocal multiply (Paddle 2.0, the oracle's
    # choice and the library default) and true division -- bit-exact under each
    for div_mode in (0, 1):
        with ops().options(warp_div_mode=div_mode):
            for shift in (-4.0, 0.0, 3.0):
                x0, y0, wx, wy = ops().warp_taps(disp.cuda(), shift)
                ex0, ey0, (ewx0, ewx1), (ewy0, ewy1) = S.warp_taps(disp[:, 0].numpy() - np.float32(shift), H, W, div_mode)
                assert np.array_equal(x0.cpu().numpy(), ex0)
                assert np.array_equal(y0.cpu().numpy(), ey0)
                assert np.array_equal(wx.cpu().numpy().view(np.uint32), np.stack([ewx0, ewx1], -1).view(np.uint32))
                assert np.array_equal(wy.cpu().numpy().view(np.uint32), np.stack([ewy0, ewy1], -1).view(np.uint32))


def test_warp_and_residual_volume_golden():
    g = golden("warp_volume")
    L, R, disp = cu(g["L"]), cu(g["R"]), cu(g["disp"])
    close_rel(ops().warp_bilinear(R, disp), g["warped"], 1e-5)
    close_rel(ops().warp_residual_volume_l1(L, R, disp, 5), g["cost"], 1e-5)
    close_rel(ops().disp_to_scale(cu(g["pred_full"]), 7, 38), g["wflow"], 1e-5)


@pytest.mark.parametrize("B,C,H,W,m", [(2, 16, 92, 308, 5), (1, 8, 184, 616, 5), (1, 16, 12, 40, 3), (2, 4, 9, 17, 1),
                                        (1, 8, 10, 33, 7)])
def test_residual_volume_vs_oracle(B, C, H, W, m):
    from oracle import lwsnet_torch as O
    L, R = rnd(8, B, C, H, W, scale=2.0), rnd(9, B, C, H, W, scale=2.0)
    disp = rnd(10, B, 1, H, W, scale=W / 8.0) + W / 6.0
    out = ops().warp_residual_volume_l1(L.cuda(), R.cuda(), disp.cuda(), m)
    close_rel(out, O.build_volume_2d3(L, R, m, disp), 1e-5)


def test_residual_volume_integer_disp_matches_shifted_l1():
    """Property: with zero disparity plane k = m-1 of the residual volume is the plain L1 distance, up to the y-row
    blend weights of the reference's fp32 coordinate round trip (|iy - y| <= 6.1e-5)."""
    B, C, H, W = 2, 16, 92, 308
    L, R = rnd(11, B, C, H, W).cuda(), rnd(12, B, C, H, W).cuda()
    out = ops().warp_residual_volume_l1(L, R, torch.zeros(B, 1, H, W, device="cuda"), 5)
    ref = (L - R).abs().sum(1)
    assert (out[:, 4] - ref).abs().max().item() < 5e-3


@pytest.mark.parametrize("B,H,W,h,w", [(2, 368, 1232, 92, 308), (1, 368, 1232, 184, 616), (1, 64, 128, 16, 32), (1, 30, 50, 7, 11)])
def test_disp_to_scale(B, H, W, h, w):
    from oracle import spec_np as S
    p = rnd(13, B, 1, H, W, scale=40.0)
    close_rel(ops().disp_to_scale(p.cuda(), h, w), S.disp_to_scale(p.numpy(), h, w), 1e-5)


# ------------------------------------------------------------------------------------------------ a6 / a7
def test_regression_golden():
    g = golden("regression")
    out24 = ops().softmax_regression(cu(g["c24"]), 0.0)
    out9 = ops().softmax_regression(cu(g["c9"]), -4.0)
    assert (out24.cpu() - torch.from_numpy(g["low24"])).abs().max().item() <= 2e-5
    assert (out9.cpu() - torch.from_numpy(g["low9"])).abs().max().item() <= 2e-5
    up = ops().scale_upsample_add(cu(g["low"]), None, 48, 160)
    close_rel(up, g["up"], 1e-5)
    close_rel(ops().scale_upsample_add(cu(g["low"]), cu(g["prev"]), 48, 160), g["up_prev"], 1e-5)


@pytest.mark.parametrize("B,D,H,W,start", [(2, 24, 46, 154, 0.0), (1, 9, 92, 308, -4.0), (1, 9, 7, 9, -4.0), (1, 48, 17, 24, 0.0),
                                            (1, 1, 4, 4, 0.0), (1, 13, 3, 5, -6.0)])
def test_regression_vs_spec(B, D, H, W, start):
    from oracle import spec_np as S
    for scale in (1.0, 10.0, 100.0):  # soft, sharp, arg-min-like volumes
        c = rnd(14, B, D, H, W, scale=scale)
        out = ops().softmax_regression(c.cuda(), start)
        ref = S.softmax_regression(c.double().numpy(), start)
        assert (out.cpu().double() - torch.from_numpy(ref)).abs().max().item() <= 2e-5


def test_regression_shift_invariance_fullsize():
    """Property: softmax regression is invariant to adding a per-pixel constant to the cost column."""
    c = (torch.round(rnd(15, 8, 24, 46, 154, scale=8.0) * 1024) / 1024).cuda()  # so that c + 64 is exact in fp32
    a = ops().softmax_regression(c, 0.0)
    b = ops().softmax_regression(c + 64.0, 0.0)
    assert (a - b).abs().max().item() <= 2e-5
    assert a.min().item() >= 0.0 and a.max().item() <= 23.0 + 1e-5


@pytest.mark.parametrize("B,h,w,H,W", [(2, 46, 154, 368, 1232), (1, 184, 616, 368, 1232), (1, 5, 7, 40, 56), (1, 6, 9, 13, 31)])
def test_scale_upsample_add(B, h, w, H, W):
    from oracle import spec_np as S
    low, prev = rnd(16, B, 1, h, w, scale=5.0), rnd(17, B, 1, H, W, scale=30.0)
    close_rel(ops().scale_upsample_add(low.cuda(), None, H, W), S.scale_upsample_add(low.numpy(), None, H, W), 1e-5)
    close_rel(ops().scale_upsample_add(low.cuda(), prev.cuda(), H, W), S.scale_upsample_add(low.numpy(), prev.numpy(), H, W), 1e-5)


@pytest.mark.parametrize("B,h,w,S", [(2, 46, 154, 8), (2, 92, 308, 4), (2, 184, 616, 2), (1, 5, 7, 2), (1, 3, 4, 4), (3, 2, 3, 8), (1, 13, 33, 2),
                                     (1, 9, 16, 4)])
def test_scale_upsample_add_integer_scale_path_bit_identical(B, h, w, S):
    """K5's periodic-tap fast path for the integer scales 2 / 4 / 8 (option k5_int = 1, default) == the generic kernel, bitwise, with and
    without the previous-stage skip, image borders and tiny images included."""
    H, W = h * S, w * S
    low = (rnd(81, B, 1, h, w, scale=5.0)).cuda()
    prev = (rnd(82, B, 1, H, W, scale=9.0)).cuda()
    for pv in (None, prev):
        with ops().options(k5_int=0):
            ref = ops().scale_upsample_add(low, pv, H, W).clone()
        with ops().options(k5_int=1):
            out = ops().scale_upsample_add(low, pv, H, W)
        assert torch.equal(out, ref), (B, h, w, S, pv is None)


@pytest.mark.parametrize("B,D,h,w,H,W,nhw,start", [
    (2, 24, 46, 154, 368, 1232, (92, 308), 0.0),     # stage 1 -> wflow of stage 2 (decimation 4)
    (1, 9, 92, 308, 368, 1232, (184, 616), -4.0),    # stage 2 -> wflow of stage 3 (decimation 2)
    (2, 9, 184, 616, 368, 1232, None, -4.0),         # stage 3: no next stage
    (1, 48, 136, 240, 1088, 1920, (272, 480), 0.0),  # configs[4]
    (1, 9, 16, 32, 64, 128, (32, 64), -4.0),         # smaller than one tile
    (1, 24, 17, 30, 136, 240, (34, 60), 0.0),        # ragged tiles
    (1, 9, 40, 56, 40, 56, None, -4.0),              # scale 1
])
def test_regression_tail_fused_is_bit_identical(B, D, h, w, H, W, nhw, start):
    """lws_regression_tail_f32 (K4 + K5 + the next stage's K2a in one pass) == the three stand-alone kernels, BITWISE, with and
    without the previous-stage skip; and against the numpy spec within the K4 / K5 tolerances."""
    from oracle import spec_np as S
    for scale in (1.0, 30.0):
        cost = rnd(18, B, D, h, w, scale=scale).cuda()
        prev = rnd(19, B, 1, H, W, scale=30.0).cuda()
        for pv in (None, prev):
            pf, wf = ops().regression_tail(cost, pv, H, W, start, 1.0, next_hw=nhw, fused=True)
            pu, wu = ops().regression_tail(cost, pv, H, W, start, 1.0, next_hw=nhw, fused=False)
            assert torch.equal(pf, pu)
            assert (wf is None and wu is None) or torch.equal(wf, wu)
        low = S.softmax_regression(cost.cpu().double().numpy(), start)
        ref = S.scale_upsample_add(low.astype(np.float32), prev.cpu().numpy(), H, W)
        # K4's bar is 2e-5 LOW-resolution pixels; the rescale by H/h turns it into 2e-5 * H/h full-resolution pixels (+ K5's 1e-5 rel)
        err = (pf.cpu().double() - torch.from_numpy(ref).double()).abs()
        assert (err <= 2e-5 * (H / h) + 1e-5 * torch.from_numpy(ref).double().abs()).all(), err.max().item()
        if nhw is not None:
            close_rel(wf, S.disp_to_scale(pf.cpu().numpy(), *nhw), 1e-5)
    assert ops().lib.lws_regression_tail_supported(5, 7, 40, 56, 0, 0) == 0
    assert ops().lib.lws_regression_tail_supported(6, 9, 13, 31, 0, 0) == -5       # non-integer scale -> stand-alone kernels
    assert ops().lib.lws_regression_tail_supported(46, 154, 368, 1232, 123, 411) == -5


# ------------------------------------------------------------------------------------------------ a5
def _stack_from_golden(prefix, C):
    from lwsnet_b200.submodules import post_3dconvs
    g = golden("conv3d_stack")
    net = post_3dconvs(4, C)
    sd = {k[len(prefix) + 3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith(prefix + "_w_")}
    net.load_state_dict(sd, strict=True)
    return net.cuda(), g


# The C = 32 and C = 8 stacks have two implementations: the tcgen05 split-fp16 (x = hi + lo*2^-11, three exact products) implicit
# GEMM (default) and the fp32 FFMA kernels (option "conv3d_tc" = 0).  Both meet SURVEY 8(c)'s 1e-4 * (1 + |y|) against the fp64
# oracle (measured r02: max |d| 5.8e-5 tensor-core vs 6.0e-5 FFMA on the C = 32 KITTI volume, fp32 oracle 3.3e-5), and both must
# stay within 2.5x of the fp32 oracle's own worst error (+1e-6 absolute) so that a regression of a few ulps per layer is caught.
CONV3D_PATHS = [("tc", "1", 1e-4), ("ffma", "0", 1e-4)]


@pytest.fixture(params=CONV3D_PATHS, ids=[p[0] for p in CONV3D_PATHS])
def conv3d_path(request):
    name, env, tol = request.param
    with ops().options(conv3d_tc=int(env)):
        yield name, tol


@pytest.mark.parametrize("name,C", [("c8", 8), ("c32", 32)])
def test_conv3d_stack_golden(name, C, conv3d_path):
    tol = conv3d_path[1]
    net, g = _stack_from_golden(name, C)
    cost = cu(g[f"{name}_cost"])
    out = net.run(cost, add_skip=True)
    ref = torch.from_numpy(g[f"{name}_out64"])
    err = (out.cpu().double() - ref).abs()
    assert (err <= tol * (1 + ref.abs())).all(), f"max err {err.max().item():.3e}"
    noskip = net(cost.unsqueeze(1))[:, 0]
    err2 = (noskip.cpu().double() + torch.from_numpy(g[f"{name}_cost"]).double() - ref).abs()
    assert (err2 <= tol * (1 + ref.abs())).all()


@pytest.mark.parametrize("C,B,D,H,W", [(32, 2, 24, 46, 154), (8, 1, 9, 92, 308), (8, 1, 9, 40, 70), (16, 1, 5, 11, 30), (32, 1, 3, 5, 9),
                                        (32, 3, 5, 7, 33), (8, 2, 9, 184, 616), (8, 3, 4, 6, 9), (8, 1, 2, 3, 5)])
def test_conv3d_stack_vs_fp64_oracle(C, B, D, H, W, conv3d_path):
    from oracle import lwsnet_torch as O
    from lwsnet_b200.submodules import post_3dconvs
    if C == 16 and conv3d_path[0] == "tc":
        pytest.skip("C = 16 has the FFMA implementation only")
    tol = conv3d_path[1]
    onet = O.post_3dconvs(4, C)
    holder = torch.nn.Module()
    holder.net = onet
    O.kaiming_normal_init_(holder, 21)
    O.randomize_bn_(holder, 22)
    net = post_3dconvs(4, C)
    net.load_state_dict(onet.state_dict(), strict=True)
    cost = rnd(23, B, D, H, W, scale=6.0).abs()
    out = net.cuda().run(cost.cuda(), add_skip=True)
    with torch.no_grad():
        ref = (onet.double()(cost.double().unsqueeze(1)) + cost.double().unsqueeze(1))[:, 0]
        floor = ((onet.float()(cost.unsqueeze(1)) + cost.unsqueeze(1))[:, 0].double() - ref).abs().max().item()
    err = (out.cpu().double() - ref).abs()
    print(f"conv3d stack C={C} [{B},{D},{H},{W}] path={conv3d_path[0]}: max err {err.max().item():.3e} (fp32 oracle max err {floor:.3e})")
    assert (err <= tol * (1 + ref.abs())).all(), f"max err {err.max().item():.3e}"
    assert err.max().item() <= 2.5 * floor + 1e-6, f"max err {err.max().item():.3e} vs fp32 oracle {floor:.3e}"


@pytest.mark.parametrize("C,B,D,H,W", [(32, 2, 24, 46, 154), (8, 2, 9, 92, 308), (8, 1, 9, 13, 150), (32, 1, 48, 9, 70), (8, 1, 3, 5, 9)])
def test_first_conv_versions_bit_identical(C, B, D, H, W):
    """The first 1 -> C conv with its tap window staged in shared memory (option first_conv = 8 / 4) performs the same FMAs in the
    same order as the version that reads its taps from global memory (first_conv = 0): identical stack output bits."""
    from oracle import lwsnet_torch as O
    from lwsnet_b200.submodules import post_3dconvs
    onet = O.post_3dconvs(4, C)
    holder = torch.nn.Module()
    holder.net = onet
    O.kaiming_normal_init_(holder, 31)
    O.randomize_bn_(holder, 32)
    net = post_3dconvs(4, C)
    net.load_state_dict(onet.state_dict(), strict=True)
    net = net.cuda()
    x = (rnd(33, B, D, H, W, scale=6.0).abs()).cuda()
    with ops().options(first_conv=0):
        ref = net.run(x, add_skip=True).clone()
    for ver in (1, 8, 4):
        with ops().options(first_conv=ver):
            assert torch.equal(net.run(x, add_skip=True), ref), ver


@pytest.mark.parametrize("B,D,H,W", [(2, 24, 46, 154), (1, 24, 3, 40), (3, 12, 17, 70), (1, 48, 9, 70)])
def test_tz_strips_bit_identical(B, D, H, W):
    """C = 32 mid layers with the tiles walking down y in strips (two of the three ky boxes stay in the shared-memory ring) against
    linear tiling (every tile loads its three boxes): the same MMAs on the same operands, identical bits."""
    from oracle import lwsnet_torch as O
    from lwsnet_b200.submodules import post_3dconvs
    onet = O.post_3dconvs(4, 32)
    holder = torch.nn.Module()
    holder.net = onet
    O.kaiming_normal_init_(holder, 51)
    O.randomize_bn_(holder, 52)
    net = post_3dconvs(4, 32)
    net.load_state_dict(onet.state_dict(), strict=True)
    net = net.cuda()
    x = (rnd(53, B, D, H, W, scale=6.0).abs()).cuda()
    with ops().options(tz_strips=0):
        ref = net.run(x, add_skip=True).clone()
    # contiguous tile ranges / segments dealt round-robin / CTA pairs (cta_group::2); +4 = strips for the closing 32 -> 1 conv too
    for mode in (1, 2, 3, 4, 5, 7):
        with ops().options(tz_strips=mode):
            assert torch.equal(net.run(x, add_skip=True), ref), mode


@pytest.mark.parametrize("B,Cf,D,H,W", [(2, 32, 24, 46, 154), (1, 32, 24, 7, 66), (3, 16, 12, 9, 130), (1, 32, 40, 5, 64), (1, 2, 3, 4, 6)])
def test_fused_volume_first_conv_bit_identical(B, Cf, D, H, W):
    """Stage 1 as ONE call (lws_cost_volume_conv3d_stack_f32: the volume is built inside the first conv kernel's shared-memory
    window) performs the same additions / FMAs in the same order as lws_cost_volume_l1_f32 + lws_conv3d_stack_f32: identical raw
    volume and identical stack output bits, ragged widths and heights included."""
    from oracle import lwsnet_torch as O
    from lwsnet_b200.submodules import post_3dconvs
    onet = O.post_3dconvs(4, 32)
    holder = torch.nn.Module()
    holder.net = onet
    O.kaiming_normal_init_(holder, 41)
    O.randomize_bn_(holder, 42)
    net = post_3dconvs(4, 32)
    net.load_state_dict(onet.state_dict(), strict=True)
    net = net.cuda()
    L, R = rnd(43, B, Cf, H, W).cuda(), rnd(44, B, Cf, H, W).cuda()
    o = ops()
    assert o.cost_volume_conv3d_stack_supported(B, Cf, H, W, D, 32, 4)
    cost_ref = o.cost_volume_l1(L, R, D, 1)
    out_ref = net.run(cost_ref, add_skip=True).clone()
    for mode in (1, 2):  # 12-disparity tiles where D % 12 == 0 / 8-disparity tiles
        with o.options(fuse_volume=mode):
            cost, out = o.cost_volume_conv3d_stack(L, R, D, net.packed(L.device), 32, 4)
        assert torch.equal(cost, cost_ref), mode
        assert torch.equal(out, out_ref), mode


def test_fused_volume_first_conv_unsupported_cases():
    o = ops()
    assert not o.cost_volume_conv3d_stack_supported(1, 32, 8, 153, 24, 32, 4)   # odd W: no 64-bit row loads
    assert not o.cost_volume_conv3d_stack_supported(1, 32, 8, 154, 9, 8, 4)     # C = 8 stack
    assert not o.cost_volume_conv3d_stack_supported(1, 32, 8, 154, 24, 32, 0)   # no mid layer: FFMA path


# ------------------------------------------------------------------------------------------------ a8 + a9
# Two implementations: channels-last tcgen05 split-fp16 (default) and the fp32 FFMA kernels (option "refine_tc" = 0).  Same bars
# for both (measured r02: tensor-core max |d| 1.07x the fp32 oracle's own max error, FFMA 0.95x).
REFINE_PATHS = [("tc", "1", 2.0, 2e-6), ("ffma", "0", 2.0, 2e-6)]


@pytest.fixture(params=REFINE_PATHS, ids=[p[0] for p in REFINE_PATHS])
def refine_path(request):
    name, env, floor_mult, scale_tol = request.param
    with ops().options(refine_tc=int(env)):
        yield name, floor_mult, scale_tol


def _check_refine(out, ref, fp32_floor, path):
    """|d| <= 1e-4 * (1 + |y|) + scale_tol * max|y| against the fp64 oracle.  The second term is the cancellation floor:
    the refinement sums ~600 products of O(100) activations into outputs that are O(1) at many pixels while max|y| is
    O(1000) with random-init weights; the fp32 oracle itself misses the pure relative bound (its max error is printed
    next to ours, and ours must stay within 2x of it on both paths)."""
    name, floor_mult, scale_tol = path
    err = (out.cpu().double() - ref).abs()
    bound = 1e-4 * (1 + ref.abs()) + scale_tol * ref.abs().max()
    msg = (f"refinement path={name}: max err {err.max().item():.3e} (|ref| max {ref.abs().max().item():.1f}, "
           f"fp32 oracle max err {fp32_floor})")
    print(msg)
    assert (err <= bound).all(), msg
    if fp32_floor is not None:
        assert err.max().item() <= floor_mult * fp32_floor + 1e-5, msg


@pytest.mark.parametrize("B,H,W,random_bn", [(1, 64, 128, True), (2, 40, 72, True), (1, 368, 1232, False), (1, 24, 30, True)])
def test_refinement_vs_fp64_oracle(B, H, W, random_bn, refine_path):
    from oracle import lwsnet_torch as O
    from util import product_from_oracle
    o64 = O.build_oracle(seed=0, random_bn=random_bn, dtype=torch.float64)
    model = product_from_oracle(O.build_oracle(seed=0, random_bn=random_bn))
    left = rnd(31, B, 3, H, W)
    pred3 = rnd(32, B, 1, H, W, scale=10.0) + 20.0
    out = model._refine(left.cuda(), pred3.cuda())
    with torch.no_grad():
        ref = o64.refine(left.double(), pred3.double())
        floor = (O.build_oracle(seed=0, random_bn=random_bn).refine(left, pred3).double() - ref).abs().max().item()
    _check_refine(out, ref, floor, refine_path)


def test_refinement_golden(refine_path):
    from oracle import lwsnet_torch as O
    from util import product_from_oracle
    g = golden("model_small")
    model = product_from_oracle(O.build_oracle(seed=0, random_bn=True))
    out = model._refine(cu(g["left"]), cu(g["pred3"]))
    ref = torch.from_numpy(g["refine64"])
    _check_refine(out, ref, None, refine_path)


@pytest.mark.parametrize("B,H,W", [(1, 40, 150), (2, 24, 30), (1, 368, 1232)])
def test_refinement_block_clp_vs_fp64(B, H, W):
    """One BN-ReLU-DW(dil 2)-PW block (block 1 of refinement1_left) through its own C-ABI entry, on the internal channels-last
    bordered layout, against fp64 torch convs on the same folded weights: |d| <= 2e-5 * (1 + |y|) (fp32 depthwise FMA chain,
    then the split-fp16 tensor-core pointwise: 22-bit operands, exact products, fp32 accumulation), border pixels exactly 0."""
    from oracle import lwsnet_torch as O
    from util import product_from_oracle
    import torch.nn.functional as F
    model = product_from_oracle(O.build_oracle(seed=0, random_bn=True))
    packed = model._refinement_packed(torch.device("cuda"))
    pk = packed.cpu()
    dw = pk[896:896 + 288].reshape(32, 1, 3, 3).double()                       # block 1 of R1_left: depthwise [32][9]
    pw = pk[896 + 288:896 + 288 + 1024].reshape(32, 32).double()               # folded pointwise [ci][co]
    bias = pk[896 + 288 + 1024:896 + 288 + 1024 + 32].double()
    x = rnd(51, B, 32, H, W, scale=3.0).abs()                                  # post-ReLU activations
    clp = torch.zeros(B, H + 32, W + 32, 32)
    clp[:, 16:16 + H, 16:16 + W, :] = x.permute(0, 2, 3, 1)
    out = ops().refinement_block_clp(clp.cuda().reshape(-1), packed, 0, 0, B, H, W).cpu().reshape(B, H + 32, W + 32, 32)
    d = F.conv2d(x.double(), dw, padding=2, dilation=2, groups=32)
    ref = torch.relu(torch.einsum("bihw,io->bohw", d, pw) + bias[None, :, None, None])
    got = out[:, 16:16 + H, 16:16 + W, :].permute(0, 3, 1, 2).double()
    err = (got - ref).abs()
    assert (err <= 2e-5 * (1 + ref.abs())).all(), f"max err {err.max().item():.3e}"
    border = out.clone()
    border[:, 16:16 + H, 16:16 + W, :] = 0
    assert border.abs().max().item() == 0.0


def _clp_input(seed, B, H, W, scale=3.0):
    x = rnd(seed, B, 32, H, W, scale=scale).abs()
    clp = torch.zeros(B, H + 32, W + 32, 32)
    clp[:, 16:16 + H, 16:16 + W, :] = x.permute(0, 2, 3, 1)
    return clp.cuda().reshape(-1)


@pytest.mark.parametrize("branch,block0,nblk,B,H,W", [
    (0, 0, 2, 2, 40, 150),     # dil 2,4: 16-row bands, ragged width
    (0, 2, 2, 1, 72, 130),     # dil 8,16: 64-row bands, last band partial (8 rows)
    (2, 0, 2, 2, 50, 260),     # refinement2 dil 8,4 (descending), 32-row bands
    (2, 2, 2, 3, 33, 70),      # dil 2,1, last block without ReLU
    (1, 0, 4, 2, 100, 200),    # the whole branch as one chain of four
    (2, 0, 4, 1, 368, 1232),   # KITTI size
    (0, 0, 2, 1, 5, 9),        # smaller than one band / one tile
])
def test_refinement_chain_bit_identical_to_blocks(branch, block0, nblk, B, H, W):
    """dwsep_chain.cu: blocks kept in L2-resident rings inside one launch == the same blocks launched one by one, BITWISE (the
    per-pixel arithmetic is identical; only the schedule and where the intermediate rows live differ).  Also checks the kernel's
    watchdog flag (a dependency wait that would have hung sets it) and that the zero border survives."""
    from oracle import lwsnet_torch as O
    from util import product_from_oracle
    model = product_from_oracle(O.build_oracle(seed=0, random_bn=True))
    packed = model._refinement_packed(torch.device("cuda"))
    x = _clp_input(61, B, H, W)
    ref = x
    for j in range(block0, block0 + nblk):
        ref = ops().refinement_block_clp(ref, packed, branch, j, B, H, W)
    for sep in (160, 8):  # default queue distance, and a tiny one that forces consumers to wait for their producers
        with ops().options(chain_sep_items=sep):
            out, ws = ops().refinement_chain_clp(x, packed, branch, block0, nblk, B, H, W, return_ws=True)
            torch.cuda.synchronize()
            ctrl = ws[:8].view(torch.int32).cpu()
        assert int(ctrl[1]) == 0, "chain kernel watchdog fired (a dependency wait timed out)"
        assert torch.equal(out, ref), f"sep={sep}: max diff {(out - ref).abs().max().item():.3e}"
    o4 = out.reshape(B, H + 32, W + 32, 32).clone()
    o4[:, 16:16 + H, 16:16 + W, :] = 0
    assert o4.abs().max().item() == 0.0


def test_refinement_chain_option_is_bit_identical():
    """lws_refinement_f32 with chains of 2 / 4 blocks == one block per launch, bitwise, at a shape large enough to use chains."""
    from oracle import lwsnet_torch as O
    from util import product_from_oracle
    model = product_from_oracle(O.build_oracle(seed=0, random_bn=True))
    left = rnd(71, 2, 3, 200, 328).cuda()
    pred3 = (rnd(72, 2, 1, 200, 328, scale=10.0) + 20.0).cuda()
    with ops().options(refine_chain=0):
        ref = model._refine(left, pred3).clone()
    for n in (2, 4):
        with ops().options(refine_chain=n, chain_min_bands=0):
            out = model._refine(left, pred3)
            torch.cuda.synchronize()
        assert torch.equal(out, ref), n


def test_refinement_segment_schedule_is_bit_identical():
    """The dense 64 -> 32 conv and the closing 32 -> 1 conv of the refinement with their strips cut into segments dealt round-robin
    (option tz_strips = 2) == contiguous tile ranges, bitwise."""
    from oracle import lwsnet_torch as O
    from util import product_from_oracle
    model = product_from_oracle(O.build_oracle(seed=0, random_bn=True))
    for (B, H, W) in ((2, 200, 328), (1, 368, 1232), (3, 40, 72)):
        left = rnd(73, B, 3, H, W).cuda()
        pred3 = (rnd(74, B, 1, H, W, scale=10.0) + 20.0).cuda()
        with ops().options(tz_strips=0):
            ref = model._refine(left, pred3).clone()
        with ops().options(tz_strips=2):
            out = model._refine(left, pred3)
            torch.cuda.synchronize()
        assert torch.equal(out, ref), (B, H, W)


# ------------------------------------------------------------------------------------------------ n1 feature pyramid
@pytest.mark.parametrize("B,H,W", [(2, 64, 128), (1, 368, 1232), (1, 40, 72), (3, 24, 160)])
def test_feature_extraction_tma_tiles_match_per_thread_loads(B, H, W):
    """The TMA-tile kernels of the feature pyramid (option fe_tma = 1, default: stride-1 convs and the 1/4 -> 1/2 transposed conv on
    maps with 16-byte aligned rows) accumulate in the same (ci, ky, kx) order as the per-thread-load kernels they replace: identical
    bits for all three feature maps, including shapes where only some layers qualify."""
    from oracle import lwsnet_torch as O
    from util import product_from_oracle
    model = product_from_oracle(O.build_oracle(seed=0, random_bn=True))
    img = rnd(45, B, 3, H, W).cuda()
    with ops().options(fe_tma=0):
        ref = [t.clone() for t in model.feature_extraction(img)]
    with ops().options(fe_tma=1):
        out = model.feature_extraction(img)
    for o, r in zip(out, ref):
        assert torch.equal(o, r)


@pytest.mark.parametrize("B,H,W,random_bn", [(1, 64, 128, True), (2, 40, 72, True), (1, 368, 1232, False), (1, 24, 40, True)])
def test_feature_extraction_vs_fp64_oracle(B, H, W, random_bn):
    """|d| <= 1e-4 * (1 + |y|) + 2e-6 * max|y| against the fp64 oracle (reference models/submodules.py:176-188)."""
    from oracle import lwsnet_torch as O
    from util import product_from_oracle
    o64 = O.build_oracle(seed=0, random_bn=random_bn, dtype=torch.float64)
    model = product_from_oracle(O.build_oracle(seed=0, random_bn=random_bn))
    img = rnd(41, B, 3, H, W)
    out = model.feature_extraction(img.cuda())
    with torch.no_grad():
        ref = o64.feature_extraction(img.double())
    shapes = [(B, 16, H // 8, W // 8), (B, 16, H // 4, W // 4), (B, 8, H // 2, W // 2)]
    for o, r, shp in zip(out, ref, shapes):
        assert tuple(o.shape) == shp
        err = (o.cpu().double() - r).abs()
        assert (err <= 1e-4 * (1 + r.abs()) + 2e-6 * r.abs().max()).all(), f"max err {err.max().item():.3e}"
    # same graph through torch/cuDNN fp32 on the GPU: an independent second opinion
    for o, t in zip(out, model.feature_extraction.forward_torch(img.cuda())):
        assert (o - t).abs().max().item() <= 1e-4 * (1 + t.abs().max().item())


# ------------------------------------------------------------------------------------------------ error behaviour
def test_cpu_tensors_are_refused():
    from lwsnet_b200._lib import LwsError
    with pytest.raises(LwsError):
        ops().cost_volume_l1(torch.zeros(1, 4, 4, 8), torch.zeros(1, 4, 4, 8), 4)


# ------------------------------------------------------------------------------------------------ n2 pre / post (bit-exact)
@pytest.mark.parametrize("B,h,w,th,tw", [(2, 375, 1242, 368, 1232), (1, 20, 33, 16, 24), (3, 16, 24, 16, 24)])
def test_preprocess_u8_bit_exact(B, h, w, th, tw):
    """inference.py:93-103 (crop, BGR->RGB, ToTensor, Normalize): byte-identical to the CPU oracle."""
    from oracle import lwsnet_torch as O
    rng = np.random.default_rng(7)
    img = rng.integers(0, 256, (B, h, w, 3), dtype=np.uint8)
    ref = torch.cat([O.preprocess_bgr_uint8(img[b], th, tw) for b in range(B)])
    out = ops().preprocess_bgr_u8(cu(img), th, tw)
    assert torch.equal(out.cpu(), ref)
    with pytest.raises(ValueError):  # the reference skips images smaller than the crop (inference.py:96-97)
        ops().preprocess_bgr_u8(cu(img), h + 1, tw)


def test_disparity_to_u8_and_jet_bit_exact():
    """inference.py:114-115: astype(uint8) + applyColorMap(JET), against numpy and the committed cv2 LUT fixture."""
    import os
    from util import GOLDEN
    lut = np.load(os.path.join(GOLDEN, "jet_lut_bgr.npy"))
    g = torch.Generator().manual_seed(11)
    disp = torch.rand(2, 4, 37, 129, generator=g) * 300.0 - 20.0  # includes negatives and > 255 (wrap modulo 256)
    disp[0, 0, 0, :6] = torch.tensor([0.0, 0.999, 1.0, 255.0, 255.999, 256.0])
    gray, color = ops().disparity_to_u8(disp.cuda())
    ref = disp.numpy().astype(np.int64).astype(np.uint8)  # C cast: truncate toward zero, then wrap
    assert np.array_equal(gray.cpu().numpy(), ref)
    assert np.array_equal(color.cpu().numpy(), lut[ref])
    g2, c2 = ops().disparity_to_u8(disp.cuda(), gray=True, color=False)
    assert c2 is None and np.array_equal(g2.cpu().numpy(), ref)


@pytest.mark.parametrize("D,H,W", [(9, 16, 32), (9, 32, 64), (4, 6, 9), (7, 10, 20)])
def test_conv3d_c8_stack_unique_and_batch_independent(D, H, W):
    """Regression for the plane-group kernel: a 126-voxel tile of a narrow volume wraps over several lines and, at the end of a
    plane, into the next plane, whose own tile accumulates its kd taps in another order.  Exactly one tile may store each voxel:
    re-running must give identical bits, and so must splitting the batch."""
    from oracle import lwsnet_torch as O
    from lwsnet_b200.submodules import post_3dconvs
    onet = O.post_3dconvs(4, 8)
    holder = torch.nn.Module()
    holder.net = onet
    O.kaiming_normal_init_(holder, 5)
    O.randomize_bn_(holder, 6)
    net = post_3dconvs(4, 8)
    net.load_state_dict(onet.state_dict(), strict=True)
    net = net.cuda()
    x = (rnd(7, 4, D, H, W, scale=6.0).abs()).cuda()
    a = net.run(x, add_skip=True)
    for _ in range(3):
        assert torch.equal(a, net.run(x, add_skip=True))
    parts = torch.cat([net.run(x[:1].contiguous(), add_skip=True), net.run(x[1:].contiguous(), add_skip=True)])
    assert torch.equal(a, parts)
