"""GPU tests of the drop-in boundary beyond the fused model path: the reference's layers called as layers
(models/models.py:158-160, :167-179), arbitrary 3D-stack widths (models/models.py:19-22), and the engine / cache hygiene around
them (weight updates under captured CUDA graphs, scratch memory per stream)."""
import numpy as np
import pytest
import torch

from util import err_stats, product_from_oracle, rnd

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pair():
    from oracle import lwsnet_torch as O
    o64 = O.build_oracle(seed=0, random_bn=True, dtype=torch.float64)
    o32 = O.build_oracle(seed=0, random_bn=True)
    return O, o32, o64, product_from_oracle(o32)


@pytest.mark.parametrize("B,H,W", [(1, 64, 128), (2, 40, 72), (1, 33, 50)])
def test_refinement_layers_called_as_layers(pair, B, H, W):
    """refinement1_left(x), refinement1_disp(x), refinement2(x) as the reference calls them (models/models.py:158-160), each against
    the fp64 oracle's module, and their composition pred3 + refinement2(concat(...)) against the fused model path."""
    O, o32, o64, prod = pair
    left = rnd(81, B, 3, H, W)
    pred3 = rnd(82, B, 1, H, W, scale=10.0) + 20.0
    with torch.no_grad():
        a64 = o64.refinement1_left(left.double())
        b64 = o64.refinement1_disp(pred3.double())
        a32, b32 = o32.refinement1_left(left), o32.refinement1_disp(pred3)
    a = prod.refinement1_left(left.cuda())
    b = prod.refinement1_disp(pred3.cuda())
    for got, r64, r32 in ((a, a64, a32), (b, b64, b32)):
        assert tuple(got.shape) == (B, 32, H, W)
        err = (got.cpu().double() - r64).abs()
        floor = (r32.double() - r64).abs().max().item()
        assert (err <= 1e-4 * (1 + r64.abs()) + 2e-6 * r64.abs().max()).all(), err.max().item()
        assert err.max().item() <= 3 * floor + 1e-5
    cat64 = torch.cat([a64, b64], 1)
    with torch.no_grad():
        r64 = o64.refinement2(cat64)
        r32 = o32.refinement2(cat64.float())
    r = prod.refinement2(cat64.float().cuda())
    assert tuple(r.shape) == (B, 1, H, W)
    err = (r.cpu().double() - r64).abs()
    floor = (r32.double() - r64).abs().max().item()
    assert (err <= 1e-4 * (1 + r64.abs()) + 2e-6 * r64.abs().max()).all(), err.max().item()
    assert err.max().item() <= 3 * floor + 1e-5
    # the reference's composition == the fused path (both within the fp32 floor of the fp64 oracle)
    composed = pred3.cuda() + prod.refinement2(torch.cat([a, b], 1))
    fused = prod._refine(left.cuda(), pred3.cuda())
    with torch.no_grad():
        ref = o64.refine(left.double(), pred3.double())
        floor = (o32.refine(left, pred3).double() - ref).abs().max().item()
    for got in (composed, fused):
        assert (got.cpu().double() - ref).abs().max().item() <= 6 * floor + 1e-5


def test_refinement_layer_argument_errors(pair):
    from lwsnet_b200._lib import LwsError
    from lwsnet_b200.submodules import refinement1, refinement2
    O, o32, o64, prod = pair
    with pytest.raises(ValueError):
        prod.refinement1_left(torch.zeros(1, 1, 16, 16, device="cuda"))
    with pytest.raises(LwsError):
        prod.refinement2(torch.zeros(1, 64, 16, 16))  # CPU tensor
    with pytest.raises(LwsError):
        refinement1(5, 32)
    with pytest.raises(LwsError):
        refinement2(32, 32)


def test_disparity_regression_class_is_a_plain_weighted_sum():
    """models/models.py:167-179: out = sum_j input_j * disp_j, input NOT renormalised; exact zeros and negative entries allowed."""
    from lwsnet_b200 import disparity_regression
    g = torch.Generator().manual_seed(3)
    p = torch.softmax(torch.randn(2, 9, 13, 37, generator=g) * 4, 1)
    p[0, 3] = 0.0           # exact zeros (log-based evaluation would give -inf)
    p[1, :, 0, 0] = 0.0     # an all-zero column -> 0, not NaN
    p[1, 2, 1, 1] = -0.25   # not a probability: still a plain weighted sum
    for start, end, stride in ((-4, 5, 1), (0, 9, 1), (0, 9, 2)):
        reg = disparity_regression(start, end, stride)
        out = reg(p.cuda())
        disp = torch.arange(start * stride, end * stride, stride, dtype=torch.float64).view(1, 9, 1, 1)
        ref = (p.double() * disp).sum(1, keepdim=True)
        assert tuple(out.shape) == (2, 1, 13, 37)
        assert torch.isfinite(out).all()
        assert (out.cpu().double() - ref).abs().max().item() <= 2e-5
    assert reg(p.cuda())[1, 0, 0, 0].item() == 0.0
    with pytest.raises(ValueError):
        disparity_regression(0, 24)(p.cuda())
    # fused op == class applied to the softmax (the reference's two steps, models/models.py:142,151-152)
    from lwsnet_b200 import ops
    cost = torch.randn(2, 9, 13, 37, generator=g) * 6
    fused = ops.softmax_regression(cost.cuda(), -4.0)
    two_step = disparity_regression(-4, 5)(torch.softmax(-cost, 1).cuda())
    assert (fused - two_step).abs().max().item() <= 2e-5


@pytest.mark.parametrize("C,B,D,H,W", [(24, 1, 9, 20, 40), (4, 2, 5, 11, 30), (40, 1, 3, 9, 17), (12, 1, 9, 16, 24)])
def test_conv3d_stack_any_width(C, B, D, H, W):
    """models/models.py:19-22 takes any channels_3d * growth_rate: widths without a tensor-core kernel run on the FFMA kernel in
    groups of 8 output channels (zero-padded), |d| <= 1e-4 * (1 + |y|) against the fp64 oracle."""
    from oracle import lwsnet_torch as O
    from lwsnet_b200.submodules import post_3dconvs
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
    err = (out.cpu().double() - ref).abs()
    assert (err <= 1e-4 * (1 + ref.abs())).all(), f"max err {err.max().item():.3e}"


def test_engine_drops_graphs_when_weights_change(pair):
    """captured graphs bake in the packed-weight addresses.  After load_state_dict / set_state_dict / repack the engine
    must re-capture: the replayed result has to follow the new weights (and never read the freed old blob)."""
    from oracle import lwsnet_torch as O
    from lwsnet_b200.runner import StereoEngine
    _, o32, o64, _ = pair
    prod = product_from_oracle(o32)
    eng = StereoEngine(prod, micro_batch=2)
    left, right = O.synthetic_pair(2, 64, 128, seed=5, max_disp=20.0)
    left, right = left.cuda(), right.cuda()
    out_a = eng.infer_device(left, right).clone()
    other = O.build_oracle(seed=7, random_bn=True)
    prod.load_state_dict({k: v.float() for k, v in other.state_dict().items()}, strict=True)
    out_b = eng.infer_device(left, right).clone()
    fresh = torch.cat(product_from_oracle(other)(left, right), dim=1)
    assert torch.equal(out_b, fresh)
    assert not torch.equal(out_a, out_b)
    # weight surgery through .data does not bump torch's version counter: repack() is the documented way
    with torch.no_grad():
        for p in prod.refinement2.parameters():
            p.data.mul_(0.5)
    prod.repack()
    out_c = eng.infer_device(left, right).clone()
    assert torch.equal(out_c, torch.cat(prod(left, right), dim=1))
    assert not torch.equal(out_c[:, 3], out_b[:, 3])


def test_workspaces_are_per_stream(pair):
    """scratch buffers are keyed by (device, stream): two streams of one device never share conv / refinement scratch."""
    from lwsnet_b200 import ops
    O, o32, o64, prod = pair
    left, right = O.synthetic_pair(1, 64, 128, seed=9, max_disp=20.0)
    left, right = left.cuda(), right.cuda()
    ref = [t.clone() for t in prod(left, right)]
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    outs = {}
    for rep in range(3):
        for s in (s1, s2):
            with torch.cuda.stream(s):
                outs[s] = prod(left, right)
    torch.cuda.synchronize()
    for s in (s1, s2):
        for a, b in zip(outs[s], ref):
            assert torch.equal(a, b)
    keys = [k for k in ops._workspaces if k[2] == "refine"]
    assert len({k[1] for k in keys}) >= 3  # default stream + the two side streams


def test_split_fp16_range_is_guarded(pair):
    """Split-fp16 operands carry activations * 2^-6 in fp16.  Defined behaviour beyond +-4.19e6: the conversions SATURATE
    (F2FP.SATFINITE), so the tensor-core path never turns finite inputs into inf / NaN; the exact-fp32 kernels (options
    conv3d_tc = 0 / refine_tc = 0) take the whole fp32 range and stay accurate."""
    from lwsnet_b200 import ops
    O, o32, o64, prod = pair
    left = rnd(91, 1, 3, 32, 64)
    big = (rnd(92, 1, 1, 32, 64).abs() * 1e3 + 1e3)  # 1e3-px "disparities": activations reach ~1e5-1e6, inside the fp16 operand range
    with torch.no_grad():
        ref = o64.refine(left.double(), big.double())
        floor = (o32.refine(left, big).double() - ref).abs().max().item()
    out = prod._refine(left.cuda(), big.cuda())
    assert torch.isfinite(out).all()
    assert (out.cpu().double() - ref).abs().max().item() <= 6 * floor + 1e-5
    huge = big * 1e6  # 1e9: far beyond the fp16 operand range
    out_sat = prod._refine(left.cuda(), huge.cuda())
    assert torch.isfinite(out_sat).all(), "tensor-core path must saturate, not overflow to inf / NaN"
    with ops.options(refine_tc=0):
        out_exact = prod._refine(left.cuda(), huge.cuda())
        torch.cuda.synchronize()
    with torch.no_grad():
        ref = o64.refine(left.double(), huge.double())
        floor = (o32.refine(left, huge).double() - ref).abs().max().item()
    assert torch.isfinite(out_exact).all()
    assert (out_exact.cpu().double() - ref).abs().max().item() <= 3 * floor + 1e-5 * ref.abs().max().item()
    # same for the 3D stacks: a cost volume of 1e9 through the C = 8 tensor-core stack stays finite; the FFMA path is exact-fp32
    cost = torch.full((1, 9, 16, 32), 1e9, device="cuda")
    st = prod.volume_postprocess[1]
    assert torch.isfinite(st.run(cost, add_skip=True)).all()
    with ops.options(conv3d_tc=0):
        assert torch.isfinite(st.run(cost, add_skip=True)).all()
