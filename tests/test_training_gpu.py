"""n4: backward of K1 / K2 / K4 and the multi-stage smooth-L1 loss of train.py:127-166, against fp64 torch autograd
of the oracle's formulation (oracle/lwsnet_torch.py restates models/models.py op by op, so its autograd graph is the reference's)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from util import rnd

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return (a.double().cpu() - b.double()).abs().max().item() / max(1.0, b.double().abs().max().item())


def _frac_bad(a, b, tol):
    """Fraction of elements further than tol * max|b| from the reference.  |x| is not differentiable at 0: where a sample lands
    within the fp32 coordinate noise (6e-5) of the left feature, fp32 and fp64 legitimately pick different
    signs, which moves a few isolated gradient entries by 2 * g."""
    d = (a.double().cpu() - b.double()).abs()
    return (d > tol * max(1.0, b.double().abs().max().item())).double().mean().item()


@pytest.mark.parametrize("B,C,H,W,D,stride", [(2, 16, 12, 40, 24, 1), (1, 8, 6, 33, 12, 2), (1, 3, 4, 9, 7, 1), (1, 16, 46, 154, 24, 1)])
def test_cost_volume_backward(B, C, H, W, D, stride):
    from oracle import lwsnet_torch as O
    from lwsnet_b200 import training as T
    L, R = rnd(1, B, C, H, W, scale=2.0), rnd(2, B, C, H, W, scale=2.0)
    g = rnd(3, B, D // stride, H, W)
    Lr, Rr = L.double().requires_grad_(), R.double().requires_grad_()
    O.build_volume_2d(Lr, Rr, D, stride=stride).backward(g.double())
    Lc, Rc = L.cuda().requires_grad_(), R.cuda().requires_grad_()
    out = T.cost_volume_l1(Lc, Rc, D, stride)
    out.backward(g.cuda())
    assert _rel(Lc.grad, Lr.grad) <= 1e-5 and _rel(Rc.grad, Rr.grad) <= 1e-5


@pytest.mark.parametrize("B,C,H,W,m", [(2, 16, 12, 40, 5), (1, 8, 10, 33, 5), (1, 4, 9, 17, 3), (1, 8, 46, 77, 5)])
def test_warp_residual_volume_backward(B, C, H, W, m):
    """Gradients w.r.t. the left / right features and the disparity.  Disparities are kept 0.2 px away from integer sampling
    abscissae, where the bilinear sample is not differentiable and fp32 / fp64 may pick different taps."""
    from oracle import lwsnet_torch as O
    from lwsnet_b200 import training as T
    L, R = rnd(4, B, C, H, W, scale=2.0), rnd(5, B, C, H, W, scale=2.0)
    frac = torch.rand(B, 1, H, W, generator=torch.Generator().manual_seed(6)) * 0.6 + 0.2
    disp = torch.floor(rnd(7, B, 1, H, W, scale=W / 8.0) + W / 6.0) + frac
    g = rnd(8, B, 2 * m - 1, H, W)
    Lr, Rr, dr = L.double().requires_grad_(), R.double().requires_grad_(), disp.double().requires_grad_()
    O.build_volume_2d3(Lr, Rr, m, dr).backward(g.double())
    Lc, Rc, dc = L.cuda().requires_grad_(), R.cuda().requires_grad_(), disp.cuda().requires_grad_()
    T.warp_residual_volume_l1(Lc, Rc, dc, m).backward(g.cuda())
    # all but a handful of isolated entries (sign ties of |l - s|, see _frac_bad) within 2e-5 of the largest gradient
    assert _frac_bad(Lc.grad, Lr.grad, 2e-5) <= 2e-3
    assert _frac_bad(Rc.grad, Rr.grad, 2e-5) <= 2e-3
    assert _frac_bad(dc.grad, dr.grad, 2e-5) <= 2e-3


@pytest.mark.parametrize("B,D,H,W,start", [(2, 24, 12, 40, 0.0), (1, 9, 7, 9, -4.0), (1, 48, 5, 6, 0.0)])
def test_softmax_regression_backward(B, D, H, W, start):
    from lwsnet_b200 import training as T
    for scale in (1.0, 10.0):
        c = rnd(9, B, D, H, W, scale=scale)
        g = rnd(10, B, 1, H, W)
        cr = c.double().requires_grad_()
        v = torch.arange(D, dtype=torch.float64).view(1, D, 1, 1) + start
        (F.softmax(-cr, 1) * v).sum(1, keepdim=True).backward(g.double())
        cc = c.cuda().requires_grad_()
        T.softmax_regression(cc, start).backward(g.cuda())
        assert (cc.grad.double().cpu() - cr.grad).abs().max().item() <= 2e-5 * max(1.0, cr.grad.abs().max().item())


def test_multistage_smooth_l1_loss_matches_train_py():
    """train.py:137-155 restated with torch: mask = gt < maxdisp, F.smooth_l1_loss(pred[mask], gt[mask], mean) * weight, summed."""
    from lwsnet_b200 import training as T
    B, H, W = 2, 37, 91
    gt = torch.rand(B, H, W, generator=torch.Generator().manual_seed(11)) * 250.0  # some pixels >= maxdisp = 192: masked out
    preds = [gt.unsqueeze(1) + rnd(12 + s, B, 1, H, W, scale=2.0 ** s) for s in range(4)]
    weights = [0.25, 0.5, 1.0, 1.0]
    pr = [p.double().requires_grad_() for p in preds]
    mask = gt < 192.0
    ref = [w * F.smooth_l1_loss(p.squeeze(1)[mask], gt.double()[mask], reduction="mean") for w, p in zip(weights, pr)]
    sum(ref).backward()
    pc = [p.cuda().requires_grad_() for p in preds]
    losses, count = T.multistage_smooth_l1_loss(pc, gt.cuda(), 192.0, weights)
    assert int(count.item()) == int(mask.sum())
    for s in range(4):
        assert abs(losses[s].item() - ref[s].item()) <= 1e-5 * max(1.0, abs(ref[s].item()))
    losses.sum().backward()
    for s in range(4):
        assert (pc[s].grad.double().cpu() - pr[s].grad).abs().max().item() <= 1e-6 * max(1.0, pr[s].grad.abs().max().item())
    # deterministic: identical bits on a second evaluation; empty mask -> zero loss, zero gradient, no NaN
    l2, _ = T.multistage_smooth_l1_loss([p.detach() for p in pc], gt.cuda(), 192.0, weights)
    assert torch.equal(l2, losses.detach())
    pe = [p.detach().clone().requires_grad_() for p in pc]
    le, ce = T.multistage_smooth_l1_loss(pe, gt.cuda() + 1000.0, 192.0, weights)
    le.sum().backward()
    assert ce.item() == 0 and le.abs().max().item() == 0.0 and all(p.grad.abs().max().item() == 0.0 for p in pe)


def test_differentiable_stage_skeleton_end_to_end():
    """A stage-2-like chain through all three differentiable kernels and the loss: cost volume -> regression -> (as disparity) ->
    warp + residual volume -> regression -> loss; gradients w.r.t. the feature maps against fp64 autograd of the oracle's functions."""
    from oracle import lwsnet_torch as O
    from lwsnet_b200 import training as T
    B, C, H, W = 1, 8, 10, 36
    L, R = rnd(20, B, C, H, W, scale=1.5), rnd(21, B, C, H, W, scale=1.5)
    gt = torch.rand(B, H, W, generator=torch.Generator().manual_seed(22)) * 12.0

    def chain(cv, wrv, reg, loss, L, R, gt):
        d0 = reg(cv(L, R, 12) * 0.35, 0.0) + 0.37          # soft volume: a smooth, non-integer disparity
        d1 = d0 + reg(wrv(L, R, d0, 3) * 0.35, -2.0)
        return loss([d0, d1], gt)

    def reg64(c, start):
        v = torch.arange(c.shape[1], dtype=c.dtype).view(1, -1, 1, 1) + start
        return (F.softmax(-c, 1) * v).sum(1, keepdim=True)

    def loss64(ps, gt):
        m = gt < 192.0
        return sum(w * F.smooth_l1_loss(p.squeeze(1)[m], gt[m], reduction="mean") for w, p in zip((0.5, 1.0), ps))

    Lr, Rr = L.double().requires_grad_(), R.double().requires_grad_()
    ref = chain(lambda a, b, D: O.build_volume_2d(a, b, D), lambda a, b, d, m: O.build_volume_2d3(a, b, m, d), reg64, loss64, Lr, Rr,
                gt.double())
    ref.backward()
    Lc, Rc = L.cuda().requires_grad_(), R.cuda().requires_grad_()
    out = chain(T.cost_volume_l1, T.warp_residual_volume_l1, T.softmax_regression,
                lambda ps, g: T.multistage_smooth_l1_loss(ps, g, 192.0, (0.5, 1.0))[0].sum(), Lc, Rc, gt.cuda())
    out.backward()
    assert abs(out.item() - ref.item()) <= 1e-4 * max(1.0, abs(ref.item()))
    for a, b in ((Lc.grad, Lr.grad), (Rc.grad, Rr.grad)):
        assert (a.double().cpu() - b).abs().max().item() <= 2e-3 * b.abs().max().item() + 1e-6
