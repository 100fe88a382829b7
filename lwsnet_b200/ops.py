"""Torch-tensor front end of the lws_b200 C ABI (include/lws.h).

torch is used for device memory, streams and nothing else: every function below validates its arguments, allocates
the output with torch and enqueues exactly the C-ABI call on torch's current stream.  CPU tensors are rejected —
there is no CPU or PyTorch fallback for any hot-path op.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence

import torch

from ._lib import LwsError, check, get_option, lib, options, set_option  # noqa: F401  (options API re-exported)

LAUNCHES = [0]  # kernels of liblws_b200 enqueued through this module (bench.py reports it as gpu_launches)
_workspaces: dict = {}
_retired: dict = {}  # outgrown buffers stay alive while a captured CUDA graph may still hold their addresses (see release_retired)


def _ptr(t: Optional[torch.Tensor], name: str, dtype=torch.float32):
    if t is None:
        return None
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise LwsError(f"{name}: lwsnet_b200 ops run on CUDA tensors only (got device {t.device}); there is no CPU path")
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name}: tensor must be contiguous")
    return ctypes.c_void_p(t.data_ptr())


def _stream(t: torch.Tensor):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"expected a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise LwsError(f"lwsnet_b200 ops run on CUDA tensors only (got device {t.device}); there is no CPU path")
    if t.dtype != torch.float32:
        raise TypeError(f"expected float32, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def workspace(device: torch.device, tag: str, nbytes: int) -> torch.Tensor:
    """Grow-only scratch buffer per (device, STREAM, tag): calls on one stream are ordered, so they may share scratch memory;
    two engines / threads working on different streams of one device get different buffers and cannot race."""
    stream = torch.cuda.current_stream(device).cuda_stream
    key = (device.index, stream, tag)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        if buf is not None:
            _retired.setdefault((device.index, stream), []).append(buf)
        buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf


def release_retired(device: Optional[torch.device] = None) -> None:
    """Free outgrown scratch buffers.  Call when the CUDA graphs that captured them have been dropped (StereoEngine does)."""
    for key in list(_retired):
        if device is None or key[0] == device.index:
            del _retired[key]


def release_workspaces() -> None:
    """Drop all scratch buffers (only safe once no captured CUDA graph that used them will be replayed)."""
    _workspaces.clear()
    _retired.clear()


# ------------------------------------------------------------------------------------------------ a1
def cost_volume_l1(feat_l: torch.Tensor, feat_r: torch.Tensor, maxdisp: int, stride: int = 1) -> torch.Tensor:
    """LWSNet._build_volume_2d (reference models/models.py:58-76)."""
    assert maxdisp % stride == 0  # the reference's own assert (models.py:63)
    feat_l, feat_r = _f32c(feat_l), _f32c(feat_r)
    if feat_l.shape != feat_r.shape or feat_l.dim() != 4:
        raise ValueError(f"feat_l/feat_r must be equal-shape [B,C,H,W], got {tuple(feat_l.shape)} {tuple(feat_r.shape)}")
    B, C, H, W = feat_l.shape
    cost = torch.empty((B, maxdisp // stride, H, W), dtype=torch.float32, device=feat_l.device)
    if cost.numel() == 0:
        return cost
    with torch.cuda.device(feat_l.device):
        check(lib.lws_cost_volume_l1_f32(_ptr(feat_l, "feat_l"), _ptr(feat_r, "feat_r"), _ptr(cost, "cost"), B, C, H, W,
                                         maxdisp, stride, _stream(feat_l)), "lws_cost_volume_l1_f32")
    LAUNCHES[0] += 1
    return cost


# ------------------------------------------------------------------------------------------------ a2
def disp_to_scale(pred_full: torch.Tensor, h: int, w: int) -> torch.Tensor:
    """wflow of LWSNet.forward (reference models/models.py:119-121)."""
    pred_full = _f32c(pred_full)
    B, one, H, W = pred_full.shape
    if one != 1:
        raise ValueError("pred_full must be [B,1,H,W]")
    out = torch.empty((B, 1, h, w), dtype=torch.float32, device=pred_full.device)
    with torch.cuda.device(pred_full.device):
        check(lib.lws_disp_to_scale_f32(_ptr(pred_full, "pred_full"), _ptr(out, "wflow"), B, H, W, h, w,
                                        _stream(pred_full)), "lws_disp_to_scale_f32")
    LAUNCHES[0] += 1
    return out


# ------------------------------------------------------------------------------------------------ a3
def warp_bilinear(x: torch.Tensor, disp: torch.Tensor) -> torch.Tensor:
    """LWSNet.warp (reference models/models.py:28-55)."""
    x, disp = _f32c(x), _f32c(disp)
    N, C, H, W = x.shape
    if tuple(disp.shape) != (N, 1, H, W):
        raise ValueError(f"disp must be [N,1,H,W]={(N, 1, H, W)}, got {tuple(disp.shape)}")
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        check(lib.lws_warp_bilinear_f32(_ptr(x, "x"), _ptr(disp, "disp"), _ptr(out, "out"), N, C, H, W, _stream(x)),
              "lws_warp_bilinear_f32")
    LAUNCHES[0] += 1
    return out


def warp_taps(disp: torch.Tensor, shift: float = 0.0):
    """Test hook: tap indices / weights of warp(., disp - shift).  Returns x0 [N,H,W] i32, y0 [H] i32, wx [N,H,W,2], wy [H,2]."""
    disp = _f32c(disp)
    N, one, H, W = disp.shape
    dev = disp.device
    x0 = torch.empty((N, H, W), dtype=torch.int32, device=dev)
    y0 = torch.empty((H,), dtype=torch.int32, device=dev)
    wx = torch.empty((N, H, W, 2), dtype=torch.float32, device=dev)
    wy = torch.empty((H, 2), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib.lws_warp_taps_f32(_ptr(disp, "disp"), float(shift), _ptr(x0, "x0", torch.int32),
                                    _ptr(y0, "y0", torch.int32), _ptr(wx, "wx"), _ptr(wy, "wy"), N, H, W,
                                    _stream(disp)), "lws_warp_taps_f32")
    return x0, y0, wx, wy


# ------------------------------------------------------------------------------------------------ a4
def warp_residual_volume_l1(feat_l: torch.Tensor, feat_r: torch.Tensor, disp: torch.Tensor, maxdisp: int,
                            stride: int = 1) -> torch.Tensor:
    """LWSNet._build_volume_2d3 (reference models/models.py:78-104)."""
    feat_l, feat_r, disp = _f32c(feat_l), _f32c(feat_r), _f32c(disp)
    B, C, H, W = feat_l.shape
    if feat_r.shape != feat_l.shape or tuple(disp.shape) != (B, 1, H, W):
        raise ValueError("shape mismatch between feat_l, feat_r and disp")
    cost = torch.empty((B, 2 * maxdisp - 1, H, W), dtype=torch.float32, device=feat_l.device)
    with torch.cuda.device(feat_l.device):
        check(lib.lws_warp_residual_volume_l1_f32(_ptr(feat_l, "feat_l"), _ptr(feat_r, "feat_r"), _ptr(disp, "disp"),
                                                  _ptr(cost, "cost"), B, C, H, W, maxdisp, stride, _stream(feat_l)),
              "lws_warp_residual_volume_l1_f32")
    LAUNCHES[0] += 1
    return cost


# ------------------------------------------------------------------------------------------------ a5
def _host_f32(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to("cpu", torch.float32).contiguous()


def _ptr_array(tensors: Sequence[torch.Tensor]):
    arr = (ctypes.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr


def pack_conv3d_stack(conv_weights: Sequence[torch.Tensor], bns: Sequence[Sequence[torch.Tensor]], eps: float, C: int,
                      layers: int) -> torch.Tensor:
    """Fold BN into the convs (host).  bns[i] = (weight, bias, _mean, _variance) of the BN in front of conv i."""
    n = layers + 2
    if len(conv_weights) != n or len(bns) != n:
        raise ValueError(f"expected {n} convs and BNs")
    cw = [_host_f32(w) for w in conv_weights]
    for i, w in enumerate(cw):
        cin, cout = (1 if i == 0 else C), (1 if i == n - 1 else C)
        if tuple(w.shape) != (cout, cin, 3, 3, 3):
            raise ValueError(f"conv {i}: expected weight {(cout, cin, 3, 3, 3)}, got {tuple(w.shape)}")
    parts = [[_host_f32(b[j]) for b in bns] for j in range(4)]
    packed = torch.zeros(int(lib.lws_conv3d_stack_packed_floats(C, layers)), dtype=torch.float32)
    check(lib.lws_pack_conv3d_stack_weights(_ptr_array(cw), _ptr_array(parts[0]), _ptr_array(parts[1]),
                                            _ptr_array(parts[2]), _ptr_array(parts[3]), float(eps), C, layers,
                                            ctypes.c_void_p(packed.data_ptr())), "lws_pack_conv3d_stack_weights")
    return packed


def conv3d_stack(cost: torch.Tensor, packed: torch.Tensor, C: int, layers: int, add_skip: bool = True) -> torch.Tensor:
    """post_3dconvs (+ skip) (reference models/submodules.py:190-221, models/models.py:136-138).  cost [B,D,H,W]."""
    cost = _f32c(cost)
    B, D, H, W = cost.shape
    out = torch.empty_like(cost)
    nbytes = int(lib.lws_conv3d_stack_workspace_bytes(B, D, H, W, C, layers))
    ws = workspace(cost.device, "conv3d", nbytes)
    with torch.cuda.device(cost.device):
        check(lib.lws_conv3d_stack_f32(_ptr(cost, "cost"), _ptr(packed, "packed"), _ptr(out, "out"),
                                       ctypes.c_void_p(ws.data_ptr()), nbytes, B, D, H, W, C, layers, int(add_skip),
                                       _stream(cost)), "lws_conv3d_stack_f32")
    LAUNCHES[0] += int(lib.lws_conv3d_stack_launches(C, layers))
    return out


def cost_volume_conv3d_stack_supported(B: int, Cf: int, H: int, W: int, maxdisp: int, C: int, layers: int) -> bool:
    """True when the fused stage-1 call applies (C = 32 tensor-core path, even W, tap window fits in shared memory)."""
    return int(lib.lws_cost_volume_conv3d_stack_supported(B, Cf, H, W, maxdisp, C, layers)) == 0


def cost_volume_conv3d_stack(L: torch.Tensor, R: torch.Tensor, maxdisp: int, packed: torch.Tensor, C: int, layers: int,
                             add_skip: bool = True):
    """Stage 1 in one call: _build_volume_2d (reference models/models.py:58-76) built inside the first conv kernel of
    post_3dconvs (+ skip, models/models.py:136-138).  Returns (raw volume [B,D,H,W], stack output [B,D,H,W])."""
    L, R = _f32c(L), _f32c(R)
    B, Cf, H, W = L.shape
    D = int(maxdisp)
    cost = torch.empty((B, D, H, W), dtype=torch.float32, device=L.device)
    out = torch.empty_like(cost)
    nbytes = int(lib.lws_conv3d_stack_workspace_bytes(B, D, H, W, C, layers))
    ws = workspace(L.device, "conv3d", nbytes)
    with torch.cuda.device(L.device):
        check(lib.lws_cost_volume_conv3d_stack_f32(_ptr(L, "L"), _ptr(R, "R"), _ptr(packed, "packed"), _ptr(cost, "cost"),
                                                   _ptr(out, "out"), ctypes.c_void_p(ws.data_ptr()), nbytes, B, Cf, H, W, D, C,
                                                   layers, int(add_skip), _stream(L)), "lws_cost_volume_conv3d_stack_f32")
    LAUNCHES[0] += int(lib.lws_conv3d_stack_launches(C, layers))
    return cost, out


def conv3d_bnrelu_layer(x: torch.Tensor, w_folded: torch.Tensor, bias: torch.Tensor,
                        out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """One BN-folded C->C layer: ReLU(conv3x3x3(x, w) + bias); x [B,C,D,H,W], w_folded [C][27][C]."""
    B, C, D, H, W = x.shape
    out = out if out is not None else torch.empty_like(x)
    with torch.cuda.device(x.device):
        check(lib.lws_conv3d_bnrelu_layer_f32(_ptr(x, "x"), _ptr(w_folded, "w"), _ptr(bias, "bias"), _ptr(out, "out"),
                                              B, C, D, H, W, _stream(x)), "lws_conv3d_bnrelu_layer_f32")
    LAUNCHES[0] += 1
    return out


# ------------------------------------------------------------------------------------------------ a6 / a7
def softmax_regression(cost: torch.Tensor, start: float, step: float = 1.0) -> torch.Tensor:
    """disparity_regression(start, end)(softmax(-cost, axis=1)) (reference models/models.py:142,151-152,167-179)."""
    cost = _f32c(cost)
    B, D, H, W = cost.shape
    low = torch.empty((B, 1, H, W), dtype=torch.float32, device=cost.device)
    with torch.cuda.device(cost.device):
        check(lib.lws_softmax_regression_f32(_ptr(cost, "cost"), _ptr(low, "low"), B, D, H, W, float(start),
                                             float(step), _stream(cost)), "lws_softmax_regression_f32")
    LAUNCHES[0] += 1
    return low


def scale_upsample_add(low: torch.Tensor, prev: Optional[torch.Tensor], H: int, W: int,
                       out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """low * H / h -> bilinear upsample to (H, W) (+ prev)  (reference models/models.py:145-148,153-156)."""
    low = _f32c(low)
    B, one, h, w = low.shape
    if prev is not None:
        prev = _f32c(prev)
        if tuple(prev.shape) != (B, 1, H, W):
            raise ValueError("prev must be [B,1,H,W]")
    pred = out if out is not None else torch.empty((B, 1, H, W), dtype=torch.float32, device=low.device)
    with torch.cuda.device(low.device):
        check(lib.lws_scale_upsample_add_f32(_ptr(low, "low"), _ptr(prev, "prev"), _ptr(pred, "pred"), B, h, w, H, W,
                                             _stream(low)), "lws_scale_upsample_add_f32")
    LAUNCHES[0] += 1
    return pred


def regression_tail(cost: torch.Tensor, prev: Optional[torch.Tensor], H: int, W: int, start: float, step: float = 1.0,
                    next_hw: Optional[tuple] = None, out: Optional[torch.Tensor] = None, fused: Optional[bool] = None):
    """Tail of one stage-loop iteration (reference models/models.py:142-156) and the head of the next (:119-121):
    pred = upsample(softmax_regression(cost) * H / h) (+ prev)  and, when ``next_hw`` is given, the next stage's
    wflow = resize(pred, next_hw) * h_next / H.  Either one fused launch (lws_regression_tail_f32, where the shapes allow it) or
    the three stand-alone kernels; both give the same bits.  ``fused=None`` follows the library option "fused_tail" (default 0: the
    stand-alone kernels are faster at the engine's micro-batch).  Returns (pred [B,1,H,W], wflow [B,1,hn,wn] or None)."""
    cost = _f32c(cost)
    B, D, h, w = cost.shape
    if prev is not None:
        prev = _f32c(prev)
        if tuple(prev.shape) != (B, 1, H, W):
            raise ValueError("prev must be [B,1,H,W]")
    hn, wn = (int(next_hw[0]), int(next_hw[1])) if next_hw is not None else (0, 0)
    if fused is None:
        fused = get_option("fused_tail") != 0
    if fused and lib.lws_regression_tail_supported(h, w, H, W, hn, wn) == 0:
        pred = out if out is not None else torch.empty((B, 1, H, W), dtype=torch.float32, device=cost.device)
        wflow = torch.empty((B, 1, hn, wn), dtype=torch.float32, device=cost.device) if next_hw is not None else None
        with torch.cuda.device(cost.device):
            check(lib.lws_regression_tail_f32(_ptr(cost, "cost"), _ptr(prev, "prev"), _ptr(pred, "pred"), _ptr(wflow, "wflow"), B, D, h,
                                              w, H, W, hn, wn, float(start), float(step), _stream(cost)), "lws_regression_tail_f32")
        LAUNCHES[0] += 1
        return pred, wflow
    low = softmax_regression(cost, start, step)
    pred = scale_upsample_add(low, prev, H, W, out=out)
    return pred, (disp_to_scale(pred, hn, wn) if next_hw is not None else None)


def disparity_regression(prob: torch.Tensor, start: float, step: float = 1.0) -> torch.Tensor:
    """disparity_regression(start, end, stride).forward(prob) (reference models/models.py:167-179) on an already soft-maxed
    volume: sum_j prob[:, j] * (start + j * step), without renormalising `prob`."""
    prob = _f32c(prob)
    B, D, H, W = prob.shape
    out = torch.empty((B, 1, H, W), dtype=torch.float32, device=prob.device)
    if out.numel() == 0:
        return out
    with torch.cuda.device(prob.device):
        check(lib.lws_disparity_regression_f32(_ptr(prob, "prob"), _ptr(out, "out"), B, D, H, W, float(start), float(step),
                                               _stream(prob)), "lws_disparity_regression_f32")
    LAUNCHES[0] += 1
    return out


# ------------------------------------------------------------------------------------------------ a8 + a9
def pack_refinement1(tensors: Sequence[torch.Tensor], in_channels: int, eps: float) -> torch.Tensor:
    """Module-local BN folding of a stand-alone refinement1 (25 tensors, order: include/lws.h)."""
    hs = [_host_f32(t) for t in tensors]
    n = int(lib.lws_refinement1_packed_floats(in_channels))
    if n == 0:
        raise LwsError(f"refinement1: in_channels must be 1 or 3 (the reference's two uses), got {in_channels}")
    packed = torch.zeros(n, dtype=torch.float32)
    check(lib.lws_pack_refinement1_weights(_ptr_array(hs), len(hs), in_channels, float(eps), ctypes.c_void_p(packed.data_ptr())),
          "lws_pack_refinement1_weights")
    return packed


def refinement1(x: torch.Tensor, packed: torch.Tensor) -> torch.Tensor:
    """refinement1(in, 32)(x) as a layer (reference models/submodules.py:282-300, called at models/models.py:158-159)."""
    x = _f32c(x)
    B, cin, H, W = x.shape
    out = torch.empty((B, 32, H, W), dtype=torch.float32, device=x.device)
    nbytes = int(lib.lws_refinement1_workspace_bytes(B, H, W))
    ws = workspace(x.device, "refine_part", nbytes)
    with torch.cuda.device(x.device):
        check(lib.lws_refinement1_f32(_ptr(x, "x"), _ptr(packed, "packed"), _ptr(out, "out"), ctypes.c_void_p(ws.data_ptr()), nbytes,
                                      B, cin, H, W, _stream(x)), "lws_refinement1_f32")
    LAUNCHES[0] += 5
    return out


def pack_refinement2(tensors: Sequence[torch.Tensor], eps: float) -> torch.Tensor:
    """Module-local BN folding of a stand-alone refinement2 (30 tensors, order: include/lws.h)."""
    hs = [_host_f32(t) for t in tensors]
    packed = torch.zeros(int(lib.lws_refinement2_packed_floats()), dtype=torch.float32)
    check(lib.lws_pack_refinement2_weights(_ptr_array(hs), len(hs), float(eps), ctypes.c_void_p(packed.data_ptr())),
          "lws_pack_refinement2_weights")
    return packed


def refinement2(x: torch.Tensor, packed: torch.Tensor) -> torch.Tensor:
    """refinement2(64, 32)(x) as a layer (reference models/submodules.py:302-327, called at models/models.py:160): no skip add."""
    x = _f32c(x)
    B, c, H, W = x.shape
    if c != 64:
        raise ValueError("refinement2 expects [B,64,H,W]")
    out = torch.empty((B, 1, H, W), dtype=torch.float32, device=x.device)
    nbytes = int(lib.lws_refinement2_workspace_bytes(B, H, W))
    ws = workspace(x.device, "refine_part", nbytes)
    with torch.cuda.device(x.device):
        check(lib.lws_refinement2_f32(_ptr(x, "x"), _ptr(packed, "packed"), _ptr(out, "out"), ctypes.c_void_p(ws.data_ptr()), nbytes,
                                      B, H, W, _stream(x)), "lws_refinement2_f32")
    LAUNCHES[0] += 7
    return out


def pack_refinement(tensors: Sequence[torch.Tensor], eps: float) -> torch.Tensor:
    """Fold the BNs of refinement1_left / refinement1_disp / refinement2 (host); tensor order: include/lws.h."""
    hs = [_host_f32(t) for t in tensors]
    packed = torch.zeros(int(lib.lws_refinement_packed_floats()), dtype=torch.float32)
    check(lib.lws_pack_refinement_weights(_ptr_array(hs), len(hs), float(eps), ctypes.c_void_p(packed.data_ptr())),
          "lws_pack_refinement_weights")
    return packed


def refinement(left: torch.Tensor, pred3: torch.Tensor, packed: torch.Tensor,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """pred3 + refinement2(concat[refinement1_left(left), refinement1_disp(pred3)]) (reference models/models.py:158-162)."""
    left, pred3 = _f32c(left), _f32c(pred3)
    B, three, H, W = left.shape
    if three != 3 or tuple(pred3.shape) != (B, 1, H, W):
        raise ValueError("left must be [B,3,H,W] and pred3 [B,1,H,W]")
    pred4 = out if out is not None else torch.empty_like(pred3)
    nbytes = int(lib.lws_refinement_workspace_bytes(B, H, W))
    ws = workspace(left.device, "refine", nbytes)
    with torch.cuda.device(left.device):
        check(lib.lws_refinement_f32(_ptr(left, "left"), _ptr(pred3, "pred3"), _ptr(packed, "packed"),
                                     _ptr(pred4, "pred4"), ctypes.c_void_p(ws.data_ptr()), nbytes, B, H, W,
                                     _stream(left)), "lws_refinement_f32")
    LAUNCHES[0] += int(lib.lws_refinement_launches(B, H, W))
    return pred4


def refinement_block_clp(in_clp: torch.Tensor, packed: torch.Tensor, branch: int, block: int, B: int, H: int, W: int,
                         out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """One BN-ReLU-DW(dil)-PW block (reference models/submodules.py:236-261) on the refinement's internal channels-last
    bordered layout [B, H+32, W+32, 32]; branch 0/1/2 = refinement1_left / refinement1_disp / refinement2, block 0..3."""
    n = int(lib.lws_refinement_clp_floats(B, H, W))
    if in_clp.numel() != n:
        raise ValueError(f"in_clp must hold {n} floats")
    out = out if out is not None else torch.empty_like(in_clp)
    with torch.cuda.device(in_clp.device):
        check(lib.lws_refinement_block_clp_f32(_ptr(in_clp, "in_clp"), _ptr(out, "out_clp"), _ptr(packed, "packed"), branch,
                                               block, B, H, W, _stream(in_clp)), "lws_refinement_block_clp_f32")
    LAUNCHES[0] += 1
    return out


def refinement_chain_clp(in_clp: torch.Tensor, packed: torch.Tensor, branch: int, block0: int, nblk: int, B: int, H: int, W: int,
                         out: Optional[torch.Tensor] = None, return_ws: bool = False):
    """Blocks [block0, block0 + nblk) of a refinement branch in ONE launch (dwsep_chain.cu): the tensors between the blocks stay
    in L2-resident row rings.  Same layout and bit-identical results as nblk calls of refinement_block_clp."""
    n = int(lib.lws_refinement_clp_floats(B, H, W))
    if in_clp.numel() != n:
        raise ValueError(f"in_clp must hold {n} floats")
    out = out if out is not None else torch.empty_like(in_clp)
    nbytes = int(lib.lws_refinement_chain_workspace_bytes(branch, block0, nblk, B, H, W))
    if nbytes == 0:
        raise ValueError("bad chain arguments")
    ws = workspace(in_clp.device, "chain", nbytes)
    with torch.cuda.device(in_clp.device):
        check(lib.lws_refinement_chain_clp_f32(_ptr(in_clp, "in_clp"), _ptr(out, "out_clp"), _ptr(packed, "packed"), branch, block0,
                                               nblk, ctypes.c_void_p(ws.data_ptr()), nbytes, B, H, W, _stream(in_clp)),
              "lws_refinement_chain_clp_f32")
    LAUNCHES[0] += 1
    return (out, ws) if return_ws else out


# ------------------------------------------------------------------------------------------------ n1 feature pyramid
def pack_feature_extraction(tensors: Sequence[torch.Tensor], eps: float) -> torch.Tensor:
    """Fold the BNs of feature_extraction (host); tensor order: include/lws.h."""
    hs = [_host_f32(t) for t in tensors]
    packed = torch.zeros(int(lib.lws_feature_extraction_packed_floats()), dtype=torch.float32)
    check(lib.lws_pack_feature_extraction_weights(_ptr_array(hs), len(hs), float(eps),
                                                  ctypes.c_void_p(packed.data_ptr())), "lws_pack_feature_extraction_weights")
    return packed


def feature_extraction(img: torch.Tensor, packed: torch.Tensor):
    """feature_extraction.forward (reference models/submodules.py:176-188) -> [f 1/8 (16ch), f 1/4 (16ch), f 1/2 (8ch)]."""
    img = _f32c(img)
    B, three, H, W = img.shape
    if three != 3 or H % 8 or W % 8:
        raise ValueError("img must be [B,3,H,W] with H, W multiples of 8")
    dev = img.device
    f8 = torch.empty((B, 16, H // 8, W // 8), dtype=torch.float32, device=dev)
    f4 = torch.empty((B, 16, H // 4, W // 4), dtype=torch.float32, device=dev)
    f2 = torch.empty((B, 8, H // 2, W // 2), dtype=torch.float32, device=dev)
    nbytes = int(lib.lws_feature_extraction_workspace_bytes(B, H, W))
    ws = workspace(dev, "fe", nbytes)
    with torch.cuda.device(dev):
        check(lib.lws_feature_extraction_f32(_ptr(img, "img"), _ptr(packed, "packed"), _ptr(f8, "f8"), _ptr(f4, "f4"),
                                             _ptr(f2, "f2"), ctypes.c_void_p(ws.data_ptr()), nbytes, B, H, W,
                                             _stream(img)), "lws_feature_extraction_f32")
    LAUNCHES[0] += 12
    return [f8, f4, f2]


# ------------------------------------------------------------------------------------------------ n2 pre / post
IMAGENET_MEAN = (0.485, 0.456, 0.406)  # reference dataloader/dataloader.py:10-11, used by inference.py:80-82
IMAGENET_STD = (0.229, 0.224, 0.225)
_luts: dict = {}


def normalize_lut(device: torch.device) -> torch.Tensor:
    """[3,256] fp32 table (RGB order) of Normalize(ToTensor(v)) = ((v / 255) - mean[c]) / std[c], evaluated in fp32 with the
    reference's own operation order (inference.py:80-82,102-103), so the device preprocessing is bit-identical to the host's."""
    key = device.index
    t = _luts.get(key)
    if t is None:
        v = torch.arange(256, dtype=torch.float32) / 255.0
        mean = torch.tensor(IMAGENET_MEAN, dtype=torch.float32).view(3, 1)
        std = torch.tensor(IMAGENET_STD, dtype=torch.float32).view(3, 1)
        t = ((v.view(1, 256) - mean) / std).contiguous().to(device)
        _luts[key] = t
    return t


def preprocess_bgr_u8(img: torch.Tensor, th: int = 368, tw: int = 1232, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """inference.py:93-103 on the device: img [B,h,w,3] uint8 HWC BGR (cv2.imread layout) -> bottom-right th x tw crop,
    BGR->RGB, ToTensor, Normalize -> [B,3,th,tw] fp32.  Raises ValueError when the image is smaller than the crop (the
    reference skips such images, inference.py:96-97)."""
    if img.dim() != 4 or img.shape[-1] != 3:
        raise ValueError(f"img must be [B,h,w,3] uint8 HWC, got {tuple(img.shape)}")
    B, h, w, _ = img.shape
    if h < th or w < tw:
        raise ValueError(f"image {h}x{w} is smaller than the crop {th}x{tw}")
    res = out if out is not None else torch.empty((B, 3, th, tw), dtype=torch.float32, device=img.device)
    if tuple(res.shape) != (B, 3, th, tw):
        raise ValueError("out must be [B,3,th,tw]")
    if B == 0:
        return res
    with torch.cuda.device(img.device):
        check(lib.lws_preprocess_bgr_u8(_ptr(img, "img", torch.uint8), _ptr(normalize_lut(img.device), "lut"), _ptr(res, "out"),
                                        B, h, w, th, tw, _stream(img)), "lws_preprocess_bgr_u8")
    LAUNCHES[0] += 1
    return res


def disparity_to_u8(disp: torch.Tensor, gray: bool = True, color: bool = True, out_gray: Optional[torch.Tensor] = None,
                    out_color: Optional[torch.Tensor] = None):
    """inference.py:114-115 on the device: gray = disp.astype(uint8) (truncate, wrap modulo 256), color = cv2.applyColorMap(
    convertScaleAbs(gray, alpha=1, beta=0), COLORMAP_JET) as [..., 3] BGR.  Returns (gray or None, color or None)."""
    disp = _f32c(disp)
    if not (gray or color):
        raise ValueError("ask for gray and/or color")
    g = (out_gray if out_gray is not None else torch.empty(disp.shape, dtype=torch.uint8, device=disp.device)) if gray else None
    c = (out_color if out_color is not None else torch.empty(tuple(disp.shape) + (3,), dtype=torch.uint8, device=disp.device)) \
        if color else None
    if disp.numel() == 0:
        return g, c
    with torch.cuda.device(disp.device):
        check(lib.lws_disparity_to_u8(_ptr(disp, "disp"), _ptr(g, "gray", torch.uint8), _ptr(c, "color", torch.uint8),
                                      disp.numel(), _stream(disp)), "lws_disparity_to_u8")
    LAUNCHES[0] += 1
    return g, c
