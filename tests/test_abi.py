"""CPU: the C-ABI library loads, exports every symbol include/lws.h declares, and its host-only entry points
(weight packing, size queries, argument validation that returns before any CUDA call) behave."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "lws.h")).read()
    return sorted(set(re.findall(r"LWS_API\s+[\w\s\*]+?\b(lws_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from lwsnet_b200 import _lib
    names = header_functions()
    assert len(names) >= 17
    for n in names:
        assert hasattr(_lib.lib, n), f"{n} declared in lws.h but not exported by {_lib.LIB_PATH}"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(names)
    assert "sm_100a" in _lib.version()


def test_status_strings_and_validation_without_gpu():
    from lwsnet_b200 import _lib
    lib = _lib.lib
    assert lib.lws_status_string(0) == b"LWS_OK"
    assert lib.lws_status_string(-1) == b"LWS_ERR_BAD_SHAPE"
    assert lib.lws_status_string(-4) == b"LWS_ERR_WORKSPACE_TOO_SMALL"
    # argument validation happens before any CUDA call
    assert lib.lws_cost_volume_l1_f32(None, None, None, 1, 1, 1, 1, 1, 1, None) == -3
    dummy = ctypes.c_void_p(16)
    assert lib.lws_cost_volume_l1_f32(dummy, dummy, dummy, 1, 4, 4, 8, 13, 2, None) == -1  # maxdisp % stride
    assert lib.lws_conv3d_stack_f32(dummy, dummy, dummy, dummy, 0, 1, 4, 4, 4, 8, 4, 1, None) == -4
    assert lib.lws_conv3d_stack_f32(dummy, dummy, dummy, dummy, 1 << 40, 1, 4, 4, 4, 0, 4, 1, None) == -1  # C must be >= 1
    # explicit options instead of environment switches
    v = ctypes.c_int(-1)
    assert lib.lws_get_option(b"conv3d_tc", ctypes.byref(v)) == 0 and v.value == 1
    assert lib.lws_set_option(b"conv3d_tc", 0) == 0 and lib.lws_get_option(b"conv3d_tc", ctypes.byref(v)) == 0 and v.value == 0
    assert lib.lws_set_option(b"conv3d_tc", 1) == 0
    assert lib.lws_set_option(b"conv3d_tc", 7) == -1 and lib.lws_set_option(b"no_such_option", 1) == -5
    assert lib.lws_set_option(None, 1) == -3
    assert lib.lws_conv3d_stack_launches(8, 4) == 6 and lib.lws_conv3d_stack_launches(24, 4) == 16
    assert lib.lws_refinement1_packed_floats(2) == 0 and lib.lws_refinement1_packed_floats(3) > 0
    assert lib.lws_refinement_f32(dummy, dummy, dummy, dummy, dummy, 0, 1, 8, 8, None) == -4
    assert lib.lws_conv3d_stack_workspace_bytes(2, 24, 46, 154, 32, 4) >= 2 * 2 * 32 * 24 * 46 * 154 * 4
    assert lib.lws_refinement_workspace_bytes(1, 368, 1232) >= 128 * 368 * 1232 * 4
    # the fused stage-1 call (volume built inside the first conv kernel): host-side applicability and argument checks
    sup = lib.lws_cost_volume_conv3d_stack_supported
    assert sup(2, 32, 46, 154, 24, 32, 4) == 0                     # KITTI stage 1
    assert sup(2, 32, 46, 153, 24, 32, 4) == -5                    # odd W: no 64-bit row loads
    assert sup(2, 32, 46, 154, 24, 8, 4) == -5                     # C = 8 stack has no fused kernel
    assert sup(2, 32, 46, 154, 24, 32, 0) == -1 and sup(0, 32, 46, 154, 24, 32, 4) == -1
    assert sup(2, 32, 46, 154, 100, 32, 4) == -5                   # TMA box of 128 + 2 (D + 2) rows would exceed 256
    fused = lib.lws_cost_volume_conv3d_stack_f32
    assert fused(None, dummy, dummy, dummy, dummy, dummy, 1 << 40, 2, 32, 46, 154, 24, 32, 4, 1, None) == -3
    assert fused(dummy, dummy, dummy, dummy, dummy, dummy, 0, 2, 32, 46, 154, 24, 32, 4, 1, None) == -4
    assert fused(dummy, dummy, dummy, dummy, dummy, dummy, 1 << 40, 2, 32, 46, 153, 24, 32, 4, 1, None) == -5
    for key, default in ((b"fuse_volume", 1), (b"tz_strips", 5), (b"k5_int", 1), (b"fe_tma", 1), (b"tz_debug", 0), (b"chain_debug", 0)):
        assert lib.lws_get_option(key, ctypes.byref(v)) == 0 and v.value == default, key


def test_pack_conv3d_stack_folds_bn():
    """Host packing: w'[ci][tap][co] = w[co][ci][tap] * s_{i+1}[co], bias = t_{i+1}, [s0,t0] first."""
    from lwsnet_b200 import ops
    from lwsnet_b200.submodules import BN_EPS
    C, layers = 8, 4
    g = torch.Generator().manual_seed(0)
    convs, bns = [], []
    for i in range(layers + 2):
        cin, cout = (1 if i == 0 else C), (1 if i == layers + 1 else C)
        convs.append(torch.randn(cout, cin, 3, 3, 3, generator=g))
        bns.append((torch.rand(cin, generator=g) + 0.5, torch.randn(cin, generator=g), torch.randn(cin, generator=g),
                    torch.rand(cin, generator=g) + 0.5))
    packed = ops.pack_conv3d_stack(convs, bns, BN_EPS, C, layers).numpy()
    s = [(b[0].double() / torch.sqrt(b[3].double() + BN_EPS)).numpy() for b in bns]
    t = [(b[1].double() - b[2].double() * torch.from_numpy(s[i])).numpy() for i, b in enumerate(bns)]
    assert np.allclose(packed[0], s[0][0]) and np.allclose(packed[1], t[0][0])
    off = 4
    for i in range(layers + 2):
        cin, cout = (1 if i == 0 else C), (1 if i == layers + 1 else C)
        w = convs[i].double().numpy().reshape(cout, cin, 27)
        scale = s[i + 1] if i < layers + 1 else np.ones(1)
        exp = (w * scale[:, None, None]).transpose(1, 2, 0).reshape(-1)
        n = cin * 27 * cout
        assert np.allclose(packed[off:off + n], exp, rtol=1e-6, atol=1e-7), i
        off += (n + 3) // 4 * 4
        if i < layers + 1:
            assert np.allclose(packed[off:off + cout], t[i + 1], rtol=1e-6, atol=1e-7)
        off += (cout + 3) // 4 * 4
    # C = 8 and C = 32 carry split-fp16 tensor-core operand tables behind the generic layout (one 9*192*32-float slot per layer)
    assert packed.size == off + layers * 9 * 192 * 32
    slot = packed[off:off + 9 * 192 * 32]                                       # first 8 -> 8 layer
    wf = packed[4 + 216 + 8:4 + 216 + 8 + 8 * 27 * 8].reshape(8, 27, 8)         # its folded weights [ci][tap][co]
    tab = slot[:9 * 384].view(np.float16).reshape(9, 2, 48, 8).astype(np.float64)   # [tap kd*3+kh][K chunk][row][ci]
    inv_sw, inv_sw_lo = float(slot[9 * 384]), float(slot[9 * 384 + 1])
    sw = 1.0 / inv_sw
    assert 256 <= np.abs(wf).max() * sw < 512 and inv_sw_lo == inv_sw / 2048
    for kd in range(3):
        for kh in range(3):
            t = tab[kd * 3 + kh]
            for kw in range(3):
                w = wf[:, kd * 9 + kh * 3 + kw, :].T.astype(np.float64)       # [co][ci]
                hi, lo = t[0, kw * 8:kw * 8 + 8], t[0, 24 + kw * 8:24 + kw * 8 + 8]
                assert np.array_equal(t[1, 24 + kw * 8:24 + kw * 8 + 8], hi)   # corr rows: [wl | wh]
                assert np.count_nonzero(t[1, kw * 8:kw * 8 + 8]) == 0          # main rows: [wh | 0]
                assert np.abs((hi + lo / 2048) / sw - w).max() <= 2.0 ** -22 * np.abs(wf).max()


def test_pack_refinement_folds_bn():
    from oracle import lwsnet_torch as O
    from lwsnet_b200 import ops
    from lwsnet_b200.submodules import BN_EPS, refinement1, refinement2, refinement_tensor_list
    m = O.build_oracle(seed=3, random_bn=True)
    r1l, r1d, r2 = refinement1(3, 32), refinement1(1, 32), refinement2(64, 32)
    r1l.load_state_dict(m.refinement1_left.state_dict())
    r1d.load_state_dict(m.refinement1_disp.state_dict())
    r2.load_state_dict(m.refinement2.state_dict())
    packed = ops.pack_refinement(refinement_tensor_list(r1l, r1d, r2), BN_EPS).numpy()
    from lwsnet_b200._lib import lib
    assert packed.size == lib.lws_refinement_packed_floats() == 36096 + 12 * 2048 + 6 * 192 * 32 + 2 * 2048 + 1540
    # first section: conv0 of R1_left [3][9][32] scaled by block-1 BN; then its bias
    bn = m.refinement1_left[1][0]
    s = (bn.weight.double() / torch.sqrt(bn._variance.double() + BN_EPS)).detach().numpy()
    t = (bn.bias.double() - bn._mean.double() * torch.from_numpy(s)).detach().numpy()
    w = m.refinement1_left[0].weight.detach().double().numpy().reshape(32, 3, 9)
    exp = (w * s[:, None, None]).transpose(1, 2, 0).reshape(-1)
    assert np.allclose(packed[:864], exp, rtol=1e-6, atol=1e-7)
    assert np.allclose(packed[864:896], t, rtol=1e-6, atol=1e-7)
    # last section: conv_last [32][9][1], unscaled
    wl = m.refinement2[5].weight.detach().numpy().reshape(32 * 9)
    assert np.allclose(packed[36096 - 288:36096], wl)
    # tensor-core operand tables: first pointwise table = split-fp16 folded weights [co][ci]: hi = fp16(w * sw),
    # lo = fp16((w * sw - hi) * 2^11) with sw the power of two that puts max|w| into [256, 512); then the epilogue scales
    pwf = packed[896 + 288:896 + 288 + 1024].reshape(32, 32)           # block 1 of R1_left, [ci][co]
    slot = packed[36096:36096 + 2048]
    tc = slot[:1024].view(np.float16).reshape(64, 32).astype(np.float64)
    c0, c1 = float(slot[1024]), float(slot[1025])
    sw = 1.0 / (c0 * 2.0 ** -6)
    assert 256 <= np.abs(pwf).max() * sw < 512 and np.log2(sw) == int(np.log2(sw)) and c1 == c0 / 2048
    rec = (tc[:32] + tc[32:] / 2048) / sw
    assert np.abs(rec - pwf.T.astype(np.float64)).max() <= 2.0 ** -22 * np.abs(pwf).max()   # 22-bit significand
    with pytest.raises(Exception):
        ops.pack_refinement(refinement_tensor_list(r1l, r1d, r2)[:-1], BN_EPS)


def test_pack_feature_extraction_folds_bn():
    from oracle import lwsnet_torch as O
    from lwsnet_b200 import ops
    from lwsnet_b200._lib import lib
    from lwsnet_b200.submodules import BN_EPS, feature_extraction
    m = O.build_oracle(seed=5, random_bn=True)
    fe = feature_extraction()
    fe.load_state_dict(m.feature_extraction.state_dict())
    tl = fe.tensor_list()
    assert len(tl) == 56
    packed = ops.pack_feature_extraction(tl, BN_EPS).numpy()
    assert packed.size == lib.lws_feature_extraction_packed_floats()
    # layer 0: dres0.0 conv [4,3,3,3] scaled by its BN -> [3][9][4]
    conv, bn = m.feature_extraction.dres0[0][0], m.feature_extraction.dres0[0][1]
    s = (bn.weight.double() / torch.sqrt(bn._variance.double() + BN_EPS)).detach().numpy()
    exp = (conv.weight.detach().double().numpy().reshape(4, 3, 9) * s[:, None, None]).transpose(1, 2, 0).reshape(-1)
    assert np.allclose(packed[:108], exp, rtol=1e-6, atol=1e-7)
    t = (bn.bias.double() - bn._mean.double() * torch.from_numpy(s)).detach().numpy()
    assert np.allclose(packed[108:112], t, rtol=1e-6, atol=1e-7)
    # last layer: classif1.2 plain conv [8,8,3,3] -> [8][9][8], zero bias
    wl = m.feature_extraction.classif1[2].weight.detach().numpy().reshape(8, 8, 9).transpose(1, 2, 0).reshape(-1)
    assert np.allclose(packed[-(576 + 8):-8], wl)
    assert np.all(packed[-8:] == 0)
    assert lib.lws_feature_extraction_workspace_bytes(1, 368, 1232) > 0
    assert lib.lws_feature_extraction_workspace_bytes(1, 370, 1232) == 0  # H, W must be multiples of 8


def test_jet_table_in_kernel_source_equals_cv2_fixture():
    """The __constant__ COLORMAP_JET table compiled into prepost.cu is the committed cv2 fixture, entry for entry."""
    src = open(os.path.join(ROOT, "lwsnet_b200", "csrc", "prepost.cu")).read()
    body = src[src.index("kJetBgr[256][3] = {"):]
    body = body[:body.index("};")]
    table = np.array([[int(v) for v in m] for m in re.findall(r"\{(\d+),(\d+),(\d+)\}", body)], np.uint8)
    assert np.array_equal(table, np.load(os.path.join(ROOT, "tests", "golden", "jet_lut_bgr.npy")))


def test_paddle_binding_compiles():
    """paddle_binding/lws_paddle_ops.cc (the Paddle custom-op binding a maintainer of the reference would build) type-checks against
    include/lws.h and the minimal paddle/extension.h stand-in, and binds every compute entry point of the header."""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = os.path.join(root, "paddle_binding", "lws_paddle_ops.cc")
    gxx = shutil.which("g++")
    assert gxx, "g++ not found"
    subprocess.run([gxx, "-std=c++17", "-fsyntax-only", "-Wall", "-Werror", "-I", os.path.join(root, "paddle_binding", "stub"),
                    "-I", os.path.join(root, "include"), src], check=True)
    text = open(src).read()
    compute = [n for n in header_functions() if n.endswith(("_f32", "_u8")) and not n.startswith("lws_pack")]
    internal = {"lws_warp_taps_f32", "lws_conv3d_bnrelu_layer_f32", "lws_refinement_block_clp_f32", "lws_refinement_chain_clp_f32"}
    missing = [n for n in compute if n not in internal and n + "(" not in text]
    assert not missing, f"entry points without a Paddle op: {missing}"
