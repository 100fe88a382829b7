"""ctypes binding of include/lws.h.  Import fails loudly if liblws_b200.so is not built: there is no fallback."""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_longlong, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LWS_B200_LIB", os.path.join(_HERE, "lib", "liblws_b200.so"))

if not os.path.isfile(LIB_PATH):
    raise ImportError(
        f"lwsnet_b200: {LIB_PATH} not found. Build it with `make -C lwsnet_b200/csrc` "
        "(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU / PyTorch fallback.")

lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)

_fp = c_void_p  # device pointers are passed as integers
_pp = POINTER(c_void_p)

# name -> (restype, argtypes); kept in sync with include/lws.h (tests/test_abi.py parses the header and checks this)
SIGNATURES = {
    "lws_status_string": (c_char_p, [c_int]),
    "lws_version": (c_char_p, []),
    "lws_set_option": (c_int, [c_char_p, c_int]),
    "lws_get_option": (c_int, [c_char_p, POINTER(c_int)]),
    "lws_cost_volume_l1_f32": (c_int, [_fp, _fp, _fp, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "lws_disp_to_scale_f32": (c_int, [_fp, _fp, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "lws_warp_bilinear_f32": (c_int, [_fp, _fp, _fp, c_int, c_int, c_int, c_int, c_void_p]),
    "lws_warp_taps_f32": (c_int, [_fp, c_float, _fp, _fp, _fp, _fp, c_int, c_int, c_int, c_void_p]),
    "lws_warp_residual_volume_l1_f32": (c_int, [_fp, _fp, _fp, _fp, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "lws_conv3d_stack_packed_floats": (c_size_t, [c_int, c_int]),
    "lws_pack_conv3d_stack_weights": (c_int, [_pp, _pp, _pp, _pp, _pp, c_float, c_int, c_int, _fp]),
    "lws_conv3d_stack_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_int]),
    "lws_conv3d_stack_launches": (c_int, [c_int, c_int]),
    "lws_conv3d_stack_f32": (c_int, [_fp, _fp, _fp, _fp, c_size_t, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                     c_void_p]),
    "lws_cost_volume_conv3d_stack_supported": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_int]),
    "lws_cost_volume_conv3d_stack_f32": (c_int, [_fp, _fp, _fp, _fp, _fp, _fp, c_size_t, c_int, c_int, c_int, c_int, c_int, c_int,
                                                 c_int, c_int, c_void_p]),
    "lws_conv3d_bnrelu_layer_f32": (c_int, [_fp, _fp, _fp, _fp, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "lws_softmax_regression_f32": (c_int, [_fp, _fp, c_int, c_int, c_int, c_int, c_float, c_float, c_void_p]),
    "lws_scale_upsample_add_f32": (c_int, [_fp, _fp, _fp, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "lws_regression_tail_supported": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int]),
    "lws_regression_tail_f32": (c_int, [_fp, _fp, _fp, _fp, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_float,
                                        c_void_p]),
    "lws_refinement_packed_floats": (c_size_t, []),
    "lws_pack_refinement_weights": (c_int, [_pp, c_int, c_float, _fp]),
    "lws_refinement_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "lws_refinement_launches": (c_int, [c_int, c_int, c_int]),
    "lws_refinement_clp_floats": (c_size_t, [c_int, c_int, c_int]),
    "lws_refinement_block_clp_f32": (c_int, [_fp, _fp, _fp, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "lws_refinement_chain_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_int]),
    "lws_refinement_chain_clp_f32": (c_int, [_fp, _fp, _fp, c_int, c_int, c_int, _fp, c_size_t, c_int, c_int, c_int, c_void_p]),
    "lws_refinement_f32": (c_int, [_fp, _fp, _fp, _fp, _fp, c_size_t, c_int, c_int, c_int, c_void_p]),
    "lws_refinement1_packed_floats": (c_size_t, [c_int]),
    "lws_pack_refinement1_weights": (c_int, [_pp, c_int, c_int, c_float, _fp]),
    "lws_refinement1_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "lws_refinement1_f32": (c_int, [_fp, _fp, _fp, _fp, c_size_t, c_int, c_int, c_int, c_int, c_void_p]),
    "lws_refinement2_packed_floats": (c_size_t, []),
    "lws_pack_refinement2_weights": (c_int, [_pp, c_int, c_float, _fp]),
    "lws_refinement2_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "lws_refinement2_f32": (c_int, [_fp, _fp, _fp, _fp, c_size_t, c_int, c_int, c_int, c_void_p]),
    "lws_disparity_regression_f32": (c_int, [_fp, _fp, c_int, c_int, c_int, c_int, c_float, c_float, c_void_p]),
    "lws_cost_volume_l1_bwd_f32": (c_int, [_fp, _fp, _fp, _fp, _fp, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "lws_warp_residual_volume_l1_bwd_f32": (c_int, [_fp, _fp, _fp, _fp, _fp, _fp, _fp, c_int, c_int, c_int, c_int, c_int, c_int,
                                                    c_void_p]),
    "lws_softmax_regression_bwd_f32": (c_int, [_fp, _fp, _fp, c_int, c_int, c_int, c_int, c_float, c_float, c_void_p]),
    "lws_smooth_l1_loss_workspace_bytes": (c_size_t, [c_longlong]),
    "lws_smooth_l1_multistage_loss_f32": (c_int, [_pp, _fp, POINTER(c_float), c_int, c_longlong, c_float, _fp, _pp, _fp, c_size_t,
                                                  c_void_p]),
    "lws_preprocess_bgr_u8": (c_int, [_fp, _fp, _fp, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "lws_disparity_to_u8": (c_int, [_fp, _fp, _fp, c_longlong, c_void_p]),
    "lws_feature_extraction_packed_floats": (c_size_t, []),
    "lws_pack_feature_extraction_weights": (c_int, [_pp, c_int, c_float, _fp]),
    "lws_feature_extraction_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "lws_feature_extraction_f32": (c_int, [_fp, _fp, _fp, _fp, _fp, _fp, c_size_t, c_int, c_int, c_int, c_void_p]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)  # AttributeError here = the .so does not export what lws.h declares
    _fn.restype = _res
    _fn.argtypes = _args


class LwsError(RuntimeError):
    pass


def check(status: int, what: str) -> None:
    if status != 0:
        raise LwsError(f"{what} failed: {lib.lws_status_string(status).decode()} ({status})")


def version() -> str:
    return lib.lws_version().decode()


def set_option(key: str, value: int) -> None:
    """lws_set_option: explicit process-wide switch of the C library (include/lws.h lists the keys)."""
    check(lib.lws_set_option(key.encode(), int(value)), f"lws_set_option({key!r}, {value})")


def get_option(key: str) -> int:
    v = c_int(0)
    check(lib.lws_get_option(key.encode(), ctypes.byref(v)), f"lws_get_option({key!r})")
    return v.value


class options:
    """``with options(refine_tc=0, conv3d_tc=0): ...`` -- set library options for a block and restore them afterwards."""

    def __init__(self, **kv):
        self.kv, self.old = kv, {}

    def __enter__(self):
        for k, v in self.kv.items():
            self.old[k] = get_option(k)
            set_option(k, v)
        return self

    def __exit__(self, *exc):
        for k, v in self.old.items():
            set_option(k, v)
