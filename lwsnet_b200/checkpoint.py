"""`.pdparams` checkpoint import / export.

The reference restores weights with ``model.set_state_dict(paddle.load(path))`` (reference inference.py:45, train.py:84-85)
and writes them with ``paddle.save(model.state_dict(), path)``.  In Paddle 2.0 a ``.pdparams`` file is a plain pickle
(protocol 2) of ``{structured_key: numpy.ndarray}`` plus two bookkeeping entries:

* ``"StructuredToParameterName@@"``: {structured_key: internal parameter name} -- ignored on load;
* ``"UnpackBigParamInfor@@"``: present only when a tensor was split into flat slices to stay under pickle-2's 4 GB limit:
  {key: {"OriginShape": shape, "slices": [part keys]}} -- re-assembled on load.

Later Paddle 2.x releases pickle every tensor as the tuple ``(name, ndarray)``; both forms are accepted.  Paddle is not
needed (and not installed): the file is read with a *restricted* unpickler that only admits numpy array reconstruction, so
loading an untrusted checkpoint cannot execute code.  Keys follow the Paddle state-dict grammar of SURVEY.md Appendix E
(``weight, bias, _mean, _variance``), which is also the key grammar of ``lwsnet_b200.LWSNet.state_dict()``.
"""
from __future__ import annotations

import collections
import io
import pickle
from typing import Dict, Mapping

import numpy as np
import torch

NAME_TABLE_KEY = "StructuredToParameterName@@"
UNPACK_KEY = "UnpackBigParamInfor@@"

_ALLOWED = {
    ("numpy.core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "_reconstruct"),
    ("numpy.core.multiarray", "scalar"), ("numpy._core.multiarray", "scalar"),
    ("numpy", "ndarray"), ("numpy", "dtype"),
    ("collections", "OrderedDict"),
    ("_codecs", "encode"),  # protocol-2 pickles of numpy arrays written by Python 3 carry their bytes through it
}


class _NumpyOnlyUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if (module, name) in _ALLOWED:
            if module.startswith("numpy.core"):
                module = module.replace("numpy.core", "numpy._core") if hasattr(np, "_core") else module
            return super().find_class(module, name)
        raise pickle.UnpicklingError(f"{module}.{name} is not allowed in a .pdparams file")


def _as_array(key, v) -> np.ndarray:
    if isinstance(v, tuple) and len(v) == 2 and isinstance(v[0], str):  # Paddle >= 2.1: (tensor name, ndarray)
        v = v[1]
    if isinstance(v, torch.Tensor):
        v = v.detach().cpu().numpy()
    if not isinstance(v, np.ndarray):
        raise TypeError(f"checkpoint entry {key!r}: expected an ndarray, got {type(v).__name__}")
    return v


def load_pdparams(path_or_file) -> Dict[str, torch.Tensor]:
    """Read a Paddle ``.pdparams`` file -> ordered {structured key: CPU torch tensor} (dtype preserved)."""
    if hasattr(path_or_file, "read"):
        raw = _NumpyOnlyUnpickler(path_or_file, encoding="latin1").load()
    else:
        with open(path_or_file, "rb") as f:
            raw = _NumpyOnlyUnpickler(f, encoding="latin1").load()
    if not isinstance(raw, Mapping):
        raise TypeError(f".pdparams must hold a dict, got {type(raw).__name__}")
    raw = dict(raw)
    raw.pop(NAME_TABLE_KEY, None)
    for key, info in (raw.pop(UNPACK_KEY, None) or {}).items():
        parts = [np.asarray(_as_array(p, raw.pop(p))).reshape(-1) for p in info["slices"]]
        raw[key] = np.concatenate(parts).reshape(tuple(info["OriginShape"]))
    out = collections.OrderedDict()
    for key, v in raw.items():
        out[key] = torch.from_numpy(np.ascontiguousarray(_as_array(key, v)))
    return out


def save_pdparams(state_dict: Mapping[str, torch.Tensor], path_or_file) -> None:
    """Write ``state_dict`` in the Paddle 2.0 ``.pdparams`` layout (pickle protocol 2, numpy arrays, name table)."""
    obj = {k: _as_array(k, v) for k, v in state_dict.items()}
    obj[NAME_TABLE_KEY] = {k: k for k in state_dict}
    if hasattr(path_or_file, "write"):
        pickle.dump(obj, path_or_file, protocol=2)
    else:
        with open(path_or_file, "wb") as f:
            pickle.dump(obj, f, protocol=2)


def convert_state(state: Mapping, reference: Mapping[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Check a loaded checkpoint against the model's own keys / shapes (Paddle's set_state_dict only warns on a mismatch;
    a silently half-loaded stereo network is useless, so this raises) and cast to the model's dtypes."""
    state = {k: v for k, v in state.items() if k not in (NAME_TABLE_KEY, UNPACK_KEY)}
    missing = [k for k in reference if k not in state]
    unexpected = [k for k in state if k not in reference]
    if missing or unexpected:
        raise KeyError(f"checkpoint does not match LWSNet: missing {missing[:4]}{'...' if len(missing) > 4 else ''} "
                       f"({len(missing)}), unexpected {unexpected[:4]}{'...' if len(unexpected) > 4 else ''} ({len(unexpected)})")
    out = collections.OrderedDict()
    for k, ref in reference.items():
        t = torch.from_numpy(np.ascontiguousarray(_as_array(k, state[k])))
        if tuple(t.shape) != tuple(ref.shape):
            raise ValueError(f"{k}: checkpoint shape {tuple(t.shape)} != model shape {tuple(ref.shape)}")
        out[k] = t.to(ref.dtype)
    return out


def dumps(state_dict) -> bytes:
    buf = io.BytesIO()
    save_pdparams(state_dict, buf)
    return buf.getvalue()
