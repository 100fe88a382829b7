"""CPU: `.pdparams` import/export (reference inference.py:45 `model.set_state_dict(paddle.load(path))`)."""
import io
import pickle

import numpy as np
import pytest
import torch


def _oracle_state():
    from oracle import lwsnet_torch as O
    return O.build_oracle(seed=3, random_bn=True).state_dict()


def test_pdparams_roundtrip_into_model(tmp_path):
    from oracle import lwsnet_torch as O
    from lwsnet_b200 import LWSNet
    from lwsnet_b200.checkpoint import NAME_TABLE_KEY, load_pdparams, save_pdparams
    sd = _oracle_state()
    path = tmp_path / "checkpoint.pdparams"
    save_pdparams(sd, str(path))
    raw = pickle.load(open(path, "rb"))  # Paddle 2.0 layout: plain pickle of ndarrays + the name table
    assert NAME_TABLE_KEY in raw and isinstance(raw["refinement2.5.weight"], np.ndarray) and len(raw) == 227
    loaded = load_pdparams(str(path))
    assert list(loaded) == list(sd) and len(loaded) == 226
    m = LWSNet(O.default_args())
    m.load_pdparams(str(path))
    for k, v in m.state_dict().items():
        assert torch.equal(v, sd[k]), k


def test_set_state_dict_accepts_paddle_forms_and_rejects_mismatch():
    from oracle import lwsnet_torch as O
    from lwsnet_b200 import LWSNet
    sd = _oracle_state()
    m = LWSNet(O.default_args())
    # Paddle >= 2.1 pickles (tensor name, ndarray) tuples; 2.0 pickles bare ndarrays
    mixed = {k: (("param_%d" % i, v.numpy()) if i % 2 else v.numpy()) for i, (k, v) in enumerate(sd.items())}
    mixed["StructuredToParameterName@@"] = {}
    m.set_state_dict(mixed)
    assert torch.equal(m.state_dict()["volume_postprocess.0.1.2.weight"], sd["volume_postprocess.0.1.2.weight"])
    bad = dict(mixed)
    bad.pop("refinement2.5.weight")
    with pytest.raises(KeyError):
        m.set_dict(bad)
    bad = dict(mixed)
    bad["refinement2.5.weight"] = np.zeros((1, 16, 3, 3), np.float32)
    with pytest.raises(ValueError):
        m.set_state_dict(bad)


def test_big_param_slices_are_reassembled():
    from lwsnet_b200.checkpoint import load_pdparams
    w = np.arange(24, dtype=np.float32).reshape(2, 3, 4)
    obj = {"a.weight@@.0": w.reshape(-1)[:10], "a.weight@@.1": w.reshape(-1)[10:], "b": np.ones(3, np.float32),
           "UnpackBigParamInfor@@": {"a.weight": {"OriginShape": (2, 3, 4), "slices": ["a.weight@@.0", "a.weight@@.1"]}},
           "StructuredToParameterName@@": {"a.weight": "conv_0.w_0", "b": "bn_0.b_0"}}
    out = load_pdparams(io.BytesIO(pickle.dumps(obj, protocol=2)))
    assert set(out) == {"a.weight", "b"} and torch.equal(out["a.weight"], torch.from_numpy(w))


def test_unpickler_refuses_code():
    from lwsnet_b200.checkpoint import load_pdparams

    class Evil:
        def __reduce__(self):
            import os
            return (os.system, ("true",))

    with pytest.raises(pickle.UnpicklingError):
        load_pdparams(io.BytesIO(pickle.dumps({"x": Evil()}, protocol=2)))
