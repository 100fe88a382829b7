"""Shared helpers for the test-suite (tests may import oracle/; the product never does)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def rnd(seed, *shape, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


def product_from_oracle(oracle_model, args=None):
    """CUDA product model carrying exactly the oracle's weights (identical state-dict key grammar)."""
    from oracle.lwsnet_torch import default_args
    from lwsnet_b200 import LWSNet
    m = LWSNet(args or default_args())
    missing, unexpected = m.load_state_dict({k: v.float() for k, v in oracle_model.state_dict().items()}, strict=True)
    assert not missing and not unexpected
    return m.cuda()


def err_stats(a, b):
    d = (a.double() - b.double()).abs().flatten()
    return dict(max=d.max().item(), mean=d.mean().item(), p999=torch.quantile(d[:: max(1, d.numel() // 1_000_000)], 0.999).item(),
                frac_le_1e3=(d <= 1e-3).double().mean().item())
