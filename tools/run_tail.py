"""The fused stage-tail kernel alone at the three KITTI stage shapes (for ncu)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lwsnet_b200 import ops
B = 24
torch.cuda.set_device(0)
for (D, h, w, nhw, has_prev) in ((24, 46, 154, (92, 308), False), (9, 92, 308, (184, 616), True), (9, 184, 616, None, True)):
    cost = torch.randn(B, D, h, w, device="cuda") * 8
    prev = torch.randn(B, 1, 368, 1232, device="cuda")
    for _ in range(2):
        ops.regression_tail(cost, prev if has_prev else None, 368, 1232, 0.0, 1.0, next_hw=nhw)
torch.cuda.synchronize()
