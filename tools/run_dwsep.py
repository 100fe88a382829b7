"""One BN-ReLU-DW-PW block (dil 4, 8 KITTI pairs) and one conv0 (3->32) launch for ncu source-level captures."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lwsnet_b200 import ops
from lwsnet_b200.synthetic import default_args, random_init_model
torch.cuda.set_device(0)
dev = torch.device("cuda")
model = random_init_model(0, default_args(), dev)
rp = model._refinement_packed(dev)
B, H, W = 8, 368, 1232
n = int(ops.lib.lws_refinement_clp_floats(B, H, W))
a = torch.zeros(n, device=dev)
a.view(B, H + 32, W + 32, 32)[:, 16:-16, 16:-16, :].uniform_(0.0, 3.0)
b = torch.empty_like(a)
for _ in range(3):
    ops.refinement_block_clp(a, rp, 2, 1, B, H, W, out=b)
torch.cuda.synchronize()
