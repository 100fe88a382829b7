"""Debug probe: C = 32 stack with tz_strips = 3 (CTA pairs) against tz_strips = 1."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lwsnet_b200 import ops
from lwsnet_b200.submodules import post_3dconvs
torch.cuda.set_device(0)
net = post_3dconvs(4, 32).cuda()
torch.manual_seed(0)
for p in net.parameters():
    p.data.normal_(0, 0.05)
for shp in ((1, 24, 3, 40), (2, 24, 46, 154), (3, 12, 17, 70)):
    x = torch.rand(*shp, device="cuda") * 6
    with ops.options(tz_strips=1):
        ref = net.run(x, add_skip=True).clone()
    with ops.options(tz_strips=3):
        out = net.run(x, add_skip=True)
    torch.cuda.synchronize()
    d = (out - ref).abs()
    print(shp, "equal", torch.equal(out, ref), "max|d|", d.max().item(), "frac differing", (d > 0).float().mean().item(), flush=True)
