import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import lwsnet_torch as O
from lwsnet_b200 import LWSNet
m = LWSNet(O.default_args()); m.load_state_dict(O.build_oracle(0, random_bn=True).state_dict()); m = m.cuda()
left, right = O.synthetic_pair(4, 64, 128, seed=21, max_disp=20.0)
left, right = left.cuda(), right.cuda()
full = m(left, right)
full2 = m(left, right)
parts = [m(left[lo:hi].contiguous(), right[lo:hi].contiguous()) for lo, hi in ((0, 1), (1, 4))]
for s in range(4):
    cat = torch.cat([p[s] for p in parts])
    d = (full[s] - cat).abs()
    d2 = (full[s] - full2[s]).abs()
    print("stage", s, "shard diff max", d.max().item(), "n", int((d > 0).sum()), "| rerun diff max", d2.max().item(), "n", int((d2 > 0).sum()))
# stack-level determinism
for si, (D, h, w) in ((1, (9, 16, 32)), (2, (9, 32, 64))):
    st = m.volume_postprocess[si]
    x = torch.rand(4, D, h, w, device="cuda") * 20
    a = st.run(x, add_skip=True); b = st.run(x, add_skip=True)
    c = torch.cat([st.run(x[:1].contiguous(), add_skip=True), st.run(x[1:].contiguous(), add_skip=True)])
    print("stack", si, "rerun n", int((a != b).sum()), "shard n", int((a != c).sum()), "max", (a - c).abs().max().item())
    idx = (a != c).nonzero()
    print(idx[:12].tolist())
