"""CUDA-event timing of one post_3dconvs stack (default: stage 1, C = 32, 24 KITTI pairs) per library option set.
Usage: python tools/time_stack.py [--c 32] [--b 24] [--iters 20] "opt=val,opt=val" "..." ...   (an empty string = defaults).
Option sets are interleaved round-robin over `--rounds` rounds so that clock drift hits all of them alike."""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lwsnet_b200 import ops
from lwsnet_b200.submodules import post_3dconvs

ap = argparse.ArgumentParser()
ap.add_argument("--c", type=int, default=32)
ap.add_argument("--b", type=int, default=24)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--rounds", type=int, default=3)
ap.add_argument("sets", nargs="*", default=[""])
a = ap.parse_args()
torch.cuda.set_device(0)
shape = {32: (24, 46, 154), 8: (9, 184, 616), 80: (9, 92, 308)}[a.c]  # 80 = the C = 8 stack of stage 2
C = 8 if a.c == 80 else a.c
net = post_3dconvs(4, C).cuda()
torch.manual_seed(0)
for p in net.parameters():
    p.data.normal_(0, 0.05)
x = torch.rand(a.b, *shape, device="cuda") * 6
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
res = {s: [] for s in a.sets}
for r in range(a.rounds):
    for s in a.sets:
        kw = {k: int(v) for k, v in (kv.split("=") for kv in s.split(",") if kv)}
        with ops.options(**kw):
            for _ in range(3):
                net.run(x, add_skip=True)
            flush.zero_()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.iters):
                net.run(x, add_skip=True)
            e1.record()
            torch.cuda.synchronize()
            res[s].append(e0.elapsed_time(e1) * 1e3 / a.iters)
for s in a.sets:
    print("[%s] C=%d B=%d stack: %s us per call (min %.1f)" % (s, C, a.b, " ".join("%.1f" % t for t in res[s]), min(res[s])))
