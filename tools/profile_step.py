"""One 4-stage forward on KITTI-shaped pairs for ncu (launch list / --set full captures).  Usage under gpurun:
  ncu --metrics gpu__time_duration.sum --clock-control none -s <skip> -c <N> --csv --log-file gpurun_out/launches.csv \
      python tools/profile_step.py --batch 2 --iters 1
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import lwsnet_torch as O  # weights / inputs only
from lwsnet_b200 import LWSNet

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=2)
ap.add_argument("--iters", type=int, default=1)
ap.add_argument("--warm", type=int, default=1)
a = ap.parse_args()
torch.cuda.set_device(0)
for kv in filter(None, os.environ.get("LWS_PROFILE_OPTS", "").split(",")):  # e.g. LWS_PROFILE_OPTS=refine_chain=0,conv3d_tc=1
    from lwsnet_b200 import ops
    ops.set_option(kv.split("=")[0], int(kv.split("=")[1]))
m = LWSNet(O.default_args())
m.load_state_dict(O.build_oracle(0).state_dict())
m = m.cuda()
l, r = O.synthetic_pair(a.batch, 368, 1232, seed=1234)
l, r = l.cuda(), r.cuda()
for _ in range(a.warm):
    m(l, r)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(a.iters):
    m(l, r)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
