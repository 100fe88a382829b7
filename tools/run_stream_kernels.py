"""Launch the streaming kernels (K1, K2 x2, K4 x3, K5) once each at batch B on KITTI shapes, for ncu captures.  Usage under gpurun:
  ncu --set full --clock-control none --import-source on -k regex:'cost_volume|warp_residual|softmax_regression|scale_upsample' \
      -o gpurun_out/stream python tools/run_stream_kernels.py --batch 64
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lwsnet_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
a = ap.parse_args()
B = a.batch
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
r = lambda *s, scale=2.0: torch.randn(s, device=dev, generator=g) * scale
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for rep in range(2):  # second round is the one to look at (first = cold code / attribute calls)
    flush.zero_()
    ops.cost_volume_l1(r(B, 16, 46, 154), r(B, 16, 46, 154), 24)
    for (h, w, C) in ((92, 308, 16), (184, 616, 8)):
        flush.zero_()
        ops.warp_residual_volume_l1(r(B, C, h, w), r(B, C, h, w), torch.rand((B, 1, h, w), device=dev, generator=g) * 20, 5)
    for (D, h, w) in ((24, 46, 154), (9, 92, 308), (9, 184, 616)):
        flush.zero_()
        ops.softmax_regression(r(B, D, h, w, scale=8.0), 0.0)
    flush.zero_()
    ops.scale_upsample_add(r(B, 1, 184, 616), r(B, 1, 368, 1232), 368, 1232)
torch.cuda.synchronize()
print("done")
