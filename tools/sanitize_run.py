"""Small-shape invocations of every kernel family for compute-sanitizer (memcheck / racecheck / initcheck):

    compute-sanitizer --tool memcheck  python tools/sanitize_run.py
    compute-sanitizer --tool racecheck python tools/sanitize_run.py
    compute-sanitizer --tool initcheck python tools/sanitize_run.py

Shapes are small (the sanitizer slows kernels down 10-100x) but cover every code path: both 3D stacks on the tensor-core and the
FFMA path incl. the grouped generic width, the refinement on both paths plus the L2-resident chain kernel, the volume / warp /
regression / upsample kernels, the feature pyramid and the uint8 pre / post kernels, and one whole forward."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lwsnet_b200 import ops
from lwsnet_b200.submodules import post_3dconvs
from lwsnet_b200.synthetic import default_args, random_init_model, synthetic_pair

torch.cuda.set_device(0)
dev = torch.device("cuda")
g = torch.Generator().manual_seed(0)
rnd = lambda *s: torch.randn(*s, generator=g).to(dev)
model = random_init_model(0, default_args(), dev)

L, R = rnd(2, 16, 12, 40), rnd(2, 16, 12, 40)
ops.cost_volume_l1(L, R, 24)
ops.cost_volume_l1(rnd(1, 6, 5, 31), rnd(1, 6, 5, 31), 8, 2)
disp = torch.rand(2, 1, 12, 40, generator=g).to(dev) * 10
ops.warp_bilinear(R, disp)
ops.warp_residual_volume_l1(L, R, disp, 5)
ops.warp_residual_volume_l1(rnd(1, 8, 10, 33), rnd(1, 8, 10, 33), torch.rand(1, 1, 10, 33, generator=g).to(dev) * 5, 5)
ops.warp_residual_volume_l1(rnd(1, 4, 9, 17), rnd(1, 4, 9, 17), torch.rand(1, 1, 9, 17, generator=g).to(dev) * 5, 3)
ops.softmax_regression(rnd(2, 24, 12, 40) * 8, 0.0)
ops.softmax_regression(rnd(1, 9, 7, 9) * 8, -4.0)
ops.disparity_regression(torch.softmax(rnd(1, 9, 7, 9), 1), -4.0)
ops.scale_upsample_add(rnd(1, 1, 5, 7), rnd(1, 1, 40, 56), 40, 56)
ops.disp_to_scale(rnd(1, 1, 64, 128), 16, 32)
for C, shape in ((32, (1, 6, 9, 20)), (8, (2, 9, 16, 32)), (8, (1, 4, 6, 9)), (16, (1, 5, 11, 30)), (24, (1, 3, 6, 10))):
    net = post_3dconvs(4, C).to(dev)
    x = rnd(*shape).abs() * 6
    for tc in (1, 0):
        with ops.options(conv3d_tc=tc):
            net.run(x, add_skip=True)
# r02 additions: the fused stage-1 volume + first conv (both disparity tile widths), every C = 32 schedule (linear, strips, segments,
# CTA pairs / cta_group::2, strips for the closing conv), the integer-scale K5 path at the three scales and the fused stage tail
net32 = post_3dconvs(4, 32).to(dev)
fl, fr = rnd(2, 16, 9, 66), rnd(2, 16, 9, 66)
for mode in (1, 2):
    with ops.options(fuse_volume=mode):
        ops.cost_volume_conv3d_stack(fl, fr, 24, net32.packed(dev), 32, 4)
x32 = rnd(2, 12, 7, 40).abs() * 6
for mode in (0, 1, 2, 3, 5, 7):
    with ops.options(tz_strips=mode):
        net32.run(x32, add_skip=True)
for S, (h, w) in ((2, (17, 24)), (4, (9, 12)), (8, (5, 7))):
    ops.scale_upsample_add(rnd(1, 1, h, w), rnd(1, 1, h * S, w * S), h * S, w * S)
    ops.scale_upsample_add(rnd(2, 1, h, w), None, h * S, w * S)
with ops.options(fused_tail=1):
    ops.regression_tail(rnd(1, 9, 16, 32) * 8, rnd(1, 1, 64, 128), 64, 128, -4.0, 1.0, next_hw=(32, 64))
left, right = synthetic_pair(1, 64, 128, seed=5, max_disp=20.0)
left, right = left.to(dev), right.to(dev)
pred3 = torch.rand(1, 1, 64, 128, generator=g).to(dev) * 30
for opts in (dict(refine_tc=1), dict(refine_tc=0), dict(refine_chain=2, chain_min_bands=0), dict(refine_chain=4, chain_min_bands=0)):
    with ops.options(**opts):
        model._refine(left, pred3)
model.refinement1_left(left)
model.refinement2(rnd(1, 64, 24, 40))
model.feature_extraction(left)
img = torch.randint(0, 256, (1, 70, 140, 3), dtype=torch.uint8, generator=g).to(dev)
ops.preprocess_bgr_u8(img, 64, 128)
ops.disparity_to_u8(pred3)
out = model(left, right)
torch.cuda.synchronize()
print("sanitize_run OK", float(out[3].abs().mean()))
