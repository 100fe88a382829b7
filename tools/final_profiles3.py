"""On the GPU box: one `ncu --set full --import-source on` capture per top kernel of an 8-pair forward, summarised to text
(metrics via tools/ncu_summary.py, hot SASS via tools/ncu_hot_sass.py); the .ncu-rep files are deleted afterwards (size cap)."""
import os
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
os.makedirs("gpurun_out", exist_ok=True)
# ncu's -k matches the base kernel name (no template arguments): the instance is picked by its position among the matching launches
kernels = {
    "tz_gemm_133": ("tz_gemm_kernel", 0),        # <1,3,3,0>: first 32 -> 32 layer of the stage-1 stack
    "tz_gemm_861": ("tz_gemm_kernel", 5),        # <8,6,1,0>: dense 64 -> 32 dilation-8 conv of refinement2
    "conv3d_c8p": ("conv3d_c8p_kernel", 5),      # a stage-3 mid layer (the first five launches are stage 2)
    "dwsep_f16_0": ("dwsep_f16_kernel", 2),      # <0>: BN-ReLU-DW(dil 4)-PW block of refinement1_left
    "dwsep_f16_3": ("dwsep_f16_kernel", 0),      # <3>: the 3 -> 32 first conv (im2col front end)
    "k2_row_c8": ("warp_residual_volume_row_kernel", 1),  # <8,9>: stage 3
}
for name, (rx, skip) in kernels.items():
    rep = f"gpurun_out/full_{tag}_{name}.ncu-rep"
    cmd = ["ncu", "--set", "full", "--import-source", "on", "--clock-control", "none", "--profile-from-start", "off", "-k",
           "regex:" + rx, "-s", str(skip), "-c", "1", "-o", rep[:-8], "-f", "python",
           "tools/profile_step.py", "--batch", "8", "--iters", "1"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    if not os.path.isfile(rep):
        print(name, "capture failed:", r.stdout[-300:], r.stderr[-300:])
        continue
    with open(f"gpurun_out/ncu_full_{tag}_{name}.txt", "w") as f:
        f.write(subprocess.run([sys.executable, "tools/ncu_summary.py", rep], capture_output=True, text=True).stdout)
        f.write("\n# hottest SASS instructions by warp-stall samples (tools/ncu_hot_sass.py)\n")
        f.write(subprocess.run([sys.executable, "tools/ncu_hot_sass.py", rep, "40"], capture_output=True, text=True).stdout)
    os.remove(rep)
    print(name, "ok")
