"""On the GPU box: `ncu --set full` captures of K5 and K4 at 24 pairs, summarised to text (the .ncu-rep files are deleted)."""
import os, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r02b"
os.makedirs("gpurun_out", exist_ok=True)
# second round of tools/run_stream_kernels.py: K4 launches 3..5, K5 launch 1 (ncu -s skips matching launches)
kernels = {"k5_upsample": ("scale_upsample_add", 1)}
for name, (rx, skip) in kernels.items():
    rep = f"gpurun_out/full_{tag}_{name}.ncu-rep"
    cmd = ["ncu", "--set", "full", "--import-source", "on", "--clock-control", "none", "-k", "regex:" + rx, "-s", str(skip), "-c", "1",
           "-o", rep[:-8], "-f", "python", "tools/run_stream_kernels.py", "--batch", "24"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    if not os.path.isfile(rep):
        print(name, "capture failed:", r.stdout[-300:], r.stderr[-300:])
        continue
    with open(f"gpurun_out/ncu_full_{tag}_{name}.txt", "w") as f:
        f.write(subprocess.run([sys.executable, "tools/ncu_summary.py", rep], capture_output=True, text=True).stdout)
        f.write("\n# hottest SASS instructions by warp-stall samples (tools/ncu_hot_sass.py)\n")
        f.write(subprocess.run([sys.executable, "tools/ncu_hot_sass.py", rep, "30"], capture_output=True, text=True).stdout)
    os.remove(rep)
    print(name, "ok")
