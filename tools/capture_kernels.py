"""On the GPU box: one `ncu --set full --import-source on` capture per requested kernel of an 8-pair forward, summarised to text (metrics
via tools/ncu_summary.py, hot SASS via tools/ncu_hot_sass.py); the .ncu-rep files are deleted afterwards (size cap).
Usage: python tools/capture_kernels.py <tag> name:regex:skip [name:regex:skip ...]   (skip = matching launches to skip, ncu -s)"""
import os, subprocess, sys
tag = sys.argv[1]
os.makedirs("gpurun_out", exist_ok=True)
for spec in sys.argv[2:]:
    name, rx, skip = spec.split(":")
    rep = f"gpurun_out/full_{tag}_{name}.ncu-rep"
    cmd = ["ncu", "--set", "full", "--import-source", "on", "--clock-control", "none", "--profile-from-start", "off", "-k", "regex:" + rx,
           "-s", skip, "-c", "1", "-o", rep[:-8], "-f", "python", "tools/profile_step.py", "--batch", "8", "--iters", "1"]
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
