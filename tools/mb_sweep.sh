#!/bin/bash
# A/B of the micro-batch (pairs per forward / CUDA-graph replay) on the device-resident configs[2] line.  Usage: tools/mb_sweep.sh <tag> [mbs]
tag=${1:-mb}; mbs=${2:-"1 2 4 8 24"}
mkdir -p gpurun_out
for mb in $mbs; do
  timeout 600 python bench.py --steps 4 --warmup 3 --micro-batch $mb --skip-probes --skip-cpu > gpurun_out/bench_${tag}_mb$mb.log 2> gpurun_out/bench_${tag}_mb$mb.err || tail -3 gpurun_out/bench_${tag}_mb$mb.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_${tag}_mb$mb.log").read().strip().splitlines()[-1])
    print("mb=$mb value %.1f e2e(u8) %.1f e2e_f32 %.1f launches %d clocks %s" % (d["value"], d["e2e"]["value"], d.get("e2e_f32", {}).get("value", 0), d["gpu_launches"], d["clocks"]))
except Exception as e:
    print("mb=$mb failed", e)
PY
done
