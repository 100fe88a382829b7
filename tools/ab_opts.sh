#!/bin/bash
# A/B of library options on the device-resident configs[2] line.  Usage: tools/ab_opts.sh <tag> "<opts A>" "<opts B>" ...   (each e.g. "--opt refine_chain=0")
tag=$1; shift
mkdir -p gpurun_out
i=0
for o in "$@"; do
  timeout 600 python bench.py --steps 4 --warmup 3 --skip-probes --skip-cpu $o > gpurun_out/bench_${tag}_$i.log 2> gpurun_out/bench_${tag}_$i.err || tail -3 gpurun_out/bench_${tag}_$i.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_${tag}_$i.log").read().strip().splitlines()[-1])
    print("[$o] value %.1f e2e(u8) %.1f e2e_f32 %.1f launches %d sm_mhz %s %s" % (d["value"], d["e2e"]["value"], d.get("e2e_f32", {}).get("value", 0), d["gpu_launches"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"]))
except Exception as e:
    print("[$o] failed", e)
PY
  i=$((i+1))
done
