#!/bin/bash
# One gpurun call: GPU parity tests, default bench line, launch list of one forward.  Usage: tools/gpu_check.sh <tag>
tag=${1:-rXX}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu_$tag.txt
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$tag.log
tail -3 gpurun_out/pytest_$tag.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$tag.log 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
cat gpurun_out/bench_$tag.log | cut -c1-600
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_$tag.csv \
    python tools/profile_step.py --batch 8 --iters 1 > gpurun_out/prof_$tag.log 2>&1
python tools/launch_summary.py gpurun_out/launches_$tag.csv
