#!/bin/bash
# Quick GPU iteration: selected parity tests + launch list of one 4-pair forward.  Usage: tools/gpu_quick.sh <tag> <pytest -k expr>
tag=${1:-q}; kexpr=${2:-conv3d}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "$kexpr" > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$tag.log
tail -4 gpurun_out/pytest_$tag.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_$tag.csv \
    python tools/profile_step.py --batch 4 --iters 1 > gpurun_out/prof_$tag.log 2>&1
python tools/launch_summary.py gpurun_out/launches_$tag.csv | head -12
