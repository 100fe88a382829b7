#!/bin/bash
# Streaming-kernel iteration: selected parity tests + the per-kernel probes at several batch sizes.  Usage: tools/gpu_probe.sh <tag> <pytest -k expr> [batches]
tag=${1:-p}; kexpr=${2:-volume}; batches=${3:-"8 16 64"}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "$kexpr" > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$tag.log
tail -4 gpurun_out/pytest_$tag.log
for b in $batches; do
  timeout 300 python bench.py --probes-only --probe-batch $b > gpurun_out/probes_${tag}_b$b.log 2> gpurun_out/probes_${tag}_b$b.err || tail -5 gpurun_out/probes_${tag}_b$b.err
  echo "--- probe batch $b"; grep -v "K3\|K6\|FE" gpurun_out/probes_${tag}_b$b.log | cut -c1-200
done
