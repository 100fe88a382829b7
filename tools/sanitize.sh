#!/bin/bash
# compute-sanitizer over tools/sanitize_run.py (memcheck, racecheck, initcheck); summaries -> gpurun_out/sanitizer_<tool>_<tag>.txt
tag=${1:-r02}
mkdir -p gpurun_out
for tool in memcheck racecheck initcheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_run.py > gpurun_out/sanitizer_${tool}_$tag.txt 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_run OK|Error|hazard" gpurun_out/sanitizer_${tool}_$tag.txt | sort | uniq -c | head -12
done
