#!/bin/bash
# Usage: tools/capture_roles.sh <tag> <kernel base name> <skip>
tag=$1; k=$2; skip=${3:-0}
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:"$k" -s $skip -c 1 \
    -o gpurun_out/roles_$tag -f python tools/profile_step.py --batch 8 --iters 1 > /dev/null 2>&1
python tools/ncu_roles.py gpurun_out/roles_$tag.ncu-rep 10 > gpurun_out/roles_$tag.txt; rm -f gpurun_out/roles_$tag.ncu-rep
cat gpurun_out/roles_$tag.txt
