#!/bin/bash
# per-kernel duration / issue-active / DRAM bytes of the launches matching a regex, per library option set (one 8-pair KITTI forward)
# Usage: tools/ncu_kernels.sh <tag> <kernel regex> "<opt=val,...>" ...
tag=$1; rx=$2; shift; shift
mkdir -p gpurun_out
i=0
for o in "$@"; do
  LWS_PROFILE_OPTS="$o" timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active \
     --clock-control none --profile-from-start off -k regex:"$rx" --csv --log-file gpurun_out/ncu_${tag}_$i.csv \
     python tools/profile_step.py --batch 8 --iters 1 > gpurun_out/ncu_${tag}_$i.log 2>&1
  echo "== $o"; python tools/ncu_kernel_table.py gpurun_out/ncu_${tag}_$i.csv
  i=$((i+1))
done
