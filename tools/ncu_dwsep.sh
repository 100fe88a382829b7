#!/bin/bash
# duration / DRAM bytes / L2 hit rate of the refinement kernels of one 8-pair forward, per library option set.
# Usage: tools/ncu_dwsep.sh <tag> "<opt=val,...>" ...
tag=$1; shift
mkdir -p gpurun_out
i=0
for o in "$@"; do
  LWS_PROFILE_OPTS="$o" timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active \
     --clock-control none --profile-from-start off -k regex:"dwsep|tz_gemm" --csv --log-file gpurun_out/ncu_${tag}_$i.csv \
     python tools/profile_step.py --batch 8 --iters 1 > gpurun_out/ncu_${tag}_$i.log 2>&1
  echo "== $o"; python tools/ncu_kernel_table.py gpurun_out/ncu_${tag}_$i.csv
  i=$((i+1))
done
