#!/bin/bash
# Round-end evidence: launch list of one 8-pair KITTI forward, ncu --set full (+ source) captures of the top kernels.
tag=${1:-r02}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_$tag.csv \
    python tools/profile_step.py --batch 8 --iters 1 > gpurun_out/prof_$tag.log 2>&1
python tools/launch_summary.py gpurun_out/launches_$tag.csv > gpurun_out/launches_$tag.txt; head -40 gpurun_out/launches_$tag.txt
for k in "tz_gemm_kernel<1, 3, 3, 0>" "tz_gemm_kernel<8, 6, 1, 0>" "conv3d_c8p_kernel<0>" "dwsep_f16_kernel<0>" "dwsep_f16_kernel<3>" "conv3d_first_ydx_v2" "warp_residual_volume_row_kernel<8"; do
  name=$(echo "$k" | tr -c 'a-zA-Z0-9' '_')
  timeout 300 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:"$(echo "$k" | sed 's/[<>,]/./g')" -s 1 -c 1 \
      -o gpurun_out/full_${tag}_$name -f python tools/profile_step.py --batch 8 --iters 1 > /dev/null 2>&1
  ls -la gpurun_out/full_${tag}_$name.ncu-rep 2>/dev/null | awk '{print $5, $9}'
done
