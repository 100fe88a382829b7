#!/bin/bash
# ncu --set full + source of the top kernels of one 8-pair forward (one launch each): for the SASS hot-loop excerpts under profiles/
tag=${1:-r02}
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none --profile-from-start off \
    -k regex:"tz_gemm_kernel|conv3d_c8p_kernel|dwsep_f16_kernel|conv3d_first|warp_residual_volume_row" -c 40 \
    -o gpurun_out/full_${tag}_top -f python tools/profile_step.py --batch 8 --iters 1 > gpurun_out/full_${tag}_top.log 2>&1
ls -la gpurun_out/full_${tag}_top.ncu-rep; tail -3 gpurun_out/full_${tag}_top.log
