tools/gpu_quick.sh q28 "conv3d or stack or stage"
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:dwsep_f16_kernel -c 1 -o gpurun_out/dw3_q28 python tools/profile_step.py --batch 4 --iters 1 > gpurun_out/dw3_q28.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'fe_conv_kernel|fe_deconv_kernel|fe_conv_s1_kernel' -c 12 -o gpurun_out/fe_q28 python tools/profile_step.py --batch 4 --iters 1 > gpurun_out/fe_q28.log 2>&1
tail -1 gpurun_out/fe_q28.log
