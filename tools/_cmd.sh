timeout 300 python -m pytest tests -m gpu -x -q -k "refine or shard or stage or end_to_end" 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:dwsep_f16_kernel -c 6 --csv --log-file gpurun_out/dw_q41.csv python tools/profile_step.py --batch 8 --iters 1 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/dw_q41.csv
