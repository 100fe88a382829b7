timeout 400 python -m pytest tests -m gpu -x -q -k "conv3d_stack or feature or fe_ or shard or stage or end_to_end or unique" 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_q43_b8.csv python tools/profile_step.py --batch 8 --iters 1 > gpurun_out/prof_q43.log 2>&1
python tools/launch_summary.py gpurun_out/launches_q43_b8.csv > gpurun_out/launches_q43_b8.txt; head -1 gpurun_out/launches_q43_b8.txt; grep "fe_\|first" gpurun_out/launches_q43_b8.txt
