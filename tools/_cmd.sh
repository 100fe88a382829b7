python tools/_dbg.py | tail -5
timeout 300 python -m pytest tests -m gpu -x -q -k "conv3d_stack or shard or stage or other_baseline" 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_q37_b8.csv python tools/profile_step.py --batch 8 --iters 1 > gpurun_out/prof_q37.log 2>&1
python tools/launch_summary.py gpurun_out/launches_q37_b8.csv > gpurun_out/launches_q37_b8.txt; head -10 gpurun_out/launches_q37_b8.txt
