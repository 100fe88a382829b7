timeout 600 python -m pytest tests -m gpu -x -q -k "stage or end_to_end or shard or golden or engine" 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_q32_b8.csv python tools/profile_step.py --batch 8 --iters 1 > gpurun_out/prof_q32.log 2>&1
python tools/launch_summary.py gpurun_out/launches_q32_b8.csv > gpurun_out/launches_q32_b8.txt; head -12 gpurun_out/launches_q32_b8.txt
