timeout 300 python -m pytest tests -m gpu -x -q -k "conv3d_stack or feature or shard or stage" 2>&1 | tail -2
python bench.py --probes-only --probe-batch 16 2>/dev/null | grep "K3 conv3d stack C=8\|FE feature" | cut -c1-140
