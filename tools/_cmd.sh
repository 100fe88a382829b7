python -m pytest tests -m gpu -x -q -k "engine or u8 or shard" 2>&1 | tail -3
for cfg in "32 8" "24 8" "32 4" "16 8"; do set -- $cfg; python bench.py --steps 4 --warmup 3 --micro-batch $1 --host-edge $2 --skip-probes --skip-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('mb', d['config']['micro_batch'], d['config']['e2e_chunks'], 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1))"; done
