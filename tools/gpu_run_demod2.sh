mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_demod.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs --latency-reps 0 --slot-latency-slots 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value', d['value'], 'e2e', d['e2e']['value']); print(json.dumps(d['from_symbols'])[:900])"
