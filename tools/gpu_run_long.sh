mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 900 python -m pytest tests/test_gpu_parity_ref.py -m gpu -x -q -k "many_layer or config1" 2>&1 | tail -5
for v in 0; do
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --latency-reps 0 --slot-latency-slots 0 --no-symbols --min-seconds 0.2 --decoder-variant $v 2>gpurun_out/b$v.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); o=d['other_configs']
print($v, o['c3_20mhz_mixed_small_tbs']['kernels_us_per_slot'], o['c3_20mhz_mixed_small_tbs']['stage_ms'], o['c3_20mhz_mixed_small_tbs']['tb_crc_ok'])
print(o['c1_full_rate_codeblocks']['kernels_ms'])
"
done
