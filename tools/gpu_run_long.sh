mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 1200 python -m pytest tests/test_gpu_parity_ref.py -m gpu -x -q -k "many_layer or config1 or config4" 2>&1 | tail -5
for v in 0 7 0 7; do timeout 300 python tools/c1_time.py $v; done
for v in 0 7; do
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --latency-reps 0 --slot-latency-slots 0 --no-symbols --min-seconds 0.2 --decoder-variant $v 2>gpurun_out/b$v.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); o=d['other_configs']
for t in o['c4_harq_rv0_rv2_rv3_64_ues']['transmissions']: print($v, t['rv'], t['kernels_ms'], t['stage_ms'], t['tb_crc_ok'])
"
done
