mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_demod.py -m gpu -x -q 2>&1 | tail -3
CMD="python bench.py --steps 3 --warmup 5 --no-cpu-baseline --no-other-configs --latency-reps 0 --slot-latency-slots 0 --min-seconds 0.05"
$CMD 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value', d['value']); f=d['from_symbols']; print(f['value_device_resident'], f['demod_stage_ms_alone'], f['roofline_demod']['frac'])"
$CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:pusch_demod_kernel -s 8 -c 1 -f -o gpurun_out/r2_demod_v2 $CMD > gpurun_out/ncu_demod.log 2>&1; echo "ncu rc=$?"
