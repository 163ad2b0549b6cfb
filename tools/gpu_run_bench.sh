mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err; echo "bench rc=$?"; tail -5 gpurun_out/r2_bench_1gpu.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_1gpu.json").read().strip().splitlines()[-1])
for k in ("value","ms_per_step","e2e","clocks","stage_ms_one_batch_in_flight","tb_latency_us","slot_latency_64_cells_us","from_symbols","cpu_baseline"):
    print(k, json.dumps(d.get(k))[:600])
for k,v in (d.get("other_configs") or {}).items():
    print(k, json.dumps(v)[:700])
PY
