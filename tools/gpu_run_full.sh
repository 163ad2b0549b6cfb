# Full GPU suite + smoke + the default bench line (what the driver runs at round end), artefacts into gpurun_out/.
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 2700 python -m pytest tests -m gpu -q --durations=6 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err; echo "bench rc=$?"; tail -5 gpurun_out/r2_bench_1gpu.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_1gpu.json").read().strip().splitlines()[-1])
for k in ("value","ms_per_step","value_tbs_left_in_hbm","e2e","clocks","stage_ms_one_batch_in_flight","tb_latency_us","slot_latency_64_cells_us","slot_latency_64_cells_resident_us","from_symbols","cpu_baseline","roofline","gpu_launches"):
    print(k, json.dumps(d.get(k))[:420])
for k,v in (d.get("other_configs") or {}).items():
    print(k, json.dumps(v)[:300])
PY
