# Full GPU suite (parity tests proper, through the C ABI) + a short bench line.
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -22 gpurun_out/pytest_gpu.log
timeout 400 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_quick.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["stage_ms_one_batch_in_flight"], d["tb_latency_us"])
PY
