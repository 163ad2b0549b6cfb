mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2a_pytest.log
timeout 400 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r2a_bench_v0.json 2> gpurun_out/r2a_bench_v0.err; echo "bench v0 rc=$?"
timeout 400 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --decoder-variant 2 > gpurun_out/r2a_bench_v2.json 2> gpurun_out/r2a_bench_v2.err; echo "bench v2 rc=$?"
python - <<'PY'
import json
for v in ("v0","v2"):
    try:
        d=json.loads(open(f"gpurun_out/r2a_bench_{v}.json").read().strip().splitlines()[-1])
        print(v, d["value"], d["ms_per_step"], d["e2e"]["value"], d["stage_ms_one_batch_in_flight"])
    except Exception as e:
        print(v, "ERR", e)
PY
