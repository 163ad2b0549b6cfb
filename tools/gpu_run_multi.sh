# Multi-GPU measurements on ONE 8-GPU box (gpurun --gpus 8): PCIe fabric ceiling with N processes copying at once
# (tools/fabric_test.py, set FABRIC=1), then the bench under torchrun at N = 8 and N = 4 (all legs incl. the 64-cell slot legs).
mkdir -p gpurun_out
export PYTHONPATH=$PWD
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
if [ -n "$FABRIC" ]; then
for n in 8 4 2; do
  timeout 300 $TR --nproc-per-node $n --master-port $((29600 + n)) tools/fabric_test.py > gpurun_out/r2_fabric_${n}gpu.jsonl 2> gpurun_out/fabric_$n.err; echo "fabric $n rc=$?"
done
fi
for n in 8 4; do
timeout 900 $TR --nproc-per-node $n --master-port $((29710 + n)) bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2_bench_${n}gpu.json 2> gpurun_out/r2_bench_${n}gpu.err; echo "bench $n rc=$?"; tail -2 gpurun_out/r2_bench_${n}gpu.err
done
python - <<'PY'
import json
for n in (8, 4):
    try:
        d=json.loads(open(f"gpurun_out/r2_bench_{n}gpu.json").read().strip().splitlines()[-1])
    except Exception as e:
        print(n, "no line", e); continue
    print(n, "value", d["value"], "ms/step", d["ms_per_step"], "hbm", json.dumps(d["value_tbs_left_in_hbm"])[:120], "e2e", json.dumps(d["e2e"])[:330])
    print("  slot", json.dumps(d.get("slot_latency_64_cells_us"))[:160])
    print("  slot resident", json.dumps(d.get("slot_latency_64_cells_resident_us"))[:160])
    print("  sym", json.dumps(d.get("from_symbols"))[200:420])
PY
