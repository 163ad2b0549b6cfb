# Multi-GPU measurements on ONE 8-GPU box (gpurun --gpus 8): PCIe fabric ceiling with N processes copying at once, then the
# bench under torchrun at N = 8 (all legs incl. the 64-cell slot-latency leg) and N = 4.
mkdir -p gpurun_out
export PYTHONPATH=$PWD
nvidia-smi -L | head -8
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 8 4 2; do
  timeout 300 $TR --nproc-per-node $n --master-port $((29600 + n)) tools/fabric_test.py > gpurun_out/r2_fabric_${n}gpu.jsonl 2> gpurun_out/fabric_$n.err; echo "fabric $n rc=$?"
  grep -E "llr_h2d_plus_tb_d2h|tb_d2h_only" gpurun_out/r2_fabric_${n}gpu.jsonl | grep '"copy_streams": 1' | cut -c1-330
done
timeout 900 $TR --nproc-per-node 8 --master-port 29711 bench.py --gpus 8 > gpurun_out/r2_bench_8gpu.json 2> gpurun_out/r2_bench_8gpu.err; echo "bench 8 rc=$?"; tail -3 gpurun_out/r2_bench_8gpu.err
timeout 900 $TR --nproc-per-node 4 --master-port 29712 bench.py --gpus 4 --latency-reps 100 > gpurun_out/r2_bench_4gpu.json 2> gpurun_out/r2_bench_4gpu.err; echo "bench 4 rc=$?"; tail -3 gpurun_out/r2_bench_4gpu.err
python - <<'PY'
import json
for n in (8, 4):
    try:
        d=json.loads(open(f"gpurun_out/r2_bench_{n}gpu.json").read().strip().splitlines()[-1])
    except Exception as e:
        print(n, "no line", e); continue
    print(n, "value", d["value"], "ms/step", d["ms_per_step"], "e2e", json.dumps(d["e2e"])[:400])
    print("  slot", json.dumps(d.get("slot_latency_64_cells_us"))[:500])
    print("  sym", json.dumps(d.get("from_symbols"))[:400])
    print("  clocks", json.dumps(d.get("clocks"))[:300])
PY
