mkdir -p gpurun_out
python tools/make_tb_file.py /tmp/c2_tbs.bin 8 18
for cfg in "64 3 300 4 4" "64 3 600 8 4" "64 4 1000 8 4"; do set -- $cfg
timeout 600 oracle/_ref/hwacc_bench --llrs /tmp/c2_tbs.bin --decoders $1 --sets $2 --slots 600 --threads $4 --workers $5 --ref-seconds 0 --agg-tbs 64 --agg-us $3 2>&1 | tee -a gpurun_out/r2_hwacc_bench5.jsonl | cut -c1-560; echo "hwacc_bench rc=$?"
done
for n in 22 64; do timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --latency-reps 0 --tbs-per-step $n --depth 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench.py tbs/step', $n, 'value', d['value'], 'e2e', d['e2e']['value'], d['stage_ms'])"; done
