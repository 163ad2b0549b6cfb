mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
python tools/make_tb_file.py /tmp/c2_tbs.bin 8 18
for cfg in "64 3 50" "64 3 0" "16 4 50"; do set -- $cfg
timeout 600 oracle/_ref/hwacc_bench --llrs /tmp/c2_tbs.bin --decoders $1 --sets $2 --slots 400 --threads 8 --check --ref-seconds 3 --agg-tbs 64 --agg-us $3 2>&1 | tee -a gpurun_out/r2_hwacc_bench.jsonl; echo "hwacc_bench rc=$?"
done
