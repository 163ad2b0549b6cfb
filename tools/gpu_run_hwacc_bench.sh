mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
python tools/make_tb_file.py /tmp/c2_tbs.bin 8 18
rm -f gpurun_out/r2_hwacc_bench.jsonl
for cfg in "64 3 0 4 4" "64 3 100 8 4" "64 3 600 8 4 --check" "64 1 600 8 4"; do set -- $cfg
timeout 600 oracle/_ref/hwacc_bench --llrs /tmp/c2_tbs.bin --decoders $1 --sets $2 --slots 1100 --threads $4 --workers $5 $6 --ref-seconds 3 --agg-tbs 64 --agg-us $3 2>/dev/null | tee -a gpurun_out/r2_hwacc_bench.jsonl | cut -c1-700; echo "hwacc_bench rc=$?"
done
