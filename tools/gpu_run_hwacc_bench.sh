# Drop-in harnesses on the reference's own classes, larger seed sweeps than the -m gpu slices, and the C++ plugin-interface
# benchmark (oracle/hwacc_bench.cpp); artefacts into gpurun_out/.
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 600 oracle/_ref/hwacc_parity 2>&1 | tail -8
timeout 600 oracle/_ref/pdsch_hwacc_parity 200 2>&1 | tail -5
timeout 900 python tests/seed_sweep_gpu.py 40000 700 2>&1 | tail -3
timeout 900 python tests/seed_sweep_gpu_tb.py 50000 160 2>&1 | tail -3
python tools/make_tb_file.py /tmp/c2_tbs.bin 8 18
rm -f gpurun_out/r2_hwacc_bench.jsonl
for cfg in "64 3 600 8 4 --check" "192 3 600 8 4"; do set -- $cfg
timeout 600 oracle/_ref/hwacc_bench --llrs /tmp/c2_tbs.bin --decoders $1 --sets $2 --slots 1100 --threads $4 --workers $5 $6 --ref-seconds 3 --agg-tbs 64 --agg-us $3 2>/dev/null | tee -a gpurun_out/r2_hwacc_bench.jsonl | cut -c1-600; echo "hwacc_bench rc=$?"
done
