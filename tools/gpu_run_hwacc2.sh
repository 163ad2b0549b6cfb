mkdir -p gpurun_out
python tools/make_tb_file.py /tmp/c2_tbs.bin 8 18
for cfg in "64 3 0 8 --check" "64 3 0 8" "64 3 0 4" "64 2 0 2" "64 4 100 8"; do set -- $cfg
timeout 600 oracle/_ref/hwacc_bench --llrs /tmp/c2_tbs.bin --decoders $1 --sets $2 --slots 400 --threads $4 $5 --ref-seconds 0 --agg-tbs 64 --agg-us $3 2>&1 | tee -a gpurun_out/r2_hwacc_bench2.jsonl | cut -c1-420; echo "hwacc_bench rc=$?"
done
