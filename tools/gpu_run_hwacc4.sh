mkdir -p gpurun_out
python tools/make_tb_file.py /tmp/c2_tbs.bin 8 18
timeout 300 oracle/_ref/hwacc_parity | tail -3
for cfg in "64 3 0 4 4 --check" "64 3 0 4 4" "64 3 50 4 4" "64 3 0 2 8" "64 2 0 4 4" "8 3 0 2 2"; do set -- $cfg
timeout 600 oracle/_ref/hwacc_bench --llrs /tmp/c2_tbs.bin --decoders $1 --sets $2 --slots 600 --threads $4 --workers $5 $6 --ref-seconds 0 --agg-tbs 64 --agg-us $3 2>&1 | tee -a gpurun_out/r2_hwacc_bench4.jsonl | cut -c1-560; echo "hwacc_bench rc=$?"
done
