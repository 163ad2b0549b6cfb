python tools/make_tb_file.py /tmp/c2_tbs.bin 8 18
LD_PRELOAD=$PWD/gpurun_variants/lib_hostprof.so timeout 600 oracle/_ref/hwacc_bench --llrs /tmp/c2_tbs.bin --decoders 64 --sets 3 --slots 400 --threads 4 --ref-seconds 0 --agg-tbs 64 --agg-us 0 2>&1 | cut -c1-400
