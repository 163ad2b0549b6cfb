# Plugin-interface bench (oracle/hwacc_bench.cpp) with and without the gather kernel for the decoders' separate soft-bit buffers.
# The copy-engine arm needs: make -C srsran_projectvtlmo_b200/csrc OUT=$PWD/gpurun_variants/lib_nogather.so EXTRA=-DH2D_GATHER_DEFAULT=0
mkdir -p gpurun_out; export PYTHONPATH=$PWD
python tools/make_tb_file.py /tmp/c2_tbs.bin 8 18 > /dev/null
: > gpurun_out/hwacc_ab.txt
for pre in "" "$PWD/gpurun_variants/lib_nogather.so" "" "$PWD/gpurun_variants/lib_nogather.so"; do
  LD_PRELOAD=$pre timeout 300 oracle/_ref/hwacc_bench --llrs /tmp/c2_tbs.bin --decoders 64 --sets 3 --slots 700 --threads 8 --workers 4 --ref-seconds 0.5 --agg-tbs 64 --agg-us 600 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('preload=${pre##*/}', d['value'], d['slot_latency_us']['p50'], d['failed_or_wrong_tbs'])" >> gpurun_out/hwacc_ab.txt
done
cat gpurun_out/hwacc_ab.txt
