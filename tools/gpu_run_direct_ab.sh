# Single-TB latency and the small-batch configs with the results written straight into page-locked host memory (default)
# and with a device-to-host copy (build: make -C srsran_projectvtlmo_b200/csrc OUT=$PWD/gpurun_variants/lib_nodirect.so EXTRA=-DDIRECT_OUT_DEFAULT=0).
mkdir -p gpurun_out; : > gpurun_out/direct_ab.txt
for lib in ${DIRECT_AB_LIBS:-default $PWD/gpurun_variants/lib_nodirect.so default $PWD/gpurun_variants/lib_nodirect.so}; do
  [ "$lib" = default ] && lib=""
  [ -z "$lib" ] && unset SRSRAN_CUDA_PUSCH_DEC_LIB || export SRSRAN_CUDA_PUSCH_DEC_LIB=$lib
  python bench.py --steps 10 --warmup 3 --min-seconds 0.3 --no-cpu-baseline --slot-latency-slots 0 --no-symbols 2>gpurun_out/direct_ab.err | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); o=d['other_configs']; c=o['c3_20mhz_mixed_small_tbs']
print('lib=${lib##*/}', 'tb_latency', {k: round(v,1) for k,v in d['tb_latency_us'].items() if k in ('p50','p99','max')}, 'c3 stages', [round(x*1e3,1) for x in c['stage_ms']], 'value', round(d['value'],1))" >> gpurun_out/direct_ab.txt || tail -3 gpurun_out/direct_ab.err >> gpurun_out/direct_ab.txt
done
cat gpurun_out/direct_ab.txt
