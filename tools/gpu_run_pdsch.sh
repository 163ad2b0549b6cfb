mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 1200 python -m pytest tests/test_gpu_pdsch_enc.py -m gpu -x -q 2>&1 | tail -6
CMD="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --latency-reps 0 --slot-latency-slots 0 --no-symbols --min-seconds 0.05"
timeout 900 $CMD 2>gpurun_out/bp.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); o=d['other_configs']['pdsch_encode_64_tbs']
print({k: o[k] for k in ('kernels_ms','stage_ms','info_gbit_per_s_kernels','ms_host_buffers','info_gbit_per_s_host_buffers','hbm_gbs')})
" || tail -20 gpurun_out/bp.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pdsch_encode_packed_kernel -s 2 -c 1 -f -o gpurun_out/r2_final_pdsch_enc $CMD > gpurun_out/ncu_pdsch.log 2>&1; echo "ncu rc=$?"
{ python tools/ncu_summary.py gpurun_out/r2_final_pdsch_enc.ncu-rep; python tools/ncu_regions.py gpurun_out/r2_final_pdsch_enc.ncu-rep 30; } > gpurun_out/r2_final_pdsch_enc_ncu_summary.txt 2>&1; rm -f gpurun_out/r2_final_pdsch_enc.ncu-rep
bash tools/ab_bench.sh 2>&1 | tail -8
bash tools/ab_bench.sh 2>&1 | tail -8
