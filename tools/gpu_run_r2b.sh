# A/B of the TMEM decoder build variants, then one ncu --set full capture of the default build's decoder.
mkdir -p gpurun_out
bash tools/ab_bench.sh > gpurun_out/r2b_ab.txt 2>&1; cat gpurun_out/r2b_ab.txt
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --latency-reps 0"
$CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ldpc_decode4 -s 3 -c 1 -f -o gpurun_out/r2_decode4t_v1 $CMD > gpurun_out/ncu_r2b.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_r2b.log
