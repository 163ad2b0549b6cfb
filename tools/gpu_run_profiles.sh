# Round-2 profile captures: launch list of the bench command + one `ncu --set full` capture per kernel (each after the plain
# command has exited 0 without ncu). Reports land in gpurun_out/; tools/ncu_summary.py turns them into profiles/*.txt.
mkdir -p gpurun_out
export PYTHONPATH=$PWD
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --latency-reps 0 --slot-latency-slots 0 --no-other-configs --min-seconds 0.02"
$B > gpurun_out/prof_plain.json 2> gpurun_out/prof_plain.err || { echo "plain bench failed"; tail -5 gpurun_out/prof_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_final_launches.csv $B > gpurun_out/ncu_launches.log 2>&1; echo "launch list rc=$?"
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:ldpc_decode4_kernel -s 6 -c 1 -o gpurun_out/r2_final_decode4t $B > gpurun_out/ncu_a.log 2>&1; echo "decode rc=$?"
$NCU -k regex:rate_dematch_kernel -s 6 -c 1 -o gpurun_out/r2_final_dematch $B > gpurun_out/ncu_b.log 2>&1; echo "dematch rc=$?"
$NCU -k regex:pusch_demod_kernel -s 4 -c 1 -o gpurun_out/r2_final_demod $B > gpurun_out/ncu_c.log 2>&1; echo "demod rc=$?"
$NCU -k regex:tb_gather_kernel -s 6 -c 1 -o gpurun_out/r2_final_tb_gather $B > gpurun_out/ncu_d.log 2>&1; echo "tb rc=$?"
C="python tools/c1_time.py"
$C 0 > gpurun_out/c1_plain.log 2>&1 || { echo "c1 failed"; exit 1; }
$NCU -k regex:ldpc_decode4_kernel -s 10 -c 1 -o gpurun_out/r2_final_decode2t_long $C 0 > gpurun_out/ncu_e.log 2>&1; echo "long rc=$?"
$NCU -k regex:ldpc_decode4_kernel -s 10 -c 1 -o gpurun_out/r2_final_decode2t_long_bulk $C 7 > gpurun_out/ncu_f.log 2>&1; echo "long bulk rc=$?"
P="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --latency-reps 0 --slot-latency-slots 0 --no-symbols --min-seconds 0.02"
$NCU -k regex:pdsch_encode_packed_kernel -s 2 -c 1 -o gpurun_out/r2_final_pdsch_enc $P > gpurun_out/ncu_g.log 2>&1; echo "pdsch rc=$?"
# Summaries are made here (the reports together exceed what gpurun brings back); only the decoder's report travels.
for r in decode4t dematch demod tb_gather decode2t_long decode2t_long_bulk pdsch_enc; do
  f=gpurun_out/r2_final_$r.ncu-rep
  [ -e $f ] || continue
  { python tools/ncu_summary.py $f; python tools/ncu_regions.py $f 30; } > gpurun_out/r2_final_${r}_ncu_summary.txt 2>&1
  [ $r = decode4t ] || rm -f $f
done
ls -la gpurun_out | tail -20
