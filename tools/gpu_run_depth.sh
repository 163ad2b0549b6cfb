# N = 8 sensitivity of the bench legs to the number of batches in flight (--depth).
export PYTHONPATH=$PWD
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for d in 6; do
timeout 600 $TR --nproc-per-node 8 --master-port $((29800 + d)) bench.py --gpus 8 --steps 20 --warmup 5 --depth $d --no-symbols --latency-reps 0 --slot-latency-slots 0 --min-seconds 1.0 > gpurun_out/depth$d.json 2>gpurun_out/depth$d.err
python -c "
import json; d=json.loads(open('gpurun_out/depth$d.json').read().strip().splitlines()[-1]); print($d, 'value', d['value'], d['ms_per_step'], 'hbm', d['value_tbs_left_in_hbm']['value'], 'e2e', d['e2e']['value'], d['e2e']['frac_of_ceiling'], d['stage_ms'])"
done
