#!/bin/bash
# A/B measurement of kernel build variants: runs bench.py once per library under gpurun_variants/ (and the default build)
# and prints the decode stage time and the headline value. GPU box only.
cd "$(dirname "$0")/.."
run() {
  python bench.py --steps 12 --warmup 4 --no-cpu-baseline --latency-reps 30 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
s=d['stage_ms_one_batch_in_flight']
print('$1', 'value %.2f' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'decode %.4f' % s['ldpc_decode'], 'dematch %.4f' % s['rate_dematch'], 'ok', d['config']['tb_crc_ok_fraction'])
"
}
run default
for f in gpurun_variants/lib_*.so; do
  [ -e "$f" ] || continue
  SRSRAN_CUDA_PUSCH_DEC_LIB=$PWD/$f run $(basename $f .so)
done
