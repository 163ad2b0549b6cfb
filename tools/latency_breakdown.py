#!/usr/bin/env python3
"""Where the single-TB latency goes: wall clock of submit_tbs / poll_tb for one config-2 TB from page-locked host LLRs and the
device-side spans of its batch (H2D incl. descriptors, dematch, decode, TB assembly, D2H). GPU box only."""
import ctypes as C
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from srsran_projectvtlmo_b200 import capi, pusch  # noqa: E402

tbs, nllr, payloads, sets = bench.make_inputs(1, 1, 18.0, 1000)
w = bench.WORKLOAD
acc = pusch.Accelerator(device=0, max_cbs_in_flight=152, nof_harq_cb_slots=152)
lib = capi.lib()
p = lib.srsran_cuda_pusch_dec_host_alloc(nllr)
buf = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int8)), shape=(nllr,))
buf[:] = sets[0][0]
cfg = [capi.TbConfig(tbs, w["bg"], 0, w["qm"], w["nref"], w["layers"], w["max_it"], w["early_stop"], 1, 0)]
args = pusch.SubmitArgs(cfg, [buf])
out = np.zeros(tbs // 8, np.uint8)
rows = []
for i in range(300):
    t0 = time.perf_counter()
    tk = pusch.submit_tbs(acc, args)
    t1 = time.perf_counter()
    pusch.poll_tb(acc, tk[0], out)
    t2 = time.perf_counter()
    rows.append([(t1 - t0) * 1e6, (t2 - t1) * 1e6, (t2 - t0) * 1e6] + [1e3 * v for v in pusch.ticket_timing(acc, tk[0])])
r = np.median(np.array(rows[20:]), axis=0)
print("median us: submit call %.1f | wait+poll %.1f | total %.1f || device spans: h2d %.1f dematch %.1f decode %.1f tb %.1f d2h %.1f (sum %.1f)"
      % (r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7], r[3:8].sum()))
acc.close()
