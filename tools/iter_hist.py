#!/usr/bin/env python3
"""Distribution of LDPC iterations per code block for the bench workload (config 2, mu as given): how much a group of four
code blocks that runs until its slowest member is done executes beyond the mean. GPU box only."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from srsran_projectvtlmo_b200 import capi, pusch  # noqa: E402

mu = float(sys.argv[1]) if len(sys.argv) > 1 else 18.0
B, ncb = 16, 152
tbs, nllr, payloads, sets = bench.make_inputs(B, 1, mu, 1000)
acc = pusch.Accelerator(device=0, max_cbs_in_flight=B * ncb, nof_harq_cb_slots=B * ncb)
w = bench.WORKLOAD
cfgs = [capi.TbConfig(tbs, w["bg"], 0, w["qm"], w["nref"], w["layers"], w["max_it"], w["early_stop"], 1, i * ncb)
        for i in range(B)]
tk = pusch.submit_tbs(acc, cfgs, [sets[0][k] for k in range(B)])
res = [pusch.poll_tb(acc, t, None) for t in tk]
its = []
lib = acc._lib
import ctypes as C
for slot in range(B * ncb):
    ok = C.c_int()
    # per code block iteration counts are not part of the TB result: use min/max/mean per TB instead
print("per TB: min", [r.iter_min for r in res], "max", [r.iter_max for r in res], "mean", [round(r.iter_mean, 2) for r in res])
