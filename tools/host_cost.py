#!/usr/bin/env python3
"""Host-side cost of one bench step (config 2, 64 TBs): wall time of submit_tbs_device and of polling 64 tickets, measured
with the GPU idle between steps (one batch in flight). GPU box only."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from srsran_projectvtlmo_b200 import capi, pusch  # noqa: E402

B, ncb = 64, 152
tbs, nllr, payloads, sets = bench.make_inputs(B, 1, 18.0, 1000)
acc = pusch.Accelerator(device=0, max_cbs_in_flight=B * ncb, nof_harq_cb_slots=2 * B * ncb)
w = bench.WORKLOAD
cfgs = [capi.TbConfig(tbs, w["bg"], 0, w["qm"], w["nref"], w["layers"], w["max_it"], w["early_stop"], 1, i * ncb)
        for i in range(B)]
dev = torch.from_numpy(sets[0]).cuda()
lst = [(dev[k].data_ptr(), nllr) for k in range(B)]
ts, tp = [], []
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 30):
    t0 = time.perf_counter()
    tk = pusch.submit_tbs(acc, cfgs, lst, device_resident=True)
    t1 = time.perf_counter()
    acc.synchronize()
    t2 = time.perf_counter()
    for t in tk:
        pusch.poll_tb(acc, t, None)
    t3 = time.perf_counter()
    ts.append((t1 - t0) * 1e3)
    tp.append((t3 - t2) * 1e3)
print("submit ms: median %.3f min %.3f | poll 64 tickets (results ready) ms: median %.3f" %
      (np.median(ts[5:]), np.min(ts[5:]), np.median(tp[5:])))
