#!/usr/bin/env python3
"""Steady state of the host-buffer leg (bench.py's e2e): Gbit/s for growing numbers of steps, with and without copying the
TB bytes out of the page-locked result buffer, plus the device-side H2D span of a batch. GPU box only."""
import ctypes as C
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from srsran_projectvtlmo_b200 import capi, pusch  # noqa: E402

B, ncb = 64, 152
tbs, nllr, payloads, sets = bench.make_inputs(B, 2, 18.0, 1000)
w = bench.WORKLOAD
acc = pusch.Accelerator(device=0, max_cbs_in_flight=B * ncb, nof_harq_cb_slots=2 * B * ncb)
lib = capi.lib()
args = []
for s in range(2):
    p = lib.srsran_cuda_pusch_dec_host_alloc(sets[s].size)
    buf = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int8)), shape=sets[s].shape)
    buf[...] = sets[s]
    cfgs = [capi.TbConfig(tbs, w["bg"], 0, w["qm"], w["nref"], w["layers"], w["max_it"], w["early_stop"], 1,
                          (s * B + i) * ncb) for i in range(B)]
    args.append(pusch.SubmitArgs(cfgs, [buf[k] for k in range(B)]))
outs = [np.zeros(tbs // 8, np.uint8) for _ in range(B)]


def run(steps, copy_out, depth=3):
    infl, h2d = [], []
    t0 = time.perf_counter()
    for i in range(steps):
        infl.append(pusch.submit_tbs(acc, args[i % 2]))
        if len(infl) >= depth:
            tk = infl.pop(0)
            pusch.poll_tbs(acc, tk, outs if copy_out else None)
            h2d.append(pusch.ticket_timing(acc, tk[0])[0])
    while infl:
        pusch.poll_tbs(acc, infl.pop(0), outs if copy_out else None)
    acc.synchronize()
    dt = time.perf_counter() - t0
    return B * tbs * steps / dt / 1e9, float(np.median(h2d)) if h2d else 0.0


for _ in range(6):
    run(4, True)
for steps in (20, 60, 200):
    for copy_out in (True, False):
        for depth in (3,):
            g, h = run(steps, copy_out, depth)
            print(f"steps {steps:4d} copy_out {copy_out!s:5s} depth {depth}: {g:6.2f} Gbit/s, median H2D span {h:.3f} ms "
                  f"({B * nllr / (h * 1e-3) / 1e9 if h else 0:.1f} GB/s)")
acc.close()
