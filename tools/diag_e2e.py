#!/usr/bin/env python3
"""Diagnostics of the host-buffer (e2e) path: raw PCIe copy bandwidth, then wall-clock and device-stage breakdown of
submit_tbs / poll_tb for the bench workload. Development tool (GPU box only); prints one JSON object."""
import ctypes as C
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

import bench  # noqa: E402
from srsran_projectvtlmo_b200 import capi, pusch  # noqa: E402


def copy_bw(nbytes, reps=5):
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    out = {}
    for name, (dst, src) in {"h2d": (d, h), "d2h": (h, d)}.items():
        best = 1e9
        for _ in range(reps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        out[name + "_gbs"] = nbytes / best / 1e9
    return out


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    steps = 6
    res = {"copy_87MB": copy_bw(87 << 20), "copy_1p3MB": copy_bw(1362816)}
    w = bench.WORKLOAD
    tbs, nllr, payloads, sets = bench.make_inputs(B, 2, 18.0, 1000)
    ncb = 152
    acc = pusch.Accelerator(device=0, max_cbs_in_flight=B * ncb, nof_harq_cb_slots=B * ncb)
    cfgs = [capi.TbConfig(tbs, w["bg"], 0, w["qm"], w["nref"], w["layers"], w["max_it"], w["early_stop"], 1, i * ncb)
            for i in range(B)]
    lib = capi.lib()
    host_sets = []
    for s in sets:
        p = lib.srsran_cuda_pusch_dec_host_alloc(s.size)
        buf = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int8)), shape=s.shape)
        buf[...] = s
        host_sets.append(buf)
    tb_out = np.zeros(tbs // 8, np.uint8)
    rows = []
    for i in range(steps):
        buf = host_sets[i % 2]
        t0 = time.perf_counter()
        tk = pusch.submit_tbs(acc, cfgs, [buf[k] for k in range(B)])
        t1 = time.perf_counter()
        acc.synchronize()
        t2 = time.perf_counter()
        ms = pusch.ticket_timing(acc, tk[0])
        for t in tk:
            pusch.poll_tb(acc, t, tb_out)
        t3 = time.perf_counter()
        rows.append({"submit_ms": (t1 - t0) * 1e3, "wait_ms": (t2 - t1) * 1e3, "poll_ms": (t3 - t2) * 1e3,
                     "stage_ms": [round(x, 3) for x in ms]})
    res["steps"] = rows
    # pageable host memory for comparison
    pag = [np.array(host_sets[0][k]) for k in range(B)]
    t0 = time.perf_counter()
    tk = pusch.submit_tbs(acc, cfgs, pag)
    t1 = time.perf_counter()
    acc.synchronize()
    t2 = time.perf_counter()
    for t in tk:
        pusch.poll_tb(acc, t, tb_out)
    res["pageable"] = {"submit_ms": (t1 - t0) * 1e3, "wait_ms": (t2 - t1) * 1e3}
    print(json.dumps(res, indent=1))
    acc.close()


if __name__ == "__main__":
    main()
