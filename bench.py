#!/usr/bin/env python3
"""PUSCH channel-decoding benchmark (BASELINE.json metric: decoded info Gbit/s, BG1 Z=384).

  python bench.py --gpus 1 --steps 10 --warmup 3            our arm (CUDA path through the C ABI)
  python bench.py --impl reference --steps 3 --warmup 1     the reference's own CPU implementation on the host cores
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   one rank per GPU, no collective on the
                                                                                  data path (TBs are independent)

A step = one pass of the hot path (rate dematch + HARQ combine -> layered LDPC -> CB CRC -> TB assembly + CRC24A) over
one batch of `--tbs-per-step` synthetic transport blocks of BASELINE config 2 (273 PRB, 256QAM, R=948/1024, 4 layers,
TBS 1 277 992, 152 code blocks BG1 Z=384, Nref 12611, rv0, early stop, <= 6 iterations), encoded by the numpy
transmitter (srsran_projectvtlmo_b200/synth.py), AWGN LLRs.

One JSON line on stdout (rank 0). `value` = device-resident inputs (LLRs already in HBM), timed with CUDA events on the
stream the kernels run on; `e2e` = the same through the host-buffer C ABI with H2D/D2H inside the timed region. Every timed
region lasts >= --min-seconds (the K-step loop is repeated; ms_per_step is per step) and aborts unless every transport
block decodes. Further keys: `from_symbols` (equalized symbols in: demodulation on the device), `slot_latency_64_cells_us`
(BASELINE config 5, sharded over the ranks by TbDispatcher), `other_configs` (BASELINE configs 1, 2 worst case, 3, 4 with
the compiled reference timed beside each; N = 1 only), `cpu_baseline`.
"""
import argparse
import ctypes as C
import gc
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOAD = {"name": "pusch_273prb_256qam_r948_4layer_rv0", "prb": 273, "qm": 8, "rate": 948, "layers": 4, "bg": 1,
            "nref": 12611, "max_it": 6, "early_stop": 1}
BG1_DEG = [19, 19, 19, 19, 3, 8, 9, 7, 10, 9, 7, 8, 7, 6, 7, 7, 6, 6, 6, 6, 6, 6, 5, 5, 6, 5, 5, 4, 5, 5, 5, 5, 5, 5, 5,
           5, 5, 4, 5, 5, 4, 5, 4, 5, 5, 4]


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d, "MEASURED_PEAKS.json"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons of ONE GPU sampled through NVML from a thread of this process while the timed region
    runs (the nvidia-smi fields of the B200_PROFILING.md recipe, read in-process: a query takes < 0.1 ms, so a 20 ms
    region yields samples, and N ranks do not start N nvidia-smi processes)."""

    def __init__(self, index, period_ms=4.0):
        self.index = index
        self.period = period_ms * 1e-3
        self.rows = []
        self.run = False
        self.h = None
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            uuid = None
            try:
                uuid = "GPU-" + str(torch_uuid(index))
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if bytes is str else uuid)
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
        except Exception:
            self.h = None

    def _loop(self):
        nv = self.nv
        while self.run:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                try:
                    pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                except Exception:
                    pw = None
                self.rows.append((sm, rs, pw))
            except Exception:
                pass
            time.sleep(self.period)

    def start(self):
        if self.h is None:
            return
        self.run = True
        self.t = threading.Thread(target=self._loop, daemon=True)
        self.t.start()

    def stop(self):
        if self.h is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"], "samples": 0}
        self.run = False
        self.t.join(timeout=1)
        nv = self.nv
        try:
            smax = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
        except Exception:
            smax = None
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                 "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        reasons = sorted(n for n, bit in names.items() if any(r[1] & bit for r in self.rows))
        sm = [float(r[0]) for r in self.rows]
        pw = [r[2] for r in self.rows if r[2] is not None]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": reasons,
                "samples": len(sm), "sm_mhz_min": float(min(sm)) if sm else None,
                "sm_mhz_first_and_last_tenth": [float(np.median(sm[:max(1, len(sm) // 10)])),
                                                float(np.median(sm[-max(1, len(sm) // 10):]))] if sm else None,
                "power_w_median": float(np.median(pw)) if pw else None, "power_w_max": float(max(pw)) if pw else None,
                "source": "NVML, sampled every few ms inside the timed value region (>= 2 s of sustained load)"}


def bind_to_gpu_cpus(index):
    """Pins this process (and the threads it starts later, CUDA's included) to the CPUs NVML names as closest to the GPU, before
    anything is allocated: page-locked buffers then land on the GPU's NUMA node and launches do not cross sockets. What
    `numactl` does for a one-process-per-GPU job; a no-op where the container's cpuset does not allow it."""
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            h = pynvml.nvmlDeviceGetHandleByUUID("GPU-" + str(torch_uuid(index)))
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(index)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return len(os.sched_getaffinity(0))
    except Exception:
        return None


def torch_uuid(index):
    import torch
    return torch.cuda.get_device_properties(index).uuid


def make_inputs(tbs_per_step, nsets, mu, seed):
    """`nsets` batches of `tbs_per_step` TBs: a few distinct payloads, an independent noise realisation per TB."""
    from srsran_projectvtlmo_b200 import synth

    w = WORKLOAD
    tbs = synth.tbs_for(w["prb"], w["qm"], w["rate"], w["layers"])
    nllr = w["prb"] * 156 * w["qm"] * w["layers"]
    rng = np.random.default_rng(seed)
    payloads, cws = [], []
    for _ in range(min(4, tbs_per_step)):
        tb = rng.integers(0, 256, tbs // 8, dtype=np.uint8)
        payloads.append(tb)
        cws.append(synth.encode_tb(tb, w["bg"], 0, w["qm"], w["nref"], w["layers"], nllr))
    sets = []
    for _ in range(nsets):
        llrs = np.empty((tbs_per_step, nllr), np.int8)
        for i in range(tbs_per_step):
            llrs[i] = synth.awgn_llrs(rng, cws[i % len(cws)], mu)
        sets.append(llrs)
    return tbs, nllr, payloads, sets


def reference_arm(args, rank, world):
    """The reference's own CPU implementation of the path on the host cores (oracle/_ref when compiled, else the port)."""
    if rank != 0:
        return
    from oracle import bindings as ob

    w = WORKLOAD
    tbs, nllr, payloads, sets = make_inputs(1, 1, args.mu, 1234)
    llr = np.ascontiguousarray(sets[0][0])
    threads = ob.host_threads()
    lib = ob.ref()
    ok = C.c_int(0)
    if lib is not None and ob.ref_flavour() is not None:
        kind, flavour = "reference", ob.ref_flavour()

        def run(reps):
            return lib.ref_pusch_bench_mt(b"auto", threads, tbs // 8, ob._pi(llr), nllr, w["bg"], w["qm"], w["nref"],
                                          w["layers"], w["max_it"], w["early_stop"], reps, C.byref(ok))
    else:
        kind, flavour, threads = "port", "scalar C restatement", 1

        def run(reps):
            return ob.port().oracle_pusch_bench(tbs // 8, ob._pi(llr), nllr, w["bg"], w["qm"], w["nref"], w["layers"],
                                                w["max_it"], w["early_stop"], reps, C.byref(ok))
    t1 = run(1)
    # A step is a bounded sample sized so that the whole --steps K --warmup W run ends within a few minutes (<= 150 s timed).
    per_step_s = min(args.ref_seconds, 150.0 / max(args.steps, 1))
    reps = max(1, int(per_step_s / max(t1, 1e-3)))
    for _ in range(args.warmup):
        run(1)
    times = [run(reps) for _ in range(args.steps)]
    t = float(np.mean(times))
    gbps = threads * reps * tbs / t / 1e9
    line = {
        "impl": "reference", "metric": "pusch_decoded_info_gbit_per_s", "value": gbps, "unit": "Gbit/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int8", "data": "synthetic",
        "config": {"workload": w["name"], "tbs_bits": tbs, "codeblocks_per_tb": 152, "mu": args.mu,
                   "tbs_per_step": threads * reps},
        "cpu_baseline": {"value": gbps, "unit": "Gbit/s", "cores": threads, "kind": kind,
                         "sample": f"{threads} threads x {reps} TBs per step, {flavour} decoder/dematcher, all TB CRC ok: "
                                   f"{ok.value == threads * reps}"},
        "e2e": {"value": gbps, "unit": "Gbit/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline(args, tbs, nllr, llr):
    """Bounded CPU sample on rank 0: the compiled reference on all host threads (kind "reference") or the port."""
    try:
        from oracle import bindings as ob
    except Exception as e:  # pragma: no cover
        return {"value": None, "unit": "Gbit/s", "cores": 0, "kind": "port", "sample": f"oracle unavailable: {e}"}
    w = WORKLOAD
    ok = C.c_int(0)
    llr = np.ascontiguousarray(llr)
    lib = ob.ref()
    if lib is not None and ob.ref_flavour() is not None:
        threads = ob.host_threads()
        t1 = lib.ref_pusch_bench_mt(b"auto", threads, tbs // 8, ob._pi(llr), nllr, w["bg"], w["qm"], w["nref"],
                                    w["layers"], w["max_it"], w["early_stop"], 1, C.byref(ok))
        reps = max(1, int(args.cpu_seconds / max(t1, 1e-3)))
        t = lib.ref_pusch_bench_mt(b"auto", threads, tbs // 8, ob._pi(llr), nllr, w["bg"], w["qm"], w["nref"],
                                   w["layers"], w["max_it"], w["early_stop"], reps, C.byref(ok))
        t1c = lib.ref_pusch_bench_mt(b"auto", 1, tbs // 8, ob._pi(llr), nllr, w["bg"], w["qm"], w["nref"], w["layers"],
                                     w["max_it"], w["early_stop"], max(1, reps // 2), C.byref(ok))
        return {"value": threads * reps * tbs / t / 1e9, "unit": "Gbit/s", "cores": threads, "kind": "reference",
                "sample": f"{threads} threads x {reps} TBs of the same workload, {ob.ref_flavour()} flavour "
                          f"({lib.ref_info().decode()})",
                "single_core_value": max(1, reps // 2) * tbs / t1c / 1e9}
    t1 = ob.port().oracle_pusch_bench(tbs // 8, ob._pi(llr), nllr, w["bg"], w["qm"], w["nref"], w["layers"], w["max_it"],
                                      w["early_stop"], 1, C.byref(ok))
    reps = max(1, int(args.cpu_seconds / max(t1, 1e-3)))
    t = ob.port().oracle_pusch_bench(tbs // 8, ob._pi(llr), nllr, w["bg"], w["qm"], w["nref"], w["layers"], w["max_it"],
                                     w["early_stop"], reps, C.byref(ok))
    return {"value": reps * tbs / t / 1e9, "unit": "Gbit/s", "cores": 1, "kind": "port",
            "sample": f"1 thread x {reps} TBs of the same workload, scalar C restatement"}


def pct(v, p):
    return float(np.percentile(v, p)) if len(v) else None


def other_configs(args, rng_seed=7):
    """BASELINE.json configurations other than the headline one, each with the compiled reference timed beside it on the
    host cores (bounded samples). Rank 0, N = 1 only. Kernel times are CUDA-event spans of the library's streams."""
    from oracle import bindings as ob
    from srsran_projectvtlmo_b200 import capi, pusch, synth
    import torch

    out = {}
    ref = ob.ref() if (ob.ref() is not None and ob.ref_flavour() is not None) else None
    threads = ob.host_threads()

    def ref_tb_gbps(tbs, nllr, llr, bg, qm, nref, nl, seconds=2.0):
        """Reference pusch_decoder_impl on all host threads, new data, this transport block."""
        if ref is None:
            return None
        ok = C.c_int(0)
        llr = np.ascontiguousarray(llr)
        t1 = ref.ref_pusch_bench_mt(b"auto", threads, tbs // 8, ob._pi(llr), nllr, bg, qm, nref, nl, 6, 1, 1, C.byref(ok))
        reps = max(1, int(seconds / max(t1, 1e-4)))
        t = ref.ref_pusch_bench_mt(b"auto", threads, tbs // 8, ob._pi(llr), nllr, bg, qm, nref, nl, 6, 1, reps, C.byref(ok))
        return threads * reps * tbs / t / 1e9

    def run_dev(acc, cfgs, dev, nllrs, reps):
        stage = np.zeros(5)
        res = None
        for _ in range(reps):
            tk = pusch.submit_tbs(acc, cfgs, [(d.data_ptr(), n) for d, n in zip(dev, nllrs)], device_resident=True)
            stage += np.array(pusch.ticket_timing(acc, tk[0]))
            res = pusch.poll_tbs(acc, tk)
        return stage / reps, res

    # ---- C1: ldpc_decoder benchmark input, BG1 Z=384 rate 1/3 (46 layers), 6 iterations, no CRC, random +-10 ----------------
    acc = pusch.Accelerator(device=0, max_cbs_in_flight=64 * 152, nof_harq_cb_slots=64 * 152)
    acc.set_decoder_variant(args.decoder_variant)
    try:
        n = 592
        mt = np.random.RandomState(0)
        llr = ((mt.randint(0, 2 ** 32, (n, 25344), dtype=np.uint64) & 1) * 20 - 10).astype(np.int8)
        bits = np.zeros((n, 1056), np.uint8)
        its = np.zeros(n, np.int32)

        def go():
            st = acc._lib.srsran_cuda_ldpc_decode_batch(acc.h, bits.ctypes.data_as(capi.u8p), llr.ctypes.data_as(capi.i8p), n,
                                                        25344, 1, 384, 0, 0, 6, C.c_float(0.8), its.ctypes.data_as(capi.intp))
            assert st == 0
        for _ in range(9):  # every batch context of the handle (8) allocates its staging on first use
            go()
        t0 = time.perf_counter()
        for _ in range(8):
            go()
        dt = (time.perf_counter() - t0) / 8
        stage = np.zeros(5)
        for _ in range(5):
            go()
            stage += np.array(pusch.last_unit_timing(acc))
        stage /= 5
        kern = float(stage[2])
        c1 = {"workload": "BG1 Z=384, 25344 soft bits (46 layers), 6 iterations, no CRC, +-10 coin-flip input (ldpc_decoder_benchmark.cpp:143-145)",
              "codeblocks": n, "kernels_ms": kern, "stage_ms": stage.tolist(),
              "info_gbit_per_s_kernels": n * 8448 / (kern * 1e-3) / 1e9,
              "edge_updates_per_s_kernels": n * 6 * 384 * 316 / (kern * 1e-3),
              "ms_host_buffers": dt * 1e3, "info_gbit_per_s": n * 8448 / dt / 1e9,
              "note": "unit-level ldpc_decoder interface (synchronous, pageable host buffers: 15 MB of H2D and the host-side "
                      "staging are inside ms_host_buffers); kernels_ms = the decoder kernel's span on the device"}
        if ref is not None:
            one = np.ascontiguousarray(llr[0])
            secs = ref.ref_ldpc_decode_bench(b"auto", ob._pi(one), 25344, 1, 384, 0, 6, 2000)
            c1["reference"] = {"value": 8448 * 2000 / secs / 1e9, "unit": "Gbit/s info", "cores": 1,
                               "what": "reference ldpc_decoder (auto = AVX-512/AVX2), one thread, same input, 2000 repetitions"}
        out["c1_full_rate_codeblocks"] = c1

        # ---- C2 worst case: random +-10 soft bits, never converges (6 iterations x 4 layers x 152 code blocks) ---------------
        rng = np.random.default_rng(rng_seed)
        B, ncb, tbs, nllr = 64, 152, 1277992, 1362816
        host = (rng.integers(0, 2, (B, nllr)) * 20 - 10).astype(np.int8)
        dev = torch.from_numpy(host).cuda()
        cfgs = [capi.TbConfig(tbs, 1, 0, 8, 12611, 4, 6, 1, 1, i * ncb) for i in range(B)]
        stage, res = run_dev(acc, cfgs, [dev[i] for i in range(B)], [nllr] * B, 5)
        kern = stage[1] + stage[2] + stage[3]
        out["c2_worst_case_random_llrs"] = {
            "workload": "64 TBs of config 2, +-10 coin-flip soft bits (pusch_decoder_hwacc_benchmark.cpp:377-380)",
            "kernels_ms": kern, "stage_ms": stage.tolist(), "info_gbit_per_s_kernels": B * tbs / (kern * 1e-3) / 1e9,
            "tb_crc_ok": int(sum(r.tb_crc_ok for r in res)), "iter_mean": float(np.mean([r.iter_mean for r in res])),
            "reference": {"value": ref_tb_gbps(tbs, nllr, host[0], 1, 8, 12611, 4), "unit": "Gbit/s", "cores": threads}}
        del dev
    finally:
        acc.close()

    # ---- C3: 20 MHz slot of 64 small TBs, mixed base graphs / lifting sizes --------------------------------------------------
    acc = pusch.Accelerator(device=0, max_cbs_in_flight=4096, nof_harq_cb_slots=4096)
    acc.set_decoder_variant(args.decoder_variant)
    try:
        rng = np.random.default_rng(rng_seed + 1)
        cases = [(52, 2, 120, 1, 2, 0.9), (52, 2, 449, 1, 1, 2.0), (52, 4, 378, 1, 1, 3.0), (52, 4, 658, 1, 1, 6.0),
                 (25, 2, 120, 1, 2, 0.9), (10, 4, 490, 1, 2, 4.0), (4, 2, 308, 1, 2, 1.5), (1, 2, 120, 1, 2, 0.9)]
        cfgs, dev, nllrs, bits_total, slot, ref_time = [], [], [], 0, 0, 0.0
        for ue in range(64):
            prb, qm, R, nl, bg, mu = cases[ue % len(cases)]
            tbs = synth.tbs_for(prb, qm, R, nl)
            nllr = prb * 156 * qm * nl
            tb = rng.integers(0, 256, tbs // 8, dtype=np.uint8)
            l = synth.awgn_llrs(rng, synth.encode_tb(tb, bg, 0, qm, 25344, nl, nllr), mu)
            cfgs.append(capi.TbConfig(tbs, bg, 0, qm, 25344, nl, 6, 1, 1, slot))
            slot += len(pusch.segment(tbs, bg, qm, nl, nllr))
            dev.append(torch.from_numpy(l).cuda())
            nllrs.append(nllr)
            bits_total += tbs
            if ref is not None and ue < len(cases):
                g = ref_tb_gbps(tbs, nllr, l, bg, qm, 25344, nl, seconds=0.3)
                ref_time += 8 * tbs / (g * 1e9)  # each shape appears 8 times in the slot
        stage, res = run_dev(acc, cfgs, dev, nllrs, 5)
        kern = stage[1] + stage[2] + stage[3]
        # Throughput form: a slot of small TBs occupies a fraction of the GPU (its duration is the longest code block's chain of
        # layers), so the slots of several 20 MHz cells are decoded side by side: NCELL such slots (disjoint HARQ slots) in ONE
        # submit_tbs call, kernel spans of that batch.
        NCELL = 8
        many_cfgs = [capi.TbConfig(c_.tbs_bits, c_.base_graph, 0, c_.modulation, 25344, c_.nof_layers, 6, 1, 1,
                                   c_.harq_first_slot + k * slot) for k in range(NCELL) for c_ in cfgs]
        many_stage, many_res = run_dev(acc, many_cfgs, dev * NCELL, nllrs * NCELL, 4)
        many_kern = many_stage[1] + many_stage[2] + many_stage[3]
        out["c3_20mhz_mixed_small_tbs"] = {
            "workload": "64 UEs, 1..52 PRB QPSK/16QAM, BG1 + BG2, Z in {8, 48, 96, 208, 288, 320, 352}",
            "kernels_us_per_slot": kern * 1e3, "stage_ms": stage.tolist(), "info_gbit_per_s_kernels": bits_total / (kern * 1e-3) / 1e9,
            "cells_in_one_batch": {"cells": NCELL, "tbs": NCELL * 64, "kernels_us": many_kern * 1e3, "stage_ms": many_stage.tolist(),
                                   "info_gbit_per_s_kernels": NCELL * bits_total / (many_kern * 1e-3) / 1e9,
                                   "tb_crc_ok": int(sum(r.tb_crc_ok for r in many_res)),
                                   "what": "the slots of eight 20 MHz cells (512 small TBs, disjoint HARQ slots) in one submit_tbs call: "
                                           "the batch lasts about as long as one cell's slot, because its duration is the longest "
                                           "code block's chain of layers, not the amount of work"},
            "tb_crc_ok": int(sum(r.tb_crc_ok for r in res)),
            "reference": {"value": (bits_total / ref_time / 1e9) if ref_time else None, "unit": "Gbit/s", "cores": threads,
                          "what": "reference pusch_decoder_impl on all host threads, per shape, summed over the slot"}}
    finally:
        acc.close()

    # ---- C4: HARQ rv0 -> rv2 -> rv3, 64 UEs with config-2 sized TBs, soft combining in HBM -----------------------------------
    acc = pusch.Accelerator(device=0, max_cbs_in_flight=64 * 152, nof_harq_cb_slots=64 * 152)
    acc.set_decoder_variant(args.decoder_variant)
    try:
        rng = np.random.default_rng(rng_seed + 2)
        B, ncb, tbs, nllr = 64, 152, 1277992, 1362816
        tb = rng.integers(0, 256, tbs // 8, dtype=np.uint8)
        tx = []
        for i, rv in enumerate([0, 2, 3]):
            cw = synth.encode_tb(tb, 1, rv, 8, 12611, 4, nllr)
            host = np.stack([synth.awgn_llrs(rng, cw, 9.0) for _ in range(4)])
            if i == 0:
                host0 = host[0].copy()
            dev = torch.from_numpy(host).cuda()
            cfgs = [capi.TbConfig(tbs, 1, rv, 8, 12611, 4, 6, 1, int(i == 0), k * ncb) for k in range(B)]
            stage, res = run_dev(acc, cfgs, [dev[k % 4] for k in range(B)], [nllr] * B, 1)
            kern = stage[1] + stage[2] + stage[3]
            tx.append({"rv": rv, "kernels_ms": kern, "stage_ms": stage.tolist(), "tb_crc_ok": int(sum(r.tb_crc_ok for r in res)),
                       "observations": int(sum(r.nof_observations for r in res)),
                       "info_gbit_per_s_kernels": B * tbs / (kern * 1e-3) / 1e9})
            del dev
        out["c4_harq_rv0_rv2_rv3_64_ues"] = {
            "workload": "64 UEs x config-2 TB at mu = 9 (rv0 fails, combining decodes), GPU-resident soft buffers, early stop",
            "transmissions": tx,
            "reference": {"value": ref_tb_gbps(tbs, nllr, host0, 1, 8, 12611, 4), "unit": "Gbit/s", "cores": threads,
                          "what": "reference pusch_decoder_impl, all host threads, the rv0 transmission of this TB at mu = 9 (fails: 6 iterations)"}}
    finally:
        acc.close()

    # ---- PDSCH mirror: 64 TBs of config-2 size encoded in one launch -------------------------------------
    from srsran_projectvtlmo_b200 import pdsch
    enc = pdsch.EncoderAccelerator(device=0)
    try:
        rng = np.random.default_rng(rng_seed + 3)
        B, tbs, nbits = 64, 1277992, 1362816
        tb = [rng.integers(0, 256, tbs // 8, dtype=np.uint8) for _ in range(B)]
        cfgs = [pdsch.pdsch_encoder_configuration(1, 0, 8, 12611, 4, nbits // 8) for _ in range(B)]
        want = np.packbits(synth.encode_tb(tb[5], 1, 0, 8, 12611, 4, nbits))
        stage, wall = np.zeros(3), 0.0
        for rep in range(6):
            t0 = time.perf_counter()
            _, pk = pdsch.encode_tbs(enc, cfgs, tb, want_bits=False, want_packed=True)
            if rep:
                wall += time.perf_counter() - t0
                stage += np.array(enc.last_timing())
            assert np.array_equal(pk[5], want), "PDSCH code word differs from the transmitter restatement"
        stage /= 5
        wall /= 5
        pe = {"workload": "64 TBs x 273 PRB / 256QAM / 4 layers / Nref 12611 (152 code blocks of BG1 Z = 384 each), rv 0: TB CRC + "
                          "code-block CRCs + LDPC encoding + rate matching, packed code words back to the host",
              "kernels_ms": float(stage[1]), "stage_ms": stage.tolist(), "info_gbit_per_s_kernels": B * tbs / (stage[1] * 1e-3) / 1e9,
              "ms_host_buffers": wall * 1e3, "info_gbit_per_s_host_buffers": B * tbs / wall / 1e9,
              "hbm_bytes_algorithmic": B * (tbs // 8 + nbits + nbits // 8),
              "hbm_gbs": B * (tbs // 8 + nbits + nbits // 8) / (stage[1] * 1e-3) / 1e9,
              "note": "the kernel writes the code word one bit per byte (the reference's pdsch_encoder output format, "
                      "consumed by the modulation mapper) and a packing kernel re-reads it: E + E / 8 bytes out per code block"}
        if ref is not None:
            # The reference's Tx segmenter refuses this TBS (ldpc_segmenter_impl.cpp:77): the 2-layer TB of the same
            # allocation stands in for it, all host threads, one pdsch_encoder_impl each.
            from concurrent.futures import ThreadPoolExecutor
            tbs2, nbits2 = 638984, 681408
            tb2 = rng.integers(0, 256, tbs2 // 8, dtype=np.uint8)

            def work(n):
                for _ in range(n):
                    ob.ref_encode_tb(tb2, 1, 0, 8, 25223, 2, nbits2 // 8)
            work(2)
            n_each = 40
            t0 = time.perf_counter()
            with ThreadPoolExecutor(threads) as ex:
                list(ex.map(work, [n_each] * threads))
            dt = time.perf_counter() - t0
            pe["reference"] = {"value": threads * n_each * tbs2 / dt / 1e9, "unit": "Gbit/s info", "cores": threads,
                               "what": "reference pdsch_encoder_impl (segmenter + AVX2 ldpc_encoder + rate matcher), one instance per "
                                       "host thread, the 2-layer TB of the same allocation"}
        out["pdsch_encode_64_tbs"] = pe
    finally:
        enc.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--tbs-per-step", type=int, default=64)
    ap.add_argument("--mu", type=float, default=18.0, help="AWGN operating point of the synthetic LLRs")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--ref-seconds", type=float, default=20.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--latency-reps", type=int, default=1010)
    ap.add_argument("--bind-cpus", type=int, default=1, help="1: pin the process to the GPU's closest CPUs (NVML), 0: leave it")
    ap.add_argument("--depth", type=int, default=3, help="batches in flight (1..7)")
    ap.add_argument("--clock-period-ms", type=float, default=4.0, help="NVML clock sampling period inside the timed region")
    ap.add_argument("--min-seconds", type=float, default=2.0,
                    help="every timed region lasts at least this long: the K-step loop is repeated (ms_per_step is per step)")
    ap.add_argument("--slot-latency-slots", type=int, default=1010, help="slots of the 64-cell latency leg (0: skip)")
    ap.add_argument("--no-other-configs", action="store_true", help="skip BASELINE configs 1, 2 (worst case), 3, 4 (N = 1 only)")
    ap.add_argument("--no-symbols", action="store_true", help="skip the legs fed with equalized symbols (device-side demodulation)")
    ap.add_argument("--snr-db", type=float, default=30.5, help="operating point of the symbol-fed legs (256QAM, R = 0.926)")
    ap.add_argument("--decoder-variant", type=int, default=0, help="0 auto (tensor-memory packed decoder where it applies), 1 general kernel only, 2 packed decoder with the messages in shared memory (round-1 kernel), 3 one code block per CTA packed kernel everywhere, 4 pairs of code blocks per CTA, 5 without the many-layer pair form, 6 that form for every batch, 7 bulk-copy input staging")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    from srsran_projectvtlmo_b200 import capi, pusch, synth
    from srsran_projectvtlmo_b200.dispatch import HarqKey, TbDispatcher

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    all_cpus = os.sched_getaffinity(0)
    ncpus_bound = bind_to_gpu_cpus(local_rank) if args.bind_cpus else None
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    w = WORKLOAD
    B = args.tbs_per_step
    tbs, nllr, payloads, sets = make_inputs(B, 2, args.mu, 1000 + rank)
    ncb = 152
    # Two sets of HARQ slots used in turn, like HARQ processes rotating from one slot (TTI) to the next: consecutive batches
    # then share no soft-buffer state and the library lets them overlap on the GPU.
    acc = pusch.Accelerator(device=local_rank, max_cbs_in_flight=B * ncb, nof_harq_cb_slots=2 * B * ncb)
    acc.set_decoder_variant(args.decoder_variant)
    cfg_sets = [[capi.TbConfig(tbs, w["bg"], 0, w["qm"], w["nref"], w["layers"], w["max_it"], w["early_stop"], 1,
                               (s * B + i) * ncb) for i in range(B)] for s in range(2)]
    cfgs = cfg_sets[0]

    # Device-resident inputs (value leg) and page-locked host inputs (e2e leg).
    dev_sets = [torch.from_numpy(s).cuda() for s in sets]
    lib = capi.lib()
    host_sets = []

    def pinned(shape, dtype):
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = lib.srsran_cuda_pusch_dec_host_alloc(nbytes)
        if not p:
            raise SystemExit("pinned host allocation failed")
        return p, np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(nbytes,)).view(dtype).reshape(shape)

    pinned_ptrs = []
    for s in sets:
        p, buf = pinned(s.shape, np.int8)
        buf[...] = s
        host_sets.append((p, buf))
        pinned_ptrs.append(p)

    # Argument lists are built once: the timed loops below only make the library calls.
    dev_lists = [[(d[k].data_ptr(), nllr) for k in range(B)] for d in dev_sets]
    host_lists = [[buf[k] for k in range(B)] for _, buf in host_sets]
    dev_args = [pusch.SubmitArgs(cfg_sets[s], dev_lists[s], device_resident=True) for s in range(2)]
    host_args = [pusch.SubmitArgs(cfg_sets[s], host_lists[s]) for s in range(2)]

    def step_device(i):
        return pusch.submit_tbs(acc, dev_args[i % 2])

    def step_host(i):
        return pusch.submit_tbs(acc, host_args[i % 2])

    tb_outs = [np.zeros(tbs // 8, np.uint8) for _ in range(B)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def check_batch(tickets, what):
        """Correctness gate: EVERY transport block of the batch must pass its CRC and equal the transmitted payload."""
        res = pusch.poll_tbs(acc, tickets, tb_outs)
        for k, r in enumerate(res):
            if not r.tb_crc_ok:
                raise SystemExit(f"{what}: transport block {k} failed its CRC at the benchmark operating point")
            if not np.array_equal(tb_outs[k], payloads[k % len(payloads)]):
                raise SystemExit(f"{what}: decoded transport block {k} differs from the transmitted payload")
        return res

    DEPTH = max(1, min(args.depth, 7))  # batches in flight (the handle has 8 batch contexts)

    # ---- warm-up + correctness gate: every TB of every warm-up step must decode to its payload; every context is touched ----
    for i in range(max(args.warmup, 8)):
        check_batch(step_device(i), "warm-up (device-resident)")
    for i in range(max(args.warmup, 8)):
        check_batch(step_host(i), "warm-up (host buffers)")

    def timed_region(step_fn, settle_fn, device_timer):
        """K steps back to back (<= DEPTH in flight), repeated until the region has lasted --min-seconds. Returns
        (seconds, steps run). `device_timer`: the library's CUDA-event stopwatch, else the wall clock."""
        total, nsteps, reps = 0.0, 0, 0
        while total < args.min_seconds or reps == 0:
            inflight = []
            if device_timer:
                acc.timer_start()
            else:
                t0 = time.perf_counter()
            for i in range(args.steps):
                inflight.append(step_fn(i))
                if len(inflight) >= DEPTH:
                    settle_fn(inflight.pop(0))
            if device_timer:
                total += acc.timer_stop() * 1e-3
            while inflight:
                settle_fn(inflight.pop(0))
            if not device_timer:
                acc.synchronize()
                total += time.perf_counter() - t0
            nsteps += args.steps
            reps += 1
        return total, nsteps, reps

    # ---- value: device-resident inputs, device stopwatch -------------------------------------------------------------------
    # The inputs of consecutive steps alternate between two sets (2 x 87 MB of LLRs + 246 MB of soft buffers per step):
    # larger than the 126 MB L2, so no explicit flush is needed between timed steps.
    sampler = ClockSampler(local_rank, args.clock_period_ms)
    gc.collect()
    gc.disable()  # no collector pauses of the submitting thread inside the timed regions
    barrier()
    sampler.start()
    launches0 = acc.launch_count
    stage = np.zeros(5)
    counters = {"ok": 0, "tbs": 0}
    iters = []

    def settle(tk):
        res = pusch.poll_tbs(acc, tk)
        stage[:] += np.array(pusch.ticket_timing(acc, tk[0]))
        for r in res:
            counters["ok"] += r.tb_crc_ok
            iters.append(r.iter_mean)
        counters["tbs"] += len(res)

    value_s, value_steps, value_reps = timed_region(step_device, settle, True)
    launches = acc.launch_count - launches0
    clocks = sampler.stop()
    barrier()
    if counters["ok"] != counters["tbs"]:
        raise SystemExit(f"value leg: {counters['tbs'] - counters['ok']} of {counters['tbs']} transport blocks failed their CRC")
    step_ms = value_s * 1e3 / value_steps
    stage_ms = (stage / value_steps).tolist()
    value_ok, value_tbs = counters["ok"], counters["tbs"]

    # ---- per-stage times with ONE batch in flight (not part of `value`): in the timed region above consecutive batches
    # overlap on the GPU, so the span of a short stage there includes the time it shared the SMs with its neighbour's decoder.
    iso = np.zeros(5)
    niso = 6
    for i in range(niso):
        tk = step_device(i)
        pusch.poll_tbs(acc, tk)
        iso += np.array(pusch.ticket_timing(acc, tk[0]))
    iso_ms = (iso / niso).tolist()

    # ---- the same device-resident leg with the decoded TBs LEFT IN HBM (a consumer on the device side of the link): shows how
    # much of `value` the return path of the decoded bits costs - at eight GPUs it is the box's PCIe fabric that limits the
    # step, not the kernels. CRC verdicts and statistics still come back and are still enforced.
    acc.set_tb_host_copy(False)
    counters = {"ok": 0, "tbs": 0}
    for i in range(4):
        settle(step_device(i))
    counters = {"ok": 0, "tbs": 0}
    saved_min = args.min_seconds
    args.min_seconds = min(args.min_seconds, 1.0)
    barrier()
    hbm_s, hbm_steps, _ = timed_region(step_device, settle, True)
    barrier()
    args.min_seconds = saved_min
    acc.set_tb_host_copy(True)
    if counters["ok"] != counters["tbs"]:
        raise SystemExit("value leg (TBs left in HBM): transport blocks failed their CRC")
    hbm_local = torch.tensor([hbm_s * 1e3 / hbm_steps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(hbm_local, op=dist.ReduceOp.MAX)
    hbm_step_ms = float(hbm_local.item())

    # ---- e2e: host LLRs in pinned memory, H2D + kernels + D2H of TB bytes and results, <= DEPTH batches in flight ------
    counters = {"ok": 0, "tbs": 0}

    def drain(tickets):
        """Completion of one batch through the host API: TB bytes copied out, results read."""
        res = pusch.poll_tbs(acc, tickets, tb_outs)
        counters["ok"] += sum(r.tb_crc_ok for r in res)
        counters["tbs"] += len(res)

    barrier()
    e2e_s, e2e_steps, _ = timed_region(step_host, drain, False)
    barrier()
    if counters["ok"] != counters["tbs"]:
        raise SystemExit(f"e2e leg: {counters['tbs'] - counters['ok']} of {counters['tbs']} transport blocks failed their CRC")
    for i in range(2):  # and the bytes, outside the timed region (a 10 MB compare per step would be timed otherwise)
        check_batch(step_host(i), "e2e leg")

    # ---- PCIe ceiling of the e2e leg, measured here and now (all ranks at the same time): the step's H2D payload from the
    # same pinned buffers, the step's D2H payload back --------------------------------------------------------------------
    barrier()
    scratch = torch.empty(B * nllr, dtype=torch.int8, device="cuda")
    host_t = torch.from_numpy(host_sets[0][1].reshape(-1))
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(10):
        scratch.copy_(host_t, non_blocking=True)
    ev1.record()
    torch.cuda.synchronize()
    h2d_gbs = 10 * B * nllr / (ev0.elapsed_time(ev1) * 1e-3) / 1e9
    del scratch
    barrier()

    # ---- symbol-fed legs: equalized symbols + noise variances in, TBs out (soft demodulation, descrambling and UL-SCH
    # demultiplexing on the device) -------------------------------------------------------------------
    symbols = None
    if not args.no_symbols:
        rps = [w["prb"] * 12 if s != 2 else 0 for s in range(14)]  # DM-RS in symbol 2, two CDM groups without data: 13 x 3276 REs
        nsym = sum(rps) * w["layers"]
        assert nsym * w["qm"] == nllr
        srng = np.random.default_rng(2000 + rank)
        sigma2 = 10 ** (-args.snr_db / 10)
        sym_sets, nv_sets, dms = [], [], []
        ndist = min(4, B)
        cw_scr = []
        for k in range(ndist):
            cw = synth.encode_tb(payloads[k], w["bg"], 0, w["qm"], w["nref"], w["layers"], nllr)
            cw_scr.append(synth.modulate(cw ^ synth.scrambling_sequence(((0x4601 + k) << 15) + 100 + k, nllr), w["qm"]))
        dms = [pusch.demod_config(w["qm"], 0x4601 + (i % ndist), 100 + (i % ndist), w["layers"], rps) for i in range(B)]
        for s in range(2):
            ps, sbuf = pinned((B, nsym), np.complex64)
            pn, nbuf = pinned((B, nsym), np.float32)
            pinned_ptrs += [ps, pn]
            for i in range(B):
                noise = (srng.standard_normal(nsym, dtype=np.float32) + 1j * srng.standard_normal(nsym, dtype=np.float32))
                sbuf[i] = cw_scr[i % ndist] + np.float32(np.sqrt(sigma2 / 2)) * noise
                nbuf[i] = np.float32(sigma2)
            sym_sets.append(sbuf)
            nv_sets.append(nbuf)
        dsym = [torch.from_numpy(s.view(np.float32)).cuda() for s in sym_sets]
        dnv = [torch.from_numpy(s).cuda() for s in nv_sets]
        sym_dev_args = [pusch.SubmitSymbolArgs(cfg_sets[s], dms, [dsym[s][i].data_ptr() for i in range(B)],
                                               [dnv[s][i].data_ptr() for i in range(B)], device_resident=True) for s in range(2)]
        sym_host_args = [pusch.SubmitSymbolArgs(cfg_sets[s], dms, [sym_sets[s][i] for i in range(B)],
                                                [nv_sets[s][i] for i in range(B)]) for s in range(2)]
        for i in range(4):
            check_batch(pusch.submit_tbs_symbols(acc, sym_dev_args[i % 2]), "warm-up (device-resident symbols)")
            check_batch(pusch.submit_tbs_symbols(acc, sym_host_args[i % 2]), "warm-up (host symbols)")
        counters = {"ok": 0, "tbs": 0}
        sym_iters, dm_ms = [], []

        def settle_sym(tk):
            res = pusch.poll_tbs(acc, tk)
            dm_ms.append(pusch.ticket_demod_ms(acc, tk[0]))
            counters["ok"] += sum(r.tb_crc_ok for r in res)
            counters["tbs"] += len(res)
            sym_iters.extend(r.iter_mean for r in res)

        barrier()
        sv_s, sv_steps, _ = timed_region(lambda i: pusch.submit_tbs_symbols(acc, sym_dev_args[i % 2]), settle_sym, True)
        barrier()
        if counters["ok"] != counters["tbs"]:
            raise SystemExit("symbol-fed value leg: a transport block failed its CRC")
        dm_one = []
        for i in range(4):  # demodulation stage of a batch processed alone
            tk = pusch.submit_tbs_symbols(acc, sym_dev_args[i % 2])
            pusch.poll_tbs(acc, tk)
            dm_one.append(pusch.ticket_demod_ms(acc, tk[0]))
        counters = {"ok": 0, "tbs": 0}
        barrier()
        se_s, se_steps, _ = timed_region(lambda i: pusch.submit_tbs_symbols(acc, sym_host_args[i % 2]), drain, False)
        barrier()
        if counters["ok"] != counters["tbs"]:
            raise SystemExit("symbol-fed e2e leg: a transport block failed its CRC")
        symbols = {"sv_ms": sv_s * 1e3 / sv_steps, "se_s": se_s, "se_steps": se_steps, "nsym": nsym,
                   "iter_mean": float(np.mean(sym_iters)), "demod_ms_overlapped": float(np.mean(dm_ms)),
                   "demod_ms_alone": float(np.mean(dm_one))}
        del dsym, dnv

    gc.enable()

    # ---- single-TB latency through the host API on an otherwise idle GPU ---------------------------------------------
    lat = []
    one = host_sets[0][1][0]
    tb_out = tb_outs[0]
    for i in range(args.latency_reps):
        t1 = time.perf_counter()
        tk = pusch.submit_tbs(acc, cfgs[:1], [one])
        pusch.poll_tb(acc, tk[0], tb_out)
        lat.append((time.perf_counter() - t1) * 1e6)
    lat = np.array(lat[10:]) if len(lat) > 20 else np.array(lat)

    # ---- BASELINE config 5: 64 cells x one config-2 TB per 30 kHz slot, sharded over the N GPUs by the HARQ-sticky
    # dispatcher (every rank evaluates the same pure function, no communication), every rank submitting its share of the same
    # slot at the same time; one slot at a time, host LLRs in, TB bytes out. Slot n + 1 retransmits nothing (all TBs decode),
    # so every HARQ process is released and re-hashed: the shares vary from slot to slot like in a live cell mix. --------------
    def slot_leg(resident):
        """One slot at a time; `resident`: the soft bits are already in HBM (born there by a device-side demodulator)."""
        lats, shares = [], []
        disp = TbDispatcher(world, balance=True)
        barrier()
        for s in range(args.slot_latency_slots):
            disp.begin_slot()
            mine = [c for c in range(64) if disp.assign(HarqKey(c, 0x4601 + (s % 16), s % 8), True, ncb) == rank]
            for c in range(64):
                disp.release(HarqKey(c, 0x4601 + (s % 16), s % 8))
            shares.append(len(mine))
            if not mine:
                lats.append(0.0)
                continue
            t1 = time.perf_counter()
            # Two pieces: the H2D copy of the second overlaps the kernels of the first.
            half = (len(mine) + 1) // 2
            tk = []
            for piece in (mine[:half], mine[half:]):
                if piece:
                    src = dev_lists[s % 2] if resident else host_sets[s % 2][1]
                    tk += pusch.submit_tbs(acc, [cfg_sets[s % 2][k % B] for k in piece], [src[k % B] for k in piece],
                                           device_resident=resident)
            res = pusch.poll_tbs(acc, tk)
            lats.append((time.perf_counter() - t1) * 1e6)
            if not all(r.tb_crc_ok for r in res):
                raise SystemExit("64-cell slot leg: a transport block failed its CRC")
        barrier()
        return (np.array(lats[10:]) if len(lats) > 20 else np.array(lats)), shares

    slot_lat, share_sizes, slot_lat_res = np.array([]), [], np.array([])
    if args.slot_latency_slots > 0:
        slot_lat, share_sizes = slot_leg(False)
        slot_lat_res, _ = slot_leg(True)

    # ---- max over ranks ------------------------------------------------------------------------------------------------
    vals = [step_ms, e2e_s / e2e_steps, -h2d_gbs,
            symbols["sv_ms"] if symbols else 0.0, (symbols["se_s"] / symbols["se_steps"]) if symbols else 0.0,
            pct(slot_lat, 50) or 0.0, pct(slot_lat, 99) or 0.0, float(slot_lat.max()) if slot_lat.size else 0.0,
            pct(slot_lat_res, 50) or 0.0, pct(slot_lat_res, 99) or 0.0, float(slot_lat_res.max()) if slot_lat_res.size else 0.0]
    red = torch.tensor(vals, dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
    step_ms_max, e2e_step_s_max, h2d_gbs_min = float(red[0]), float(red[1]), -float(red[2])
    n_gpus = world
    value = n_gpus * B * tbs / (step_ms_max * 1e-3) / 1e9
    e2e = n_gpus * B * tbs / e2e_step_s_max / 1e9

    if rank == 0:
        peaks, peak_src = measured_peaks()
        mean_it = float(np.mean(iters)) if iters else 0.0
        # Algorithmic work of the decoder: edge updates U = iterations * Z * sum(deg of processed layers),
        # 4 bytes of shared-memory traffic each (soft read + write, message read + write). Layers processed here: 4.
        edges = sum(BG1_DEG[:4])
        U = mean_it * 384 * edges * ncb * B
        dec_s = iso_ms[2] * 1e-3
        sm_clk = (clocks.get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)) * 1e6
        smem_peak = 128.0 * 148 * sm_clk / 1e9
        smem_ach = 4.0 * U / dec_s / 1e9 if dec_s > 0 else 0.0
        # Dematch: read E, write the reference's write set W = 12611 bytes per code block.
        dm_bytes = B * (nllr + ncb * 12611)
        dm_s = iso_ms[1] * 1e-3
        hbm_ach = dm_bytes / dm_s / 1e9 if dm_s > 0 else 0.0
        prof = {}
        pf = ROOT / "profiles" / "r2_decode_ncu_counters.json"
        if pf.exists():
            prof = json.loads(pf.read_text())
        e2e_ceiling = B * tbs / (B * nllr / (h2d_gbs_min * 1e9)) / 1e9 * n_gpus
        line = {
            "metric": "pusch_decoded_info_gbit_per_s", "value": value, "unit": "Gbit/s", "n_gpus": n_gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms_max, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int8", "data": "synthetic",
            "config": {"workload": w["name"], "tbs_bits": tbs, "codeblocks_per_tb": ncb, "tbs_per_step_per_gpu": B,
                       "lifting_size": 384, "base_graph": 1, "max_iterations": w["max_it"], "early_stop": True,
                       "mu": args.mu, "arithmetic": "int8 LLR algebra of the reference, bit-exact; the decoder holds the (integer) values in binary16 lanes",
                       "decoder_variant": args.decoder_variant, "cpus_bound_to_gpu": ncpus_bound, "mean_iterations": mean_it,
                       "tb_crc_ok_fraction": value_ok / max(value_tbs, 1),
                       "correctness_gate": "every TB of every warm-up step and of two post-region steps equals its payload; every TB of the timed regions passed its CRC (else the run aborts)",
                       "timed_region": {"seconds": value_s, "steps_run": value_steps, "repeats_of_the_K_step_loop": value_reps,
                                        "min_seconds": args.min_seconds},
                       "timing": f"device stopwatch (CUDA events on the library streams) over all steps, <= {DEPTH} batches in flight; "
                                 "inputs alternate between two sets larger than L2 (no flush needed)"},
            "value_tbs_left_in_hbm": {
                "value": n_gpus * B * tbs / (hbm_step_ms * 1e-3) / 1e9, "unit": "Gbit/s", "ms_per_step": hbm_step_ms,
                "seconds": hbm_s, "steps_run": hbm_steps,
                "what": "the `value` leg with the decoded transport blocks left in HBM (srsran_cuda_pusch_dec_set_tb_host_copy(0): a "
                        "consumer on the device side of the link); CRC verdicts and statistics still return and are enforced. The "
                        "difference to `value` is the cost of returning 10.4 MB of decoded bits per step and GPU over PCIe"},
            "e2e": {"value": e2e, "unit": "Gbit/s", "h2d_bytes_per_step": B * nllr,
                    "d2h_bytes_per_step": B * (tbs // 8 + 3 + 8 + ncb * 16), "seconds": e2e_s, "steps_run": e2e_steps,
                    "pcie_h2d_gbs_measured": h2d_gbs_min, "ceiling_from_h2d": e2e_ceiling,
                    "frac_of_ceiling": e2e / e2e_ceiling if e2e_ceiling else None,
                    "ceiling_note": "the step's H2D payload copied from the same pinned buffers by every rank at once (min over ranks): int8 soft bits are 1.066 B per decoded info bit",
                    "note": f"pinned host LLRs -> submit_tbs -> poll_tbs (TB bytes + results), <= {DEPTH} batches in flight, wall clock"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "stage_ms": {"h2d_descriptors": stage_ms[0], "rate_dematch": stage_ms[1], "ldpc_decode": stage_ms[2],
                         "tb_assemble_crc": stage_ms[3], "d2h_results": stage_ms[4],
                         "note": "event spans inside the timed region; batches overlap there, so spans sum to more than ms_per_step"},
            "stage_ms_one_batch_in_flight": {"h2d_descriptors": iso_ms[0], "rate_dematch": iso_ms[1], "ldpc_decode": iso_ms[2],
                                             "tb_assemble_crc": iso_ms[3], "d2h_results": iso_ms[4]},
            "roofline": {"kernel": "ldpc_decode4_kernel<384,384,2,1> (messages in tensor memory, two CTAs per SM)", "bound": "smem",
                         "achieved": smem_ach, "peak": smem_peak,
                         "unit": "GB/s", "frac": smem_ach / smem_peak if smem_peak else None,
                         "traffic": prof.get("dram_bytes_per_launch_64_tbs"),
                         "traffic_note": "dram__bytes_read + dram__bytes_write of one launch (64 TBs), from_profile: "
                                         + prof.get("source", "none committed"),
                         "peak_source": "128 B/clk/SM x 148 SMs x SM clock sampled during the run (B300_MICROARCH.md: "
                                        "smem crossbar 128 B/cyc/SM)",
                         "algorithmic": f"4 B per edge update (SURVEY 8d), U = {mean_it:.2f} it x 384 x {edges} edges x {ncb * B} CBs",
                         "share_of_step": iso_ms[2] / sum(iso_ms[1:4]) if sum(iso_ms[1:4]) else None,
                         "share_note": "decode span / kernel spans (dematch + decode + TB assembly) of a batch processed alone"},
            # Not a distance to a bound but a utilisation figure: the kernel's OWN instruction count per edge update against
            # the schedulers' issue rate (one warp instruction per clock and scheduler).
            "issue_utilisation": {"kernel": "ldpc_decode4_kernel", "what": "executed thread-instructions per second / (148 SMs x 4 schedulers x 32 lanes x clock)",
                                  "instructions_per_edge_update": prof.get("thread_instructions_per_edge_update"),
                                  "value": ((prof["thread_instructions_per_edge_update"] * U / dec_s) / (148 * 4 * 32 * sm_clk))
                                  if prof.get("thread_instructions_per_edge_update") and dec_s > 0 else None,
                                  "from_profile": prof.get("source"),
                                  "ncu_issue_slots_busy": prof.get("issue_slots_busy"), "ncu_alu_pipe": prof.get("alu_pipe"),
                                  "ncu_fma_pipe_fp16": prof.get("fma_pipe_fp16"), "ncu_occupancy": prof.get("achieved_occupancy")},
            "roofline_dematch": {"kernel": "rate_dematch_kernel", "bound": "hbm", "achieved": hbm_ach,
                                 "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                                 "frac": hbm_ach / peaks["hbm_gbs"] if peaks.get("hbm_gbs") else None, "traffic": None,
                                 "peak_source": peak_src, "algorithmic": "E + 12611 B per code block",
                                 "timed": "batches processed one at a time after the timed region (stage_ms_one_batch_in_flight)"},
            "tb_latency_us": {"p50": pct(lat, 50), "p99": pct(lat, 99), "max": float(lat.max()) if lat.size else None,
                              "n": int(lat.size),
                              "what": "one TB, host LLRs -> TB bytes, idle GPU, wall clock; slot budget 500 us"},
            "slot_latency_64_cells_us": {
                "p50": float(red[5]), "p99": float(red[6]), "max": float(red[7]), "n": int(slot_lat.size), "budget_us": 500,
                "tbs_per_gpu_per_slot_mean": float(np.mean(share_sizes)) if share_sizes else None,
                "what": "BASELINE config 5: 64 cells x one config-2 TB per slot, sharded by TbDispatcher (sticky HARQ, least-loaded) over the N GPUs, all ranks at once; host LLRs in (two pieces per share), TB results out; max over ranks of each rank's percentile"},
            "slot_latency_64_cells_resident_us": {
                "p50": float(red[8]), "p99": float(red[9]), "max": float(red[10]), "n": int(slot_lat_res.size), "budget_us": 500,
                "what": "the same slot with the soft bits already in HBM (born there by a device-side demodulator, SURVEY 8(f) row 2): "
                        "descriptors in, TB bytes and results out; what is left of the slot latency when the 1.36 MB of soft bits per "
                        "TB do not cross the PCIe fabric"},
        }
        if symbols:
            sv = n_gpus * B * tbs / (float(red[3]) * 1e-3) / 1e9
            se = n_gpus * B * tbs / float(red[4]) / 1e9
            line["from_symbols"] = {
                "what": "equalized symbols (complex64) + noise variances (float32) in, TBs out: soft demodulation + descrambling + UL-SCH demultiplexing on the device in front of the same path",
                "value_device_resident": sv, "e2e_host_symbols": se, "unit": "Gbit/s", "snr_db": args.snr_db,
                "mean_iterations": symbols["iter_mean"],
                "h2d_bytes_per_step": B * symbols["nsym"] * 12, "d2h_bytes_per_step": B * (tbs // 8 + 3 + 8 + ncb * 16),
                "demod_stage_ms_alone": symbols["demod_ms_alone"], "demod_stage_ms_overlapped": symbols["demod_ms_overlapped"],
                "roofline_demod": {"kernel": "pusch_demod_kernel + scr_seq_kernel", "bound": "hbm",
                                   "achieved": B * symbols["nsym"] * (12 + w["qm"]) / (symbols["demod_ms_alone"] * 1e-3) / 1e9,
                                   "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                                   "frac": (B * symbols["nsym"] * (12 + w["qm"]) / (symbols["demod_ms_alone"] * 1e-3) / 1e9) / peaks["hbm_gbs"]
                                   if peaks.get("hbm_gbs") else None,
                                   "algorithmic": "12 B read (symbol + noise variance) + Qm B written per symbol"},
                "note": "binary32 symbols are 12 B per 8 soft bits: over PCIe this input is 1.5x larger than the int8 soft bits, so the host-fed number is lower than `e2e`; the point of the row is the device-resident case (the equalizer's output already in HBM), where no soft bit crosses PCIe"}
        if n_gpus == 1 and not args.no_other_configs:
            line["other_configs"] = other_configs(args)
        if n_gpus == 1 and not args.no_cpu_baseline:
            os.sched_setaffinity(0, all_cpus)  # the CPU baseline gets every host thread, not only the GPU's neighbours
            line["cpu_baseline"] = cpu_baseline(args, tbs, nllr, sets[0][0])
        print(json.dumps(line), flush=True)

    for p in pinned_ptrs:
        lib.srsran_cuda_pusch_dec_host_free(p)
    acc.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
