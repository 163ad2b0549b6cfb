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
stream the kernels run on; `e2e` = the same through the host-buffer C ABI with H2D/D2H inside the timed region.
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
                self.rows.append((sm, rs))
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
        reasons = sorted(n for n, bit in names.items() if any(r & bit for _, r in self.rows))
        sm = [float(c) for c, _ in self.rows]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": reasons,
                "samples": len(sm), "source": "NVML, sampled inside the timed value region"}


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
    reps = max(1, int(args.ref_seconds / max(t1, 1e-3)))
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
    ap.add_argument("--latency-reps", type=int, default=200)
    ap.add_argument("--bind-cpus", type=int, default=1, help="1: pin the process to the GPU's closest CPUs (NVML), 0: leave it")
    ap.add_argument("--depth", type=int, default=3, help="batches in flight (1..5)")
    ap.add_argument("--clock-period-ms", type=float, default=4.0, help="NVML clock sampling period inside the timed region")
    ap.add_argument("--decoder-variant", type=int, default=0, help="0 auto, 1 general kernel only, 2 packed groups with 2 threads per check, 3 one code block per CTA packed kernel everywhere, 4 pairs of code blocks per CTA (two CTAs per SM)")
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

    from srsran_projectvtlmo_b200 import capi, pusch

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
    for s in sets:
        p = lib.srsran_cuda_pusch_dec_host_alloc(s.size)
        if not p:
            raise SystemExit("pinned host allocation failed")
        buf = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int8)), shape=s.shape)
        buf[...] = s
        host_sets.append((p, buf))

    # Argument lists are built once: the timed loops below only make the library calls.
    dev_lists = [[(d[k].data_ptr(), nllr) for k in range(B)] for d in dev_sets]
    host_lists = [[buf[k] for k in range(B)] for _, buf in host_sets]

    dev_args = [pusch.SubmitArgs(cfg_sets[s], dev_lists[s], device_resident=True) for s in range(2)]
    host_args = [pusch.SubmitArgs(cfg_sets[s], host_lists[s]) for s in range(2)]

    def step_device(i):
        return pusch.submit_tbs(acc, dev_args[i % 2])

    def step_host(i):
        return pusch.submit_tbs(acc, host_args[i % 2])

    tb_out = np.zeros(tbs // 8, np.uint8)

    tb_outs = [np.zeros(tbs // 8, np.uint8) for _ in range(B)]

    def drain(tickets):
        """Completion of one batch through the host API: TB bytes copied out, results read."""
        return sum(r.tb_crc_ok for r in pusch.poll_tbs(acc, tickets, tb_outs))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    DEPTH = args.depth  # batches in flight (the handle has 6 batch contexts)

    # ---- warm-up + correctness gate: every TB of the warm-up must decode to its payload; every batch context is touched ----
    for i in range(max(args.warmup, 6)):  # every batch context of the handle allocates its buffers on first use
        tk = step_device(i)
        for k, t in enumerate(tk):
            r = pusch.poll_tb(acc, t, tb_out)
            if i == 0 and r.tb_crc_ok and not np.array_equal(tb_out, payloads[k % len(payloads)]):
                raise SystemExit("decoded TB differs from the transmitted payload")
    for i in range(max(args.warmup, 6)):
        drain(step_host(i))

    # ---- value: device-resident inputs, K steps back to back (<= DEPTH in flight), device stopwatch ---------------------
    # The inputs of consecutive steps alternate between two sets (2 x 87 MB of LLRs + 246 MB of soft buffers per step):
    # larger than the 126 MB L2, so no explicit flush is needed between timed steps.
    sampler = ClockSampler(local_rank, args.clock_period_ms)
    gc.collect()
    gc.disable()  # no collector pauses of the submitting thread inside the timed regions
    barrier()
    sampler.start()
    launches0 = acc.launch_count
    stage = np.zeros(5)
    ok_tbs = 0
    iters = []

    def settle(tk):
        nonlocal ok_tbs, stage
        res = pusch.poll_tbs(acc, tk)
        stage += np.array(pusch.ticket_timing(acc, tk[0]))
        for r in res:
            ok_tbs += r.tb_crc_ok
            iters.append(r.iter_mean)

    acc.timer_start()
    inflight = []
    for i in range(args.steps):
        inflight.append(step_device(i))
        if len(inflight) >= DEPTH:
            settle(inflight.pop(0))
    total_ms = acc.timer_stop()
    while inflight:
        settle(inflight.pop(0))
    launches = acc.launch_count - launches0
    clocks = sampler.stop()
    barrier()
    step_ms = total_ms / args.steps
    stage_ms = (stage / args.steps).tolist()

    # ---- per-stage times with ONE batch in flight (not part of `value`): in the timed region above consecutive batches
    # overlap on the GPU, so the span of a short stage there includes the time it shared the SMs with its neighbour's decoder.
    iso = np.zeros(5)
    niso = min(args.steps, 6)
    for i in range(niso):
        tk = step_device(i)
        pusch.poll_tbs(acc, tk)
        iso += np.array(pusch.ticket_timing(acc, tk[0]))
    iso_ms = (iso / niso).tolist()

    # ---- e2e: host LLRs in pinned memory, H2D + kernels + D2H of TB bytes and results, <= DEPTH batches in flight ------
    barrier()
    t0 = time.perf_counter()
    inflight = []
    for i in range(args.steps):
        inflight.append(step_host(i))
        if len(inflight) >= DEPTH:
            drain(inflight.pop(0))
    while inflight:
        drain(inflight.pop(0))
    acc.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    gc.enable()

    # ---- single-TB latency through the host API on an otherwise idle GPU ---------------------------------------------
    lat = []
    one = host_sets[0][1][0]
    for i in range(args.latency_reps):
        t1 = time.perf_counter()
        tk = pusch.submit_tbs(acc, cfgs[:1], [one])
        pusch.poll_tb(acc, tk[0], tb_out)
        lat.append((time.perf_counter() - t1) * 1e6)
    lat = np.array(lat[10:]) if len(lat) > 20 else np.array(lat)

    # ---- max over ranks ------------------------------------------------------------------------------------------------
    red = torch.tensor([step_ms, e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
    step_ms_max, e2e_s_max = float(red[0]), float(red[1])
    n_gpus = world
    value = n_gpus * B * tbs / (step_ms_max * 1e-3) / 1e9
    e2e = n_gpus * B * tbs * args.steps / e2e_s_max / 1e9

    if rank == 0:
        peaks, peak_src = measured_peaks()
        mean_it = float(np.mean(iters)) if iters else 0.0
        # Algorithmic work of the decoder: edge updates U = iterations * Z * sum(deg of processed layers),
        # 4 bytes of shared-memory traffic per edge update. Layers processed here: 4 (E = 8960/8992 <= 8976 + ...).
        edges = sum(BG1_DEG[:4])
        U = mean_it * 384 * edges * ncb * B
        dec_s = stage_ms[2] * 1e-3
        sm_clk = (clocks.get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)) * 1e6
        smem_peak = 128.0 * 148 * sm_clk / 1e9
        smem_ach = 4.0 * U / dec_s / 1e9 if dec_s > 0 else 0.0
        # Dematch: read E, write the reference's write set W = 12611 bytes per code block.
        dm_bytes = B * (nllr + ncb * 12611)
        dm_s = iso_ms[1] * 1e-3
        hbm_ach = dm_bytes / dm_s / 1e9 if dm_s > 0 else 0.0
        line = {
            "metric": "pusch_decoded_info_gbit_per_s", "value": value, "unit": "Gbit/s", "n_gpus": n_gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms_max, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int8", "data": "synthetic",
            "config": {"workload": w["name"], "tbs_bits": tbs, "codeblocks_per_tb": ncb, "tbs_per_step_per_gpu": B,
                       "lifting_size": 384, "base_graph": 1, "max_iterations": w["max_it"], "early_stop": True,
                       "mu": args.mu, "arithmetic": "int8 LLR algebra of the reference, bit-exact; the decoder holds the (integer) values in binary16 lanes", "decoder_variant": args.decoder_variant, "cpus_bound_to_gpu": ncpus_bound, "mean_iterations": mean_it, "tb_crc_ok_fraction": ok_tbs / (B * args.steps),
                       "timing": f"device stopwatch (CUDA events on the library streams) over all steps, <= {DEPTH} batches in flight; "
                                 "inputs alternate between two sets larger than L2 (no flush needed)"},
            "e2e": {"value": e2e, "unit": "Gbit/s", "h2d_bytes_per_step": B * nllr,
                    "d2h_bytes_per_step": B * (tbs // 8 + 3 + 8 + ncb * 16),
                    "note": f"pinned host LLRs -> submit_tbs -> poll_tb (TB bytes + results), <= {DEPTH} batches in flight, wall clock"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "stage_ms": {"h2d_descriptors": stage_ms[0], "rate_dematch": stage_ms[1], "ldpc_decode": stage_ms[2],
                         "tb_assemble_crc": stage_ms[3], "d2h_results": stage_ms[4],
                         "note": "event spans inside the timed region; batches overlap there, so spans sum to more than ms_per_step"},
            "stage_ms_one_batch_in_flight": {"h2d_descriptors": iso_ms[0], "rate_dematch": iso_ms[1], "ldpc_decode": iso_ms[2],
                                             "tb_assemble_crc": iso_ms[3], "d2h_results": iso_ms[4]},
            "roofline": {"kernel": "ldpc_decode4_kernel", "bound": "smem", "achieved": smem_ach, "peak": smem_peak,
                         "unit": "GB/s", "frac": smem_ach / smem_peak if smem_peak else None,
                         "traffic": 94.7e6 * B / 64, "traffic_note": "dram__bytes_read+write of one launch (ncu, 64 TBs): "
                         "the soft buffers are read once (89 MB), decoded bits written (6 MB)",
                         "peak_source": "128 B/clk/SM x 148 SMs x SM clock sampled during the run (B300_MICROARCH.md: "
                                        "smem crossbar 128 B/cyc/SM)",
                         "algorithmic": f"4 B per edge update, U = {mean_it:.2f} it x 384 x {edges} edges x {ncb * B} CBs",
                         "share_of_step": iso_ms[2] / sum(iso_ms[1:4]) if sum(iso_ms[1:4]) else None,
                         "share_note": "decode span / kernel spans (dematch + decode + TB assembly) of a batch processed alone"},
            # The decoder is bound by instruction issue, not by shared-memory bandwidth: the layer body executes 46.8 warp
            # instructions per (thread, edge) for four code blocks = 11.7 per code-block edge update (ncu instruction counts
            # of the layer lines, profiles/r1_v8_packed_decode_ncu_summary.txt), split over the ALU pipe (HMNMX2, HSET2, PRMT,
            # LOP3), the FMA pipe (HFMA2, HADD2, IMAD) and the LSU; a scheduler issues at most one warp instruction per clock.
            "roofline_issue": {"kernel": "ldpc_decode4_kernel", "bound": "instruction-issue",
                               "achieved": 11.7 * U / dec_s / 1e12 if dec_s > 0 else 0.0,
                               "peak": 148 * 4 * 32 * sm_clk / 1e12, "unit": "T thread-instr/s",
                               "frac": (11.7 * U / dec_s) / (148 * 4 * 32 * sm_clk) if dec_s > 0 else None,
                               "algorithmic": "11.7 instructions per edge update x U (mean iterations, not executed ones)",
                               "ncu_issue_slots_busy": 0.53, "ncu_alu_pipe_busy": 0.44, "ncu_fma_pipe_busy": 0.56},
            "roofline_dematch": {"kernel": "rate_dematch_kernel", "bound": "hbm", "achieved": hbm_ach,
                                 "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                                 "frac": hbm_ach / peaks["hbm_gbs"] if peaks.get("hbm_gbs") else None, "traffic": None,
                                 "peak_source": peak_src, "algorithmic": "E + 12611 B per code block",
                                 "timed": "batches processed one at a time after the timed region (stage_ms_one_batch_in_flight)"},
            "tb_latency_us": {"p50": float(np.percentile(lat, 50)) if lat.size else None,
                              "p99": float(np.percentile(lat, 99)) if lat.size else None,
                              "max": float(lat.max()) if lat.size else None, "n": int(lat.size),
                              "what": "one TB, host LLRs -> TB bytes, idle GPU, wall clock; slot budget 500 us"},
        }
        if n_gpus == 1 and not args.no_cpu_baseline:
            os.sched_setaffinity(0, all_cpus)  # the CPU baseline gets every host thread, not only the GPU's neighbours
            line["cpu_baseline"] = cpu_baseline(args, tbs, nllr, sets[0][0])
        print(json.dumps(line), flush=True)

    for p, _ in host_sets:
        lib.srsran_cuda_pusch_dec_host_free(p)
    acc.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
