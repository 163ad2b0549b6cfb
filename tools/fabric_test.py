"""Host <-> device fabric ceiling of the box: N processes (one per GPU, torchrun), each copying the benchmark's per-step
payloads between pinned host memory and its own GPU AT THE SAME TIME - the 87 MB of int8 soft bits host -> device and the
10.4 MB of transport-block bytes device -> host that one step of `bench.py`'s host-buffer leg moves (and the 131 MB of
equalized symbols + noise variances of its symbol-fed leg). Prints one JSON line per configuration (rank 0): per-GPU and
aggregate GB/s (min / mean over ranks, device-timed per rank with CUDA events, all ranks between barriers).

  python tools/fabric_test.py                         # one GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/fabric_test.py

Not a test (pytest collects tests/ only: pytest.ini)."""
import json
import os
import time

import torch
import torch.distributed as dist

H2D_LLR = 64 * 1362816          # int8 soft bits of 64 config-2 TBs
H2D_SYM = 64 * 170352 * 12      # complex64 symbols + float32 noise variances of the same TBs
D2H_TB = 64 * 159749 + 64 * 32  # TB bytes + result records


def bind_cpus(dev):
    """Pins the process to the CPUs closest to its GPU (NVML), like bench.py does."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(dev)
        n = os.cpu_count()
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, (n + 63) // 64)
        cpus = [i for i in range(n) if (mask[i // 64] >> (i % 64)) & 1]
        if cpus:
            os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0


def run(world, rank, dev, h2d_bytes, d2h_bytes, reps, streams):
    hs = torch.empty(h2d_bytes, dtype=torch.int8).pin_memory()
    ds = torch.empty(h2d_bytes, dtype=torch.int8, device=dev)
    hr = torch.empty(max(d2h_bytes, 1), dtype=torch.int8).pin_memory()
    dr = torch.empty(max(d2h_bytes, 1), dtype=torch.int8, device=dev)
    up = [torch.cuda.Stream(dev) for _ in range(streams)]
    down = torch.cuda.Stream(dev)
    piece = h2d_bytes // streams

    def step():
        for i, s in enumerate(up):
            with torch.cuda.stream(s):
                ds[i * piece:(i + 1) * piece].copy_(hs[i * piece:(i + 1) * piece], non_blocking=True)
        if d2h_bytes:
            with torch.cuda.stream(down):
                hr.copy_(dr, non_blocking=True)

    for _ in range(3):
        step()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(torch.cuda.current_stream(dev))
    for s in up + [down]:
        s.wait_event(e0)
    for _ in range(reps):
        step()
    for s in up + [down]:
        torch.cuda.current_stream(dev).wait_stream(s)
    e1.record(torch.cuda.current_stream(dev))
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    wall = time.perf_counter() - t0
    mine = torch.tensor([h2d_bytes * reps / ms / 1e6, d2h_bytes * reps / ms / 1e6, wall], dtype=torch.float64)
    if world > 1:
        allv = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allv, mine)
    else:
        allv = [mine]
    return torch.stack(allv)


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    ncpu = bind_cpus(local)
    if world > 1:
        dist.init_process_group("gloo")
    for name, h2d, d2h in (("llr_h2d_only", H2D_LLR, 0), ("llr_h2d_plus_tb_d2h", H2D_LLR, D2H_TB),
                           ("symbols_h2d_plus_tb_d2h", H2D_SYM, D2H_TB), ("tb_d2h_only", 1 << 20, D2H_TB)):
        for streams in (1, 2):
            v = run(world, rank, dev, h2d, d2h, 40, streams)
            if rank == 0:
                print(json.dumps({"config": name, "n_gpus": world, "copy_streams": streams, "h2d_bytes": h2d, "d2h_bytes": d2h,
                                  "h2d_gbs_per_gpu_min": round(float(v[:, 0].min()), 2),
                                  "h2d_gbs_per_gpu_mean": round(float(v[:, 0].mean()), 2),
                                  "h2d_gbs_aggregate": round(float(v[:, 0].sum()), 2),
                                  "d2h_gbs_per_gpu_min": round(float(v[:, 1].min()), 2),
                                  "d2h_gbs_aggregate": round(float(v[:, 1].sum()), 2),
                                  "cpus_bound": ncpu}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
