#!/usr/bin/env python3
"""Summarises an .ncu-rep (read here, no GPU): headline counters + executed instructions by opcode for one kernel."""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
keys = ("Duration", "Elapsed Cycles", "SM Busy", "Issue Slots Busy", "Executed Ipc Active", "Registers Per Thread",
        "Theoretical Occupancy", "Achieved Occupancy", "Eligible Warps", "No Eligible", "Warp Cycles Per Issued",
        "Executed Instructions", "Mem Pipes Busy", "Dynamic Shared Memory Per Block", "DRAM Throughput",
        "Memory Throughput", "Block Size", "Grid Size", "L2 Cache Throughput")
print(det.splitlines()[0] if det else "")
for line in det.splitlines():
    if any(k in line for k in keys) or "void " in line or "Kernel" in line[:12]:
        print(line.rstrip()[:150])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[-1]
want = ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma", "pipe_lsu.avg.pct",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "issue_stalled", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active")
for h, v in zip(hdr, vals):
    if any(w in h for w in want) and "realtime" not in h and (".ratio" in h or ".sum" in h or ".avg.pct" in h) and "not_issued" not in h:
        if "issue_stalled" in h and "per_issue_active" not in h:
            continue
        print(f"{h:95s} {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if "Address" in r and "Source" in r)
hdr = rows[hi]
iS, iE, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ops, samp, tot = collections.Counter(), collections.Counter(), 0
for r in rows[hi + 1:]:
    try:
        n, sm = int(r[iE]), int(r[iSm])
    except (ValueError, IndexError):
        continue
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[iS].strip())
    op = m.group(2) if m else r[iS][:10]
    parts = op.split(".")
    key = ".".join(parts[:2]) if parts[0] in ("VIMNMX", "VIADDMNMX", "VIMNMX3", "IMAD", "LDS", "STS", "LDG", "STG", "LDL", "STL", "BAR") else parts[0]
    ops[key] += n
    samp[key] += sm
    tot += n
print("executed warp instructions:", tot)
for k, v in ops.most_common(28):
    print(f"  {k:22s} {v:12d} {100 * v / tot:5.1f}%  stall samples {samp[k]}")
