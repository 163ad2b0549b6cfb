#!/usr/bin/env python3
"""Stall samples and executed instructions of one kernel in an .ncu-rep, per source line (top N) and per file.
Usage: tools/ncu_regions.py report.ncu-rep [N]   (needs -lineinfo and --import-source on)"""
import collections
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    cur, hdr, lines = None, None, []
    for r in csv.reader(txt.splitlines()):
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or r[0] in ("", "Function Name"):
            continue
        try:
            ln = int(r[0])
        except ValueError:
            continue

        def num(name):
            try:
                return int(r[hdr.index(name)])
            except ValueError:
                return 0
        lines.append((num("# Samples"), num("Instructions Executed"), cur, ln, r[1][:90]))
    ts = sum(l[0] for l in lines) or 1
    ti = sum(l[1] for l in lines) or 1
    per_file = collections.Counter()
    for l in lines:
        per_file[l[2]] += l[0]
    print(f"samples {ts}, warp instructions {ti}")
    for f, v in per_file.most_common():
        print(f"  {f:32s} {100 * v / ts:5.1f}% of samples")
    for l in sorted(lines, reverse=True)[:top]:
        print(f"{100 * l[0] / ts:5.1f}% smp {100 * l[1] / ti:5.1f}% ins  {l[2]}:{l[3]}  {l[4]}")


if __name__ == "__main__":
    main()
