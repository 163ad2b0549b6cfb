#!/usr/bin/env python3
"""Writes a file of synthetic transport blocks (payload + int8 soft bits) for oracle/_ref/hwacc_bench: BASELINE config 2 by
default (273 PRB, 256QAM, R = 948/1024, 4 layers, Nref 12611; the reference's own transmitter cannot encode this TB - its Tx
segmenter asserts on it, ldpc_segmenter_impl.cpp:77 - so the numpy transmitter of srsran_projectvtlmo_b200/synth.py, which
tests/test_oracle_cpu.py checks against the reference encoder, produces it).
Usage: make_tb_file.py OUT [nof_tbs] [mu] [prb qm rate layers bg nref]"""
import struct
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from srsran_projectvtlmo_b200 import synth  # noqa: E402


def main():
    out = sys.argv[1]
    ntb = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    mu = float(sys.argv[3]) if len(sys.argv) > 3 else 18.0
    prb, qm, rate, nl, bg, nref = (int(v) for v in sys.argv[4:10]) if len(sys.argv) >= 10 else (273, 8, 948, 4, 1, 12611)
    tbs = synth.tbs_for(prb, qm, rate, nl)
    nllr = prb * 156 * qm * nl
    rng = np.random.default_rng(77)
    with open(out, "wb") as f:
        f.write(struct.pack("<8I", 0x50425443, ntb, tbs, bg, qm, nl, nref, nllr))
        for _ in range(ntb):
            tb = rng.integers(0, 256, tbs // 8, dtype=np.uint8)
            llr = synth.awgn_llrs(rng, synth.encode_tb(tb, bg, 0, qm, nref, nl, nllr), mu)
            f.write(tb.tobytes())
            f.write(llr.tobytes())
    print(f"wrote {out}: {ntb} TBs of {tbs} bits, {nllr} soft bits each")


if __name__ == "__main__":
    main()
