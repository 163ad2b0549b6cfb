#!/usr/bin/env python3
"""Distribution of LDPC iterations per code block for the bench workload (config 2, mu as given), computed with the compiled
reference on the CPU: how much a group of four code blocks that runs until its slowest member is done executes beyond the
mean (round 1, mu = 18: iterations 2 / 3 / 4 for 48 / 249 / 7 of 304 code blocks, mean 2.87, mean of group maxima 3.08).
Build container only (needs oracle/_ref)."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))  # test-side analysis: the only place besides tests that may use oracle/
import bench  # noqa: E402
from oracle import bindings as ob  # noqa: E402
from srsran_projectvtlmo_b200 import pusch  # noqa: E402

mu = float(sys.argv[1]) if len(sys.argv) > 1 else 18.0
ntb = 2
tbs, nllr, payloads, sets = bench.make_inputs(ntb, 1, mu, 1000)
segs = pusch.segment(tbs, 1, 8, 4, nllr)
its = []
for tb in range(ntb):
    llr = sets[0][tb]
    for m in segs:
        soft = np.zeros(25344, np.int8)
        ob.ref_dematch(soft, llr[m.cw_offset:m.cw_offset + m.rm_length], True, 0, 8, 12611, m.nof_filler_bits)
        r = ob.ref_decode(soft, 1, 384, m.nof_filler_bits, 2, 6)
        its.append(r[0] if isinstance(r, tuple) else r)
its = np.array(its)
print("histogram of iterations:", np.bincount(its.clip(0)).tolist())
print("mean per code block %.3f, mean of the maxima of groups of four %.3f" % (its.mean(), its.reshape(-1, 4).max(1).mean()))
