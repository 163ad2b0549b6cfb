#!/usr/bin/env python3
"""Generates tests/golden/pdsch_enc.npz from the UNMODIFIED reference's software PDSCH encoder (pdsch_encoder_impl: segmenter
Tx + LDPC encoder + rate matcher, compiled under oracle/_ref by oracle/Makefile) on seeded transport blocks.

    python tests/golden/make_golden_pdsch_enc.py       # build container only (needs /root/reference through oracle/_ref)

Inputs are stored next to the outputs (packed code words), so the fixture does not depend on numpy's random stream. Test
infrastructure only."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import bindings as ob  # noqa: E402
from srsran_projectvtlmo_b200 import synth  # noqa: E402

# (prb, Qm, R x 1024, layers, BG, Nref, rv): single and multi code block, both base graphs, every modulation order, limited
# buffer rate matching, every redundancy version, wrap-around (E > Ncb), code-block lengths that are not byte multiples.
CASES = [(52, 4, 658, 1, 1, 0, 0), (25, 2, 120, 1, 2, 0, 2), (24, 8, 948, 2, 1, 12611, 3), (106, 6, 873, 2, 1, 0, 1),
         (52, 2, 449, 1, 1, 25344, 0), (10, 4, 490, 1, 2, 0, 3), (4, 2, 308, 1, 2, 0, 0), (1, 2, 120, 1, 2, 0, 1),
         (273, 2, 308, 1, 1, 0, 2), (273, 2, 193, 2, 2, 0, 0), (133, 8, 948, 3, 1, 12611, 0), (51, 6, 567, 3, 1, 9000, 2),
         (273, 8, 948, 2, 1, 25223, 0)]


def main():
    rng = np.random.default_rng(4242)
    out = {"cases": np.array(CASES, np.int64)}
    for i, (prb, qm, R, nl, bg, nref, rv) in enumerate(CASES):
        tbs = synth.tbs_for(prb, qm, R, nl)
        nbits = prb * 156 * qm * nl
        tb = rng.integers(0, 256, tbs // 8, dtype=np.uint8)
        cw = ob.ref_encode_tb(tb, bg, rv, qm, nref, nl, nbits // qm)
        out[f"tb{i}"] = tb
        out[f"cw{i}"] = np.packbits(cw)
    np.savez_compressed(Path(__file__).resolve().parent / "pdsch_enc.npz", **out)
    print("wrote pdsch_enc.npz:", len(CASES), "cases")


if __name__ == "__main__":
    main()
