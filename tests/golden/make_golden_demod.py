#!/usr/bin/env python3
"""Generates tests/golden/demod.npz from the COMPILED REFERENCE (oracle/_ref, built from /root/reference by oracle/Makefile):
demodulation_mapper blocks for every modulation (SIMD bodies and scalar tails, special values) and whole PUSCH codewords
through pusch_demodulator_impl (stub equalizer) + ulsch_demultiplex_impl. Run in the build container:
    python tests/golden/make_golden_demod.py"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
from oracle import bindings as ob  # noqa: E402
from tests.test_oracle_demod_cpu import random_block  # noqa: E402


def main():
    assert ob.ref() is not None, "compiled reference not available"
    rng = np.random.default_rng(2025)
    meta, syms, nvs, llrs = [], [], [], []
    for qm in (1, 1, 2, 4, 6, 8):
        for trial in range(12):
            pi2 = (qm == 1 and trial % 2 == 1)
            n = int(rng.integers(1, 160))
            sym, nv = random_block(rng, n, spread=float(rng.choice([0.4, 0.9])), special=trial % 3 == 0)
            meta.append((qm, int(pi2), n))
            syms.append(sym)
            nvs.append(nv)
            llrs.append(ob.ref_demodulate_soft(sym, nv, qm, pi2))
    cw_meta, cw_sym, cw_nv, cw_llr = [], [], [], []
    for (qm, nl, nprb, s0, ns, dmrs, cdm) in [(8, 4, 24, 0, 14, 1 << 2, 2), (6, 2, 35, 0, 14, 1 << 2, 2),
                                              (6, 1, 57, 2, 12, (1 << 2) | (1 << 11), 1), (4, 1, 52, 0, 14, 1 << 2, 2),
                                              (2, 1, 25, 0, 14, (1 << 2) | (1 << 7) | (1 << 11), 1), (2, 1, 1, 0, 14, 1 << 2, 2)]:
        rps = ob.pusch_re_per_symbol(nprb, s0, ns, dmrs, cdm)
        n = int(rps.sum()) * nl
        sym, nv = random_block(rng, n, special=True)
        rnti, n_id = int(rng.integers(1, 65520)), int(rng.integers(0, 1024))
        cw_meta.append((qm, rnti, n_id, nl, nprb, s0, ns, dmrs, cdm))
        cw_sym.append(sym)
        cw_nv.append(nv)
        cw_llr.append(ob.ref_pusch_demodulate(sym, nv, qm, rnti, n_id, nl, nprb, s0, ns, dmrs, cdm))
    out = Path(__file__).resolve().parent / "demod.npz"
    np.savez_compressed(out, blk_meta=np.array(meta, np.int64), blk_sym=np.concatenate(syms), blk_nv=np.concatenate(nvs),
                        blk_llr=np.concatenate(llrs), cw_meta=np.array(cw_meta, np.int64), cw_sym=np.concatenate(cw_sym),
                        cw_nv=np.concatenate(cw_nv), cw_llr=np.concatenate(cw_llr))
    print("wrote", out, out.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
