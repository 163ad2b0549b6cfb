#!/usr/bin/env python3
"""Generates tests/golden/*.npz from the UNMODIFIED reference compiled under oracle/_ref (oracle/Makefile).

Run in the build container (needs /root/reference, through oracle/_ref/libref_oracle.so):

    python tests/golden/make_golden.py

The reference ships no golden vectors for this path in this checkout (all *_test_data.tar.gz are absent, SURVEY.md
section 4), so these fixtures are outputs of the reference itself ("avx512" decoder and dematcher, "auto" CRC, i.e. the
flavour production selects on an AVX-512 host) on seeded inputs. Inputs are stored next to the outputs so that the
fixtures do not depend on the numpy random stream. Everything here is test infrastructure.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import bindings as ob  # noqa: E402
from srsran_projectvtlmo_b200 import synth  # noqa: E402
from tests.helpers import awgn_llrs, random_cb_case  # noqa: E402

OUT = Path(__file__).resolve().parent


def crc_fixture(rng):
    """crc_calculator_test.cpp:218-248 sizes (bytes) plus bit lengths that are not byte multiples."""
    msgs, nbits, sums = [], [], []
    for n in [8 * 8, 16 * 8, 32 * 8, 257 * 8, 997 * 8, 6012 * 8, 1, 7, 24, 129, 997, 8448 - 16, 3824 + 16]:
        data = rng.integers(0, 256, (n + 7) // 8, dtype=np.uint8)
        if n % 8:
            data[-1] &= (0xFF00 >> (n % 8)) & 0xFF
        row = []
        for poly in (ob.CRC24A, ob.CRC24B, ob.CRC16):
            # crc_calculator_clmul_impl mishandles 13..15 whole bytes (SURVEY.md Appendix A trap 8): "lut" is the
            # flavour that implements the plain polynomial division for every length.
            row.append(ob.ref().ref_crc(b"lut", poly, ob._p8(data), n))
        msgs.append(data)
        nbits.append(n)
        sums.append(row)
    return {"crc_msgs": np.concatenate(msgs), "crc_nbits": np.array(nbits, np.uint32),
            "crc_sums": np.array(sums, np.uint32)}


def dematch_fixture(rng):
    """HARQ chains rv 0,2,3,1 (new data on the first only) on dirty buffers; the full buffer after every transmission."""
    cases = [  # bg, Z, E, Qm, F, Nref
        (1, 384, 8960, 8, 16, 12611), (1, 384, 8992, 8, 16, 12611), (2, 208, 16224, 2, 136, 25344),
        (1, 352, 16224, 2, 680, 25344), (2, 8, 312, 2, 32, 25344), (1, 96, 6336 * 2, 6, 40, 0), (2, 52, 2600, 4, 0, 2000),
        (1, 2, 300, 1, 3, 0), (2, 15, 744, 8, 7, 0),
    ]
    meta, bufs, llrs, outs = [], [], [], []
    for (bg, z, E, qm, F, nref) in cases:
        N = ob.ns(bg) * z
        buf = rng.integers(-120, 121, N, dtype=np.int8)
        bufs.append(buf.copy())
        for i, rv in enumerate([0, 2, 3, 1]):
            llr = rng.integers(-120, 121, E, dtype=np.int8)
            ob.ref_dematch(buf, llr, i == 0, rv, qm, nref, F, kind="avx512")
            llrs.append(llr)
            outs.append(buf.copy())
        meta.append((bg, z, E, qm, F, nref))
    return {"dm_meta": np.array(meta, np.uint32), "dm_init": np.concatenate(bufs), "dm_llrs": np.concatenate(llrs),
            "dm_out": np.concatenate(outs)}


def decoder_fixture(rng):
    """ldpc_decoder: noisy inputs, fillers, shortened tails, with / without CRC, 1..8 iterations."""
    meta, llrs, bits = [], [], []
    zs = [2, 3, 7, 16, 30, 52, 88, 120, 176, 208, 256, 320, 384]
    for bg in (1, 2):
        for z in zs:
            K, N = ob.kb(bg) * z, ob.ns(bg) * z
            msg, F, crc_poly = random_cb_case(rng, bg, z)
            cw = synth.ldpc_encode(msg, bg, z)
            llr = awgn_llrs(rng, cw, float(rng.choice([2, 3, 4, 6, 8])))
            llr[K - 2 * z - F:K - 2 * z] = 127
            nlen = int(rng.integers(K + 2 * z, N + 1)) if rng.random() < 0.7 else N
            llr[nlen:] = 0
            max_it = int(rng.integers(1, 9))
            out = np.full((K + 7) // 8, 0x5A, np.uint8)
            it, out = ob.ref_decode(llr, bg, z, F, crc_poly, max_it, out, kind="avx512")
            meta.append((bg, z, F, crc_poly, max_it, it & 0xFFFFFFFF, llr.size, out.size))
            llrs.append(llr)
            bits.append(out)
    # BASELINE config 1: the benchmark's own input, (mt19937(0)() & 1) * 20 - 10 (ldpc_decoder_benchmark.cpp:143-145).
    mt = np.random.RandomState(0)  # MT19937, seed 0: same 32-bit stream as std::mt19937(0)
    raw = mt.randint(0, 2 ** 32, 25344, dtype=np.uint64)
    llr = ((raw & 1) * 20 - 10).astype(np.int8)
    out = np.zeros(1056, np.uint8)
    it, out = ob.ref_decode(llr, 1, 384, 0, ob.CRC_NONE, 6, out, kind="avx512")
    meta.append((1, 384, 0, ob.CRC_NONE, 6, it & 0xFFFFFFFF, llr.size, out.size))
    llrs.append(llr)
    bits.append(out)
    return {"dec_meta": np.array(meta, np.uint32), "dec_llrs": np.concatenate(llrs), "dec_bits": np.concatenate(bits)}


def tb_fixture(rng):
    """pusch_decoder_impl over the rv sequence 0,2,3,1 with soft combining (pusch_decoder_vectortest.cpp:261-397)."""
    cases = [  # prb, Qm, R, layers, bg, Nref, mu, early stop, max iterations
        (52, 4, 658, 1, 1, 25344, 1.2, 1, 6), (25, 2, 120, 1, 2, 25344, 0.6, 1, 6), (52, 4, 378, 1, 1, 25344, 0.9, 0, 2),
        (24, 8, 948, 2, 1, 12611, 5.0, 1, 6),
    ]
    meta, tbs_in, llrs, tbs_out, softs = [], [], [], [], []
    for key, (prb, qm, R, nl, bg, nref, mu, es, max_it) in enumerate(cases):
        tbs = synth.tbs_for(prb, qm, R, nl)
        nllr = prb * 156 * qm * nl
        tb = rng.integers(0, 256, tbs // 8, dtype=np.uint8)
        dec = ob.RefPusch(dec="avx512", dem="avx512", crc="auto")
        tbs_in.append(tb)
        for i, rv in enumerate([0, 2, 3, 1]):
            cw = ob.ref_encode_tb(tb, bg, rv, qm, nref, nl, nllr // qm)
            llr = awgn_llrs(rng, cw, mu)
            out, res, crcs, soft = dec.decode(key, tbs // 8, llr, bg, rv, qm, nref, nl, max_it, es, i == 0, want_soft=True)
            meta.append((prb, qm, R, nl, bg, nref, es, max_it, rv, tbs, nllr, res.tb_crc_ok, res.nof_codeblocks,
                         res.nof_observations, res.iter_min, res.iter_max, int(round(res.iter_mean * 1000)),
                         soft.size, int(crcs.sum())))
            llrs.append(llr)
            tbs_out.append(out)
            softs.append(soft)
    return {"tb_meta": np.array(meta, np.uint32), "tb_payload": np.concatenate(tbs_in), "tb_llrs": np.concatenate(llrs),
            "tb_out": np.concatenate(tbs_out), "tb_soft": np.concatenate(softs)}


def main():
    assert ob.ref() is not None and ob.ref_flavour() == "avx512", "needs the compiled reference on an AVX-512 host"
    rng = np.random.default_rng(20261018)
    np.savez_compressed(OUT / "crc.npz", **crc_fixture(rng))
    np.savez_compressed(OUT / "dematch.npz", **dematch_fixture(rng))
    np.savez_compressed(OUT / "decoder.npz", **decoder_fixture(rng))
    np.savez_compressed(OUT / "tb.npz", **tb_fixture(rng))
    for f in sorted(OUT.glob("*.npz")):
        print(f.name, f.stat().st_size)


if __name__ == "__main__":
    main()
