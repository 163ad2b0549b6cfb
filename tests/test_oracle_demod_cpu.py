"""CPU tests of the oracle for the PUSCH soft-demodulation chain (soft demodulation, descrambling,
UL-SCH demultiplexing): the plain-C port (oracle/oracle_demod_port.c) against golden vectors generated from the compiled
reference (tests/golden/demod.npz, made by tests/golden/make_golden_demod.py) and, where the compiled reference is
available, differentially on fresh inputs."""
from pathlib import Path

import numpy as np
import pytest

from oracle import bindings as ob

GOLDEN = Path(__file__).resolve().parent / "golden"
needs_ref = pytest.mark.skipif(ob.ref() is None or ob.ref_flavour() is None or not hasattr(ob.ref(), "ref_demodulate_soft"),
                               reason="compiled reference (oracle/_ref) not available on this host")


def random_block(rng, n, spread=0.8, special=False):
    sym = (rng.normal(0, spread, n) + 1j * rng.normal(0, spread, n)).astype(np.complex64)
    nv = rng.uniform(0.003, 0.4, n).astype(np.float32)
    if special:
        k = max(1, n // 10)
        nv[rng.integers(0, n, k)] = rng.choice(np.array([0, -1, np.inf, np.nan, 1e-12], np.float32), k)
        sym[rng.integers(0, n, k)] = 0
        sym[rng.integers(0, n, k)] *= np.float32(1e-6)
        sym[rng.integers(0, n, k)] *= np.float32(40)
    return sym, nv


def test_scrambling_sequence_known_answer():
    """TS 38.211 5.2.1 with c_init = 0: x2 is all zero, so c(n) = x1(n + 1600), the m-sequence of x^31 + x^3 + 1 from 1."""
    x = [1] + [0] * 30
    for n in range(1600 + 256):
        x.append(x[n + 3] ^ x[n])
    want = np.array(x[1600:1600 + 256], np.uint8)
    assert np.array_equal(ob.port_scrambling_sequence(0, 256), want)
    # linearity in c_init (the property the device kernel's jump tables rely on)
    a, b = 0x1234567, 0x7654321
    sa, sb, sab, s0 = (ob.port_scrambling_sequence(c, 4096) for c in (a, b, a ^ b, 0))
    assert np.array_equal(sa ^ sb ^ s0, sab)


def test_golden_demodulation_mapper_blocks():
    g = np.load(GOLDEN / "demod.npz")
    meta = g["blk_meta"]
    o_s = o_l = 0
    for qm, pi2, n in meta:
        sym = g["blk_sym"][o_s:o_s + n]
        nv = g["blk_nv"][o_s:o_s + n]
        want = g["blk_llr"][o_l:o_l + n * qm]
        got = ob.port_demodulate_soft(sym, nv, int(qm), bool(pi2))
        assert np.array_equal(got, want), (qm, pi2, n, np.nonzero(got != want)[0][:8])
        o_s += n
        o_l += n * qm
    assert len(meta) >= 60


def test_golden_pusch_demodulate_codewords():
    g = np.load(GOLDEN / "demod.npz")
    o_s = o_l = 0
    for (qm, rnti, n_id, nl, nprb, s0, ns, dmrs, cdm) in g["cw_meta"]:
        rps = ob.pusch_re_per_symbol(int(nprb), int(s0), int(ns), int(dmrs), int(cdm))
        n = int(rps.sum()) * int(nl)
        sym, nv = g["cw_sym"][o_s:o_s + n], g["cw_nv"][o_s:o_s + n]
        want = g["cw_llr"][o_l:o_l + n * qm]
        got = ob.port_pusch_demodulate(sym, nv, int(qm), int(rnti), int(n_id), int(nl), rps)
        assert np.array_equal(got, want), (qm, nl, nprb, np.nonzero(got != want)[0][:8])
        o_s += n
        o_l += n * qm


@needs_ref
def test_port_vs_reference_scrambling():
    rng = np.random.default_rng(21)
    for _ in range(20):
        c = int(rng.integers(0, 1 << 31))
        n = int(rng.integers(1, 30000))
        assert np.array_equal(ob.port_scrambling_sequence(c, n), ob.ref_scrambling_sequence(c, n)), (c, n)


@needs_ref
@pytest.mark.parametrize("qm", [1, 2, 4, 6, 8])
def test_port_vs_reference_demodulation_mapper(qm):
    """Every modulation, block lengths that end inside a SIMD batch (scalar tails), symbols at and around zero, noise
    variances that are zero / negative / infinite / NaN, symbols far outside the constellation (clipping)."""
    rng = np.random.default_rng(22 + qm)
    for trial in range(400):
        n = int(rng.integers(1, 300))
        sym, nv = random_block(rng, n, spread=float(rng.choice([0.3, 0.8, 1.5])), special=trial % 4 == 0)
        for pi2 in ([False, True] if qm == 1 else [False]):
            a = ob.port_demodulate_soft(sym, nv, qm, pi2)
            b = ob.ref_demodulate_soft(sym, nv, qm, pi2)
            assert np.array_equal(a, b), (qm, pi2, n, np.nonzero(a != b)[0][:8])


@needs_ref
def test_port_vs_reference_pusch_demodulator_chain():
    """pusch_demodulator_impl (stub equalizer) + ulsch_demultiplex_impl without UCI: block partition per OFDM symbol,
    demapper tails, descrambling."""
    rng = np.random.default_rng(23)
    for (qm, nl, nprb, s0, ns, dmrs, cdm) in [(8, 4, 273, 0, 14, 1 << 2, 2), (6, 2, 106, 0, 14, 1 << 2, 2),
                                              (6, 1, 57, 2, 12, (1 << 2) | (1 << 11), 1), (4, 1, 52, 0, 14, 1 << 2, 2),
                                              (2, 1, 25, 0, 14, (1 << 2) | (1 << 7) | (1 << 11), 1), (2, 1, 1, 0, 14, 1 << 2, 2),
                                              (8, 1, 11, 1, 9, 1 << 3, 1)]:
        rps = ob.pusch_re_per_symbol(nprb, s0, ns, dmrs, cdm)
        n = int(rps.sum()) * nl
        sym, nv = random_block(rng, n, special=True)
        rnti, n_id = int(rng.integers(1, 65520)), int(rng.integers(0, 1024))
        a = ob.port_pusch_demodulate(sym, nv, qm, rnti, n_id, nl, rps)
        b = ob.ref_pusch_demodulate(sym, nv, qm, rnti, n_id, nl, nprb, s0, ns, dmrs, cdm)
        assert np.array_equal(a, b), (qm, nl, nprb, np.nonzero(a != b)[0][:8])
