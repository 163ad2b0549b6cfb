"""GPU parity tests of the device-side soft demodulation / descrambling / UL-SCH demultiplexing
through the C ABI: bit-exact int8 soft bits against the compiled reference (demodulation_mapper, pusch_demodulator_impl
behind a stub equalizer + ulsch_demultiplex_impl) where it is available on the box, else against the port that
tests/test_oracle_demod_cpu.py pins to it; then symbols in -> transport blocks out against the reference's decoder fed with
the reference's own soft bits."""
import numpy as np
import pytest

from oracle import bindings as ob
from srsran_projectvtlmo_b200 import pusch, synth
from tests.test_oracle_demod_cpu import random_block

pytestmark = pytest.mark.gpu

HAVE_REF = ob.ref() is not None and ob.ref_flavour() is not None and hasattr(ob.ref(), "ref_demodulate_soft")


def _check_demod(symbols, noise_vars, qm, pi2=False):
    return (ob.ref_demodulate_soft if HAVE_REF else ob.port_demodulate_soft)(symbols, noise_vars, qm, pi2)


@pytest.mark.parametrize("qm", [1, 2, 4, 6, 8])
def test_demodulation_mapper_blocks(acc, qm):
    """One block per call, lengths ending inside a SIMD batch, special symbols and noise variances."""
    rng = np.random.default_rng(300 + qm)
    dm = pusch.demodulation_mapper_cuda(acc)
    for trial in range(60):
        n = int(rng.integers(1, 700))
        sym, nv = random_block(rng, n, spread=float(rng.choice([0.3, 0.8, 1.5])), special=trial % 3 == 0)
        for pi2 in ([False, True] if qm == 1 else [False]):
            got = dm.demodulate_soft(sym, nv, qm, pi2)
            want = _check_demod(sym, nv, qm, pi2)
            assert np.array_equal(got, want), (qm, pi2, n, np.nonzero(got != want)[0][:8])


CODEWORDS = [
    # qm, layers, prb, first symbol, symbols, DM-RS mask, CDM groups without data
    (8, 4, 273, 0, 14, 1 << 2, 2),
    (6, 2, 106, 0, 14, 1 << 2, 2),
    (6, 1, 57, 2, 12, (1 << 2) | (1 << 11), 1),
    (4, 1, 52, 0, 14, 1 << 2, 2),
    (2, 1, 25, 0, 14, (1 << 2) | (1 << 7) | (1 << 11), 1),
    (2, 1, 1, 0, 14, 1 << 2, 2),
    (8, 1, 11, 1, 9, 1 << 3, 1),
    (1, 1, 3, 0, 14, 1 << 2, 2),
]


@pytest.mark.parametrize("case", CODEWORDS)
def test_pusch_demodulate_codeword(acc, case):
    """Whole codeword: block partition per OFDM symbol, demapper tails, scrambling sequence (chunked generation), demux."""
    qm, nl, nprb, s0, ns, dmrs, cdm = case
    rng = np.random.default_rng(400 + CODEWORDS.index(case))
    rps = ob.pusch_re_per_symbol(nprb, s0, ns, dmrs, cdm)
    n = int(rps.sum()) * nl
    for rep in range(2):
        sym, nv = random_block(rng, n, special=rep == 1)
        rnti, n_id = int(rng.integers(1, 65520)), int(rng.integers(0, 1024))
        got = pusch.pusch_demodulate(acc, sym, nv, pusch.demod_config(qm, rnti, n_id, nl, rps))
        if HAVE_REF:
            want = ob.ref_pusch_demodulate(sym, nv, qm, rnti, n_id, nl, nprb, s0, ns, dmrs, cdm)
        else:
            want = ob.port_pusch_demodulate(sym, nv, qm, rnti, n_id, nl, rps)
        assert np.array_equal(got, want), (case, rep, np.nonzero(got != want)[0][:8])


def _modulate(bits, qm):
    """TS 38.211 5.1 mapping of a bit array to unit-power QAM symbols."""
    b = bits.reshape(-1, qm).astype(np.float64)
    if qm == 2:
        return ((1 - 2 * b[:, 0]) + 1j * (1 - 2 * b[:, 1])) / np.sqrt(2)
    if qm == 4:
        return ((1 - 2 * b[:, 0]) * (2 - (1 - 2 * b[:, 2])) + 1j * (1 - 2 * b[:, 1]) * (2 - (1 - 2 * b[:, 3]))) / np.sqrt(10)
    if qm == 6:
        re = (1 - 2 * b[:, 0]) * (4 - (1 - 2 * b[:, 2]) * (2 - (1 - 2 * b[:, 4])))
        im = (1 - 2 * b[:, 1]) * (4 - (1 - 2 * b[:, 3]) * (2 - (1 - 2 * b[:, 5])))
        return (re + 1j * im) / np.sqrt(42)
    re = (1 - 2 * b[:, 0]) * (8 - (1 - 2 * b[:, 2]) * (4 - (1 - 2 * b[:, 4]) * (2 - (1 - 2 * b[:, 6]))))
    im = (1 - 2 * b[:, 1]) * (8 - (1 - 2 * b[:, 3]) * (4 - (1 - 2 * b[:, 5]) * (2 - (1 - 2 * b[:, 7]))))
    return (re + 1j * im) / np.sqrt(170)


@pytest.mark.parametrize("case", [(273, 8, 948, 4, 1, 12611, 32.0), (106, 6, 873, 2, 1, 0, 25.0), (52, 4, 658, 1, 1, 25344, 17.0),
                                  (25, 2, 120, 1, 2, 25344, 1.0)])
def test_symbols_in_transport_blocks_out(case):
    """The whole device chain (demodulate -> descramble -> demultiplex -> rate dematch -> LDPC -> CRC -> TB assembly) from
    equalized symbols, several TBs per call, against the reference decoder fed with the reference demodulator's soft bits:
    TB bytes, CRC verdict, iteration statistics, and every soft-buffer byte."""
    prb, qm, R, nl, bg, nref, snr_db = case
    rng = np.random.default_rng(500 + prb)
    tbs = synth.tbs_for(prb, qm, R, nl)
    rps = ob.pusch_re_per_symbol(prb, 0, 14, 1 << 2, 2)
    nsym = int(rps.sum()) * nl
    nllr = nsym * qm
    nseg = len(pusch.segment(tbs, bg, qm, nl, nllr))
    ntb = 3
    acc = pusch.Accelerator(device=0, max_cbs_in_flight=ntb * nseg, nof_harq_cb_slots=ntb * nseg)
    try:
        cfgs, dms, syms, nvs, payloads, want_llrs = [], [], [], [], [], []
        for i in range(ntb):
            payload = rng.integers(0, 256, tbs // 8, dtype=np.uint8)
            cw = synth.encode_tb(payload, bg, 0, qm, nref, nl, nllr)
            rnti, n_id = 0x4601 + i, 100 + i
            scr = ob.port_scrambling_sequence((rnti << 15) + n_id, nllr)
            x = _modulate(cw ^ scr, qm)
            sigma2 = 10 ** (-(snr_db - 1.5 * i) / 10)
            y = (x + np.sqrt(sigma2 / 2) * (rng.normal(size=nsym) + 1j * rng.normal(size=nsym))).astype(np.complex64)
            nv = np.full(nsym, sigma2, np.float32) * rng.uniform(0.9, 1.1, nsym).astype(np.float32)
            cfgs.append(pusch.TbConfig(tbs, bg, 0, qm, nref, nl, 6, 1, 1, i * nseg))
            dms.append(pusch.demod_config(qm, rnti, n_id, nl, rps))
            syms.append(y)
            nvs.append(nv)
            payloads.append(payload)
            if HAVE_REF:
                want_llrs.append(ob.ref_pusch_demodulate(y, nv, qm, rnti, n_id, nl, prb, 0, 14, 1 << 2, 2))
            else:
                want_llrs.append(ob.port_pusch_demodulate(y, nv, qm, rnti, n_id, nl, rps))
        tickets = pusch.submit_tbs_symbols(acc, pusch.SubmitSymbolArgs(cfgs, dms, syms, nvs))
        outs = [np.zeros(tbs // 8, np.uint8) for _ in range(ntb)]
        res = pusch.poll_tbs(acc, tickets, outs, block=True)
        assert pusch.ticket_demod_ms(acc, tickets[0]) > 0
        ok = 0
        for i in range(ntb):
            if HAVE_REF:
                tb_r, res_r, crcs_r, soft = ob.RefPusch().decode(0, tbs // 8, want_llrs[i], bg, 0, qm, nref, nl, 6, True, True,
                                                                  want_soft=True)
                metas = ob.ref_segment(tbs, bg, qm, nl, nllr)
                offs = np.concatenate([[0], np.cumsum([m.full_length for m in metas])])
                softs = [soft[offs[k]:offs[k + 1]] for k in range(len(metas))]
            else:
                port = ob.PortPusch()
                tb_r, res_r = port.decode(0, tbs // 8, want_llrs[i], bg, 0, qm, nref, nl, 6, True, True)
                _, softs = port.harq_state(0, ob.port_segment(tbs, bg, qm, nl, nllr))
            assert res[i].tb_crc_ok == res_r.tb_crc_ok, (case, i)
            assert (res[i].iter_min, res[i].iter_max, res[i].nof_observations) == \
                (res_r.iter_min, res_r.iter_max, res_r.nof_observations), (case, i)
            for cb in range(nseg):
                got = acc.read_softbuffer(i * nseg + cb, softs[cb].size)
                assert np.array_equal(got, softs[cb]), (case, i, cb)
            if res_r.tb_crc_ok:
                assert np.array_equal(outs[i], payloads[i]), (case, i)
                ok += 1
        assert ok >= 1, "no TB decoded: the operating point does not exercise the chain end to end"
    finally:
        acc.close()
