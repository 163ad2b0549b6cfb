"""GPU parity tests: the CUDA path (through the C ABI) against the oracle on identical seeded inputs. Bit-exact bar:
decoded bits, CRC verdict, iteration count and combined soft-buffer bytes (BASELINE.json north_star)."""
import ctypes as C

import numpy as np
import pytest

from oracle import bindings as ob
from srsran_projectvtlmo_b200 import capi, pusch, synth
from tests.helpers import ALL_Z, awgn_llrs, random_cb_case

pytestmark = pytest.mark.gpu


# ---------------------------------------------------------------------------------------------------------------------
# CRC kernel (crc_calculator interface) - sizes of the reference's own test, crc_calculator_test.cpp:218-248
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("poly", [pusch.CRC24A, pusch.CRC24B, pusch.CRC16])
def test_crc_calculator(acc, poly):
    rng = np.random.default_rng(0)
    crc = pusch.crc_calculator_cuda(acc, poly)
    for nbits in (8, 16, 32, 257, 997, 6012, 1, 7, 24, 127, 128, 129, 8448, 131072 + 8, 1277992 + 24, 200000 * 8 + 3):
        data = rng.integers(0, 256, (nbits + 7) // 8, dtype=np.uint8)
        assert crc.calculate(data, nbits) == ob.port_crc(poly, data, nbits), (poly, nbits)
    bits = rng.integers(0, 2, 997, dtype=np.uint8)
    assert crc.calculate_bit(bits) == ob.port_crc(poly, np.packbits(bits), 997)
    # CRC of (payload || CRC) is zero.
    payload = rng.integers(0, 256, 500, dtype=np.uint8)
    order = 16 if poly == pusch.CRC16 else 24
    c = crc.calculate_byte(payload)
    full = np.concatenate([payload, np.frombuffer(int(c).to_bytes(order // 8, "big"), np.uint8)])
    assert crc.calculate_byte(full) == 0


# ---------------------------------------------------------------------------------------------------------------------
# Rate dematcher kernel (ldpc_rate_dematcher interface): write-set-exact, HARQ chains, dirty buffers
# ---------------------------------------------------------------------------------------------------------------------
def _dematch_chain(acc, rng, bg, z, qm, F, nref, dirty, e_max_factor=3.0, special=False):
    dm = pusch.ldpc_rate_dematcher_cuda(acc)
    N = ob.ns(bg) * z
    a = rng.integers(-120, 121, N, dtype=np.int8) if dirty else np.zeros(N, np.int8)
    if dirty and special:
        a[rng.integers(0, N, N // 8)] = rng.choice(np.array([127, -127], np.int8), N // 8)
    b = a.copy()
    for i, rv in enumerate([0, 2, 3, 1]):
        E = int(rng.integers(1, max(2, int(e_max_factor * N) // qm + 1))) * qm
        llr = rng.integers(-120, 121, E, dtype=np.int8)
        if special:
            llr[rng.integers(0, E, max(1, E // 16))] = rng.choice(np.array([127, -127], np.int8), max(1, E // 16))
        ob.port_dematch(a, llr, i == 0, rv, qm, nref, F)
        dm.rate_dematch(b, llr, i == 0, rv, qm, nref, F)
        assert np.array_equal(a, b), (bg, z, qm, F, nref, rv, E, np.nonzero(a != b)[0][:10])


def test_rate_dematcher_random_chains(acc):
    rng = np.random.default_rng(1)
    for trial in range(150):
        bg = int(rng.integers(1, 3))
        z = int(rng.choice(ALL_Z))
        N = ob.ns(bg) * z
        qm = int(rng.choice([1, 2, 4, 6, 8]))
        sys_ = (ob.kb(bg) - 2) * z
        F = int(rng.integers(0, sys_ // 2 + 1)) if rng.random() < 0.7 else 0
        nref = 0 if rng.random() < 0.5 else int(rng.integers(sys_ + z, N + 50))
        _dematch_chain(acc, rng, bg, z, qm, F, nref, dirty=rng.random() < 0.5)


def test_rate_dematcher_nonfinite_inputs(acc):
    """+-127 in non-filler positions: the reference's SIMD loop and scalar tail differ; the kernel follows both."""
    rng = np.random.default_rng(2)
    for trial in range(40):
        bg = int(rng.integers(1, 3))
        z = int(rng.choice([8, 20, 36, 96, 208, 384]))
        qm = int(rng.choice([1, 2, 4, 6, 8]))
        sys_ = (ob.kb(bg) - 2) * z
        F = int(rng.integers(0, sys_ // 3 + 1))
        _dematch_chain(acc, rng, bg, z, qm, F, 0, dirty=True, e_max_factor=2.2, special=True)


def test_rate_dematcher_baseline_configs(acc):
    """The E / Nref combinations of SURVEY.md section 8 (273 PRB 256QAM 4 layers with Nref = 12611, 52 PRB QPSK wrap)."""
    rng = np.random.default_rng(3)
    dm = pusch.ldpc_rate_dematcher_cuda(acc)
    for (bg, z, E, qm, F, nref) in [(1, 384, 8960, 8, 16, 12611), (1, 384, 8992, 8, 16, 12611), (2, 208, 16224, 2, 136, 25344),
                                    (1, 352, 16224, 2, 680, 25344), (2, 8, 312, 2, 32, 25344), (1, 384, 8960, 8, 16, 0)]:
        N = ob.ns(bg) * z
        a = rng.integers(-120, 121, N, dtype=np.int8)
        b = a.copy()
        for i, rv in enumerate([0, 2, 3, 1]):
            llr = rng.integers(-120, 121, E, dtype=np.int8)
            ob.port_dematch(a, llr, i == 0, rv, qm, nref, F)
            dm.rate_dematch(b, llr, i == 0, rv, qm, nref, F)
            assert np.array_equal(a, b), (bg, z, E, rv)


# ---------------------------------------------------------------------------------------------------------------------
# LDPC decoder kernel (ldpc_decoder interface): all 2 x 51 graphs
# ---------------------------------------------------------------------------------------------------------------------
def _decode_case(acc, rng, bg, z, max_it=None):
    dec = pusch.ldpc_decoder_cuda(acc)
    K, N = ob.kb(bg) * z, ob.ns(bg) * z
    msg, F, crc_poly = random_cb_case(rng, bg, z)
    cw = synth.ldpc_encode(msg, bg, z)
    mu = float(rng.choice([2, 3, 4, 5, 6, 8]))
    llr = awgn_llrs(rng, cw, mu)
    llr[K - 2 * z - F:K - 2 * z] = 127
    nlen = int(rng.integers(K + 2 * z, N + 1)) if rng.random() < 0.7 else N
    llr[nlen:] = 0
    max_it = int(rng.integers(1, 9)) if max_it is None else max_it
    o1 = np.full((K + 7) // 8, 0x5A, np.uint8)
    o2 = o1.copy()
    it1, o1, layers = ob.port_decode(llr, bg, z, F, crc_poly, max_it, o1)
    it2 = dec.decode(o2, llr, crc_poly, bg, z, F, max_it)
    assert (it1 if it1 >= 0 else None) == it2, (bg, z, F, crc_poly, max_it, it1, it2, layers)
    assert np.array_equal(o1, o2), (bg, z, F, crc_poly, max_it, layers)
    return it2


@pytest.mark.parametrize("bg", [1, 2])
def test_ldpc_decoder_all_lifting_sizes(acc, bg):
    rng = np.random.default_rng(10 + bg)
    successes = 0
    for z in ALL_Z:
        for _ in range(2):
            successes += _decode_case(acc, rng, bg, z) is not None
    assert successes > 20  # early stop is exercised


def test_ldpc_decoder_noise_free_one_iteration(acc):
    """Mirror of ldpc_enc_dec_test.cpp:287-317: noise-free +-10 LLRs, 1 iteration, the message is recovered."""
    rng = np.random.default_rng(20)
    dec = pusch.ldpc_decoder_cuda(acc)
    for bg in (1, 2):
        for z in (2, 7, 16, 52, 104, 176, 384):
            K = ob.kb(bg) * z
            msg = rng.integers(0, 2, K, dtype=np.uint8)
            cw = synth.ldpc_encode(msg, bg, z)
            llr = (10 - 20 * cw.astype(np.int16)).astype(np.int8)
            out = np.zeros((K + 7) // 8, np.uint8)
            assert dec.decode(out, llr, pusch.CRC_NONE, bg, z, 0, 1) is None
            assert np.array_equal(np.unpackbits(out)[:K], msg)


def test_ldpc_decoder_all_zero_input(acc):
    """Mirror of ldpc_enc_dec_test.cpp:334-358: all-zero LLRs -> all ones without CRC; untouched with CRC."""
    dec = pusch.ldpc_decoder_cuda(acc)
    for bg, z in ((1, 384), (2, 3), (1, 36)):
        K, N = ob.kb(bg) * z, ob.ns(bg) * z
        out = np.full((K + 7) // 8, 0x5A, np.uint8)
        assert dec.decode(out, np.zeros(N, np.int8), pusch.CRC_NONE, bg, z, 0, 3) is None
        assert np.all(np.unpackbits(out)[:K] == 1)
        out = np.full((K + 7) // 8, 0x5A, np.uint8)
        assert dec.decode(out, np.zeros(N, np.int8), pusch.CRC16, bg, z, 0, 3) is None
        assert np.all(out == 0x5A)


def test_ldpc_decoder_config1_benchmark_input(acc):
    """BASELINE config 1: BG1 Z=384, 25344 LLRs = +-10 from mt19937(0), no CRC, 6 iterations, batch of code blocks."""
    rng = np.random.default_rng(30)
    dec = pusch.ldpc_decoder_cuda(acc)
    ncb = 8
    llr = (rng.integers(0, 2, (ncb, 25344)) * 20 - 10).astype(np.int8)
    out = np.zeros((ncb, 1056), np.uint8)
    its = dec.decode_batch(out, llr, ncb, pusch.CRC_NONE, 1, 384, 0, 6)
    assert np.all(its == -1)
    for i in range(ncb):
        it, o, layers = ob.port_decode(llr[i], 1, 384, 0, 0, 6)
        assert layers == 46
        assert np.array_equal(o, out[i])


# ---------------------------------------------------------------------------------------------------------------------
# TB level (pusch_decoder interface): rate dematch + LDPC + CRC + TB assembly, HARQ with GPU-resident soft buffers
# ---------------------------------------------------------------------------------------------------------------------
TB_CASES = [
    # prb, qm, R, layers, bg, nref
    (273, 8, 948, 4, 1, 12611),
    (52, 2, 120, 1, 2, 25344),
    (52, 2, 449, 1, 1, 25344),
    (52, 4, 378, 1, 1, 25344),
    (52, 4, 658, 1, 1, 25344),
    (25, 2, 120, 1, 2, 25344),
    (10, 4, 490, 1, 2, 25344),
    (4, 2, 308, 1, 2, 25344),
    (1, 2, 120, 1, 2, 25344),
]


def _tb_sequence(acc, port, rng, prb, qm, R, nl, bg, nref, mu, slot0, early_stop, max_it, rvs=(0, 2, 3, 1)):
    tbs = synth.tbs_for(prb, qm, R, nl)
    nllr = prb * 156 * qm * nl
    tb = rng.integers(0, 256, tbs // 8, dtype=np.uint8)
    gdec = pusch.pusch_decoder_cuda(acc)
    metas = pusch.segment(tbs, bg, qm, nl, nllr)
    tb_g = np.zeros(tbs // 8, np.uint8)
    for i, rv in enumerate(rvs):
        cw = synth.encode_tb(tb, bg, rv, qm, nref, nl, nllr)
        llr = awgn_llrs(rng, cw, mu)
        tb_p, res_p = port.decode(slot0, tbs // 8, llr, bg, rv, qm, nref, nl, max_it, early_stop, i == 0)
        cfg = pusch.pusch_decoder_configuration(bg, rv, qm, nref, nl, max_it, early_stop, i == 0)
        gdec.new_data(tb_g, slot0, None, cfg)
        gdec.on_new_softbits(llr)
        res_g = gdec.on_end_softbits()
        key = (prb, qm, R, nl, rv, mu)
        assert res_g.tb_crc_ok == res_p.tb_crc_ok, key
        assert res_g.nof_codeblocks_total == res_p.nof_codeblocks == len(metas), key
        assert res_g.nof_observations == res_p.nof_observations, key
        assert (res_g.iter_min, res_g.iter_max) == (res_p.iter_min, res_p.iter_max), key
        assert abs(res_g.iter_mean - res_p.iter_mean) < 1e-4, key
        if res_p.tb_crc_ok:
            assert np.array_equal(tb_g, tb_p), key
            assert np.array_equal(tb_g, tb), key
        # Combined soft buffers, byte for byte.
        hb = C.cast(port.harq[slot0][0], C.POINTER(ob_harq_struct())).contents
        for cb, m in enumerate(metas):
            soft_p = np.ctypeslib.as_array(hb.soft, shape=(len(metas) * 25344,))[cb * 25344:cb * 25344 + m.full_length]
            soft_g = acc.read_softbuffer(slot0 + cb, m.full_length)
            assert np.array_equal(soft_p, soft_g), (key, cb)
            assert acc.read_cb_crc(slot0 + cb) == bool(hb.crc[cb]), (key, cb)
        if res_p.tb_crc_ok:
            break
    return res_g


def ob_harq_struct():
    class H(C.Structure):
        _fields_ = [("nof_cbs", C.c_uint32), ("crc", C.c_uint8 * 162), ("soft", C.POINTER(C.c_int8)),
                    ("data", C.POINTER(C.c_uint8))]

    return H


@pytest.mark.parametrize("case", TB_CASES)
def test_pusch_tb_harq_sequences(acc, case):
    prb, qm, R, nl, bg, nref = case
    rng = np.random.default_rng(prb * 1000 + qm)
    # Low SNR first (several retransmissions, soft combining), then new TBs in the SAME HARQ slots: the reference never
    # clears a slot, so the new transmission decodes with the stale LLRs the old one left (SURVEY.md 8(a) trap 3).
    base = {8: 14.0, 4: 5.0, 2: 1.6}[qm]
    slot0 = 1000 + 400 * TB_CASES.index(case)  # slots no other test touches: the port starts from the same zeros
    port = ob.PortPusch()
    _tb_sequence(acc, port, rng, prb, qm, R, nl, bg, nref, base * 0.55, slot0, True, 6)
    _tb_sequence(acc, port, rng, prb, qm, R, nl, bg, nref, base * 1.6, slot0, True, 6)
    _tb_sequence(acc, port, rng, prb, qm, R, nl, bg, nref, base * 0.8, slot0, False, 4)
    _tb_sequence(acc, port, rng, prb, qm, R, nl, bg, nref, base * 1.0, slot0 + 200, True, 6)


def test_pusch_tb_random_llrs_never_converge(acc):
    """Worst case of BASELINE config 2: random +-10 LLRs (pusch_decoder_hwacc_benchmark.cpp:377-380), 6 iterations."""
    rng = np.random.default_rng(5)
    tbs = synth.tbs_for(273, 8, 948, 4)
    nllr = 273 * 156 * 8 * 4
    llr = (rng.integers(0, 2, nllr) * 20 - 10).astype(np.int8)
    port = ob.PortPusch()
    tb_p, res_p = port.decode(1, tbs // 8, llr, 1, 0, 8, 12611, 4, 6, True, True)
    gdec = pusch.pusch_decoder_cuda(acc)
    tb_g = np.zeros(tbs // 8, np.uint8)
    gdec.new_data(tb_g, 400, None, pusch.pusch_decoder_configuration(1, 0, 8, 12611, 4, 6, True, True))
    gdec.on_new_softbits(llr)
    res_g = gdec.on_end_softbits()
    assert res_g.tb_crc_ok == res_p.tb_crc_ok == 0
    assert res_g.nof_codeblocks_total == 152
    assert (res_g.iter_min, res_g.iter_max, res_g.nof_observations) == (res_p.iter_min, res_p.iter_max, 152)


def test_pusch_batch_of_tbs_device_resident_and_host(acc):
    """submit_tbs with several TBs in one launch set (host LLRs) gives the same results as one TB at a time."""
    rng = np.random.default_rng(6)
    prb, qm, R, nl, bg, nref = 52, 4, 658, 1, 1, 25344
    tbs = synth.tbs_for(prb, qm, R, nl)
    nllr = prb * 156 * qm * nl
    ntb = 6
    tbs_bytes, llrs = [], []
    for i in range(ntb):
        tb = rng.integers(0, 256, tbs // 8, dtype=np.uint8)
        cw = synth.encode_tb(tb, bg, 0, qm, nref, nl, nllr)
        tbs_bytes.append(tb)
        llrs.append(awgn_llrs(rng, cw, 4.0 if i % 2 == 0 else 1.2))
    nseg = len(pusch.segment(tbs, bg, qm, nl, nllr))
    cfgs = [pusch.TbConfig(tbs, bg, 0, qm, nref, nl, 6, 1, 1, 600 + i * nseg) for i in range(ntb)]
    tickets = pusch.submit_tbs(acc, cfgs, llrs)
    for i, t in enumerate(tickets):
        out = np.zeros(tbs // 8, np.uint8)
        res = pusch.poll_tb(acc, t, out)
        port = ob.PortPusch()
        tb_p, res_p = port.decode(0, tbs // 8, llrs[i], bg, 0, qm, nref, nl, 6, True, True)
        assert res.tb_crc_ok == res_p.tb_crc_ok
        assert (res.iter_min, res.iter_max) == (res_p.iter_min, res_p.iter_max)
        if res_p.tb_crc_ok:
            assert np.array_equal(out, tbs_bytes[i])


def test_pusch_soft_bits_in_separate_page_locked_buffers(acc):
    """Soft bits handed over in page-locked buffers that are not adjacent (one per decoder instance behind the plugin
    interface) are read by the gather kernel (set_h2d_gather): same results as with copy-engine jobs and as the oracle;
    ragged sizes, one of them not a multiple of 16 bytes."""
    rng = np.random.default_rng(66)
    lib = capi.lib()
    shapes = [(52, 4, 658, 1, 1), (25, 2, 120, 1, 2), (106, 6, 873, 2, 1), (3, 2, 308, 1, 2), (52, 2, 449, 1, 1), (1, 2, 120, 1, 2),
              (52, 4, 378, 1, 1)]
    cfgs, bufs, llrs, payloads, ptrs, slot = [], [], [], [], [], 5000
    for prb, qm, R, nl, bg in shapes:
        tbs = synth.tbs_for(prb, qm, R, nl)
        nllr = prb * 156 * qm * nl - (qm * nl if prb == 3 else 0)  # 3 PRB: one symbol fewer -> 930 soft bits
        tb = rng.integers(0, 256, tbs // 8, dtype=np.uint8)
        llr = awgn_llrs(rng, synth.encode_tb(tb, bg, 0, qm, 25344, nl, nllr), 3.0)
        p = lib.srsran_cuda_pusch_dec_host_alloc(nllr)
        assert p
        ptrs.append(p)
        b = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int8)), shape=(nllr,))
        b[...] = llr
        cfgs.append(pusch.TbConfig(tbs, bg, 0, qm, 25344, nl, 6, 1, 1, slot))
        slot += len(pusch.segment(tbs, bg, qm, nl, nllr))
        bufs.append(b); llrs.append(llr); payloads.append(tb)
    assert any(b.size % 16 for b in bufs)
    try:
        acc.set_direct_io(False, True)  # a batch this small would otherwise be read by the dematcher itself
        results = {}
        for ctas in (0, 8, 32):
            acc.set_h2d_gather(ctas)
            n0 = acc.launch_count
            tickets = pusch.submit_tbs(acc, cfgs, bufs)
            launches = acc.launch_count - n0
            outs = [np.zeros(c.tbs_bits // 8, np.uint8) for c in cfgs]
            results[ctas] = ([pusch.poll_tb(acc, t, o) for t, o in zip(tickets, outs)], outs, launches)
        assert results[32][2] == results[0][2] + 1 and results[8][2] == results[0][2] + 1  # the gather kernel did run
        port = ob.PortPusch()
        for i, c in enumerate(cfgs):
            _, res_p = port.decode(i, c.tbs_bits // 8, llrs[i], c.base_graph, 0, c.modulation, 25344, c.nof_layers, 6, True, True)
            for ctas in results:
                res, outs, _ = results[ctas]
                assert res[i].tb_crc_ok == res_p.tb_crc_ok
                assert (res[i].iter_min, res[i].iter_max) == (res_p.iter_min, res_p.iter_max)
                if res_p.tb_crc_ok:
                    assert np.array_equal(outs[i], payloads[i])
        assert sum(r.tb_crc_ok for r in results[32][0]) >= 4
        # ... and the same batch with the dematcher reading the page-locked buffers itself (the default for small batches),
        # with and without the results written straight to host memory.
        for din, dout in ((True, True), (True, False), (False, False)):
            acc.set_direct_io(din, dout)
            n0 = acc.launch_count
            tickets = pusch.submit_tbs(acc, cfgs, bufs)
            assert acc.launch_count - n0 == results[0][2] + (0 if din else 1)
            for i, t in enumerate(tickets):
                out = np.zeros(cfgs[i].tbs_bits // 8, np.uint8)
                res = pusch.poll_tb(acc, t, out)
                ref = results[0][0][i]
                assert (res.tb_crc_ok, res.iter_min, res.iter_max, res.nof_observations) == \
                    (ref.tb_crc_ok, ref.iter_min, ref.iter_max, ref.nof_observations)
                if ref.tb_crc_ok:
                    assert np.array_equal(out, payloads[i])
    finally:
        acc.set_h2d_gather(32)
        acc.set_direct_io(True, True)
        for p in ptrs:
            lib.srsran_cuda_pusch_dec_host_free(p)


def test_pusch_streamed_tb_with_scattered_harq_slots(acc):
    """stream_begin / stream_push / stream_submit with non-consecutive HARQ slots (rx_buffer absolute code-block ids):
    same TB result as the oracle over rv0 -> rv2, per-code-block CRC flags and iteration observations included."""
    rng = np.random.default_rng(62)
    prb, qm, R, nl, bg, nref = 106, 6, 873, 2, 1, 0
    tbs = synth.tbs_for(prb, qm, R, nl)
    nllr = prb * 156 * qm * nl
    nseg = len(pusch.segment(tbs, bg, qm, nl, nllr))
    ids = [3000 + 7 * i + (i % 3) for i in range(nseg)]  # scattered, increasing
    payload = rng.integers(0, 256, tbs // 8, dtype=np.uint8)
    port = ob.PortPusch()
    for rv, new_data, mu in ((0, 1, 1.3), (2, 0, 1.3)):
        llr = awgn_llrs(rng, synth.encode_tb(payload, bg, rv, qm, nref, nl, nllr), mu)
        cuts = sorted(rng.choice(np.arange(1, nllr), 9, replace=False).tolist())
        blocks = [np.ascontiguousarray(b) for b in np.split(llr, cuts)]
        cfg = pusch.TbConfig(tbs, bg, rv, qm, nref, nl, 6, 1, new_data, 0)
        ticket = pusch.submit_tb_streamed(acc, cfg, blocks, ids)
        out = np.zeros(tbs // 8, np.uint8)
        res = pusch.poll_tb(acc, ticket, out)
        crc, its = pusch.tb_cb_outputs(acc, ticket, nseg)
        tb_p, res_p = port.decode(0, tbs // 8, llr, bg, rv, qm, nref, nl, 6, True, bool(new_data))
        assert res.tb_crc_ok == res_p.tb_crc_ok
        assert (res.iter_min, res.iter_max, res.nof_observations) == (res_p.iter_min, res_p.iter_max, res_p.nof_observations)
        obs = its[its != 0xffffffff]
        assert obs.size == res_p.nof_observations
        if obs.size:
            assert (obs.min(), obs.max()) == (res_p.iter_min, res_p.iter_max)
        for i, slot in enumerate(ids):
            assert acc.read_cb_crc(slot) == bool(crc[i])
        if res_p.tb_crc_ok:
            assert np.array_equal(out, payload)
            assert crc.all()


def test_pusch_poll_tbs_whole_batch(acc):
    """poll_tbs (one call per batch, all or nothing when not blocking) returns what poll_tb returns ticket by ticket; a
    retransmission batch (rv2, same HARQ slots) exercises the descriptor fast path with a non-zero soft-buffer extent."""
    rng = np.random.default_rng(61)
    prb, qm, R, nl, bg, nref = 106, 6, 873, 2, 1, 0
    tbs = synth.tbs_for(prb, qm, R, nl)
    nllr = prb * 156 * qm * nl
    nseg = len(pusch.segment(tbs, bg, qm, nl, nllr))
    ntb = 5
    payload = [rng.integers(0, 256, tbs // 8, dtype=np.uint8) for _ in range(ntb)]
    ports = [ob.PortPusch() for _ in range(ntb)]
    for rv, new_data, mu in ((0, 1, 1.4), (2, 0, 1.4)):
        llrs = [awgn_llrs(rng, synth.encode_tb(payload[i], bg, rv, qm, nref, nl, nllr), mu) for i in range(ntb)]
        cfgs = [pusch.TbConfig(tbs, bg, rv, qm, nref, nl, 6, 1, new_data, 2000 + i * nseg) for i in range(ntb)]
        tickets = pusch.submit_tbs(acc, cfgs, llrs)
        outs = [np.zeros(tbs // 8, np.uint8) for _ in range(ntb)]
        res = None
        while res is None:
            res = pusch.poll_tbs(acc, tickets, outs, block=False)
        assert len(res) == ntb
        for i in range(ntb):
            tb_p, res_p = ports[i].decode(0, tbs // 8, llrs[i], bg, rv, qm, nref, nl, 6, True, bool(new_data))
            assert res[i].tb_crc_ok == res_p.tb_crc_ok
            assert (res[i].iter_min, res[i].iter_max, res[i].nof_observations) == \
                (res_p.iter_min, res_p.iter_max, res_p.nof_observations)
            assert abs(res[i].iter_mean - res_p.iter_mean) < 1e-4
            if res_p.tb_crc_ok:
                assert np.array_equal(outs[i], payload[i])


def test_pusch_transport_blocks_left_in_hbm(acc):
    """set_tb_host_copy(0): the decoded transport blocks stay in device memory (tb_data_device) for a consumer on the device
    side of the link; results and CRC verdicts still come back, poll_tbs leaves its output buffers untouched."""
    rng = np.random.default_rng(62)
    prb, qm, R, nl, bg, nref = 106, 6, 873, 2, 1, 0
    tbs = synth.tbs_for(prb, qm, R, nl)
    nllr = prb * 156 * qm * nl
    nseg = len(pusch.segment(tbs, bg, qm, nl, nllr))
    ntb = 4
    payload = [rng.integers(0, 256, tbs // 8, dtype=np.uint8) for _ in range(ntb)]
    llrs = [awgn_llrs(rng, synth.encode_tb(payload[i], bg, 0, qm, nref, nl, nllr), 16.0) for i in range(ntb)]
    cfgs = [pusch.TbConfig(tbs, bg, 0, qm, nref, nl, 6, 1, 1, 3000 + i * nseg) for i in range(ntb)]
    acc.set_tb_host_copy(False)
    try:
        tickets = pusch.submit_tbs(acc, cfgs, llrs)
        acc.synchronize()
        import ctypes

        cudart = ctypes.CDLL("libcudart.so.12")  # the runtime the library itself is linked against
        cudart.cudaMemcpy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
        for i, tk in enumerate(tickets):
            ptr = pusch.tb_data_device(acc, tk)
            assert ptr
            got = np.zeros(tbs // 8, np.uint8)
            assert cudart.cudaMemcpy(got.ctypes.data, ptr, got.size, 2) == 0  # device -> host
            assert np.array_equal(got, payload[i]), i
            assert pusch.tb_data(acc, tk, tbs // 8) is None  # nothing was copied to the host result buffer
        outs = [np.full(tbs // 8, 0x5A, np.uint8) for _ in range(ntb)]
        res = pusch.poll_tbs(acc, tickets, outs)
        assert all(r.tb_crc_ok for r in res)
        assert all(np.all(o == 0x5A) for o in outs)
    finally:
        acc.set_tb_host_copy(True)
    # back to the default: the next batch is copied out again
    tickets = pusch.submit_tbs(acc, cfgs, llrs)
    outs = [np.zeros(tbs // 8, np.uint8) for _ in range(ntb)]
    res = pusch.poll_tbs(acc, tickets, outs)
    assert all(r.tb_crc_ok for r in res) and all(np.array_equal(outs[i], payload[i]) for i in range(ntb))


# ---------------------------------------------------------------------------------------------------------------------
# hal::hw_accelerator_pusch_dec call sequence, as driven by pusch_decoder_hw_impl::on_end_softbits (:132-342)
# ---------------------------------------------------------------------------------------------------------------------
def test_hw_accelerator_call_sequence(acc):
    rng = np.random.default_rng(7)
    prb, qm, R, nl, bg, nref = 52, 4, 658, 1, 1, 25344
    tbs = synth.tbs_for(prb, qm, R, nl)
    nllr = prb * 156 * qm * nl
    tb = rng.integers(0, 256, tbs // 8, dtype=np.uint8)
    hw = pusch.hw_accelerator_pusch_dec_cuda(acc)
    assert hw.is_external_harq_supported()
    metas = pusch.segment(tbs, bg, qm, nl, nllr)
    nseg = len(metas)
    port = ob.PortPusch()
    slot0 = 900
    for i, rv in enumerate((0, 2)):
        cw = synth.encode_tb(tb, bg, rv, qm, nref, nl, nllr)
        llr = awgn_llrs(rng, cw, 1.6)
        tb_p, res_p = port.decode(3, tbs // 8, llr, bg, rv, qm, nref, nl, 6, True, i == 0)
        hw.reserve_queue()
        for cb, m in enumerate(metas):
            K = m.full_length // 3
            cfg = pusch.CbConfig(bg, qm, nseg, rv, m.rm_length, m.lifting_size, m.full_length, nref,
                                 K - m.nof_crc_bits - m.nof_filler_bits, m.nof_filler_bits, 6, 1, int(i == 0), 24,
                                 pusch.CB_CRC24B, slot0 + cb)
            hw.configure_operation(cfg, cb)
            assert hw.enqueue_operation(llr[m.cw_offset:m.cw_offset + m.rm_length], None, cb)
        hb = C.cast(port.harq[3][0], C.POINTER(ob_harq_struct())).contents
        for cb, m in enumerate(metas):
            K = m.full_length // 3
            bits = np.zeros(K // 8, np.uint8)
            soft = np.zeros(m.full_length, np.int8)
            while not hw.dequeue_operation(bits, soft, cb):
                pass
            crc_ok, iters = hw.read_operation_outputs(cb, slot0 + cb)
            soft_p = np.ctypeslib.as_array(hb.soft, shape=(nseg * 25344,))[cb * 25344:cb * 25344 + m.full_length]
            assert np.array_equal(soft, soft_p)
            data_p = np.ctypeslib.as_array(hb.data, shape=(nseg * (25344 // 8 + 8),))[cb * 3176:cb * 3176 + K // 8]
            # The software path skips code blocks whose CRC was already ok; the accelerator decodes what it is given.
            if i == 0:
                assert crc_ok == bool(hb.crc[cb])
                assert np.array_equal(bits, data_p)
        hw.free_queue()


# ---------------------------------------------------------------------------------------------------------------------
# Packed decoder (four code blocks per CTA): groups of 1..4 same-shape code blocks, mixed shapes in one batch, both
# decoder variants against the oracle's pusch_codeblock_decoder restatement
# ---------------------------------------------------------------------------------------------------------------------
def _cb_llrs(rng, bg, z, F, crc_poly, E, qm, rv, nref, mu):
    K = ob.kb(bg) * z
    crc_len = {1: 24, 2: 24, 3: 16}[crc_poly]
    npay = K - F - crc_len
    msg = np.zeros(K, np.uint8)
    msg[:npay] = rng.integers(0, 2, npay, dtype=np.uint8)
    c = ob.port_crc(crc_poly, np.packbits(msg[:npay]), npay)
    msg[npay:npay + crc_len] = [(c >> (crc_len - 1 - i)) & 1 for i in range(crc_len)]
    cw = synth.ldpc_encode(msg, bg, z)
    return awgn_llrs(rng, synth.rate_match(cw, bg, z, F, E, rv, qm, nref), mu)


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4])
def test_packed_decoder_groups_via_hal(acc, variant):
    rng = np.random.default_rng(40 + variant)
    hw = pusch.hw_accelerator_pusch_dec_cuda(acc)
    acc.set_decoder_variant(variant)
    try:
        for rnd in range(6):
            early_stop = int(rng.integers(0, 2))
            max_it = int(rng.integers(1, 7))
            ops = []
            slot = 7000
            # several runs of same-shape code blocks back to back: groups of 4 plus a remainder, shapes alternate
            for _ in range(int(rng.integers(2, 5))):
                bg = int(rng.integers(1, 3))
                z = int(rng.choice([144, 160, 208, 256, 288, 320, 384]))
                K, N = ob.kb(bg) * z, ob.ns(bg) * z
                qm = int(rng.choice([2, 4, 6, 8]))
                crc_poly = pusch.CRC24B
                F = int(rng.integers(0, 5)) * 8
                # few layers: E just above the systematic part
                E = (int((K - 2 * z - F) * rng.uniform(1.04, 1.9 if rng.random() < 0.4 else 1.25)) // qm) * qm
                nref = 0 if rng.random() < 0.5 else int(N * 0.6)
                mu = float(rng.choice([3.0, 6.0, 12.0, 20.0]))
                for _ in range(int(rng.integers(1, 8))):
                    llr = _cb_llrs(rng, bg, z, F, crc_poly, E, qm, 0, nref, mu)
                    if rng.random() < 0.1:
                        llr[:] = 0
                    ops.append((bg, z, qm, F, E, nref, llr, slot))
                    slot += 1
            hw.reserve_queue()
            for i, (bg, z, qm, F, E, nref, llr, s) in enumerate(ops):
                K, N = ob.kb(bg) * z, ob.ns(bg) * z
                cfg = pusch.CbConfig(bg, qm, len(ops), 0, E, z, N, nref, K - 24 - F, F, max_it, early_stop, 1, 24,
                                     pusch.CB_CRC24B, s)
                hw.configure_operation(cfg, i)
                assert hw.enqueue_operation(llr, None, i)
            for i, (bg, z, qm, F, E, nref, llr, s) in enumerate(ops):
                K, N = ob.kb(bg) * z, ob.ns(bg) * z
                bits = np.full(K // 8, 0x5A, np.uint8)
                soft = np.zeros(N, np.int8)
                while not hw.dequeue_operation(bits, soft, i):
                    pass
                crc_ok, iters = hw.read_operation_outputs(i, s)
                want_soft = np.zeros(N, np.int8)
                want_bits = np.full(K // 8, 0x5A, np.uint8)
                # the slots are reused round after round: start the oracle from what the slot held
                want_soft[:] = _packed_prev.get(s, np.zeros(25344, np.int8))[:N]
                it = ob.port().oracle_cb_decode(ob._p8(want_bits), ob._pi(want_soft), N, ob._pi(llr), E, 1, 0, qm, nref, F,
                                                bg, z, pusch.CRC24B, early_stop, max_it)
                full = _packed_prev.get(s, np.zeros(25344, np.int8)).copy()
                full[:N] = want_soft
                _packed_prev[s] = full
                key = (variant, rnd, i, bg, z, qm, F, E, nref, early_stop, max_it)
                assert np.array_equal(soft, want_soft), key
                assert crc_ok == (it >= 0), key
                assert iters == (it if it >= 0 else max_it), key
                if not (early_stop and not llr.any()):
                    assert np.array_equal(bits, want_bits), key
            hw.free_queue()
    finally:
        acc.set_decoder_variant(0)


@pytest.mark.parametrize("variant", [0, 4])
def test_packed_decoder_many_iterations_saturated_inputs(acc, variant):
    """Groups of four / two with up to 40 iterations, many layers and inputs that drive soft values to +-infinity in both
    directions (strong wrong parity bits, +-127 and +-120 everywhere): a promoted value is updated dozens of times, its
    binary16 representation grows up to the IEEE infinities and must behave like the reference's sticky +-127."""
    rng = np.random.default_rng(70 + variant)
    hw = pusch.hw_accelerator_pusch_dec_cuda(acc)
    acc.set_decoder_variant(variant)
    try:
        for (bg, z, max_it) in [(1, 144, 40), (2, 208, 25), (1, 160, 12)]:
            K, N = ob.kb(bg) * z, ob.ns(bg) * z
            F, qm, nref = 16, 2, 0
            E = (int((K - 2 * z - F) * 1.8) // qm) * qm
            ops = []
            for c in range(7):
                if c % 3 == 0:
                    llr = _cb_llrs(rng, bg, z, F, pusch.CRC24B, E, qm, 0, nref, 28.0)
                    flip = rng.choice(np.arange(E // 2, E), 30, replace=False)
                    llr[flip] = -llr[flip]
                elif c % 3 == 1:
                    llr = rng.choice(np.array([-127, 127, -120, 120, 119, -119], np.int8), E)
                else:
                    llr = np.where(rng.random(E) < 0.5, 120, -120).astype(np.int8)
                ops.append(llr)
            hw.reserve_queue()
            for i, llr in enumerate(ops):
                cfg = pusch.CbConfig(bg, qm, len(ops), 0, E, z, N, nref, K - 24 - F, F, max_it, 1, 1, 24, pusch.CB_CRC24B,
                                     7400 + i)
                hw.configure_operation(cfg, i)
                assert hw.enqueue_operation(llr, None, i)
            for i, llr in enumerate(ops):
                bits = np.full(K // 8, 0x5A, np.uint8)
                soft = np.zeros(N, np.int8)
                while not hw.dequeue_operation(bits, soft, i):
                    pass
                crc_ok, iters = hw.read_operation_outputs(i, 7400 + i)
                want_soft = np.zeros(N, np.int8)
                want_bits = np.full(K // 8, 0x5A, np.uint8)
                it = ob.port().oracle_cb_decode(ob._p8(want_bits), ob._pi(want_soft), N, ob._pi(llr), E, 1, 0, qm, nref, F,
                                                bg, z, pusch.CRC24B, 1, max_it)
                key = (variant, bg, z, max_it, i)
                assert np.array_equal(soft, want_soft), key
                assert crc_ok == (it >= 0), key
                assert iters == (it if it >= 0 else max_it), key
                assert np.array_equal(bits, want_bits), key
            hw.free_queue()
    finally:
        acc.set_decoder_variant(0)


_packed_prev = {}


# ---------------------------------------------------------------------------------------------------------------------
# Drop-in proof: the reference's own pusch_decoder_hw_impl driving this repo's C++ hal accelerator (CUDA), compared with
# the reference's software pusch_decoder_impl on the same transport blocks (oracle/hwacc_harness.cpp)
# ---------------------------------------------------------------------------------------------------------------------
def test_reference_hw_decoder_over_cuda_accelerator():
    import subprocess
    from pathlib import Path

    exe = Path(__file__).resolve().parent.parent / "oracle" / "_ref" / "hwacc_parity"
    if not exe.exists():
        pytest.skip("oracle/_ref/hwacc_parity not built (needs /root/reference at build time)")
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    print(r.stdout)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "PASSED" in r.stdout


def test_hw_accelerator_queue_full_retry():
    """enqueue_operation returns False when the queue is full; the caller dequeues what it has and retries, exactly like
    pusch_decoder_hw_impl::on_end_softbits (pusch_decoder_hw_impl.cpp:189-338). A 3-code-block TB through a queue of 2."""
    rng = np.random.default_rng(9)
    small = pusch.Accelerator(device=0, max_cbs_in_flight=2, nof_harq_cb_slots=16)
    try:
        hw = pusch.hw_accelerator_pusch_dec_cuda(small)
        prb, qm, R, nl, bg, nref = 52, 4, 658, 1, 1, 25344
        tbs = synth.tbs_for(prb, qm, R, nl)
        nllr = prb * 156 * qm * nl
        tb = rng.integers(0, 256, tbs // 8, dtype=np.uint8)
        llr = awgn_llrs(rng, synth.encode_tb(tb, bg, 0, qm, nref, nl, nllr), 6.0)
        metas = pusch.segment(tbs, bg, qm, nl, nllr)
        assert len(metas) == 3
        port = ob.PortPusch()
        tb_p, res_p = port.decode(0, tbs // 8, llr, bg, 0, qm, nref, nl, 6, True, True)
        assert res_p.tb_crc_ok
        hw.reserve_queue()
        enq = deq = 0
        got = {}
        while deq != len(metas):
            while enq != len(metas):
                m = metas[enq]
                K = m.full_length // 3
                hw.configure_operation(pusch.CbConfig(bg, qm, len(metas), 0, m.rm_length, m.lifting_size, m.full_length, nref,
                                                      K - m.nof_crc_bits - m.nof_filler_bits, m.nof_filler_bits, 6, 1, 1, 24,
                                                      pusch.CB_CRC24B, enq), enq)
                if not hw.enqueue_operation(llr[m.cw_offset:m.cw_offset + m.rm_length], None, enq):
                    break  # queue full
                enq += 1
            assert enq > deq
            while deq != enq:
                K = metas[deq].full_length // 3
                bits = np.zeros(K // 8, np.uint8)
                while not hw.dequeue_operation(bits, None, deq):
                    pass
                got[deq] = (bits, hw.read_operation_outputs(deq, deq))
                deq += 1
        hw.free_queue()
        assert enq == 3 and all(got[i][1][0] for i in range(3))
        # the payload bits of the three code blocks concatenate to the TB (+ 24 CRC bits)
        payload = np.concatenate([np.unpackbits(got[i][0])[:metas[i].full_length // 3 - 24 - metas[i].nof_filler_bits]
                                  for i in range(3)])[:tbs]
        assert np.array_equal(np.packbits(payload), tb)
    finally:
        small.close()


def test_pusch_ragged_batch_mixed_shapes(acc):
    """One submit of transport blocks with different base graphs, lifting sizes, code-block counts and code rates (BASELINE
    config 3 + a multi-code-block TB): every decoder kernel runs in the same batch, on concurrent streams."""
    rng = np.random.default_rng(10)
    cases = [(52, 2, 120, 1, 2, 0.9), (52, 2, 449, 1, 1, 2.0), (52, 4, 378, 1, 1, 3.0), (52, 4, 658, 1, 1, 6.0),
             (25, 2, 120, 1, 2, 0.4), (10, 4, 490, 1, 2, 4.0), (4, 2, 308, 1, 2, 1.5), (1, 2, 120, 1, 2, 0.9),
             (106, 8, 948, 2, 1, 18.0), (30, 6, 567, 1, 1, 6.0), (52, 4, 658, 1, 1, 2.0)]
    cfgs, llrs, ref = [], [], []
    slot = 5000
    for k in range(22):
        prb, qm, R, nl, bg, mu = cases[k % len(cases)]
        tbs = synth.tbs_for(prb, qm, R, nl)
        nllr = prb * 156 * qm * nl
        nref = 25344 if k % 3 else 12611
        tb = rng.integers(0, 256, tbs // 8, dtype=np.uint8)
        llr = awgn_llrs(rng, synth.encode_tb(tb, bg, 0, qm, nref, nl, nllr), mu)
        nseg = len(pusch.segment(tbs, bg, qm, nl, nllr))
        es = int(k % 4 != 3)
        cfgs.append(pusch.TbConfig(tbs, bg, 0, qm, nref, nl, 5, es, 1, slot))
        slot += nseg
        llrs.append(llr)
        port = ob.PortPusch()
        ref.append((tb, port.decode(0, tbs // 8, llr, bg, 0, qm, nref, nl, 5, bool(es), True)))
    tickets = pusch.submit_tbs(acc, cfgs, llrs)
    nok = 0
    for k, t in enumerate(tickets):
        out = np.zeros(cfgs[k].tbs_bits // 8, np.uint8)
        res = pusch.poll_tb(acc, t, out)
        tb, (tb_p, res_p) = ref[k]
        assert res.tb_crc_ok == res_p.tb_crc_ok, k
        assert (res.nof_codeblocks_total, res.nof_observations, res.iter_min, res.iter_max) == \
               (res_p.nof_codeblocks, res_p.nof_observations, res_p.iter_min, res_p.iter_max), k
        if res_p.tb_crc_ok:
            nok += 1
            assert np.array_equal(out, tb), k
    assert 5 <= nok <= 22
