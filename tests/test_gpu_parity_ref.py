"""GPU parity tests on the configurations the published numbers are quoted on (BASELINE.json configs 1, 2, 4), compared
DIRECTLY with the compiled, unmodified reference (oracle/_ref: pusch_decoder_impl / ldpc_decoder_avx512 / rate dematcher /
CRC) where it is available on the box, else with the oracle port (which test_oracle_cpu.py pins to the reference).
Modelled on pusch_decoder_vectortest.cpp:261-397 (rv sequences with soft combining; TB bytes, CRC, statistics), extended to
every combined soft-buffer byte and every code-block CRC flag. Also bounded slices of the seed sweeps."""
import numpy as np
import pytest

from oracle import bindings as ob
from srsran_projectvtlmo_b200 import pusch, synth
from tests.helpers import awgn_llrs

pytestmark = pytest.mark.gpu

HAVE_REF = ob.ref() is not None and ob.ref_flavour() is not None


class _Checker:
    """pusch_decoder_impl of the compiled reference (or the port) with persistent HARQ buffers; returns the per-code-block
    CRC flags and soft buffers after every transmission."""

    def __init__(self, threads=1):
        self.ref = ob.RefPusch(nof_threads=threads) if HAVE_REF else None
        self.port = None if HAVE_REF else ob.PortPusch()

    def decode(self, key, tb_bytes, llr, bg, rv, qm, nref, nl, max_it, early_stop, new_data):
        if self.ref is not None:
            tb, res, crcs, soft = self.ref.decode(key, tb_bytes, llr, bg, rv, qm, nref, nl, max_it, early_stop, new_data,
                                                  want_soft=True)
            metas = ob.ref_segment(tb_bytes * 8, bg, qm, nl, llr.size)
            offs = np.concatenate([[0], np.cumsum([m.full_length for m in metas])])
            softs = [soft[offs[i]:offs[i + 1]] for i in range(len(metas))]
            return tb, res, [bool(c) for c in crcs], softs
        tb, res = self.port.decode(key, tb_bytes, llr, bg, rv, qm, nref, nl, max_it, early_stop, new_data)
        crcs, softs = self.port.harq_state(key, ob.port_segment(tb_bytes * 8, bg, qm, nl, llr.size))
        return tb, res, crcs, softs


def _assert_tb_equal(key, res_g, tb_g, res_r, tb_r, payload):
    assert res_g.tb_crc_ok == res_r.tb_crc_ok, key
    assert res_g.nof_codeblocks_total == res_r.nof_codeblocks, key
    assert res_g.nof_observations == res_r.nof_observations, key
    assert (res_g.iter_min, res_g.iter_max) == (res_r.iter_min, res_r.iter_max), key
    assert abs(res_g.iter_mean - res_r.iter_mean) < 1e-4, key
    if res_r.tb_crc_ok:
        assert np.array_equal(tb_g, tb_r), key
        assert np.array_equal(tb_g, payload), key


def _assert_harq_equal(key, acc, slots, crcs_r, softs_r):
    for cb, slot in enumerate(slots):
        soft_g = acc.read_softbuffer(slot, softs_r[cb].size)
        assert np.array_equal(soft_g, softs_r[cb]), (key, cb, np.nonzero(soft_g != softs_r[cb])[0][:8])
        assert acc.read_cb_crc(slot) == crcs_r[cb], (key, cb)


# ---------------------------------------------------------------------------------------------------------------------
# BASELINE config 2 = the bench workload: 273 PRB, 256QAM, R = 948/1024, 4 layers, TBS 1 277 992, 152 code blocks of
# BG1 / Z = 384, Nref = 12611 (limited-buffer rate matching: the soft-buffer write set has the stale gap of SURVEY 8(a)
# trap 3). rv 0 -> 2 -> 3 with a failing rv0, first on fresh slots, then new data on the same (now stale) slots.
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("variant", [0, 2])
def test_headline_tb_vs_reference_harq_fresh_and_stale_slots(variant):
    prb, qm, R, nl, bg, nref = 273, 8, 948, 4, 1, 12611
    tbs = synth.tbs_for(prb, qm, R, nl)
    nllr = prb * 156 * qm * nl
    assert tbs == 1277992
    rng = np.random.default_rng(2024 + variant)
    chk = _Checker()
    acc = pusch.Accelerator(device=0, max_cbs_in_flight=256, nof_harq_cb_slots=256)  # fresh (zeroed) HARQ slots
    acc.set_decoder_variant(variant)
    try:
        slot0 = 40
        gdec = pusch.pusch_decoder_cuda(acc)
        nseg = len(pusch.segment(tbs, bg, qm, nl, nllr))
        assert nseg == 152
        slots = [slot0 + i for i in range(nseg)]
        # (SNR, early stop, iterations): a TB that needs combining, then a clean one into the stale slots, then one
        # without early stop that fails its first transmission again.
        for seq, (mu, early_stop, max_it) in enumerate([(8.5, True, 6), (18.0, True, 6), (9.5, False, 4)]):
            payload = rng.integers(0, 256, tbs // 8, dtype=np.uint8)
            tb_g = np.zeros(tbs // 8, np.uint8)
            transmissions = 0
            for i, rv in enumerate((0, 2, 3)):
                llr = awgn_llrs(rng, synth.encode_tb(payload, bg, rv, qm, nref, nl, nllr), mu)
                tb_r, res_r, crcs_r, softs_r = chk.decode(7, tbs // 8, llr, bg, rv, qm, nref, nl, max_it, early_stop, i == 0)
                gdec.new_data(tb_g, slot0, None, pusch.pusch_decoder_configuration(bg, rv, qm, nref, nl, max_it, early_stop, i == 0))
                gdec.on_new_softbits(llr)
                res_g = gdec.on_end_softbits()
                key = (variant, seq, rv, mu)
                _assert_tb_equal(key, res_g, tb_g, res_r, tb_r, payload)
                _assert_harq_equal(key, acc, slots, crcs_r, softs_r)
                transmissions += 1
                if res_r.tb_crc_ok:
                    break
            if seq == 0:
                assert transmissions > 1, "the first sequence is meant to exercise soft combining"
            # (The clean TB of sequence 1 may still need a retransmission: its decoder sees the stale LLRs the previous TB
            # left in the gap the Nref-limited dematcher never writes - exactly the behaviour being compared.)
    finally:
        acc.close()


# ---------------------------------------------------------------------------------------------------------------------
# BASELINE config 4: 64 UEs per slot in ONE submit_tbs call, HARQ rv0 -> rv2 -> rv3 with GPU-resident soft combining
# (new_data only on rv0), SNRs at which rv0 fails for part of the code blocks, early stop on; the combined soft buffers
# and CRC flags of every code block are verified after every transmission. UEs 0..15 carry the full 100 MHz / 4-layer TB
# (152 code blocks of Z = 384: 13 layers after rv2), the others a mix of the 52 / 106 PRB allocations of config 3.
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("variant", [0, 7])  # 7: the 13-layer retransmission groups stage their inputs by bulk copies
def test_config4_64_ues_harq_rv0_rv2_rv3_soft_buffers(variant):
    # (prb, Qm, R, layers, BG, Nref, mu): mu at the waterfall of the rate (calibrated with the compiled reference)
    shapes = [(273, 8, 948, 4, 1, 12611, 12.8)] * 16 + [(106, 6, 873, 2, 1, 0, 8.0)] * 12 + \
             [(52, 4, 658, 1, 1, 25344, 4.4)] * 12 + [(52, 4, 378, 1, 1, 25344, 2.05)] * 8 + \
             [(52, 2, 449, 1, 1, 25344, 1.5)] * 8 + [(52, 2, 120, 1, 2, 25344, 0.6)] * 4 + [(25, 2, 120, 1, 2, 25344, 0.6)] * 4
    assert len(shapes) == 64
    rng = np.random.default_rng(404)
    ues, slot = [], 0
    for ue, (prb, qm, R, nl, bg, nref, mu) in enumerate(shapes):
        tbs = synth.tbs_for(prb, qm, R, nl)
        nllr = prb * 156 * qm * nl
        nseg = len(pusch.segment(tbs, bg, qm, nl, nllr))
        ues.append(dict(prb=prb, qm=qm, R=R, nl=nl, bg=bg, nref=nref, mu=mu * float(rng.choice([0.92, 1.0, 1.1])), tbs=tbs,
                        nllr=nllr, nseg=nseg, slot0=slot, payload=rng.integers(0, 256, tbs // 8, dtype=np.uint8),
                        done=False))
        slot += nseg
    acc = pusch.Accelerator(device=0, max_cbs_in_flight=slot, nof_harq_cb_slots=slot)
    acc.set_decoder_variant(variant)
    chk = _Checker()
    try:
        failed_first = 0
        for i, rv in enumerate((0, 2, 3)):
            active = [u for u in ues if not u["done"]]
            if not active:
                break
            llrs = [awgn_llrs(rng, synth.encode_tb(u["payload"], u["bg"], rv, u["qm"], u["nref"], u["nl"], u["nllr"]), u["mu"])
                    for u in active]
            cfgs = [pusch.TbConfig(u["tbs"], u["bg"], rv, u["qm"], u["nref"], u["nl"], 6, 1, int(i == 0), u["slot0"])
                    for u in active]
            tickets = pusch.submit_tbs(acc, cfgs, llrs)
            outs = [np.zeros(u["tbs"] // 8, np.uint8) for u in active]
            res = pusch.poll_tbs(acc, tickets, outs, block=True)
            for k, u in enumerate(active):
                tb_r, res_r, crcs_r, softs_r = chk.decode(u["slot0"], u["tbs"] // 8, llrs[k], u["bg"], rv, u["qm"], u["nref"],
                                                          u["nl"], 6, True, i == 0)
                key = (rv, ues.index(u), u["prb"], u["qm"])
                _assert_tb_equal(key, res[k], outs[k], res_r, tb_r, u["payload"])
                _assert_harq_equal(key, acc, range(u["slot0"], u["slot0"] + u["nseg"]), crcs_r, softs_r)
                u["done"] = bool(res_r.tb_crc_ok)
                failed_first += int(i == 0 and not res_r.tb_crc_ok)
        assert failed_first >= 8, f"only {failed_first} of 64 first transmissions failed: retransmissions barely exercised"
    finally:
        acc.close()


# ---------------------------------------------------------------------------------------------------------------------
# BASELINE config 1: BG1, Z = 384, all 46 layers, 6 iterations, >= 148 code blocks per batch.
#   (i)  the benchmark's own input: +-10 from a coin flip, no CRC -> 46 layers x 6 iterations always
#   (ii) the AWGN variant: CRC24B-terminated random codewords, LLR = clamp(round(4 N(mu, 2 mu))), mu in {3, 4, 5, 6, 8},
#        early stop exercised
# ---------------------------------------------------------------------------------------------------------------------
def _ref_or_port_decode(llr, bg, z, F, crc_poly, max_it):
    if HAVE_REF:
        it, out = ob.ref_decode(llr, bg, z, F, crc_poly, max_it)
        return it, out
    it, out, _ = ob.port_decode(llr, bg, z, F, crc_poly, max_it)
    return it, out


@pytest.mark.parametrize("variant", [0, 7])  # 7: inputs staged by cp.async.bulk + mbarrier
def test_config1_full_rate_batch_random_input(acc, variant):
    rng = np.random.default_rng(101)
    dec = pusch.ldpc_decoder_cuda(acc)
    ncb = 160
    llr = (rng.integers(0, 2, (ncb, 25344)) * 20 - 10).astype(np.int8)
    out = np.zeros((ncb, 1056), np.uint8)
    acc.set_decoder_variant(variant)
    try:
        its = dec.decode_batch(out, llr, ncb, pusch.CRC_NONE, 1, 384, 0, 6)
    finally:
        acc.set_decoder_variant(0)
    assert np.all(its == -1)
    step = 1 if HAVE_REF else 10  # the port needs ~0.1 s per full-rate code block
    for i in range(0, ncb, step):
        it, o = _ref_or_port_decode(llr[i], 1, 384, 0, 0, 6)
        assert it < 0
        assert np.array_equal(o, out[i]), i


def test_config1_full_rate_batch_awgn_early_stop(acc):
    rng = np.random.default_rng(102)
    dec = pusch.ldpc_decoder_cuda(acc)
    ncb, bg, z = 150, 1, 384
    K, N = 22 * z, 66 * z
    llr = np.zeros((ncb, N), np.int8)
    mus = [3, 4, 5, 6, 8]
    for i in range(ncb):
        msg = np.zeros(K, np.uint8)
        msg[:K - 24] = rng.integers(0, 2, K - 24, dtype=np.uint8)
        c = ob.port_crc(pusch.CRC24B, np.packbits(msg[:K - 24]), K - 24)
        msg[K - 24:] = [(c >> (23 - b)) & 1 for b in range(24)]
        llr[i] = awgn_llrs(rng, synth.ldpc_encode(msg, bg, z), mus[i % 5] * (0.25 if i % 7 == 0 else 1.0))
    out = np.zeros((ncb, 1056), np.uint8)
    its = dec.decode_batch(out, llr, ncb, pusch.CRC24B, bg, z, 0, 6)
    step = 1 if HAVE_REF else 10
    seen = set()
    for i in range(0, ncb, step):
        it, o = _ref_or_port_decode(llr[i], bg, z, 0, pusch.CRC24B, 6)
        assert (it if it >= 0 else -1) == its[i], (i, it, its[i])
        assert np.array_equal(o, out[i]), i
        seen.add(int(its[i]))
    assert len(seen) >= 3, seen  # several distinct iteration counts incl. failures


# ---------------------------------------------------------------------------------------------------------------------
# Many-layer code blocks on the pair form of the packed decoder (two code blocks per CTA, messages in tensor memory, the
# last layers' messages in shared memory): every lifting size it accepts (Z % 32 == 0, Z >= 160), both base graphs, with
# and without CRC, odd batch sizes (the last code block runs alone on the one-code-block kernel), shortened inputs
# (trailing zeros: fewer layers than the host's bound), an all-zero code block inside a batch, saturated inputs.
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("bg,z,ncb", [(1, 384, 7), (1, 256, 6), (2, 384, 5), (1, 352, 4), (2, 160, 6), (1, 192, 4),
                                       (2, 224, 3), (1, 288, 2), (1, 320, 4)])
def test_many_layer_pairs_unit_batches(acc, bg, z, ncb):
    rng = np.random.default_rng(7000 + z + bg)
    dec = pusch.ldpc_decoder_cuda(acc)
    K, N = ob.kb(bg) * z, ob.ns(bg) * z
    # Variant 6 takes the pair form for every batch (by default it needs at least one code block per SM).
    acc.set_decoder_variant(6)
    try:
        _many_layer_unit_batches(acc, dec, rng, bg, z, ncb, K, N)
    finally:
        acc.set_decoder_variant(0)


def _many_layer_unit_batches(acc, dec, rng, bg, z, ncb, K, N):
    for crc_poly, max_it in ((pusch.CRC24B, 6), (pusch.CRC_NONE, 3), (pusch.CRC16, 2)):
        nb = {pusch.CRC24B: 24, pusch.CRC16: 16, pusch.CRC_NONE: 0}[crc_poly]
        llr = np.zeros((ncb, N), np.int8)
        for i in range(ncb):
            msg = rng.integers(0, 2, K, dtype=np.uint8)
            if nb:
                c = ob.port_crc(crc_poly, np.packbits(msg[:K - nb]), K - nb)
                msg[K - nb:] = [(c >> (nb - 1 - b)) & 1 for b in range(nb)]
            cw = synth.ldpc_encode(msg, bg, z)
            mu = [1.2, 2.0, 3.0, 0.6, 30.0][i % 5] * (1.0 if bg == 1 else 0.6)
            llr[i] = awgn_llrs(rng, cw, mu)
            if i % 5 == 4:
                llr[i] = np.where(cw > 0, -127, 127).astype(np.int8)  # saturated ("infinite") inputs
        if ncb > 2:
            llr[1, N - 5 * z - 3:] = 0  # shortened: trailing zeros, not on a layer boundary
        if ncb > 3:
            llr[2, :] = 0  # all-zero input (ldpc_decoder_impl.cpp:88-94)
        out = np.full((ncb, (K + 7) // 8), 0x5A, np.uint8)
        its = dec.decode_batch(out, llr, ncb, crc_poly, bg, z, 0, max_it)
        for i in range(ncb):
            if HAVE_REF:
                o = np.full((K + 7) // 8, 0x5A, np.uint8)
                it, o = ob.ref_decode(llr[i], bg, z, 0, crc_poly, max_it, o)
            else:
                it, o, _ = ob.port_decode(llr[i], bg, z, 0, crc_poly, max_it, np.full((K + 7) // 8, 0x5A, np.uint8))
            assert (it if it >= 0 else -1) == its[i], (bg, z, crc_poly, i, it, its[i])
            assert np.array_equal(o, out[i]), (bg, z, crc_poly, i)


@pytest.mark.parametrize("variant", [6, 5])
def test_many_layer_tb_low_rate_harq_vs_reference(variant):
    """Low-rate TBs (all 46 / 42 layers) through the TB path: rv0 -> rv2 -> rv3 with soft combining, every combined
    soft-buffer byte and CRC flag compared; variant 6 = the pair form for every batch, 5 = without it (one code block per
    CTA)."""
    # (prb, Qm, R, layers, BG, Nref, mu)
    shapes = [(273, 2, 308, 1, 1, 0, 0.7), (273, 4, 340, 2, 1, 0, 1.1), (273, 2, 193, 2, 2, 0, 0.65)]
    rng = np.random.default_rng(909 + variant)
    chk = _Checker()
    acc = pusch.Accelerator(device=0, max_cbs_in_flight=256, nof_harq_cb_slots=256)
    acc.set_decoder_variant(variant)
    try:
        gdec = pusch.pusch_decoder_cuda(acc)
        failed_first = 0
        for si, (prb, qm, R, nl, bg, nref, mu) in enumerate(shapes):
            tbs = synth.tbs_for(prb, qm, R, nl)
            nllr = prb * 156 * qm * nl
            segs = pusch.segment(tbs, bg, qm, nl, nllr)
            slot0 = 10 + 60 * si
            slots = [slot0 + i for i in range(len(segs))]
            payload = rng.integers(0, 256, tbs // 8, dtype=np.uint8)
            tb_g = np.zeros(tbs // 8, np.uint8)
            for i, rv in enumerate((0, 2, 3)):
                llr = awgn_llrs(rng, synth.encode_tb(payload, bg, rv, qm, nref, nl, nllr), mu)
                tb_r, res_r, crcs_r, softs_r = chk.decode(si, tbs // 8, llr, bg, rv, qm, nref, nl, 6, True, i == 0)
                gdec.new_data(tb_g, slot0, None, pusch.pusch_decoder_configuration(bg, rv, qm, nref, nl, 6, True, i == 0))
                gdec.on_new_softbits(llr)
                res_g = gdec.on_end_softbits()
                key = (variant, si, rv)
                _assert_tb_equal(key, res_g, tb_g, res_r, tb_r, payload)
                _assert_harq_equal(key, acc, slots, crcs_r, softs_r)
                failed_first += int(i == 0 and not res_r.tb_crc_ok)
                if res_r.tb_crc_ok:
                    break
        assert failed_first >= 1, "no first transmission failed: soft combining of many-layer code blocks not exercised"
    finally:
        acc.close()


# ---------------------------------------------------------------------------------------------------------------------
# Bounded slices of the hand-run seed sweeps (tests/seed_sweep_gpu.py, tests/seed_sweep_gpu_tb.py), so that the round-end
# GPU run exercises them: random shapes / rates / SNRs / iteration limits / decoder variants / rv orders.
# ---------------------------------------------------------------------------------------------------------------------
def test_seed_sweep_code_blocks_slice():
    from tests import seed_sweep_gpu

    checked = seed_sweep_gpu.run(first=20000, count=60)
    assert checked > 300


def test_seed_sweep_transport_blocks_slice():
    from tests import seed_sweep_gpu_tb

    done = seed_sweep_gpu_tb.run(first=30000, count=14)
    assert done > 20
