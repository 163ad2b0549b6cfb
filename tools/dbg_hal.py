import sys, numpy as np
sys.path.insert(0, '/root/repo')
from oracle import bindings as ob
from srsran_projectvtlmo_b200 import pusch, synth
from tests.test_gpu_parity import _cb_llrs
acc = pusch.Accelerator(device=0, max_cbs_in_flight=4096, nof_harq_cb_slots=8192)
hw = pusch.hw_accelerator_pusch_dec_cuda(acc)
variant = int(sys.argv[1]) if len(sys.argv) > 1 else 1
acc.set_decoder_variant(variant)
rng = np.random.default_rng(40 + variant)
prev = {}
for rnd in range(6):
    early_stop = int(rng.integers(0, 2)); max_it = int(rng.integers(1, 7))
    ops = []; slot = 7000
    for _ in range(int(rng.integers(2, 5))):
        bg = int(rng.integers(1, 3)); z = int(rng.choice([144, 160, 208, 256, 288, 320, 384]))
        K, N = ob.kb(bg) * z, ob.ns(bg) * z
        qm = int(rng.choice([2, 4, 6, 8])); F = int(rng.integers(0, 5)) * 8
        E = (int((K - 2 * z - F) * rng.uniform(1.04, 1.9 if rng.random() < 0.4 else 1.25)) // qm) * qm
        nref = 0 if rng.random() < 0.5 else int(N * 0.6)
        mu = float(rng.choice([3.0, 6.0, 12.0, 20.0]))
        for _ in range(int(rng.integers(1, 8))):
            llr = _cb_llrs(rng, bg, z, F, 2, E, qm, 0, nref, mu)
            if rng.random() < 0.1: llr[:] = 0
            ops.append((bg, z, qm, F, E, nref, llr, slot)); slot += 1
    hw.reserve_queue()
    for i, (bg, z, qm, F, E, nref, llr, s) in enumerate(ops):
        K, N = ob.kb(bg) * z, ob.ns(bg) * z
        hw.configure_operation(pusch.CbConfig(bg, qm, len(ops), 0, E, z, N, nref, K - 24 - F, F, max_it, early_stop, 1, 24, pusch.CB_CRC24B, s), i)
        assert hw.enqueue_operation(llr, None, i)
    for i, (bg, z, qm, F, E, nref, llr, s) in enumerate(ops):
        K, N = ob.kb(bg) * z, ob.ns(bg) * z
        bits = np.full(K // 8, 0x5A, np.uint8); soft = np.zeros(N, np.int8)
        while not hw.dequeue_operation(bits, soft, i): pass
        crc_ok, iters = hw.read_operation_outputs(i, s)
        ws = np.zeros(N, np.int8); wb = np.full(K // 8, 0x5A, np.uint8)
        ws[:] = prev.get(s, np.zeros(25344, np.int8))[:N]
        it = ob.port().oracle_cb_decode(ob._p8(wb), ob._pi(ws), N, ob._pi(llr), E, 1, 0, qm, nref, F, bg, z, 2, early_stop, max_it)
        full = prev.get(s, np.zeros(25344, np.int8)).copy(); full[:N] = ws; prev[s] = full
        nz = np.nonzero(ws)[0]; last = int(nz[-1]) + 1 if nz.size else 0
        L = max(-(-(last + 2*z)//z), ob.kb(bg)+4) - ob.kb(bg)
        ok = np.array_equal(bits, wb) or (early_stop and not llr.any())
        if not ok or not np.array_equal(soft, ws):
            d = np.unpackbits(bits ^ wb)
            print("MISMATCH rnd", rnd, "op", i, "of", len(ops), (bg, z, qm, F, E, nref), "es", early_stop, "max_it", max_it, "L", L, "allzero", not llr.any(), "softok", np.array_equal(soft, ws),
                  "bitdiff", int(d.sum()), "first", np.nonzero(d)[0][:5], "crc", crc_ok, it, "iters", iters)
            dec = pusch.ldpc_decoder_cuda(acc)
            ub = np.full(K // 8, 0x5A, np.uint8)
            uit = dec.decode(ub, ws, 0, bg, z, F, max_it)
            print("   unit-level decode of the same soft buffer equals oracle:", np.array_equal(ub, wb), "nonzero extent of soft", last, "stale region nonzero count", int(np.count_nonzero(ws[6684:14288])))
            # decode again through the HAL with only this op in the batch
            hw2 = pusch.hw_accelerator_pusch_dec_cuda(acc)
    hw.free_queue()
print("done")
