import sys, numpy as np
sys.path.insert(0, '/root/repo')
from oracle import bindings as ob
from srsran_projectvtlmo_b200 import pusch, synth
from tests.test_gpu_parity import _cb_llrs
acc = pusch.Accelerator(device=0, max_cbs_in_flight=64, nof_harq_cb_slots=64)
hw = pusch.hw_accelerator_pusch_dec_cuda(acc)
rng = np.random.default_rng(1)
for (bg, z, nlanes, max_it, es, mu) in [(1, 384, 4, 6, 1, 40.0), (1, 384, 3, 6, 0, 40.0), (2, 144, 2, 6, 1, 30.0), (1, 144, 1, 2, 0, 3.0), (1, 384, 4, 6, 1, 12.0)]:
    K, N = ob.kb(bg)*z, ob.ns(bg)*z
    qm, F, nref = 8, 16, 0
    E = (int((K - 2*z - F)*1.25)//qm)*qm
    llrs = [_cb_llrs(rng, bg, z, F, 2, E, qm, 0, nref, mu) for _ in range(nlanes)]
    hw.reserve_queue()
    for i, llr in enumerate(llrs):
        cfg = pusch.CbConfig(bg, qm, nlanes, 0, E, z, N, nref, K-24-F, F, max_it, es, 1, 24, pusch.CB_CRC24B, i)
        hw.configure_operation(cfg, i); assert hw.enqueue_operation(llr, None, i)
    for i, llr in enumerate(llrs):
        bits = np.zeros(K//8, np.uint8); soft = np.zeros(N, np.int8)
        while not hw.dequeue_operation(bits, soft, i): pass
        crc_ok, iters = hw.read_operation_outputs(i, i)
        ws = np.zeros(N, np.int8); wb = np.zeros(K//8, np.uint8)
        it = ob.port().oracle_cb_decode(ob._p8(wb), ob._pi(ws), N, ob._pi(llr), E, 1, 0, qm, nref, F, bg, z, 2, es, max_it)
        diff = np.unpackbits(bits ^ wb)
        cols = sorted(set((np.nonzero(diff)[0] // z).tolist()))
        print((bg, z, nlanes, max_it, es), "lane", i, "soft_ok", np.array_equal(soft, ws), "bitdiff", int(diff.sum()), "cols", cols[:30], "crc", crc_ok, it >= 0, "iters", iters, it)
    hw.free_queue()
