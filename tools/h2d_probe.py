"""Host -> device stage of one batch of 64 config-2 TBs, one batch in flight: the soft bits in one page-locked buffer
(one copy), in 64 page-locked buffers (64 copy-engine jobs) and the same read by the gather kernel. GPU box."""
import ctypes as C
import sys

import numpy as np
from srsran_projectvtlmo_b200 import capi, pusch, synth

B, prb, qm, nl, bg, R = 64, 273, 8, 4, 1, 948
tbs = synth.tbs_for(prb, qm, R, nl); nllr = prb * 156 * qm * nl
ncb = len(pusch.segment(tbs, bg, qm, nl, nllr))
rng = np.random.default_rng(3)
tb = rng.integers(0, 256, tbs // 8, dtype=np.uint8)
llr = synth.awgn_llrs(rng, synth.encode_tb(tb, bg, 0, qm, 12611, nl, nllr), 18.0)
lib = capi.lib()

def pinned(n):
    p = lib.srsran_cuda_pusch_dec_host_alloc(n)
    assert p
    return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int8)), shape=(n,))

big = pinned(B * nllr)
contiguous = [big[k * nllr:(k + 1) * nllr] for k in range(B)]
separate = [pinned(nllr) for _ in range(B)]
for b in contiguous + separate:
    b[...] = llr
acc = pusch.Accelerator(device=0, max_cbs_in_flight=B * ncb, nof_harq_cb_slots=B * ncb)
cfgs = [capi.TbConfig(tbs, bg, 0, qm, 12611, nl, 6, 1, 1, i * ncb) for i in range(B)]
for name, bufs, g in (("contiguous", contiguous, 0), ("separate, copy engine", separate, 0), ("separate, gather 16", separate, 16),
                      ("separate, gather 32", separate, 32), ("separate, gather 64", separate, 64), ("separate, gather 148", separate, 148)):
    acc.set_h2d_gather(g)
    args = pusch.SubmitArgs(cfgs, bufs)
    st, n0 = np.zeros(5), 0
    for rep in range(8):
        l0 = acc.launch_count
        tk = pusch.submit_tbs(acc, args)
        t = np.array(pusch.ticket_timing(acc, tk[0])); res = pusch.poll_tbs(acc, tk)
        assert all(r.tb_crc_ok for r in res)
        if rep >= 2:
            st += t; n0 = acc.launch_count - l0
    st /= 6
    print(f"{name:24s} h2d {st[0]*1e3:7.1f} us = {B*nllr/st[0]/1e6:5.1f} GB/s  stages {np.round(st*1e3,1)} launches/batch {n0}", flush=True)
