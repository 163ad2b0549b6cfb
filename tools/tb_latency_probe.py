"""One config-2 transport block on an idle GPU, page-locked host soft bits in -> TB bytes out: wall-clock latency and the
stage spans with the copy-engine path and with the kernels reading / writing page-locked host memory themselves
(set_direct_io). GPU box."""
import ctypes as C
import sys
import time

import numpy as np
from srsran_projectvtlmo_b200 import capi, pusch, synth

prb, qm, nl, bg, R = 273, 8, 4, 1, 948
tbs = synth.tbs_for(prb, qm, R, nl); nllr = prb * 156 * qm * nl
ncb = len(pusch.segment(tbs, bg, qm, nl, nllr))
rng = np.random.default_rng(3)
tb = rng.integers(0, 256, tbs // 8, dtype=np.uint8)
llr = synth.awgn_llrs(rng, synth.encode_tb(tb, bg, 0, qm, 12611, nl, nllr), 18.0)
lib = capi.lib()
NTB = int(sys.argv[1]) if len(sys.argv) > 1 else 1  # transport blocks per batch (back to back in one page-locked buffer)
REPS = int(sys.argv[2]) if len(sys.argv) > 2 else 520
p = lib.srsran_cuda_pusch_dec_host_alloc(NTB * nllr)
big = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int8)), shape=(NTB * nllr,))
bufs = [big[k * nllr:(k + 1) * nllr] for k in range(NTB)]
for b in bufs:
    b[...] = llr
acc = pusch.Accelerator(device=0, max_cbs_in_flight=4 * NTB * ncb, nof_harq_cb_slots=4 * NTB * ncb)
cfg = [capi.TbConfig(tbs, bg, 0, qm, 12611, nl, 6, 1, 1, k * ncb) for k in range(NTB)]
args = pusch.SubmitArgs(cfg, bufs)
outs = [np.zeros(tbs // 8, np.uint8) for _ in range(NTB)]
print(f"{NTB} TB(s) per batch, {NTB * nllr / 1e6:.2f} MB of soft bits", flush=True)
for rnd in range(2):
    for din, dout in (((False, False), (True, True)) if NTB > 1 else ((False, False), (False, True), (True, False), (True, True))):
        acc.set_direct_io(din, dout)
        lat, st = [], np.zeros(5)
        for rep in range(REPS):
            t0 = time.perf_counter()
            tk = pusch.submit_tbs(acc, args)
            res = pusch.poll_tbs(acc, tk, outs)
            t1 = time.perf_counter()
            assert all(r.tb_crc_ok for r in res) and all(np.array_equal(o, tb) for o in outs)
            if rep >= 20:
                lat.append((t1 - t0) * 1e6); st += np.array(pusch.ticket_timing(acc, tk[0]))
        lat = np.array(lat)
        print(f"direct_in={int(din)} direct_out={int(dout)}  p50 {np.percentile(lat,50):6.1f}  p99 {np.percentile(lat,99):6.1f}  max {lat.max():6.1f} us   stages us {np.round(st/len(lat)*1e3,1)}", flush=True)
