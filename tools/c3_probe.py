"""BASELINE config 3 (64 small TBs of eight shapes per slot): decode span of the whole slot and of each shape alone, to see
which code blocks bound the slot (a never-converging BG1 / Z = 352 block: 6 iterations x 27 layers on one SM). GPU box."""
import sys

import numpy as np, torch
from srsran_projectvtlmo_b200 import capi, pusch, synth
cases = [(52, 2, 120, 1, 2, 0.9), (52, 2, 449, 1, 1, 2.0), (52, 4, 378, 1, 1, 3.0), (52, 4, 658, 1, 1, 6.0),
         (25, 2, 120, 1, 2, 0.9), (10, 4, 490, 1, 2, 4.0), (4, 2, 308, 1, 2, 1.5), (1, 2, 120, 1, 2, 0.9)]
rng = np.random.default_rng(8)
for sel in [None] + list(range(8)):
    acc = pusch.Accelerator(device=0, max_cbs_in_flight=4096, nof_harq_cb_slots=4096)
    acc.set_decoder_variant(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
    cfgs, dev, nllrs, slot = [], [], [], 0
    for ue in range(64):
        ci = ue % 8
        if sel is not None and ci != sel:
            continue
        prb, qm, R, nl, bg, mu = cases[ci]
        tbs = synth.tbs_for(prb, qm, R, nl); nllr = prb * 156 * qm * nl
        tb = rng.integers(0, 256, tbs // 8, dtype=np.uint8)
        l = synth.awgn_llrs(rng, synth.encode_tb(tb, bg, 0, qm, 25344, nl, nllr), mu)
        cfgs.append(capi.TbConfig(tbs, bg, 0, qm, 25344, nl, 6, 1, 1, slot))
        segs = pusch.segment(tbs, bg, qm, nl, nllr); slot += len(segs)
        dev.append(torch.from_numpy(l).cuda()); nllrs.append(nllr)
    st = np.zeros(5)
    for rep in range(4):
        tk = pusch.submit_tbs(acc, cfgs, [(d.data_ptr(), n) for d, n in zip(dev, nllrs)], device_resident=True)
        t = np.array(pusch.ticket_timing(acc, tk[0])); res = pusch.poll_tbs(acc, tk)
        if rep: st += t
    st /= 3
    z = segs[0].lifting_size if sel is not None else 0
    print("shape", sel, cases[sel] if sel is not None else "all", "Z", z, "C", len(segs) if sel is not None else "", "decode us", round(st[2]*1e3,1), "ok", sum(r.tb_crc_ok for r in res), "its", [round(r.iter_mean,1) for r in res][:3], flush=True)
    acc.close()
