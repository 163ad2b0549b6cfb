#!/usr/bin/env python3
"""Seed sweep of the transport-block path (dematch + HARQ combining + LDPC + CB / TB CRC + TB assembly) on the GPU against
the oracle's pusch_decoder_impl restatement: random allocations, modulations, code rates, layers, limited-buffer sizes,
SNRs and orders of the redundancy versions, several transport blocks per HARQ slot set one after the other (stale soft
bits). Compares TB verdict, statistics, TB bytes, every combined soft buffer byte for byte and every code-block CRC flag.
GPU box; run by hand. Usage: seed_sweep_gpu_tb.py [first] [count]."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle import bindings as ob  # noqa: E402
from srsran_projectvtlmo_b200 import pusch, synth  # noqa: E402
from tests.test_gpu_parity import _tb_sequence  # noqa: E402


def base_graph(tbs, rate):
    """TS 38.212 section 6.2.2 (get_ldpc_base_graph, lib/ran/sch/sch_segmentation.cpp)."""
    if tbs <= 292 or (tbs <= 3824 and rate <= 0.67) or rate <= 0.25:
        return 2
    return 1


def run(first=0, count=40):
    """Seeds [first, first + count): returns the number of code blocks per transmission compared."""
    done = 0
    for seed in range(first, first + count):
        rng = np.random.default_rng(seed)
        # Fresh (zeroed) HARQ slots and a fresh oracle per seed; within a seed the transport blocks share one shape and are
        # decoded one after the other into the SAME slots, so that later ones see the stale soft bits of earlier ones.
        acc = pusch.Accelerator(device=0, max_cbs_in_flight=1024, nof_harq_cb_slots=1024)
        port = ob.PortPusch()
        acc.set_decoder_variant(int(rng.choice([0, 0, 0, 0, 1, 2, 3, 4, 5, 6, 6, 7])))
        qm = int(rng.choice([2, 4, 6, 8]))
        nl = int(rng.choice([1, 1, 2, 4]))
        prb = int(rng.choice([1, 2, 5, 13, 24, 52, 79, 106]))
        R = int(rng.choice([120, 193, 308, 449, 602, 772, 873, 948]))
        tbs = synth.tbs_for(prb, qm, R, nl)
        if tbs < 24 or tbs > 400000:
            acc.close()
            continue
        bg = base_graph(tbs, R / 1024)
        metas = pusch.segment(tbs, bg, qm, nl, prb * 156 * qm * nl)
        N = (66 if bg == 1 else 50) * metas[0].lifting_size
        nref = int(rng.choice([0, 0, 25344, int(N * 0.55), int(N * 0.8)]))
        for _ in range(int(rng.integers(1, 4))):
            # SNR around the waterfall of the rate so that some first transmissions fail and retransmissions combine
            base = {2: 1.0, 4: 3.0, 6: 7.0, 8: 12.0}[qm] * (0.5 + R / 1024)
            mu = base * float(rng.choice([0.45, 0.7, 1.0, 1.6]))
            rvs = tuple(int(v) for v in ([0] + list(rng.permutation([1, 2, 3]))))
            early = bool(rng.random() < 0.8)
            _tb_sequence(acc, port, rng, prb, qm, R, nl, bg, nref, mu, 100, early, int(rng.integers(1, 9)), rvs)
            done += len(metas)
        acc.close()
    return done


def main():
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    count = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    done = run(first, count)
    print(f"seeds {first}..{first + count - 1}: transport blocks with {done} code blocks per transmission bit-exact on the GPU")


if __name__ == "__main__":
    main()
