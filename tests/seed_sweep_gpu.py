#!/usr/bin/env python3
"""Seed sweep of the CUDA decoders against the oracle through the hal call sequence (GPU box; run by hand for more coverage
than the regular -m gpu suite): random shapes, code rates, SNRs, iteration limits, decoder variants, with saturated and
non-finite LLRs sprinkled in. Usage: seed_sweep_gpu.py [first] [count]."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle import bindings as ob  # noqa: E402
from srsran_projectvtlmo_b200 import pusch  # noqa: E402
from tests.test_gpu_parity import _cb_llrs  # noqa: E402


def run(first=0, count=100):
    """Seeds [first, first + count): returns the number of code blocks compared (bounded slices run as -m gpu tests)."""
    PREV.clear()
    acc = pusch.Accelerator(device=0, max_cbs_in_flight=4096, nof_harq_cb_slots=8192)
    hw = pusch.hw_accelerator_pusch_dec_cuda(acc)
    checked = 0
    for seed in range(first, first + count):
        rng = np.random.default_rng(seed)
        acc.set_decoder_variant(int(rng.choice([0, 0, 0, 1, 2, 3, 4, 5, 6, 6, 7])))
        early_stop = int(rng.random() < 0.8)
        max_it = int(rng.integers(1, 13))
        ops = []
        for _ in range(int(rng.integers(1, 4))):
            bg = int(rng.integers(1, 3))
            z = int(rng.choice([16, 36, 64, 104, 144, 160, 208, 256, 288, 320, 384]))
            K, N = ob.kb(bg) * z, ob.ns(bg) * z
            qm = int(rng.choice([1, 2, 4, 6, 8]))
            F = int(rng.integers(0, 5)) * 8
            E = (int((K - 2 * z - F) * rng.uniform(1.04, 2.6)) // qm) * qm
            nref = 0 if rng.random() < 0.5 else int(N * rng.uniform(0.5, 0.9))
            mu = float(rng.choice([1.5, 3.0, 6.0, 12.0, 20.0, 28.0]))
            for _ in range(int(rng.integers(1, 7))):
                llr = _cb_llrs(rng, bg, z, F, pusch.CRC24B, E, qm, 0, nref, mu)
                r = rng.random()
                if r < 0.15:
                    idx = rng.choice(E, max(1, E // 40), replace=False)
                    llr[idx] = rng.choice(np.array([-127, 127, -120, 120, 0], np.int8), idx.size)
                elif r < 0.2:
                    llr[:] = 0
                ops.append((bg, z, qm, F, E, nref, llr))
        hw.reserve_queue()
        for i, (bg, z, qm, F, E, nref, llr) in enumerate(ops):
            K, N = ob.kb(bg) * z, ob.ns(bg) * z
            cfg = pusch.CbConfig(bg, qm, len(ops), 0, E, z, N, nref, K - 24 - F, F, max_it, early_stop, 1, 24,
                                 pusch.CB_CRC24B, 100 + i)
            hw.configure_operation(cfg, i)
            assert hw.enqueue_operation(llr, None, i)
        for i, (bg, z, qm, F, E, nref, llr) in enumerate(ops):
            K, N = ob.kb(bg) * z, ob.ns(bg) * z
            bits = np.full(K // 8, 0x5A, np.uint8)
            soft = np.zeros(N, np.int8)
            while not hw.dequeue_operation(bits, soft, i):
                pass
            crc_ok, iters = hw.read_operation_outputs(i, 100 + i)
            want_soft = np.zeros(N, np.int8)
            want_bits = np.full(K // 8, 0x5A, np.uint8)
            prev = PREV.get(100 + i, np.zeros(25344, np.int8))
            want_soft[:] = prev[:N]
            it = ob.port().oracle_cb_decode(ob._p8(want_bits), ob._pi(want_soft), N, ob._pi(llr), E, 1, 0, qm, nref, F, bg,
                                            z, pusch.CRC24B, early_stop, max_it)
            full = prev.copy()
            full[:N] = want_soft
            PREV[100 + i] = full
            key = (seed, i, bg, z, qm, F, E, nref, early_stop, max_it)
            assert np.array_equal(soft, want_soft), key
            assert crc_ok == (it >= 0), key
            assert iters == (it if it >= 0 else max_it), key
            if not (early_stop and not llr.any()):
                assert np.array_equal(bits, want_bits), key
            checked += 1
        hw.free_queue()
    acc.close()
    return checked


def main():
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    count = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    checked = run(first, count)
    print(f"seeds {first}..{first + count - 1}: {checked} code blocks bit-exact on the GPU")


PREV = {}

if __name__ == "__main__":
    main()
