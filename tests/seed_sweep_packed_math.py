#!/usr/bin/env python3
"""Seed sweep of the packed-arithmetic emulation against the oracle (not collected by pytest: run by hand for more
coverage than tests/test_packed_math_cpu.py affords in the regular suite). Usage: seed_sweep_packed_math.py [first] [count]."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle import bindings as ob  # noqa: E402
from tests import test_packed_math_cpu as T  # noqa: E402


def main():
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    count = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    emu = T.load_emu()
    checked = 0
    for seed in range(first, first + count):
        rng = np.random.default_rng(seed)
        lanes_per_thread = int(rng.choice([4, 2]))
        emu.pk_host_set_lanes_per_thread(lanes_per_thread)
        bg = int(rng.integers(1, 3))
        z = int(rng.choice([16, 24, 36, 52, 80, 104, 144, 208]))
        K, N = ob.kb(bg) * z, ob.ns(bg) * z
        crc_poly = int(rng.choice([1, 2, 3]))
        mu = float(rng.choice([1.5, 2, 3, 5, 9, 16, 28]))
        nlen = int(rng.integers(K + 2 * z, N + 1))
        lanes = []
        for _ in range(int(rng.integers(1, 5))):
            llr, F = T.make_lane(rng, bg, z, crc_poly, mu)
            llr = llr[:nlen].copy()
            r = rng.random()
            if r < 0.2:
                idx = rng.choice(nlen, max(1, nlen // 50), replace=False)
                llr[idx] = rng.choice(np.array([-127, 127, -120, 120, 0], np.int8), idx.size)
            elif r < 0.3:
                llr[int(rng.integers(K, nlen)):] = 0
            lanes.append((llr, F))
        max_it = int(rng.integers(1, 13))
        layers = max(T.ref_layers(l[0], bg, z) for l in lanes)
        outs, iters = T.run_group(emu, lanes, bg, z, crc_poly, max_it, 1, layers)
        for c, (llr, F) in enumerate(lanes):
            want = np.full((K + 7) // 8, 0x5A, np.uint8)
            it, want, _ = ob.port_decode(llr, bg, z, F, crc_poly, max_it, want)
            assert iters[c] == it and np.array_equal(outs[c], want), (seed, bg, z, c, F, crc_poly, max_it, layers, mu)
            checked += 1
    print(f"seeds {first}..{first + count - 1}: {checked} code blocks bit-exact")


if __name__ == "__main__":
    main()
