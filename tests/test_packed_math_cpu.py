"""CPU test of the packed (four code blocks per thread) LDPC arithmetic the CUDA decoder executes
(srsran_projectvtlmo_b200/csrc/ldpc_packed_math.h), compiled for the host and compared with the oracle: decoded bits and
iteration counts, all lifting-size families, fillers, infinite LLRs, shortened inputs, extra (all-zero) layers."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import bindings as ob
from srsran_projectvtlmo_b200 import synth
from tests.helpers import awgn_llrs

HERE = Path(__file__).resolve().parent / "host_emul"
MULT = 52428  # (uint16)(0.8f * 65536), avx512_support.h:69-83


def load_emu():
    subprocess.run(["make", "-s", "-C", str(HERE)], check=True)
    lib = C.CDLL(str(HERE / "libpacked_math_host.so"))
    lib.pk_host_decode_group.restype = C.c_int
    return lib


@pytest.fixture(scope="module")
def emu():
    return load_emu()


def make_lane(rng, bg, z, crc_poly, mu):
    K, N = ob.kb(bg) * z, ob.ns(bg) * z
    crc_len = {1: 24, 2: 24, 3: 16}[crc_poly]
    F = int(rng.integers(0, max(1, min(K // 4, K - crc_len - 8)))) if rng.random() < 0.6 else 0
    npay = K - F - crc_len
    msg = np.zeros(K, np.uint8)
    msg[:npay] = rng.integers(0, 2, npay, dtype=np.uint8)
    c = ob.port_crc(crc_poly, np.packbits(msg[:npay]), npay)
    msg[npay:npay + crc_len] = [(c >> (crc_len - 1 - i)) & 1 for i in range(crc_len)]
    llr = awgn_llrs(rng, synth.ldpc_encode(msg, bg, z), mu)
    llr[K - 2 * z - F:K - 2 * z] = 127
    return llr, F


def run_group(emu, lanes, bg, z, crc_poly, max_it, mode, layers, mult=MULT):
    K = ob.kb(bg) * z
    n = len(lanes)
    outs = [np.full((K + 7) // 8, 0x5A, np.uint8) for _ in range(n)]
    bits_p = (C.c_void_p * 4)(*[o.ctypes.data for o in outs])
    llr_p = (C.c_void_p * 4)(*[np.ascontiguousarray(l[0]).ctypes.data for l in lanes])
    keep = [np.ascontiguousarray(l[0]) for l in lanes]
    llr_p = (C.c_void_p * 4)(*[k.ctypes.data for k in keep])
    n_in = (C.c_uint32 * 4)(*[k.size for k in keep])
    fill = (C.c_uint32 * 4)(*[l[1] for l in lanes])
    iters = (C.c_int * 4)()
    r = emu.pk_host_decode_group(bits_p, llr_p, n_in, n, bg, z, fill, crc_poly, max_it, mode, mult, layers, iters)
    assert r == 0
    return outs, list(iters)[:n]


def ref_layers(llr, bg, z):
    nz = np.nonzero(llr)[0]
    last = int(nz[-1]) + 1 if nz.size else 0
    cbl = max(last + 2 * z, (ob.kb(bg) + 4) * z)
    cbl = -(-cbl // z) * z
    return cbl // z - ob.kb(bg)


@pytest.mark.parametrize("lanes_per_thread", [4, 2])
@pytest.mark.parametrize("bg", [1, 2])
def test_packed_math_matches_oracle(emu, bg, lanes_per_thread):
    emu.pk_host_set_lanes_per_thread(lanes_per_thread)
    rng = np.random.default_rng(100 + bg)
    zs = [2, 3, 5, 7, 9, 11, 13, 15, 16, 24, 36, 52, 80, 104, 144, 208, 288, 384]
    for z in zs:
        K, N = ob.kb(bg) * z, ob.ns(bg) * z
        if K < 64:
            crc_poly = 3 if K >= 40 else None
        else:
            crc_poly = int(rng.choice([1, 2, 3]))
        if crc_poly is None:
            continue
        nl = int(rng.integers(1, 5))
        mu = float(rng.choice([2, 3, 4, 6, 8]))
        nlen = int(rng.integers(K + 2 * z, N + 1)) if rng.random() < 0.8 else N
        if z >= 208:
            nlen = min(nlen, K + 10 * z)  # keep the big ones fast
        lanes = []
        for _ in range(nl):
            llr, F = make_lane(rng, bg, z, crc_poly, mu)
            llr = llr[:nlen].copy()
            if rng.random() < 0.3:
                llr[int(rng.integers(K, nlen)):] = 0  # trailing zeros: data-dependent layer count in the reference
            lanes.append((llr, F))
        if rng.random() < 0.15:
            lanes[0] = (np.zeros(nlen, np.int8), lanes[0][1])  # all-zero lane
        max_it = int(rng.integers(1, 9))
        layers = max(ref_layers(l[0], bg, z) for l in lanes)
        layers = min(layers + int(rng.integers(0, 3)), (46 if bg == 1 else 42))  # extra all-zero layers are no-ops
        for mode in (1, 2):
            outs, iters = run_group(emu, lanes, bg, z, crc_poly, max_it, mode, layers)
            for c, (llr, F) in enumerate(lanes):
                want = np.full((K + 7) // 8, 0x5A, np.uint8)
                if mode == 1:
                    it, want, _ = ob.port_decode(llr, bg, z, F, crc_poly, max_it, want)
                else:
                    _, want, _ = ob.port_decode(llr, bg, z, F, ob.CRC_NONE, max_it, want)
                    it = max_it if ob.port_crc(crc_poly, want, K - F) == 0 else -1
                assert iters[c] == it, (bg, z, mode, c, F, crc_poly, max_it, layers)
                assert np.array_equal(outs[c], want), (bg, z, mode, c, F, crc_poly, max_it, layers)


def test_packed_math_saturating_inputs(emu):
    emu.pk_host_set_lanes_per_thread(4)
    """High-SNR inputs drive many soft values to +-infinity (promotion), low-SNR random +-120 inputs never converge."""
    rng = np.random.default_rng(7)
    for (bg, z) in [(1, 96), (2, 120), (1, 384)]:
        K, N = ob.kb(bg) * z, ob.ns(bg) * z
        nlen = K + 6 * z
        lanes = []
        for mu in (20.0, 28.0):
            llr, F = make_lane(rng, bg, z, 2, mu)
            lanes.append((llr[:nlen].copy(), F))
        lanes.append((rng.choice(np.array([-120, 120, 127, -127, 0, 1, -1], np.int8), nlen), 0))
        lanes.append((rng.integers(-120, 121, nlen, dtype=np.int8), 5))
        for max_it in (2, 6):
            outs, iters = run_group(emu, lanes, bg, z, 2, max_it, 1, max(ref_layers(l[0], bg, z) for l in lanes))
            for c, (llr, F) in enumerate(lanes):
                want = np.full((K + 7) // 8, 0x5A, np.uint8)
                it, want, _ = ob.port_decode(llr, bg, z, F, 2, max_it, want)
                assert iters[c] == it and np.array_equal(outs[c], want), (bg, z, c, max_it)


def test_packed_math_many_updates_of_infinite_values(emu):
    """Full-length code words (all 46 / 42 layers, column degrees up to 30) and many iterations: a promoted soft value is
    updated dozens of times, its binary16 representation runs up to the IEEE infinities and must stay infinite, with no
    NaN on the way; inputs that never converge keep the decoder running to the last iteration."""
    emu.pk_host_set_lanes_per_thread(4)
    rng = np.random.default_rng(9)
    for (bg, z, max_it) in [(1, 48, 40), (2, 80, 25), (1, 144, 12)]:
        K, N = ob.kb(bg) * z, ob.ns(bg) * z
        lanes = []
        llr, F = make_lane(rng, bg, z, 2, 28.0)
        bad = llr.copy()
        flip = rng.choice(np.arange(K - 2 * z, N - 2 * z), 40, replace=False)
        bad[flip] = -bad[flip]  # strong wrong parity bits: promoted values of both signs, no convergence for a while
        lanes.append((bad, F))
        lanes.append((rng.choice(np.array([-127, 127, -120, 120, 119, -119], np.int8), N - 2 * z), 0))
        sat = np.where(rng.random(N - 2 * z) < 0.5, 120, -120).astype(np.int8)
        lanes.append((sat, 0))
        lanes.append((llr, F))
        layers = 46 if bg == 1 else 42
        outs, iters = run_group(emu, lanes, bg, z, 2, max_it, 1, layers)
        for c, (l, Fc) in enumerate(lanes):
            want = np.full((K + 7) // 8, 0x5A, np.uint8)
            it, want, _ = ob.port_decode(l, bg, z, Fc, 2, max_it, want)
            assert iters[c] == it and np.array_equal(outs[c], want), (bg, z, c, max_it, it, iters[c])


def test_packed_math_no_scaling(emu):
    rng = np.random.default_rng(8)
    bg, z = 2, 64
    K = ob.kb(bg) * z
    nlen = K + 8 * z
    lanes = []
    for _ in range(4):
        llr, F = make_lane(rng, bg, z, 3, 3.0)
        lanes.append((llr[:nlen].copy(), F))
    outs, iters = run_group(emu, lanes, bg, z, 3, 5, 1, max(ref_layers(l[0], bg, z) for l in lanes), mult=0)
    for c, (llr, F) in enumerate(lanes):
        want = np.full((K + 7) // 8, 0x5A, np.uint8)
        it, want, _ = ob.port_decode(llr, bg, z, F, 3, 5, want, scaling=1.0)
        assert iters[c] == it and np.array_equal(outs[c], want), c


@pytest.mark.parametrize("bg", [1, 2])
def test_intra_codeblock_packing_matches_oracle(emu, bg):
    """Four lifted checks of ONE code block per thread (lane rotation by quarter turns): every Z divisible by 4."""
    rng = np.random.default_rng(300 + bg)
    for z in [4, 8, 12, 16, 20, 24, 28, 32, 36, 40, 44, 48, 52, 56, 60, 64, 72, 80, 88, 96, 104, 112, 120, 128, 144,
              160, 176, 192, 208, 224, 240, 256, 288, 320, 352, 384]:
        K, N = ob.kb(bg) * z, ob.ns(bg) * z
        crc_poly = int(rng.choice([1, 2, 3])) if K >= 64 else 3
        mu = float(rng.choice([2, 3, 4, 6, 8, 20]))
        nlen = int(rng.integers(K + 2 * z, N + 1)) if rng.random() < 0.8 else N
        if z >= 208:
            nlen = min(nlen, K + 12 * z)
        llr, F = make_lane(rng, bg, z, crc_poly, mu)
        llr = llr[:nlen].copy()
        if rng.random() < 0.3:
            llr[int(rng.integers(K, nlen)):] = 0
        max_it = int(rng.integers(1, 9))
        layers = min(ref_layers(llr, bg, z) + int(rng.integers(0, 2)), 46 if bg == 1 else 42)
        for mode in (1, 2):
            out = np.full((K + 7) // 8, 0x5A, np.uint8)
            it = C.c_int(0)
            keep = np.ascontiguousarray(llr)
            r = emu.pk_host_decode_q4(out.ctypes.data_as(C.c_void_p), keep.ctypes.data_as(C.c_void_p), keep.size, bg, z, F,
                                      crc_poly, max_it, mode, MULT, layers, C.byref(it))
            assert r == 0
            want = np.full((K + 7) // 8, 0x5A, np.uint8)
            if mode == 1:
                wit, want, _ = ob.port_decode(llr, bg, z, F, crc_poly, max_it, want)
            else:
                _, want, _ = ob.port_decode(llr, bg, z, F, ob.CRC_NONE, max_it, want)
                wit = max_it if ob.port_crc(crc_poly, want, K - F) == 0 else -1
            assert it.value == wit, (bg, z, mode, F, crc_poly, max_it, layers)
            assert np.array_equal(out, want), (bg, z, mode, F, crc_poly, max_it, layers)
