"""Shared input synthesis for the tests (CPU side, uses the oracle as the checker only)."""
import numpy as np

from oracle import bindings as ob

ALL_Z = [2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 18, 20, 22, 24, 26, 28, 30, 32, 36, 40, 44, 48, 52, 56, 60,
         64, 72, 80, 88, 96, 104, 112, 120, 128, 144, 160, 176, 192, 208, 224, 240, 256, 288, 320, 352, 384]


def awgn_llrs(rng, bits, mu):
    """LLR = clamp(round(4 * N(+-mu, 2 mu)), +-120) (the AWGN variant of config 1)."""
    x = (1.0 - 2.0 * bits.astype(np.float64)) * mu + rng.normal(0.0, np.sqrt(2.0 * mu), bits.size)
    return np.clip(np.round(4.0 * x), -120, 120).astype(np.int8)


def random_cb_case(rng, bg, z, with_crc=True):
    """A decoder input of N LLRs for one code block: random message, optional CRC, filler bits, AWGN, shortened tail."""
    K, N = ob.kb(bg) * z, ob.ns(bg) * z
    crc_poly = int(rng.choice([0, 1, 2, 3])) if with_crc else 0
    if K < 64 and crc_poly in (1, 2):
        crc_poly = 3
    if K < 40:
        crc_poly = 0
    crc_len = {0: 0, 1: 24, 2: 24, 3: 16}[crc_poly]
    F = int(rng.integers(0, max(1, min(K // 4, K - crc_len - 8)))) if rng.random() < 0.6 else 0
    # crc_calculator_clmul_impl mishandles 13..15 whole bytes (SURVEY.md Appendix A trap 8): stay clear of it.
    if crc_poly and (K - F) // 8 in (13, 14, 15):
        F = 0
    if crc_poly and (K - F) // 8 in (13, 14, 15):
        crc_poly, crc_len = 0, 0
    npay = K - F - crc_len
    msg = np.zeros(K, np.uint8)
    msg[:npay] = rng.integers(0, 2, npay, dtype=np.uint8)
    if crc_len:
        c = ob.port_crc(crc_poly, np.packbits(msg[:npay]), npay)
        msg[npay:npay + crc_len] = [(c >> (crc_len - 1 - i)) & 1 for i in range(crc_len)]
    return msg, F, crc_poly


def numpy_ldpc_encode(msg, bg, z):
    """Systematic 5G NR LDPC encoding from the base-graph tables (via the product's synthetic-input module)."""
    from srsran_projectvtlmo_b200 import synth

    return synth.ldpc_encode(msg, bg, z)
