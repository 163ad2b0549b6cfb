"""Synthetic PUSCH inputs: a numpy 5G NR UL-SCH transmitter (TS 38.212 sections 5.1, 5.2.2, 5.3.2, 5.4.2, 5.5) and an
AWGN LLR channel.

bench.py and the tests need VALID code words (a decoder benchmark fed with noise only never exercises early stop). The
reference synthesises them with its own Tx chain (pdsch_encoder_impl.cpp:28-78); that code is not available on the GPU
box, so this module restates the standard's transmitter from the base-graph tables. It is host-side input synthesis
only (numpy), never on the decode path. tests/test_synth.py pins it against the reference's encoder output.
"""
import re
from functools import lru_cache
from pathlib import Path

import numpy as np

_TABLES = Path(__file__).resolve().parent / "csrc" / "nr_ldpc_bg_tables.h"

ALL_Z = [2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 18, 20, 22, 24, 26, 28, 30, 32, 36, 40, 44, 48, 52, 56, 60,
         64, 72, 80, 88, 96, 104, 112, 120, 128, 144, 160, 176, 192, 208, 224, 240, 256, 288, 320, 352, 384]
CRC_POLY = {"24A": (0x1864CFB, 24), "24B": (0x1800063, 24), "16": (0x11021, 16)}
K0_FACTOR = {1: (0, 17, 33, 56), 2: (0, 13, 25, 43)}


@lru_cache(maxsize=None)
def bg_tables(bg):
    """(row_ptr, col, shift[8][E]) of base graph `bg`, parsed from the generated C table."""
    text = _TABLES.read_text()

    def arr(name):
        m = re.search(name + r"\[[^\]]*\](\[[^\]]*\])? = \{(.*?)\};", text, re.S)
        return np.array([int(t) for t in re.findall(r"\d+", m.group(2))], dtype=np.int64)

    row_ptr = arr(f"NR_BG{bg}_ROW_PTR")
    col = arr(f"NR_BG{bg}_COL")
    shift = arr(f"NR_BG{bg}_SHIFT").reshape(8, col.size)
    return row_ptr, col, shift


def ls_index(z):
    for i, a in enumerate((2, 3, 5, 7, 9, 11, 13, 15)):
        v = a
        while v <= 384:
            if v == z:
                return i
            v *= 2
    raise ValueError(f"invalid lifting size {z}")


def kb(bg):
    return 22 if bg == 1 else 10


def ns(bg):
    return 66 if bg == 1 else 50


@lru_cache(maxsize=None)
def _crc_table(name):
    gen, order = CRC_POLY[name]
    tab = np.zeros(256, np.uint32)
    for b in range(256):
        reg = b << (order - 8)
        for _ in range(8):
            reg <<= 1
            if reg & (1 << order):
                reg ^= gen
        tab[b] = reg & ((1 << order) - 1)
    return tab


def crc_bits(bits, name):
    """CRC of bit arrays of shape (..., n) (one bit per element, MSB first). Returns integer array of shape (...)."""
    gen, order = CRC_POLY[name]
    bits = np.asarray(bits, dtype=np.uint8)
    n = bits.shape[-1]
    lead = bits.shape[:-1]
    pad = (-n) % 8
    if pad:
        bits = np.concatenate([np.zeros(lead + (pad,), np.uint8), bits], axis=-1)  # leading zeros do not change the CRC
    by = np.packbits(bits, axis=-1).astype(np.uint32)
    tab = _crc_table(name)
    reg = np.zeros(lead, np.uint32)
    mask = np.uint32((1 << order) - 1)
    for i in range(by.shape[-1]):
        idx = ((reg >> np.uint32(order - 8)) ^ by[..., i]) & np.uint32(0xFF)
        reg = ((reg << np.uint32(8)) & mask) ^ tab[idx]
    return reg


def int_to_bits(v, n):
    v = np.asarray(v, dtype=np.uint32)
    sh = np.arange(n - 1, -1, -1, dtype=np.uint32)
    return ((v[..., None] >> sh) & 1).astype(np.uint8)


def ldpc_encode(msg, bg, z):
    """Systematic LDPC encoding. msg: (..., K) bits with filler bits zero. Returns (..., N) with N = 66Z / 50Z (the first
    2Z systematic bits are punctured, TS 38.212 5.3.2)."""
    row_ptr, col, shift = bg_tables(bg)
    sh = shift[ls_index(z)] % z
    k_b, n_s = kb(bg), ns(bg)
    msg = np.asarray(msg, dtype=np.uint8)
    lead = msg.shape[:-1]
    nodes = np.zeros(lead + (n_s + 2, z), np.uint8)
    nodes[..., :k_b, :] = msg.reshape(lead + (k_b, z))
    nrows = row_ptr.size - 1

    def rot(x, s):  # (P^s x)[j] = x[(j + s) % z]
        return np.roll(x, -int(s), axis=-1)

    # lambda_i = sum over the systematic columns of the 4 core rows
    lam = []
    for r in range(4):
        acc = np.zeros(lead + (z,), np.uint8)
        for e in range(row_ptr[r], row_ptr[r + 1]):
            if col[e] < k_b:
                acc ^= rot(nodes[..., col[e], :], sh[e])
        lam.append(acc)
    # First parity column: three entries, two with equal shift. Summing the four rows leaves P^y p1 = sum(lambda).
    c0 = k_b
    ent = {}
    for r in range(4):
        for e in range(row_ptr[r], row_ptr[r + 1]):
            if col[e] == c0:
                ent[r] = int(sh[e])
    vals = list(ent.values())
    odd = [v for v in vals if vals.count(v) % 2 == 1]
    assert len(ent) == 3 and len(set(odd)) == 1, "unexpected core parity structure"
    y = odd[0]
    total = lam[0] ^ lam[1] ^ lam[2] ^ lam[3]
    p1 = np.roll(total, y, axis=-1)  # undo P^y
    nodes[..., c0, :] = p1
    # Dual diagonal: row r (0..2) introduces column c0 + 1 + r with shift 0.
    prev = None
    for r in range(3):
        acc = lam[r].copy()
        if r in ent:
            acc ^= rot(p1, ent[r])
        if prev is not None:
            acc ^= prev
        nodes[..., c0 + 1 + r, :] = acc
        prev = acc
    # Extension rows: one new column each, identity.
    for r in range(4, nrows):
        acc = np.zeros(lead + (z,), np.uint8)
        newc = k_b + r
        for e in range(row_ptr[r], row_ptr[r + 1]):
            if col[e] != newc:
                acc ^= rot(nodes[..., col[e], :], sh[e])
        nodes[..., newc, :] = acc
    cw = nodes.reshape(lead + ((n_s + 2) * z,))
    return np.ascontiguousarray(cw[..., 2 * z:])


def check_codeword(cw, bg, z, punctured):
    """True if H c = 0. `punctured`: the 2Z punctured systematic bits."""
    row_ptr, col, shift = bg_tables(bg)
    sh = shift[ls_index(z)] % z
    full = np.concatenate([punctured, cw]).reshape(ns(bg) + 2, z)
    for r in range(row_ptr.size - 1):
        acc = np.zeros(z, np.uint8)
        for e in range(row_ptr[r], row_ptr[r + 1]):
            acc ^= np.roll(full[col[e]], -int(sh[e]))
        if acc.any():
            return False
    return True


def segmentation(tbs, bg):
    """(C, Z, K, K', tb_crc_len, cb_crc_len, zero_pad) per TS 38.212 5.2.2, with the reference's generalisation to payload
    sizes that do not divide evenly (ldpc_segmenter_impl.cpp:118-133: the last segment is zero padded)."""
    tb_crc = 16 if tbs <= 3824 else 24
    b = tbs + tb_crc
    kcb = 8448 if bg == 1 else 3840
    c = 1 if b <= kcb else -(-b // (kcb - 24))
    bp = b + (24 * c if c > 1 else 0)
    ref_len = 22
    if bg == 2:
        ref_len = 10 if b > 640 else 9 if b > 560 else 8 if b > 192 else 6
    z = next(v for v in ALL_Z if v * c * ref_len >= bp)
    k = kb(bg) * z
    cb_crc = 24 if c > 1 else 0
    kp = -(-bp // c)  # info + CB CRC bits per segment
    zero_pad = kp * c - bp
    return c, z, k, kp, tb_crc, cb_crc, zero_pad


def rm_lengths(c, nof_llrs, qm, nof_layers):
    nsl = (nof_llrs // qm) // nof_layers
    nshort = c - (nsl % c)
    return [((nsl // c) if i < nshort else -(-nsl // c)) * nof_layers * qm for i in range(c)]


def rate_match(cw, bg, z, nfill, e_len, rv, qm, nref):
    """Bit selection + interleaving (TS 38.212 5.4.2.1-2) of one code word cw (N bits) -> e_len bits."""
    n = ns(bg) * z
    ncb = min(nref, n) if nref else n
    k0 = int(np.floor(K0_FACTOR[bg][rv] * ncb / n)) * z
    sys_ = (kb(bg) - 2) * z
    valid = np.ones(ncb, bool)
    valid[sys_ - nfill:sys_] = False
    idx = np.nonzero(np.roll(valid, -k0))[0]
    idx = (idx + k0) % ncb
    reps = -(-e_len // idx.size)
    sel = np.tile(idx, reps)[:e_len]
    e = cw[sel]
    if qm > 1:
        e = e.reshape(qm, e_len // qm).T.reshape(-1)  # f[i*Qm + j] = e[j*E/Qm + i]
    return e


def tx_segments(tb_bytes, bg):
    """Tx segmentation (TS 38.212 5.2.2, ldpc_segmenter_tx_impl.cpp): TB bytes -> (msg, z, kp, nfill, tb_crc) with msg the
    (C, K) message bits of the code blocks (TB CRC and code-block CRCs attached, filler bits zero), kp = K - nfill."""
    tb_bits = np.unpackbits(np.asarray(tb_bytes, dtype=np.uint8))
    tbs = tb_bits.size
    c, z, k, kp, tb_crc, cb_crc, zero_pad = segmentation(tbs, bg)
    crc = crc_bits(tb_bits, "16" if tb_crc == 16 else "24A")
    stream = np.concatenate([tb_bits, int_to_bits(crc, tb_crc), np.zeros(zero_pad, np.uint8)])
    info = kp - cb_crc
    segs = stream.reshape(c, info)
    msg = np.zeros((c, k), np.uint8)
    msg[:, :info] = segs
    if c > 1:
        msg[:, info:kp] = int_to_bits(crc_bits(segs, "24B"), 24)
    return msg, z, kp, k - kp, int(crc)


def encode_tb(tb_bytes, bg, rv, qm, nref, nof_layers, nof_llrs):
    """UL-SCH transmitter: TB bytes -> code word bits (one per element, nof_llrs of them)."""
    msg, z, kp, nfill, _ = tx_segments(tb_bytes, bg)
    c = msg.shape[0]
    cws = ldpc_encode(msg, bg, z)
    out = []
    for i, e_len in enumerate(rm_lengths(c, nof_llrs, qm, nof_layers)):
        out.append(rate_match(cws[i], bg, z, nfill, e_len, rv, qm, nref))
    return np.concatenate(out)


def awgn_llrs(rng, bits, mu):
    """LLR = clamp(round(4 * N(+-mu, 2 mu)), +-120): consistent Gaussian LLRs, int8 quantised like the reference's
    demodulator output range (log_likelihood_ratio.h:46-51)."""
    x = (1.0 - 2.0 * bits.astype(np.float32)) * np.float32(mu)
    x = x + rng.standard_normal(bits.size, dtype=np.float32) * np.float32(np.sqrt(2.0 * mu))
    return np.clip(np.round(4.0 * x), -120, 120).astype(np.int8)


def tbs_for(nof_prb, qm, rate_x1024, nof_layers, nof_re_per_prb=156):
    """TS 38.214 5.1.3.2 transport block size (the reference's tbs_calculator.cpp); 156 RE/PRB = 14 symbols with one
    DM-RS symbol carrying two CDM groups, as in pusch_processor_benchmark.cpp:370-376."""
    n_re = min(156, nof_re_per_prb) * nof_prb
    n_info = n_re * (rate_x1024 / 1024.0) * qm * nof_layers
    if n_info <= 3824:
        n = max(3, int(np.floor(np.log2(n_info))) - 6)
        n_info_q = max(24, (1 << n) * int(np.floor(n_info / (1 << n))))
        table = [24, 32, 40, 48, 56, 64, 72, 80, 88, 96, 104, 112, 120, 128, 136, 144, 152, 160, 168, 176, 184, 192,
                 208, 224, 240, 256, 272, 288, 304, 320, 336, 352, 368, 384, 408, 432, 456, 480, 504, 528, 552, 576,
                 608, 640, 672, 704, 736, 768, 808, 848, 888, 928, 984, 1032, 1064, 1128, 1160, 1192, 1224, 1256, 1288,
                 1320, 1352, 1416, 1480, 1544, 1608, 1672, 1736, 1800, 1864, 1928, 2024, 2088, 2152, 2216, 2280, 2408,
                 2472, 2536, 2600, 2664, 2728, 2792, 2856, 2976, 3104, 3240, 3368, 3496, 3624, 3752, 3824]
        return next(t for t in table if t >= n_info_q)
    n = int(np.floor(np.log2(n_info - 24))) - 5
    n_info_q = max(3840, (1 << n) * int(np.round((n_info - 24) / (1 << n))))
    if rate_x1024 / 1024.0 <= 0.25:
        c = -(-(n_info_q + 24) // 3816)
    elif n_info_q > 8424:
        c = -(-(n_info_q + 24) // 8424)
    else:
        c = 1
    return 8 * c * (-(-(n_info_q + 24) // (8 * c))) - 24


# ---- transmitter side of the soft demodulator's input (TS 38.211 5.1, 5.2.1, 6.3.1.1-2): scrambling and QAM mapping --------
def scrambling_sequence(c_init, n):
    """First n bits of the pseudo-random sequence of TS 38.211 5.2.1 (x1 from 1, x2 from c_init, Nc = 1600), numpy uint8.
    Word-level form of the two recurrences (they hold between 32-bit words of the packed sequences as well)."""
    def words(init, taps, nwords):
        st = init & 0x7fffffff
        bits = []
        for _ in range(1600 + 31 * 32):
            bits.append(st & 1)
            f = bin(st & taps).count("1") & 1
            st = (st >> 1) | (f << 30)
        w = np.zeros(max(nwords, 31), np.uint32)
        b = np.array(bits[1600:], np.uint32).reshape(31, 32)
        w[:31] = (b << np.arange(32, dtype=np.uint32)).sum(axis=1, dtype=np.uint64).astype(np.uint32)
        tl = [t for t in range(4) if (taps >> t) & 1]
        for i in range(31, nwords):
            v = np.uint32(0)
            for t in tl:
                v ^= w[i - 31 + t]
            w[i] = v
        return w[:nwords]

    nwords = (n + 31) // 32
    c = words(1, 0x9, nwords) ^ words(c_init, 0xf, nwords)
    return ((c[:, None] >> np.arange(32, dtype=np.uint32)) & 1).astype(np.uint8).reshape(-1)[:n]


def modulate(bits, qm):
    """TS 38.211 5.1.3-5.1.6: QPSK / 16QAM / 64QAM / 256QAM, unit average power, complex64."""
    b = 1.0 - 2.0 * bits.reshape(-1, qm).astype(np.float32)
    if qm == 2:
        re, im, s = b[:, 0], b[:, 1], np.sqrt(2.0)
    elif qm == 4:
        re, im, s = b[:, 0] * (2 - b[:, 2]), b[:, 1] * (2 - b[:, 3]), np.sqrt(10.0)
    elif qm == 6:
        re = b[:, 0] * (4 - b[:, 2] * (2 - b[:, 4]))
        im = b[:, 1] * (4 - b[:, 3] * (2 - b[:, 5]))
        s = np.sqrt(42.0)
    else:
        re = b[:, 0] * (8 - b[:, 2] * (4 - b[:, 4] * (2 - b[:, 6])))
        im = b[:, 1] * (8 - b[:, 3] * (4 - b[:, 5] * (2 - b[:, 7])))
        s = np.sqrt(170.0)
    return ((re + 1j * im) / s).astype(np.complex64)
