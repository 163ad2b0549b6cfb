"""TEST INFRASTRUCTURE - NOT PART OF THE PRODUCT PATH.

ctypes bindings for the two CPU checkers:
  * ``port()``  -> oracle/liboracle_port.so, the plain-C restatement (oracle_port.c);
  * ``ref()``   -> oracle/_ref/libref_oracle.so, the UNMODIFIED reference compiled by oracle/Makefile (None if absent).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REFERENCE_ROOT = Path("/root/reference")

CRC_NONE, CRC24A, CRC24B, CRC16 = 0, 1, 2, 3
MAX_CB = 162

u8p = C.POINTER(C.c_uint8)
i8p = C.POINTER(C.c_int8)


class CbMeta(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in
                ("bg", "Z", "full_length", "rm_length", "nof_filler_bits", "cw_offset", "nof_crc_bits")]


class TbResult(C.Structure):
    _fields_ = [("tb_crc_ok", C.c_int32), ("nof_codeblocks", C.c_uint32), ("nof_observations", C.c_uint32),
                ("iter_min", C.c_uint32), ("iter_max", C.c_uint32), ("iter_mean", C.c_float)]


def _p8(a):
    assert a.dtype == np.uint8 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(u8p)


def _pi(a):
    assert a.dtype == np.int8 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(i8p)


def build(port_only=False):
    """Compiles the checkers. Building the checker is not using it."""
    subprocess.run(["make", "-s", "-C", str(HERE), "port"], check=True)
    if not port_only and REFERENCE_ROOT.is_dir():
        subprocess.run(["make", "-s", "-j8", "-C", str(HERE), "ref"], check=True)


_port = None
_ref = None


def port():
    global _port
    if _port is None:
        so = HERE / "liboracle_port.so"
        if not so.exists():
            build(port_only=True)
        lib = C.CDLL(str(so))
        lib.oracle_crc.restype = C.c_uint32
        lib.oracle_crc.argtypes = [C.c_int, u8p, C.c_uint32]
        lib.oracle_llr_add.restype = C.c_int8
        lib.oracle_llr_add.argtypes = [C.c_int8, C.c_int8]
        lib.oracle_llr_promotion_sum.restype = C.c_int8
        lib.oracle_llr_promotion_sum.argtypes = [C.c_int8, C.c_int8]
        lib.oracle_hard_decision.restype = C.c_int
        lib.oracle_hard_decision.argtypes = [u8p, i8p, C.c_uint32]
        lib.oracle_dematch.restype = C.c_int
        lib.oracle_dematch.argtypes = [i8p, C.c_uint32, i8p, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_uint32,
                                       C.c_uint32]
        lib.oracle_ldpc_decode.restype = C.c_int
        lib.oracle_ldpc_decode.argtypes = [u8p, i8p, C.c_uint32, C.c_int, C.c_int, C.c_uint32, C.c_int, C.c_int,
                                           C.c_float, C.POINTER(C.c_uint32)]
        lib.oracle_cb_decode.restype = C.c_int
        lib.oracle_cb_decode.argtypes = [u8p, i8p, C.c_uint32, i8p, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_uint32,
                                         C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        lib.oracle_segment_rx.restype = C.c_int
        lib.oracle_segment_rx.argtypes = [C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_uint32, C.POINTER(CbMeta)]
        lib.oracle_harq_create.restype = C.c_void_p
        lib.oracle_harq_create.argtypes = [C.c_uint32]
        lib.oracle_harq_destroy.argtypes = [C.c_void_p]
        lib.oracle_pusch_decode.restype = C.c_int
        lib.oracle_pusch_decode.argtypes = [C.c_void_p, u8p, C.c_uint32, i8p, C.c_uint32, C.c_int, C.c_int, C.c_int,
                                            C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(TbResult)]
        lib.oracle_pusch_bench.restype = C.c_double
        lib.oracle_pusch_bench.argtypes = [C.c_uint32, i8p, C.c_uint32, C.c_int, C.c_int, C.c_uint32, C.c_int, C.c_int,
                                           C.c_int, C.c_int, C.POINTER(C.c_int)]
        f32p, u32p = C.POINTER(C.c_float), C.POINTER(C.c_uint32)
        lib.oracle_scrambling_sequence.argtypes = [C.c_uint32, u8p, C.c_uint32]
        lib.oracle_demodulate_soft.argtypes = [C.c_int, C.c_int, i8p, f32p, f32p, C.c_uint32]
        lib.oracle_pusch_demodulate.restype = C.c_int
        lib.oracle_pusch_demodulate.argtypes = [C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, u32p, C.c_uint32,
                                                f32p, f32p, i8p]
        _port = lib
    return _port


def ref():
    """The compiled reference, or None when oracle/_ref/libref_oracle.so is absent (and cannot be built)."""
    global _ref
    if _ref is None:
        so = HERE / "_ref" / "libref_oracle.so"
        if not so.exists():
            if not REFERENCE_ROOT.is_dir():
                return None
            build()
        try:
            lib = C.CDLL(str(so))
        except OSError:
            return None
        lib.ref_info.restype = C.c_char_p
        lib.ref_dematch.restype = C.c_int
        lib.ref_dematch.argtypes = [C.c_char_p, i8p, C.c_uint32, i8p, C.c_uint32, C.c_int, C.c_int, C.c_int,
                                    C.c_uint32, C.c_uint32]
        lib.ref_ldpc_decode.restype = C.c_int
        lib.ref_ldpc_decode.argtypes = [C.c_char_p, u8p, i8p, C.c_uint32, C.c_int, C.c_int, C.c_uint32, C.c_int,
                                        C.c_int, C.c_float]
        for name in ("ref_crc", "ref_crc_byte", "ref_crc_bit"):
            f = getattr(lib, name)
            f.restype = C.c_uint32
            f.argtypes = [C.c_char_p, C.c_int, u8p, C.c_uint32]
        lib.ref_hard_decision.restype = C.c_int
        lib.ref_hard_decision.argtypes = [u8p, i8p, C.c_uint32]
        lib.ref_segment_rx.restype = C.c_int
        lib.ref_segment_rx.argtypes = [C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_int, C.c_uint32,
                                       C.POINTER(CbMeta)]
        lib.ref_encode_tb.restype = C.c_int
        lib.ref_encode_tb.argtypes = [u8p, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_int, C.c_uint32, u8p]
        lib.ref_ldpc_encode.restype = C.c_int
        lib.ref_ldpc_encode.argtypes = [u8p, C.c_int, C.c_int, u8p]
        lib.ref_pusch_create.restype = C.c_void_p
        lib.ref_pusch_create.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]
        lib.ref_pusch_destroy.argtypes = [C.c_void_p]
        lib.ref_pusch_decode.restype = C.c_int
        lib.ref_pusch_decode.argtypes = [C.c_void_p, C.c_uint64, u8p, C.c_uint32, i8p, C.c_uint32, C.c_int, C.c_int,
                                         C.c_int, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(TbResult),
                                         u8p, i8p]
        lib.ref_pusch_bench.restype = C.c_double
        lib.ref_pusch_bench.argtypes = [C.c_void_p, C.c_uint32, i8p, C.c_uint32, C.c_int, C.c_int, C.c_uint32, C.c_int,
                                        C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
        lib.ref_pusch_bench_mt.restype = C.c_double
        lib.ref_pusch_bench_mt.argtypes = [C.c_char_p, C.c_int, C.c_uint32, i8p, C.c_uint32, C.c_int, C.c_int,
                                           C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
        lib.ref_ldpc_decode_bench.restype = C.c_double
        lib.ref_ldpc_decode_bench.argtypes = [C.c_char_p, i8p, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        f32p = C.POINTER(C.c_float)
        if hasattr(lib, "ref_demodulate_soft"):
            lib.ref_demodulate_soft.argtypes = [C.c_int, C.c_int, i8p, f32p, f32p, C.c_uint32]
            lib.ref_scrambling_sequence.argtypes = [C.c_uint32, u8p, C.c_uint32]
            lib.ref_pusch_demodulate.restype = C.c_int
            lib.ref_pusch_demodulate.argtypes = [C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                                 C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, f32p, f32p, C.c_uint32,
                                                 i8p, C.c_uint32]
        _ref = lib
    return _ref


def ref_flavour():
    """The reference flavour whose arithmetic the port restates and the host can run: "avx512", "avx2" or None."""
    lib = ref()
    if lib is None:
        return None
    info = lib.ref_info().decode()
    if "avx512f=1" in info and "avx512bw=1" in info and "avx512vbmi=1" in info:
        return "avx512"
    if "avx2=1" in info:
        return "avx2"
    return None


# ---- numpy-level helpers (same call shapes for port and reference) -------------------------------------------------

def kb(bg):
    return 22 if bg == 1 else 10


def ns(bg):
    return 66 if bg == 1 else 50


def port_crc(poly, packed, nbits):
    return port().oracle_crc(poly, _p8(packed), nbits)


def port_dematch(softbuf, llrs, new_data, rv, qm, nref, nfill):
    r = port().oracle_dematch(_pi(softbuf), softbuf.size, _pi(llrs), llrs.size, int(new_data), rv, qm, nref, nfill)
    assert r == 0
    return softbuf


def ref_dematch(softbuf, llrs, new_data, rv, qm, nref, nfill, kind="auto"):
    r = ref().ref_dematch(kind.encode(), _pi(softbuf), softbuf.size, _pi(llrs), llrs.size, int(new_data), rv, qm, nref,
                          nfill)
    assert r == 0
    return softbuf


def port_decode(llrs, bg, z, nfill, crc_poly, max_it, out=None, scaling=0.8):
    k = kb(bg) * z
    if out is None:
        out = np.zeros((k + 7) // 8, np.uint8)
    layers = C.c_uint32(0)
    it = port().oracle_ldpc_decode(_p8(out), _pi(llrs), llrs.size, bg, z, nfill, crc_poly, max_it, scaling,
                                   C.byref(layers))
    return it, out, layers.value


def ref_decode(llrs, bg, z, nfill, crc_poly, max_it, out=None, kind="auto", scaling=0.8):
    k = kb(bg) * z
    if out is None:
        out = np.zeros((k + 7) // 8, np.uint8)
    it = ref().ref_ldpc_decode(kind.encode(), _p8(out), _pi(llrs), llrs.size, bg, z, nfill, crc_poly, max_it, scaling)
    return it, out


def port_segment(tbs_bits, bg, qm, nof_layers, nof_llrs):
    metas = (CbMeta * MAX_CB)()
    c = port().oracle_segment_rx(tbs_bits, bg, qm, nof_layers, nof_llrs, metas)
    return [metas[i] for i in range(max(c, 0))]


def ref_segment(tbs_bits, bg, qm, nof_layers, nof_llrs, rv=0, nref=0):
    metas = (CbMeta * MAX_CB)()
    c = ref().ref_segment_rx(tbs_bits, bg, rv, qm, nref, nof_layers, nof_llrs, metas)
    return [metas[i] for i in range(c)]


def ref_encode_tb(tb_bytes, bg, rv, qm, nref, nof_layers, nof_ch_symbols):
    cw = np.zeros(nof_ch_symbols * qm, np.uint8)
    ref().ref_encode_tb(_p8(tb_bytes), tb_bytes.size, bg, rv, qm, nref, nof_layers, nof_ch_symbols, _p8(cw))
    return cw


def ref_ldpc_encode(msg_bits, bg, z):
    cw = np.zeros(ns(bg) * z + 2 * z, np.uint8)[: ns(bg) * z]
    cw = np.ascontiguousarray(cw)
    ref().ref_ldpc_encode(_p8(msg_bits), bg, z, _p8(cw))
    return cw


class PortPusch:
    """pusch_decoder_impl restatement with persistent HARQ buffers keyed by an integer."""

    def __init__(self):
        self.harq = {}

    def decode(self, key, tb_bytes, llrs, bg, rv, qm, nref, nof_layers, max_it, early_stop, new_data):
        lib = port()
        metas = port_segment(tb_bytes * 8, bg, qm, nof_layers, llrs.size)
        if key not in self.harq or self.harq[key][1] != len(metas):
            if key in self.harq:
                lib.oracle_harq_destroy(self.harq[key][0])
            self.harq[key] = (lib.oracle_harq_create(len(metas)), len(metas))
        tb = np.zeros(tb_bytes, np.uint8)
        res = TbResult()
        r = lib.oracle_pusch_decode(self.harq[key][0], _p8(tb), tb_bytes, _pi(llrs), llrs.size, bg, rv, qm, nref,
                                    nof_layers, max_it, int(early_stop), int(new_data), C.byref(res))
        assert r == 0
        return tb, res

    def harq_state(self, key, metas):
        """(CRC flags, soft buffers) per code block of the HARQ buffer `key`: the first full_length soft bits each."""

        class H(C.Structure):
            _fields_ = [("nof_cbs", C.c_uint32), ("crc", C.c_uint8 * 162), ("soft", C.POINTER(C.c_int8)),
                        ("data", C.POINTER(C.c_uint8))]

        hb = C.cast(self.harq[key][0], C.POINTER(H)).contents
        all_soft = np.ctypeslib.as_array(hb.soft, shape=(len(metas) * 25344,))
        return ([bool(hb.crc[cb]) for cb in range(len(metas))],
                [all_soft[cb * 25344:cb * 25344 + m.full_length].copy() for cb, m in enumerate(metas)])

    def __del__(self):
        for h, _ in self.harq.values():
            port().oracle_harq_destroy(h)


class RefPusch:
    """The reference's pusch_decoder_impl over harness-owned HARQ buffers."""

    def __init__(self, dec="auto", dem="auto", crc="auto", nof_threads=1):
        self.h = ref().ref_pusch_create(dec.encode(), dem.encode(), crc.encode(), nof_threads)
        assert self.h

    def decode(self, key, tb_bytes, llrs, bg, rv, qm, nref, nof_layers, max_it, early_stop, new_data,
               want_soft=False):
        metas = ref_segment(tb_bytes * 8, bg, qm, nof_layers, llrs.size)
        tb = np.zeros(tb_bytes, np.uint8)
        res = TbResult()
        crcs = np.zeros(len(metas), np.uint8)
        soft = np.zeros(sum(m.full_length for m in metas), np.int8) if want_soft else None
        ref().ref_pusch_decode(self.h, key, _p8(tb), tb_bytes, _pi(llrs), llrs.size, bg, rv, qm, nref, nof_layers,
                               max_it, int(early_stop), int(new_data), C.byref(res), _p8(crcs),
                               _pi(soft) if want_soft else None)
        return tb, res, crcs, soft

    def __del__(self):
        if getattr(self, "h", None):
            ref().ref_pusch_destroy(self.h)
            self.h = None


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ---- PUSCH soft-demodulation chain ---------------------------------------------------------------

def _pf(a):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _demod_args(symbols, noise_vars):
    sym = np.ascontiguousarray(symbols.astype(np.complex64)).view(np.float32)
    nv = np.ascontiguousarray(noise_vars, dtype=np.float32)
    assert sym.size == 2 * nv.size
    return sym, nv


def port_demodulate_soft(symbols, noise_vars, qm, pi2=False):
    """demodulation_mapper::demodulate_soft on one block (complex64 symbols, float32 noise variances) -> int8 LLRs."""
    sym, nv = _demod_args(symbols, noise_vars)
    out = np.zeros(nv.size * qm, np.int8)
    port().oracle_demodulate_soft(qm, int(pi2), _pi(out), _pf(sym), _pf(nv), nv.size)
    return out


def ref_demodulate_soft(symbols, noise_vars, qm, pi2=False):
    sym, nv = _demod_args(symbols, noise_vars)
    out = np.zeros(nv.size * qm, np.int8)
    ref().ref_demodulate_soft(qm, int(pi2), _pi(out), _pf(sym), _pf(nv), nv.size)
    return out


def port_scrambling_sequence(c_init, n):
    out = np.zeros(n, np.uint8)
    port().oracle_scrambling_sequence(c_init, _p8(out), n)
    return out


def ref_scrambling_sequence(c_init, n):
    out = np.zeros(n, np.uint8)
    ref().ref_scrambling_sequence(c_init, _p8(out), n)
    return out


def pusch_re_per_symbol(nof_prb, start_symbol, nof_symbols, dmrs_mask, nof_cdm_groups_without_data):
    """Data REs per layer in every OFDM symbol of a DM-RS type 1 allocation (pusch_demodulator_impl.cpp:142-150)."""
    out = []
    for s in range(start_symbol, start_symbol + nof_symbols):
        if (dmrs_mask >> s) & 1:
            out.append(nof_prb * (12 - 6 * nof_cdm_groups_without_data))
        else:
            out.append(nof_prb * 12)
    return np.array(out, np.uint32)


def port_pusch_demodulate(symbols, noise_vars, qm, rnti, n_id, nof_layers, re_per_symbol, pi2=False):
    sym, nv = _demod_args(symbols, noise_vars)
    re_per_symbol = np.ascontiguousarray(re_per_symbol, dtype=np.uint32)
    assert int(re_per_symbol.sum()) * nof_layers == nv.size
    out = np.zeros(nv.size * qm, np.int8)
    n = port().oracle_pusch_demodulate(qm, int(pi2), rnti, n_id, nof_layers,
                                       re_per_symbol.ctypes.data_as(C.POINTER(C.c_uint32)), re_per_symbol.size, _pf(sym),
                                       _pf(nv), _pi(out))
    assert n == out.size
    return out


def ref_pusch_demodulate(symbols, noise_vars, qm, rnti, n_id, nof_layers, nof_prb, start_symbol, nof_symbols, dmrs_mask,
                         nof_cdm_groups_without_data, pi2=False):
    """The reference's pusch_demodulator_impl (stub equalizer) + ulsch_demultiplex_impl without UCI."""
    sym, nv = _demod_args(symbols, noise_vars)
    out = np.zeros(nv.size * qm, np.int8)
    n = ref().ref_pusch_demodulate(qm, int(pi2), rnti, n_id, nof_layers, nof_prb, start_symbol, nof_symbols, dmrs_mask,
                                   nof_cdm_groups_without_data, _pf(sym), _pf(nv), nv.size, _pi(out), out.size)
    assert n == out.size, n
    return out
