"""ctypes binding of the C ABI declared in include/srsran_cuda_pusch_dec.h.

This is the only way Python reaches the product: the shared library libsrsran_cuda_pusch_dec.so (hand-written sm_100a
kernels + host runtime, built in-tree by ``csrc/Makefile``). There is NO CPU fallback: if the library is missing or no
CUDA device is usable, every entry point raises.
"""
import ctypes as C
import subprocess
from pathlib import Path

import os

HERE = Path(__file__).resolve().parent
# SRSRAN_CUDA_PUSCH_DEC_LIB: another build of the same library (kernel A/B measurements); never a different implementation.
LIB_PATH = Path(os.environ.get("SRSRAN_CUDA_PUSCH_DEC_LIB", HERE / "libsrsran_cuda_pusch_dec.so"))

OK, ERR_NO_DEVICE, ERR_INVALID, ERR_NO_MEMORY, ERR_CUDA, ERR_STATE, ERR_BUSY = 0, -1, -2, -3, -4, -5, -6
CRC_NONE, CRC24A, CRC24B, CRC16 = 0, 1, 2, 3
CB_CRC16, CB_CRC24B, CB_CRC24A = 0, 1, 2
MAX_NOF_SEGMENTS = 162
MAX_CB_LENGTH = 25344

u8p = C.POINTER(C.c_uint8)
i8p = C.POINTER(C.c_int8)
u32p = C.POINTER(C.c_uint32)
intp = C.POINTER(C.c_int)


class CbConfig(C.Structure):
    """srsran_cuda_pusch_dec_cb_config == hal::hw_pusch_decoder_configuration."""
    _fields_ = [(n, C.c_uint32) for n in (
        "base_graph", "modulation", "nof_segments", "rv", "cw_length", "lifting_size", "Ncb", "Nref",
        "nof_segment_bits", "nof_filler_bits", "max_nof_ldpc_iterations", "use_early_stop", "new_data", "cb_crc_len",
        "cb_crc_type", "absolute_cb_id")]


class TbConfig(C.Structure):
    """srsran_cuda_pusch_dec_tb_config == pusch_decoder::configuration + TBS + HARQ slot."""
    _fields_ = [(n, C.c_uint32) for n in (
        "tbs_bits", "base_graph", "rv", "modulation", "Nref", "nof_layers", "nof_ldpc_iterations", "use_early_stop",
        "new_data", "harq_first_slot")]


class TbResult(C.Structure):
    _fields_ = [("tb_crc_ok", C.c_int32), ("nof_codeblocks_total", C.c_uint32), ("nof_observations", C.c_uint32),
                ("iter_min", C.c_uint32), ("iter_max", C.c_uint32), ("iter_mean", C.c_float)]


class CbMeta(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in
                ("base_graph", "lifting_size", "full_length", "rm_length", "nof_filler_bits", "cw_offset",
                 "nof_crc_bits")]


class PdschEncConfig(C.Structure):
    """srsran_cuda_pdsch_enc_config == hal::hw_pdsch_encoder_configuration (include/srsran_cuda_pdsch_enc.h)."""
    _fields_ = [(n, C.c_uint32) for n in (
        "nof_tb_bits", "nof_tb_crc_bits", "base_graph", "modulation", "nof_segments", "nof_short_segments", "rv",
        "cw_length_a", "cw_length_b", "lifting_size", "Ncb", "Nref", "nof_segment_bits", "nof_filler_bits",
        "rm_length")] + [("tb_crc", C.c_uint8 * 3), ("cb_mode", C.c_uint8)]


class PdschEncTbConfig(C.Structure):
    """srsran_cuda_pdsch_enc_tb_config == pdsch_encoder::configuration + TBS."""
    _fields_ = [(n, C.c_uint32) for n in
                ("tbs_bits", "base_graph", "rv", "modulation", "Nref", "nof_layers", "nof_ch_symbols")]


class DemodConfig(C.Structure):
    """srsran_cuda_pusch_demod_config: what pusch_demodulator::configuration fixes about a codeword's soft demodulation."""
    _fields_ = [("modulation", C.c_uint32), ("pi2_bpsk", C.c_uint32), ("rnti", C.c_uint32), ("n_id", C.c_uint32),
                ("nof_layers", C.c_uint32), ("nof_ofdm_symbols", C.c_uint32), ("re_per_symbol", C.c_uint32 * 14)]


f32p = C.POINTER(C.c_float)

# Every symbol include/srsran_cuda_pusch_dec.h declares: name -> (restype, argtypes).
SYMBOLS = {
    "srsran_cuda_pusch_dec_create": (C.c_int, [C.c_int, C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p)]),
    "srsran_cuda_pusch_dec_destroy": (None, [C.c_void_p]),
    "srsran_cuda_pusch_dec_last_error": (C.c_char_p, [C.c_void_p]),
    "srsran_cuda_pusch_dec_launch_count": (C.c_uint64, [C.c_void_p]),
    "srsran_cuda_pusch_dec_set_combine_flavour": (C.c_int, [C.c_void_p, C.c_uint32]),
    "srsran_cuda_pusch_dec_set_decoder_variant": (C.c_int, [C.c_void_p, C.c_uint32]),
    "srsran_cuda_pusch_dec_host_alloc": (C.c_void_p, [C.c_size_t]),
    "srsran_cuda_pusch_dec_host_free": (None, [C.c_void_p]),
    "srsran_cuda_pusch_dec_reserve_queue": (C.c_int, [C.c_void_p]),
    "srsran_cuda_pusch_dec_free_queue": (C.c_int, [C.c_void_p]),
    "srsran_cuda_pusch_dec_configure": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(CbConfig)]),
    "srsran_cuda_pusch_dec_enqueue": (C.c_int, [C.c_void_p, C.c_uint32, i8p, C.c_uint32, i8p, C.c_uint32]),
    "srsran_cuda_pusch_dec_dequeue": (C.c_int, [C.c_void_p, C.c_uint32, u8p, C.c_uint32, i8p, C.c_uint32]),
    "srsran_cuda_pusch_dec_read_outputs": (C.c_int, [C.c_void_p, C.c_uint32, intp, u32p]),
    "srsran_cuda_pusch_dec_free_harq": (C.c_int, [C.c_void_p, C.c_uint32]),
    "srsran_cuda_pusch_dec_is_external_harq_supported": (C.c_int, [C.c_void_p]),
    "srsran_cuda_pusch_dec_segment": (C.c_int, [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                                 C.POINTER(CbMeta)]),
    "srsran_cuda_pusch_dec_submit_tb": (C.c_int, [C.c_void_p, C.POINTER(TbConfig), i8p, C.c_uint32]),
    "srsran_cuda_pusch_dec_submit_tb_cb_ids": (C.c_int, [C.c_void_p, C.POINTER(TbConfig), i8p, C.c_uint32, u32p, C.c_uint32]),
    "srsran_cuda_pusch_dec_tb_cb_outputs": (C.c_int, [C.c_void_p, C.c_int, u8p, u32p, C.c_uint32]),
    "srsran_cuda_pusch_dec_stream_begin": (C.c_int, [C.c_void_p, C.c_uint32]),
    "srsran_cuda_pusch_dec_stream_push": (C.c_int, [C.c_void_p, C.c_int, i8p, C.c_uint32]),
    "srsran_cuda_pusch_dec_stream_submit": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(TbConfig), u32p, C.c_uint32]),
    "srsran_cuda_pusch_dec_poll_tb": (C.c_int, [C.c_void_p, C.c_int, C.c_int, u8p, C.POINTER(TbResult)]),
    "srsran_cuda_pusch_dec_poll_tbs": (C.c_int, [C.c_void_p, C.c_uint32, intp, C.c_int, C.POINTER(u8p),
                                                  C.POINTER(TbResult)]),
    "srsran_cuda_pusch_dec_tb_data": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(u8p)]),
    "srsran_cuda_pusch_dec_tb_data_device": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(u8p)]),
    "srsran_cuda_pusch_dec_set_tb_host_copy": (C.c_int, [C.c_void_p, C.c_int]),
    "srsran_cuda_pusch_dec_set_h2d_gather": (C.c_int, [C.c_void_p, C.c_uint32]),
    "srsran_cuda_pusch_dec_set_direct_io": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "srsran_cuda_pusch_dec_submit_tbs_device": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(TbConfig),
                                                           C.POINTER(C.c_void_p), u32p, intp]),
    "srsran_cuda_pusch_dec_submit_tbs": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(TbConfig),
                                                    C.POINTER(C.c_void_p), u32p, intp]),
    "srsran_cuda_pusch_dec_ticket_timing": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_float)]),
    "srsran_cuda_pusch_dec_last_unit_timing": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "srsran_cuda_pusch_dec_submit_tbs_cb_ids": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(TbConfig), C.POINTER(C.c_void_p),
                                                           u32p, intp, u32p, u32p, intp]),
    "srsran_cuda_pusch_dec_wait_ticket": (C.c_int, [C.c_void_p, C.c_int]),
    "srsran_cuda_pusch_dec_peek_tb": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(TbResult)]),
    "srsran_cuda_demodulate_soft": (C.c_int, [C.c_void_p, i8p, f32p, f32p, C.c_uint32, C.c_uint32, C.c_uint32]),
    "srsran_cuda_pusch_demodulate": (C.c_int, [C.c_void_p, i8p, f32p, f32p, C.POINTER(DemodConfig)]),
    "srsran_cuda_pusch_dec_submit_tbs_symbols": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(TbConfig),
                                                            C.POINTER(DemodConfig), C.POINTER(C.c_void_p),
                                                            C.POINTER(C.c_void_p), intp, C.c_int]),
    "srsran_cuda_pusch_dec_ticket_demod_ms": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_float)]),
    "srsran_cuda_pusch_dec_timer_start": (C.c_int, [C.c_void_p]),
    "srsran_cuda_pusch_dec_timer_stop": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "srsran_cuda_pusch_dec_synchronize": (C.c_int, [C.c_void_p]),
    "srsran_cuda_ldpc_rate_dematch": (C.c_int, [C.c_void_p, i8p, C.c_uint32, i8p, C.c_uint32, C.c_int, C.c_uint32,
                                                 C.c_uint32, C.c_uint32, C.c_uint32]),
    "srsran_cuda_ldpc_decode": (C.c_int, [C.c_void_p, u8p, i8p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                           C.c_uint32, C.c_uint32, C.c_float, intp]),
    "srsran_cuda_ldpc_decode_batch": (C.c_int, [C.c_void_p, u8p, i8p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                                 C.c_uint32, C.c_uint32, C.c_uint32, C.c_float, intp]),
    "srsran_cuda_crc_calculate": (C.c_int, [C.c_void_p, C.c_uint32, u8p, C.c_uint32, u32p]),
    "srsran_cuda_pusch_dec_read_softbuffer": (C.c_int, [C.c_void_p, C.c_uint32, i8p, C.c_uint32]),
    "srsran_cuda_pusch_dec_write_softbuffer": (C.c_int, [C.c_void_p, C.c_uint32, i8p, C.c_uint32]),
    "srsran_cuda_pusch_dec_read_cb_crc": (C.c_int, [C.c_void_p, C.c_uint32, intp]),
    # ---- include/srsran_cuda_pdsch_enc.h ----
    "srsran_cuda_pdsch_enc_create": (C.c_void_p, [C.c_int, C.c_uint32]),
    "srsran_cuda_pdsch_enc_destroy": (None, [C.c_void_p]),
    "srsran_cuda_pdsch_enc_last_error": (C.c_char_p, [C.c_void_p]),
    "srsran_cuda_pdsch_enc_create_error": (C.c_char_p, []),
    "srsran_cuda_pdsch_enc_configure": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(PdschEncConfig)]),
    "srsran_cuda_pdsch_enc_enqueue": (C.c_int, [C.c_void_p, C.c_uint32, u8p, C.c_uint32]),
    "srsran_cuda_pdsch_enc_dequeue": (C.c_int, [C.c_void_p, C.c_uint32, u8p, C.c_uint32, u8p, C.c_uint32]),
    "srsran_cuda_pdsch_enc_encode_tbs": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(PdschEncTbConfig), C.POINTER(u8p),
                                                   C.POINTER(u8p), C.POINTER(u8p)]),
    "srsran_cuda_pdsch_enc_encode_tbs_resident": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(PdschEncTbConfig), C.POINTER(u8p),
                                                            C.POINTER(u8p), C.POINTER(C.c_uint64)]),
    "srsran_cuda_pdsch_enc_last_timing": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "srsran_cuda_pdsch_enc_launch_count": (C.c_uint64, [C.c_void_p]),
}


class CudaPuschDecError(RuntimeError):
    pass


def build(verbose=False):
    """Compiles the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", str(HERE / "csrc")], capture_output=not verbose, text=True)
    if r.returncode != 0:
        raise CudaPuschDecError("building libsrsran_cuda_pusch_dec.so failed:\n" + (r.stdout or "") + (r.stderr or ""))
    return LIB_PATH


_lib = None


def lib():
    """Loads the shared library (never a fallback) and binds every declared symbol."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise CudaPuschDecError(
                f"{LIB_PATH} is missing: build it with `make -C srsran_projectvtlmo_b200/csrc` "
                "(or __graft_entry__.build()). There is no CPU fallback.")
        handle = C.CDLL(str(LIB_PATH))
        for name, (restype, argtypes) in SYMBOLS.items():
            fn = getattr(handle, name)  # AttributeError if a declared symbol is not exported
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
    return _lib


def check(handle, status, what):
    if status < 0:
        msg = lib().srsran_cuda_pusch_dec_last_error(handle)
        raise CudaPuschDecError(f"{what} failed with status {status}: {msg.decode() if msg else ''}")
    return status
