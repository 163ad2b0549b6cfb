"""Host-side mirror (Python) of the reference interfaces of the PUSCH channel-decoding path, over the C ABI.

The production host side is C++ (srsran_projectvtlmo_b200/host/*.h, compiled against the reference headers); this module
exposes the same operations with the same names and argument meaning to Python so that the parity tests read like the
reference's own tests:

  hw_accelerator_pusch_dec_cuda  ~ srsran::hal::hw_accelerator_pusch_dec      (hw_accelerator_pusch_dec.h:80-115)
  ldpc_decoder_cuda              ~ srsran::ldpc_decoder                       (ldpc_decoder.h:37-75)
  ldpc_rate_dematcher_cuda       ~ srsran::ldpc_rate_dematcher                (ldpc_rate_dematcher.h:35-56)
  crc_calculator_cuda            ~ srsran::crc_calculator                     (crc_calculator.h:62-84)
  pusch_decoder_cuda             ~ srsran::pusch_decoder + pusch_decoder_buffer (pusch_decoder.h:54-99)

Everything computes on the GPU through libsrsran_cuda_pusch_dec.so; nothing here falls back to the CPU.
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import capi
from .capi import CB_CRC16, CB_CRC24A, CB_CRC24B, CRC16, CRC24A, CRC24B, CRC_NONE, CbConfig, TbConfig, TbResult


def _i8(a):
    a = np.ascontiguousarray(a, dtype=np.int8)
    return a, a.ctypes.data_as(capi.i8p)


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a, a.ctypes.data_as(capi.u8p)


class Accelerator:
    """Owns one srsran_cuda_pusch_dec handle (one GPU, its HARQ slots and batch contexts)."""

    def __init__(self, device=0, max_cbs_in_flight=4096, nof_harq_cb_slots=4096):
        self._lib = capi.lib()
        h = C.c_void_p()
        st = self._lib.srsran_cuda_pusch_dec_create(device, max_cbs_in_flight, nof_harq_cb_slots, C.byref(h))
        if st != capi.OK:
            msg = self._lib.srsran_cuda_pusch_dec_last_error(None)
            raise capi.CudaPuschDecError(
                f"srsran_cuda_pusch_dec_create failed ({st}): {msg.decode() if msg else ''} - no CPU fallback exists")
        self.h = h
        self.device = device
        self.nof_harq_cb_slots = nof_harq_cb_slots

    def close(self):
        if getattr(self, "h", None):
            self._lib.srsran_cuda_pusch_dec_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st, what):
        return capi.check(self.h, st, what)

    @property
    def launch_count(self):
        return int(self._lib.srsran_cuda_pusch_dec_launch_count(self.h))

    def synchronize(self):
        self._check(self._lib.srsran_cuda_pusch_dec_synchronize(self.h), "synchronize")

    def timer_start(self):
        self._check(self._lib.srsran_cuda_pusch_dec_timer_start(self.h), "timer_start")

    def timer_stop(self):
        ms = C.c_float(0)
        self._check(self._lib.srsran_cuda_pusch_dec_timer_stop(self.h, C.byref(ms)), "timer_stop")
        return float(ms.value)

    def read_softbuffer(self, absolute_cb_id, n):
        out = np.zeros(n, np.int8)
        self._check(self._lib.srsran_cuda_pusch_dec_read_softbuffer(self.h, absolute_cb_id,
                                                                    out.ctypes.data_as(capi.i8p), n), "read_softbuffer")
        return out

    def write_softbuffer(self, absolute_cb_id, data):
        data, p = _i8(data)
        self._check(self._lib.srsran_cuda_pusch_dec_write_softbuffer(self.h, absolute_cb_id, p, data.size),
                    "write_softbuffer")

    def read_cb_crc(self, absolute_cb_id):
        v = C.c_int(0)
        self._check(self._lib.srsran_cuda_pusch_dec_read_cb_crc(self.h, absolute_cb_id, C.byref(v)), "read_cb_crc")
        return bool(v.value)

    def set_decoder_variant(self, variant):
        self._check(self._lib.srsran_cuda_pusch_dec_set_decoder_variant(self.h, variant), "set_decoder_variant")

    def set_tb_host_copy(self, enable):
        """False: decoded transport blocks stay in HBM (tb_data_device); results and CRC verdicts still come back."""
        self._check(self._lib.srsran_cuda_pusch_dec_set_tb_host_copy(self.h, int(bool(enable))), "set_tb_host_copy")

    def set_h2d_gather(self, nof_ctas):
        """CTAs of the kernel that reads soft bits arriving in many separate page-locked pieces (0: copy-engine jobs only)."""
        self._check(self._lib.srsran_cuda_pusch_dec_set_h2d_gather(self.h, nof_ctas), "set_h2d_gather")

    def set_direct_io(self, direct_in=True, direct_out=True):
        """Small batches: soft bits read from / results written to page-locked host memory by the kernels themselves."""
        self._check(self._lib.srsran_cuda_pusch_dec_set_direct_io(self.h, int(bool(direct_in)), int(bool(direct_out))), "set_direct_io")

    def set_combine_flavour(self, simd_block):
        self._check(self._lib.srsran_cuda_pusch_dec_set_combine_flavour(self.h, simd_block), "set_combine_flavour")


def segment(tbs_bits, base_graph, modulation, nof_layers, nof_llrs):
    """ldpc_segmenter_rx::segment metadata (host arithmetic inside the library)."""
    metas = (capi.CbMeta * capi.MAX_NOF_SEGMENTS)()
    c = capi.lib().srsran_cuda_pusch_dec_segment(tbs_bits, base_graph, modulation, nof_layers, nof_llrs, metas)
    if c < 0:
        raise capi.CudaPuschDecError(f"segmentation failed ({c})")
    return [metas[i] for i in range(c)]


class hw_accelerator_pusch_dec_cuda:
    """hal::hw_accelerator_pusch_dec over the C ABI (same method names and call order as pusch_decoder_hw_impl uses)."""

    def __init__(self, acc: Accelerator):
        self.acc = acc
        self._lib = acc._lib

    def reserve_queue(self):
        self.acc._check(self._lib.srsran_cuda_pusch_dec_reserve_queue(self.acc.h), "reserve_queue")

    def free_queue(self):
        self.acc._check(self._lib.srsran_cuda_pusch_dec_free_queue(self.acc.h), "free_queue")

    def configure_operation(self, config: CbConfig, cb_index=0):
        self.acc._check(self._lib.srsran_cuda_pusch_dec_configure(self.acc.h, cb_index, C.byref(config)),
                        "configure_operation")

    def enqueue_operation(self, data, aux_data=None, cb_index=0):
        data, p = _i8(data)
        st = self._lib.srsran_cuda_pusch_dec_enqueue(self.acc.h, cb_index, p, data.size, None, 0)
        return bool(self.acc._check(st, "enqueue_operation"))

    def dequeue_operation(self, data, aux_data=None, segment_index=0):
        assert data.dtype == np.uint8 and data.flags["C_CONTIGUOUS"]
        auxp, auxn = None, 0
        if aux_data is not None and aux_data.size:
            assert aux_data.dtype == np.int8 and aux_data.flags["C_CONTIGUOUS"]
            auxp, auxn = aux_data.ctypes.data_as(capi.i8p), aux_data.size
        st = self._lib.srsran_cuda_pusch_dec_dequeue(self.acc.h, segment_index, data.ctypes.data_as(capi.u8p),
                                                     data.size, auxp, auxn)
        return bool(self.acc._check(st, "dequeue_operation"))

    def read_operation_outputs(self, cb_index=0, absolute_cb_id=0):
        crc, it = C.c_int(0), C.c_uint32(0)
        self.acc._check(self._lib.srsran_cuda_pusch_dec_read_outputs(self.acc.h, cb_index, C.byref(crc), C.byref(it)),
                        "read_operation_outputs")
        return bool(crc.value), int(it.value)

    def free_harq_context_entry(self, absolute_cb_id):
        self.acc._check(self._lib.srsran_cuda_pusch_dec_free_harq(self.acc.h, absolute_cb_id), "free_harq_context_entry")

    def is_external_harq_supported(self):
        return bool(self._lib.srsran_cuda_pusch_dec_is_external_harq_supported(self.acc.h))


class ldpc_rate_dematcher_cuda:
    """ldpc_rate_dematcher::rate_dematch(output /*in-out*/, input, new_data, metadata)."""

    def __init__(self, acc: Accelerator):
        self.acc = acc

    def rate_dematch(self, output, input, new_data, rv, modulation, Nref, nof_filler_bits):
        assert output.dtype == np.int8 and output.flags["C_CONTIGUOUS"]
        input, pi = _i8(input)
        st = self.acc._lib.srsran_cuda_ldpc_rate_dematch(self.acc.h, output.ctypes.data_as(capi.i8p), output.size, pi,
                                                         input.size, int(new_data), rv, modulation, Nref,
                                                         nof_filler_bits)
        self.acc._check(st, "rate_dematch")
        return output


class ldpc_decoder_cuda:
    """ldpc_decoder::decode(output bits, input LLRs, crc, cfg): returns the iteration count or None (std::nullopt)."""

    def __init__(self, acc: Accelerator):
        self.acc = acc

    def decode(self, output, input, crc_poly, base_graph, lifting_size, nof_filler_bits=0, max_iterations=6,
               scaling_factor=0.8):
        assert output.dtype == np.uint8 and output.flags["C_CONTIGUOUS"]
        input, pi = _i8(input)
        it = C.c_int(0)
        st = self.acc._lib.srsran_cuda_ldpc_decode(self.acc.h, output.ctypes.data_as(capi.u8p), pi, input.size,
                                                   base_graph, lifting_size, nof_filler_bits, crc_poly, max_iterations,
                                                   scaling_factor, C.byref(it))
        self.acc._check(st, "ldpc_decode")
        return None if it.value < 0 else it.value

    def decode_batch(self, output, input, nof_cbs, crc_poly, base_graph, lifting_size, nof_filler_bits=0,
                     max_iterations=6, scaling_factor=0.8):
        assert output.dtype == np.uint8 and output.flags["C_CONTIGUOUS"]
        input, pi = _i8(input)
        its = np.zeros(nof_cbs, np.int32)
        st = self.acc._lib.srsran_cuda_ldpc_decode_batch(
            self.acc.h, output.ctypes.data_as(capi.u8p), pi, nof_cbs, input.size // nof_cbs, base_graph, lifting_size,
            nof_filler_bits, crc_poly, max_iterations, scaling_factor, its.ctypes.data_as(capi.intp))
        self.acc._check(st, "ldpc_decode_batch")
        return its


class crc_calculator_cuda:
    """crc_calculator::calculate / calculate_byte / calculate_bit for one generator polynomial."""

    def __init__(self, acc: Accelerator, poly):
        self.acc = acc
        self.poly = poly

    def get_generator_poly(self):
        return self.poly

    def calculate(self, packed, nof_bits):
        packed, p = _u8(packed)
        out = C.c_uint32(0)
        self.acc._check(self.acc._lib.srsran_cuda_crc_calculate(self.acc.h, self.poly, p, nof_bits, C.byref(out)),
                        "crc_calculate")
        return int(out.value)

    def calculate_byte(self, data):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        return self.calculate(data, 8 * data.size)

    def calculate_bit(self, bits):
        bits = np.ascontiguousarray(bits, dtype=np.uint8)
        return self.calculate(np.packbits(bits & 1), bits.size)


@dataclass
class pusch_decoder_configuration:
    """pusch_decoder::configuration (pusch_decoder.h:57-77)."""
    base_graph: int = 1
    rv: int = 0
    mod: int = 1
    Nref: int = 0
    nof_layers: int = 1
    nof_ldpc_iterations: int = 6
    use_early_stop: bool = True
    new_data: bool = True


class pusch_decoder_cuda:
    """pusch_decoder + pusch_decoder_buffer: new_data() -> on_new_softbits()* -> on_end_softbits() -> notifier.

    `harq_first_slot` plays the role of the unique_rx_buffer: the HARQ soft bits, decoded bits and CRC flags of code
    block i live in HBM slot harq_first_slot + i and persist across transmissions.
    """

    def __init__(self, acc: Accelerator):
        self.acc = acc
        self._chunks = []
        self._cfg = None

    def new_data(self, transport_block, harq_first_slot, notifier, cfg: pusch_decoder_configuration):
        assert transport_block.dtype == np.uint8
        self._tb = transport_block
        self._slot = harq_first_slot
        self._notifier = notifier
        self._cfg = cfg
        self._chunks = []
        return self

    def on_new_softbits(self, softbits):
        self._chunks.append(np.ascontiguousarray(softbits, dtype=np.int8))

    def on_end_softbits(self):
        llrs = self._chunks[0] if len(self._chunks) == 1 else np.concatenate(self._chunks)
        cfg = self._cfg
        tbc = TbConfig(self._tb.size * 8, cfg.base_graph, cfg.rv, cfg.mod, cfg.Nref, cfg.nof_layers,
                       cfg.nof_ldpc_iterations, int(cfg.use_early_stop), int(cfg.new_data), self._slot)
        lib = self.acc._lib
        ticket = lib.srsran_cuda_pusch_dec_submit_tb(self.acc.h, C.byref(tbc), llrs.ctypes.data_as(capi.i8p), llrs.size)
        self.acc._check(ticket, "submit_tb")
        res = TbResult()
        st = lib.srsran_cuda_pusch_dec_poll_tb(self.acc.h, ticket, 1, self._tb.ctypes.data_as(capi.u8p), C.byref(res))
        self.acc._check(st, "poll_tb")
        if self._notifier is not None:
            self._notifier(res)
        return res


class SubmitArgs:
    """The marshalled arguments of one submit_tbs call (configurations, LLR addresses and lengths), reusable: a caller
    that submits the same set of TB configurations and buffers slot after slot builds them once."""

    def __init__(self, configs, llrs_list, device_resident=False):
        n = len(configs)
        self.n = n
        self.device_resident = device_resident
        self.cfg_arr = (TbConfig * n)(*configs)
        self.ptrs = (C.c_void_p * n)()
        self.lens = (C.c_uint32 * n)()
        self.keep = []
        for i, a in enumerate(llrs_list):
            if device_resident:
                self.ptrs[i], self.lens[i] = a[0], a[1]
            else:
                self.keep.append(a)
                self.ptrs[i], self.lens[i] = a.ctypes.data, a.size
        self.tickets = (C.c_int * n)()


def submit_tbs(acc: Accelerator, configs, llrs_list=None, device_resident=False):
    """Batch submit: `llrs_list` holds numpy int8 arrays (host) or (device_ptr, n) tuples (device-resident); or `configs`
    is a prepared SubmitArgs."""
    a = configs if isinstance(configs, SubmitArgs) else SubmitArgs(configs, llrs_list, device_resident)
    fn = acc._lib.srsran_cuda_pusch_dec_submit_tbs_device if a.device_resident else acc._lib.srsran_cuda_pusch_dec_submit_tbs
    acc._check(fn(acc.h, a.n, a.cfg_arr, a.ptrs, a.lens, a.tickets), "submit_tbs")
    return list(a.tickets)


def submit_tb_streamed(acc: Accelerator, config, llr_blocks, cb_ids=None):
    """One TB whose LLRs arrive in blocks (pusch_decoder_buffer::on_new_softbits): every block's host -> device copy starts
    when it is pushed. `llr_blocks`: numpy int8 arrays that stay alive until the ticket completes; `cb_ids`: the rx
    buffer's absolute code-block ids (HARQ slots), or None for config.harq_first_slot + i. Returns the ticket."""
    total = sum(b.size for b in llr_blocks)
    sid = acc._check(acc._lib.srsran_cuda_pusch_dec_stream_begin(acc.h, total), "stream_begin")
    for b in llr_blocks:
        acc._check(acc._lib.srsran_cuda_pusch_dec_stream_push(acc.h, sid, b.ctypes.data_as(capi.i8p), b.size), "stream_push")
    ids, n = None, 0
    if cb_ids is not None:
        n = len(cb_ids)
        ids = (C.c_uint32 * n)(*cb_ids)
    return acc._check(acc._lib.srsran_cuda_pusch_dec_stream_submit(acc.h, sid, C.byref(config), ids, n), "stream_submit")


def tb_cb_outputs(acc: Accelerator, ticket, nof_cbs):
    """(crc flags, iteration observations) per code block of a completed TB; 0xffffffff = no observation."""
    crc = np.zeros(nof_cbs, np.uint8)
    its = np.zeros(nof_cbs, np.uint32)
    acc._check(acc._lib.srsran_cuda_pusch_dec_tb_cb_outputs(acc.h, ticket, crc.ctypes.data_as(capi.u8p),
                                                           its.ctypes.data_as(capi.u32p), nof_cbs), "tb_cb_outputs")
    return crc, its


def poll_tb(acc: Accelerator, ticket, tb_out=None, block=True):
    res = TbResult()
    p = tb_out.ctypes.data_as(capi.u8p) if tb_out is not None else None
    st = acc._lib.srsran_cuda_pusch_dec_poll_tb(acc.h, ticket, int(block), p, C.byref(res))
    acc._check(st, "poll_tb")
    return (res if st == 1 else None)


def poll_tbs(acc: Accelerator, tickets, tb_outs=None, block=True):
    """All tickets of a batch in one call; `tb_outs`: list of numpy uint8 arrays (or None entries) or None. Returns the
    list of results, or None when `block` is False and some TB is still in flight."""
    n = len(tickets)
    tk = (C.c_int * n)(*tickets)
    res = (TbResult * n)()
    ptrs = None
    if tb_outs is not None:
        ptrs = (capi.u8p * n)()
        for i, a in enumerate(tb_outs):
            if a is not None:
                ptrs[i] = a.ctypes.data_as(capi.u8p)
    st = acc._lib.srsran_cuda_pusch_dec_poll_tbs(acc.h, n, tk, int(block), ptrs, res)
    acc._check(st, "poll_tbs")
    return list(res) if st == n else None


def tb_data(acc: Accelerator, ticket, nbytes):
    """Zero-copy view (numpy, read-only use) of a completed TB in the batch's page-locked result buffer, or None."""
    p = capi.u8p()
    acc._check(acc._lib.srsran_cuda_pusch_dec_tb_data(acc.h, ticket, C.byref(p)), "tb_data")
    if not p:
        return None
    return np.ctypeslib.as_array(p, shape=(nbytes,))


def tb_data_device(acc: Accelerator, ticket):
    """Device address (int) of a completed transport block left in HBM (set_tb_host_copy(False)), or None."""
    p = capi.u8p()
    acc._check(acc._lib.srsran_cuda_pusch_dec_tb_data_device(acc.h, ticket, C.byref(p)), "tb_data_device")
    return C.cast(p, C.c_void_p).value


def ticket_timing(acc: Accelerator, ticket):
    """Device-side stage durations [h2d, dematch, decode, tb_crc, d2h] in ms of the batch `ticket` belongs to."""
    ms = (C.c_float * 5)()
    acc._check(acc._lib.srsran_cuda_pusch_dec_ticket_timing(acc.h, ticket, ms), "ticket_timing")
    return [float(v) for v in ms]


# ---- soft demodulation + descrambling + UL-SCH demultiplexing on the device -------------------------
DemodConfig = capi.DemodConfig


def demod_config(modulation, rnti, n_id, nof_layers, re_per_symbol, pi2_bpsk=False):
    """srsran_cuda_pusch_demod_config from the fields of pusch_demodulator::configuration that shape the soft bits:
    `re_per_symbol` = data REs per layer in every OFDM symbol of the allocation."""
    d = DemodConfig()
    d.modulation, d.pi2_bpsk, d.rnti, d.n_id, d.nof_layers = modulation, int(pi2_bpsk), rnti, n_id, nof_layers
    d.nof_ofdm_symbols = len(re_per_symbol)
    for i, v in enumerate(re_per_symbol):
        d.re_per_symbol[i] = int(v)
    return d


def _f32(symbols, noise_vars):
    sym = np.ascontiguousarray(symbols, dtype=np.complex64).view(np.float32)
    nv = np.ascontiguousarray(noise_vars, dtype=np.float32)
    assert sym.size == 2 * nv.size
    return sym, nv


class demodulation_mapper_cuda:
    """Mirror of srsran::demodulation_mapper (include/srsran/phy/upper/channel_modulation/demodulation_mapper.h): one call =
    one block."""

    def __init__(self, acc: Accelerator):
        self.acc = acc

    def demodulate_soft(self, symbols, noise_vars, modulation, pi2_bpsk=False):
        sym, nv = _f32(symbols, noise_vars)
        out = np.zeros(nv.size * modulation, np.int8)
        st = self.acc._lib.srsran_cuda_demodulate_soft(self.acc.h, out.ctypes.data_as(capi.i8p),
                                                       sym.ctypes.data_as(capi.f32p), nv.ctypes.data_as(capi.f32p),
                                                       nv.size, modulation, int(pi2_bpsk))
        self.acc._check(st, "demodulate_soft")
        return out


def pusch_demodulate(acc: Accelerator, symbols, noise_vars, config):
    """pusch_demodulator_impl::demodulate minus the equalizer + ulsch_demultiplex without UCI, one codeword, host buffers."""
    sym, nv = _f32(symbols, noise_vars)
    out = np.zeros(nv.size * config.modulation, np.int8)
    n = acc._lib.srsran_cuda_pusch_demodulate(acc.h, out.ctypes.data_as(capi.i8p), sym.ctypes.data_as(capi.f32p),
                                              nv.ctypes.data_as(capi.f32p), C.byref(config))
    acc._check(n, "pusch_demodulate")
    assert n == out.size
    return out


class SubmitSymbolArgs:
    """Marshalled arguments of one submit_tbs_symbols call. `symbols_list` / `noise_list`: numpy arrays (host: complex64 /
    float32, kept alive here) or device addresses (ints) when `device_resident`."""

    def __init__(self, configs, demod_configs, symbols_list, noise_list, device_resident=False):
        n = len(configs)
        self.n = n
        self.device_resident = device_resident
        self.cfg_arr = (TbConfig * n)(*configs)
        self.dm_arr = (DemodConfig * n)(*demod_configs)
        self.sym = (C.c_void_p * n)()
        self.nv = (C.c_void_p * n)()
        self.keep = []
        for i in range(n):
            if device_resident:
                self.sym[i], self.nv[i] = symbols_list[i], noise_list[i]
            else:
                s, v = _f32(symbols_list[i], noise_list[i])
                self.keep.append((s, v))
                self.sym[i], self.nv[i] = s.ctypes.data, v.ctypes.data
        self.tickets = (C.c_int * n)()


def submit_tbs_symbols(acc: Accelerator, args: SubmitSymbolArgs):
    """Equalized symbols + noise variances in, transport blocks out: demodulation, descrambling, rate dematching, LDPC
    decoding and TB assembly in one set of launches. Returns the tickets."""
    acc._check(acc._lib.srsran_cuda_pusch_dec_submit_tbs_symbols(acc.h, args.n, args.cfg_arr, args.dm_arr, args.sym,
                                                                 args.nv, args.tickets, int(args.device_resident)),
               "submit_tbs_symbols")
    return list(args.tickets)


def ticket_demod_ms(acc: Accelerator, ticket):
    ms = C.c_float(0)
    acc._check(acc._lib.srsran_cuda_pusch_dec_ticket_demod_ms(acc.h, ticket, C.byref(ms)), "ticket_demod_ms")
    return ms.value


def last_unit_timing(acc: Accelerator):
    """Stage durations (ms) of the last unit-level batch (ldpc_decoder_cuda.decode_batch, ...): see ticket_timing."""
    ms = (C.c_float * 5)()
    acc._check(acc._lib.srsran_cuda_pusch_dec_last_unit_timing(acc.h, ms), "last_unit_timing")
    return list(ms)
