"""Host-side mirror of the reference's PDSCH encoder interfaces over the CUDA accelerator.

Same names, argument meaning and error behaviour as the reference, so that the parity tests read like the reference's own:

* ``hw_accelerator_pdsch_enc_cuda``  - hal::hw_accelerator_pdsch_enc
  (include/srsran/hal/phy/upper/channel_processors/hw_accelerator_pdsch_enc.h:76-97): reserve_queue / free_queue /
  configure_operation / enqueue_operation / dequeue_operation / get_cb_mode / get_max_tb_size.
* ``pdsch_encoder_cuda``             - pdsch_encoder (include/srsran/phy/upper/channel_processors/pdsch_encoder.h):
  encode(codeword, transport_block, config), one transport block per call.
* ``encode_tbs``                     - a slot's worth of transport blocks in ONE launch.

Everything computes on the device through include/srsran_cuda_pdsch_enc.h; there is no CPU fallback.
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import capi
from .capi import PdschEncConfig as hw_pdsch_encoder_configuration  # noqa: F401  (hal::hw_pdsch_encoder_configuration)
from .capi import PdschEncTbConfig


class EncoderAccelerator:
    """One srsran_cuda_pdsch_enc handle (hw_accelerator_pdsch_enc_factory::create)."""

    def __init__(self, device=0, max_ops=256):
        self._lib = capi.lib()
        self.h = self._lib.srsran_cuda_pdsch_enc_create(device, max_ops)
        if not self.h:
            raise capi.CudaPuschDecError("srsran_cuda_pdsch_enc_create: " + self._lib.srsran_cuda_pdsch_enc_create_error().decode())
        self.max_ops = max_ops
        self._pinned_ptr, self._pinned_cap = None, 0

    def _check(self, st, what):
        if st < 0:
            raise capi.CudaPuschDecError(f"{what}: status {st}: {self._lib.srsran_cuda_pdsch_enc_last_error(self.h).decode()}")
        return st

    @property
    def launch_count(self):
        return int(self._lib.srsran_cuda_pdsch_enc_launch_count(self.h))

    def last_timing(self):
        """[host->device, kernels, device->host] of the last batch, ms (CUDA events)."""
        ms = (C.c_float * 3)()
        self._check(self._lib.srsran_cuda_pdsch_enc_last_timing(self.h, ms), "last_timing")
        return [float(v) for v in ms]

    def pinned(self, nbytes):
        """A uint8 view of page-locked host memory owned by the accelerator object (grow-only; reused by the next call), so that
        the outputs of a batch leave the device in one asynchronous copy."""
        if nbytes > self._pinned_cap:
            if self._pinned_ptr:
                self._lib.srsran_cuda_pusch_dec_host_free(self._pinned_ptr)
            self._pinned_ptr = self._lib.srsran_cuda_pusch_dec_host_alloc(nbytes)
            if not self._pinned_ptr:
                self._pinned_cap = 0
                raise capi.CudaPuschDecError("pinned host allocation failed")
            self._pinned_cap = nbytes
        return np.ctypeslib.as_array(C.cast(self._pinned_ptr, capi.u8p), shape=(max(nbytes, 1),))[:nbytes]

    def close(self):
        if self.h:
            self._lib.srsran_cuda_pdsch_enc_destroy(self.h)
            self.h = None
        if self._pinned_ptr:
            self._lib.srsran_cuda_pusch_dec_host_free(self._pinned_ptr)
            self._pinned_ptr, self._pinned_cap = None, 0


class hw_accelerator_pdsch_enc_cuda:
    """hal::hw_accelerator_pdsch_enc over the C ABI."""

    def __init__(self, acc: EncoderAccelerator, cb_mode=True, max_tb_size=1 << 20):
        self.acc = acc
        self.cb_mode = cb_mode
        self.max_tb_size = max_tb_size

    def reserve_queue(self):
        pass

    def free_queue(self):
        pass

    def get_cb_mode(self):
        return self.cb_mode

    def get_max_tb_size(self):
        return self.max_tb_size

    def configure_operation(self, config: hw_pdsch_encoder_configuration, cb_index=0):
        self.acc._check(self.acc._lib.srsran_cuda_pdsch_enc_configure(self.acc.h, cb_index, C.byref(config)), "configure_operation")

    def enqueue_operation(self, data: np.ndarray, aux_data=None, cb_index=0) -> bool:
        data = np.ascontiguousarray(data, dtype=np.uint8)
        st = self.acc._check(self.acc._lib.srsran_cuda_pdsch_enc_enqueue(self.acc.h, cb_index, data.ctypes.data_as(capi.u8p), data.size),
                             "enqueue_operation")
        return st == 1

    def dequeue_operation(self, data: np.ndarray, packed_data: np.ndarray = None, segment_index=0) -> bool:
        assert data.dtype == np.uint8 and data.flags.c_contiguous
        pk = packed_data.ctypes.data_as(capi.u8p) if packed_data is not None else None
        st = self.acc._check(self.acc._lib.srsran_cuda_pdsch_enc_dequeue(self.acc.h, segment_index, data.ctypes.data_as(capi.u8p), data.size,
                                                                         pk, 0 if packed_data is None else packed_data.size),
                             "dequeue_operation")
        return st == 1


@dataclass
class pdsch_encoder_configuration:
    """pdsch_encoder::configuration; `mod` = bits per symbol."""
    base_graph: int
    rv: int
    mod: int
    Nref: int
    nof_layers: int
    nof_ch_symbols: int


def encode_tbs(acc: EncoderAccelerator, configs, tbs, want_bits=True, want_packed=True):
    """Encodes the transport blocks `tbs` (uint8 arrays of TBS / 8 bytes) in one launch. Returns (codewords, packed): lists of
    uint8 arrays (one bit per byte / packed MSB first), None where not wanted. The arrays are views of the accelerator
    object's page-locked output buffer, laid out back to back (one device-to-host copy per kind): valid until the next
    call."""
    n = len(tbs)
    cfgs = (PdschEncTbConfig * n)()
    tb_ptrs = (capi.u8p * n)()
    cw_ptrs = (capi.u8p * n)()
    pk_ptrs = (capi.u8p * n)()
    nbits = [c.nof_ch_symbols * max(c.mod, 1) for c in configs]
    tot_bits = sum(nbits) if want_bits else 0
    tot_pk = sum((b + 7) // 8 for b in nbits) if want_packed else 0
    buf = acc.pinned(tot_bits + tot_pk)
    keep, cws, pks = [], [], []
    bo, po = 0, tot_bits
    for i, (c, tb) in enumerate(zip(configs, tbs)):
        tb = np.ascontiguousarray(tb, dtype=np.uint8)
        keep.append(tb)
        cfgs[i] = PdschEncTbConfig(tb.size * 8, c.base_graph, c.rv, c.mod, c.Nref, c.nof_layers, c.nof_ch_symbols)
        tb_ptrs[i] = tb.ctypes.data_as(capi.u8p)
        cws.append(buf[bo:bo + nbits[i]] if want_bits else None)
        pks.append(buf[po:po + (nbits[i] + 7) // 8] if want_packed else None)
        cw_ptrs[i] = cws[i].ctypes.data_as(capi.u8p) if want_bits else None
        pk_ptrs[i] = pks[i].ctypes.data_as(capi.u8p) if want_packed else None
        if want_bits:
            bo += nbits[i]
        if want_packed:
            po += (nbits[i] + 7) // 8
    acc._check(acc._lib.srsran_cuda_pdsch_enc_encode_tbs(acc.h, n, cfgs, tb_ptrs, cw_ptrs, pk_ptrs), "encode_tbs")
    return cws, pks


def encode_tbs_resident(acc: EncoderAccelerator, configs, tbs):
    """Encodes a batch and leaves the code words in device memory (the input of a modulation mapper running on the GPU).
    Returns (device_address, offsets, nbits): TB i's bits, one per byte, start at device_address + offsets[i]; valid until
    the next call on the accelerator."""
    n = len(tbs)
    cfgs = (PdschEncTbConfig * n)()
    tb_ptrs = (capi.u8p * n)()
    keep = []
    for i, (c, tb) in enumerate(zip(configs, tbs)):
        tb = np.ascontiguousarray(tb, dtype=np.uint8)
        keep.append(tb)
        cfgs[i] = PdschEncTbConfig(tb.size * 8, c.base_graph, c.rv, c.mod, c.Nref, c.nof_layers, c.nof_ch_symbols)
        tb_ptrs[i] = tb.ctypes.data_as(capi.u8p)
    dev = capi.u8p()
    offs = (C.c_uint64 * n)()
    acc._check(acc._lib.srsran_cuda_pdsch_enc_encode_tbs_resident(acc.h, n, cfgs, tb_ptrs, C.byref(dev), offs), "encode_tbs_resident")
    return C.cast(dev, C.c_void_p).value, [int(o) for o in offs], [c.nof_ch_symbols * max(c.mod, 1) for c in configs]


class pdsch_encoder_cuda:
    """pdsch_encoder over the accelerator: encode(codeword, transport_block, config)."""

    def __init__(self, acc: EncoderAccelerator):
        self.acc = acc

    def encode(self, codeword: np.ndarray, transport_block: np.ndarray, config: pdsch_encoder_configuration):
        assert codeword.dtype == np.uint8 and codeword.size == config.nof_ch_symbols * max(config.mod, 1), "Wrong codeword length."
        cws, _ = encode_tbs(self.acc, [config], [transport_block], want_packed=False)
        codeword[:] = cws[0]
