"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol include/*.h declares, refuses to
run without a CUDA device (no CPU fallback), and its host-only arithmetic (segmentation) matches the oracle."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import bindings as ob
from srsran_projectvtlmo_b200 import capi, synth

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = "".join(p.read_text() for p in sorted((ROOT / "include").glob("*.h")))
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(srsran_cuda_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    names = declared_symbols()
    assert len(names) >= 30
    lib = capi.lib()  # binds every entry of capi.SYMBOLS, AttributeError if one is missing
    out = subprocess.run(["nm", "-D", "--defined-only", str(capi.LIB_PATH)], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (srsran_cuda_[a-z0-9_]+)", out))
    assert set(names) <= exported, sorted(set(names) - exported)
    assert set(names) == set(capi.SYMBOLS), sorted(set(names) ^ set(capi.SYMBOLS))
    assert lib is not None


def test_no_cpu_fallback_without_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    lib = capi.lib()
    h = C.c_void_p()
    st = lib.srsran_cuda_pusch_dec_create(0, 16, 16, C.byref(h))
    assert st == capi.ERR_NO_DEVICE and not h.value
    assert b"CUDA" in lib.srsran_cuda_pusch_dec_last_error(None)
    from srsran_projectvtlmo_b200 import pusch

    with pytest.raises(capi.CudaPuschDecError):
        pusch.Accelerator(device=0)
    # every other entry point rejects a null handle instead of computing anything
    assert lib.srsran_cuda_pusch_dec_synchronize(None) == capi.ERR_INVALID
    assert lib.srsran_cuda_ldpc_decode(None, None, None, 0, 1, 384, 0, 0, 6, 0.8, None) == capi.ERR_INVALID


def test_reference_side_adapter_fails_loudly_without_device():
    """The reference's pusch_decoder_hw_impl + our hal accelerator (oracle/_ref/hwacc_parity): exit code 2 = no device."""
    import torch

    exe = ROOT / "oracle" / "_ref" / "hwacc_parity"
    if torch.cuda.is_available() or not exe.exists():
        pytest.skip("needs the harness binary and no CUDA device")
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 2 and "no usable CUDA device" in r.stdout


def test_segmentation_matches_oracle():
    from srsran_projectvtlmo_b200 import pusch

    for (prb, qm, R, nl, bg) in [(273, 8, 948, 4, 1), (273, 8, 948, 2, 1), (273, 8, 948, 1, 1), (52, 2, 120, 1, 2),
                                 (52, 2, 449, 1, 1), (52, 4, 378, 1, 1), (52, 4, 658, 1, 1), (25, 2, 120, 1, 2),
                                 (10, 4, 490, 1, 2), (4, 2, 308, 1, 2), (1, 2, 120, 1, 2), (106, 6, 567, 2, 1)]:
        tbs = synth.tbs_for(prb, qm, R, nl)
        nllr = prb * 156 * qm * nl
        a = pusch.segment(tbs, bg, qm, nl, nllr)
        b = ob.port_segment(tbs, bg, qm, nl, nllr)
        assert len(a) == len(b) > 0
        for x, y in zip(a, b):
            assert (x.base_graph, x.lifting_size, x.full_length, x.rm_length, x.nof_filler_bits, x.cw_offset, x.nof_crc_bits) == \
                   (y.bg, y.Z, y.full_length, y.rm_length, y.nof_filler_bits, y.cw_offset, y.nof_crc_bits)
    # 273 PRB / 256QAM / 4 layers -> 152 code blocks of Z = 384, 16 filler bits
    m = pusch.segment(1277992, 1, 8, 4, 1362816)
    assert len(m) == 152 and m[0].lifting_size == 384 and m[0].nof_filler_bits == 16
    assert sorted({x.rm_length for x in m}) == [8960, 8992]
    # inconsistent inputs are rejected, not guessed
    with pytest.raises(capi.CudaPuschDecError):
        pusch.segment(1277992, 1, 8, 4, 1362816 - 8)
