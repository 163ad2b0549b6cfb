"""GPU tests of the reference-facing C++ layer beyond single-decoder parity: the upper-PHY
wiring on the reference's real rx_buffer_pool_impl with external soft bits, and many concurrent decoder instances through
the device's slot aggregator. The executables are built by oracle/Makefile (target hwacc) from the unmodified reference
sources plus srsran_projectvtlmo_b200/host/."""
import json
import subprocess
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _exe(name):
    exe = ROOT / "oracle" / "_ref" / name
    if not exe.exists():
        pytest.skip(f"oracle/_ref/{name} not built (needs /root/reference at build time)")
    return exe


def test_upper_phy_wiring_rx_buffer_pool_external_soft_bits():
    """pusch_decoder_type = "cuda" as the patched upper_phy_factories.cpp wires it (integration/0001-*.patch): decoder factory
    by configuration string, rx_buffer_pool_impl with external_soft_bits, interleaved HARQ processes, refused and expired
    reservations, recycled code-block identifiers - against the software decoder on a pool with internal soft bits."""
    r = subprocess.run([str(_exe("upper_phy_wiring"))], capture_output=True, text=True, timeout=600)
    print(r.stdout)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "PASSED" in r.stdout
    assert "refused by both pools" in r.stdout and "recycled" in r.stdout


def test_many_concurrent_decoders_through_the_slot_aggregator(tmp_path):
    """48 pusch_decoder_cuda_impl instances x 3 sets (144 transport blocks outstanding on ONE device with 8 batch contexts:
    the accelerator is busy most of the time and must answer with back-pressure, not with an error), every decoded transport
    block compared with its payload."""
    from srsran_projectvtlmo_b200 import synth
    import struct

    exe = _exe("hwacc_bench")
    prb, qm, rate, nl, bg, nref, mu = 106, 6, 873, 2, 1, 0, 11.0
    tbs = synth.tbs_for(prb, qm, rate, nl)
    nllr = prb * 156 * qm * nl
    rng = np.random.default_rng(5)
    f = tmp_path / "tbs.bin"
    with open(f, "wb") as fh:
        fh.write(struct.pack("<8I", 0x50425443, 4, tbs, bg, qm, nl, nref, nllr))
        for _ in range(4):
            tb = rng.integers(0, 256, tbs // 8, dtype=np.uint8)
            fh.write(tb.tobytes())
            fh.write(synth.awgn_llrs(rng, synth.encode_tb(tb, bg, 0, qm, nref, nl, nllr), mu).tobytes())
    r = subprocess.run([str(exe), "--llrs", str(f), "--decoders", "48", "--sets", "3", "--slots", "60", "--threads", "4",
                        "--workers", "3", "--agg-us", "30", "--ref-seconds", "0", "--check"], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["failed_or_wrong_tbs"] == 0 and line["payload_checked"] is True
    assert line["tb_latency_us"]["n"] >= 48 * 40
