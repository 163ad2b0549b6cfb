"""GPU parity tests of the PDSCH encoding accelerator (CRC attachment + LDPC encoding + rate matching
behind hal::hw_accelerator_pdsch_enc) against the compiled, unmodified reference's software pdsch_encoder_impl
(oracle/_ref) where it is available, the committed golden fixture made from it (tests/golden/pdsch_enc.npz), and the numpy
transmitter that test_oracle_cpu.py pins to the reference. Bit-exact: every code-word bit, one per byte and packed."""
from pathlib import Path

import numpy as np
import pytest

from oracle import bindings as ob
from srsran_projectvtlmo_b200 import pdsch, synth

pytestmark = pytest.mark.gpu

HAVE_REF = ob.ref() is not None and ob.ref_flavour() is not None
GOLDEN = Path(__file__).resolve().parent / "golden" / "pdsch_enc.npz"


@pytest.fixture(scope="module")
def enc():
    a = pdsch.EncoderAccelerator(device=0, max_ops=256)
    yield a
    a.close()


def _reference_codeword(tb, bg, rv, qm, nref, nl, nbits):
    if HAVE_REF:
        return ob.ref_encode_tb(tb, bg, rv, qm, nref, nl, nbits // qm)
    return synth.encode_tb(tb, bg, rv, qm, nref, nl, nbits)


def test_golden_codewords(enc):
    g = np.load(GOLDEN)
    cfgs, tbs = [], []
    for i, (prb, qm, R, nl, bg, nref, rv) in enumerate(g["cases"]):
        cfgs.append(pdsch.pdsch_encoder_configuration(int(bg), int(rv), int(qm), int(nref), int(nl), int(prb) * 156 * int(nl)))
        tbs.append(g[f"tb{i}"])
    launches = enc.launch_count
    cws, pks = pdsch.encode_tbs(enc, cfgs, tbs)
    assert enc.launch_count - launches == 4  # TB CRCs, the two encoder forms (Z % 32 == 0 or not), whole-TB packing: one batch
    for i in range(len(cfgs)):
        assert np.array_equal(np.packbits(cws[i]), g[f"cw{i}"]), i
        assert np.array_equal(pks[i], g[f"cw{i}"]), i
        assert cws[i].max() <= 1


def test_random_transport_blocks_vs_reference(enc):
    """pdsch_encoder::encode one TB at a time (pdsch_encoder_test.cpp structure: encode, compare the code word): random
    allocations over both base graphs, every modulation order and rv, limited and unlimited buffers."""
    rng = np.random.default_rng(515)
    e = pdsch.pdsch_encoder_cuda(enc)
    done = 0
    while done < 40:
        prb = int(rng.integers(1, 274))
        qm = int(rng.choice([2, 4, 6, 8]))
        R = int(rng.integers(100, 949))
        nl = int(rng.integers(1, 5))
        tbs = synth.tbs_for(prb, qm, R, nl)
        if tbs < 24 or tbs > 1100000:  # (the reference's Tx segmenter asserts on larger TBs: ldpc_segmenter_impl.cpp:77)
            continue
        bg = 2 if (tbs <= 292 or (tbs <= 3824 and R <= 686) or R <= 256) else 1
        nbits = prb * 156 * qm * nl
        n_full = (66 if bg == 1 else 50) * synth.segmentation(tbs, bg)[1]
        nref = int(rng.choice([0, 0, n_full, n_full * 2 // 3, n_full // 2 + 7]))
        rv = int(rng.integers(0, 4))
        tb = rng.integers(0, 256, tbs // 8, dtype=np.uint8)
        cw = np.zeros(nbits, np.uint8)
        e.encode(cw, tb, pdsch.pdsch_encoder_configuration(bg, rv, qm, nref, nl, nbits // qm))
        assert np.array_equal(cw, _reference_codeword(tb, bg, rv, qm, nref, nl, nbits)), (prb, qm, R, nl, bg, nref, rv)
        done += 1


def test_headline_slot_of_64_tbs_one_launch(enc):
    """64 cells x one 100 MHz / 256QAM / 4-layer TB (the downlink mirror of BASELINE config 5) in one launch; every TB checked
    against the numpy transmitter (the reference's Tx segmenter refuses this TBS: ldpc_segmenter_impl.cpp:77), two of them
    also code block by code block against the compiled reference's ldpc_encoder + rate matcher through the CB-mode seam."""
    rng = np.random.default_rng(516)
    prb, qm, R, nl, bg, nref = 273, 8, 948, 4, 1, 12611
    tbs = synth.tbs_for(prb, qm, R, nl)
    nbits = prb * 156 * qm * nl
    tb = [rng.integers(0, 256, tbs // 8, dtype=np.uint8) for _ in range(64)]
    cfgs = [pdsch.pdsch_encoder_configuration(bg, i % 4, qm, nref, nl, nbits // qm) for i in range(64)]
    cws, pks = pdsch.encode_tbs(enc, cfgs, tb)
    cws, pks = [c.copy() for c in cws], [p.copy() for p in pks]
    for i in range(0, 64, 9):
        want = synth.encode_tb(tb[i], bg, i % 4, qm, nref, nl, nbits)
        assert np.array_equal(cws[i], want), i
        assert np.array_equal(pks[i], np.packbits(want)), i
    for i in range(64):  # cheap whole-batch property: TBs with the same rv and payload would coincide; packed == bits
        assert np.array_equal(np.packbits(cws[i]), pks[i]), i


def test_code_words_left_in_device_memory(enc):
    """encode_tbs_resident: the code words stay in HBM (the modulation mapper's input when it runs on the GPU too)."""
    import ctypes

    rng = np.random.default_rng(519)
    cases = [(52, 4, 658, 1, 1, 0, 0), (106, 6, 873, 2, 1, 0, 1), (25, 2, 120, 1, 2, 0, 2)]
    cfgs, tbs, want = [], [], []
    for prb, qm, R, nl, bg, nref, rv in cases:
        nbits = prb * 156 * qm * nl
        tb = rng.integers(0, 256, synth.tbs_for(prb, qm, R, nl) // 8, dtype=np.uint8)
        cfgs.append(pdsch.pdsch_encoder_configuration(bg, rv, qm, nref, nl, nbits // qm))
        tbs.append(tb)
        want.append(_reference_codeword(tb, bg, rv, qm, nref, nl, nbits))
    dev, offs, nbits = pdsch.encode_tbs_resident(enc, cfgs, tbs)
    cudart = ctypes.CDLL("libcudart.so.12")  # the runtime the library itself is linked against
    cudart.cudaMemcpy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
    for i in range(len(cases)):
        got = np.zeros(nbits[i], np.uint8)
        assert cudart.cudaMemcpy(got.ctypes.data, dev + offs[i], got.size, 2) == 0  # device -> host
        assert np.array_equal(got, want[i]), cases[i]


def _hw_config(tb, bg, rv, qm, nref, nl, nbits, cb_mode):
    """What pdsch_encoder_hw_impl::set_hw_enc_tb_configuration (pdsch_encoder_hw_impl.cpp:196-328) fills in."""
    msg, z, kp, nfill, tb_crc = synth.tx_segments(tb, bg)
    c = msg.shape[0]
    tbs = tb.size * 8
    lens = synth.rm_lengths(c, nbits, qm, nl)
    nsl = (nbits // qm) // nl
    cfg = pdsch.hw_pdsch_encoder_configuration()
    cfg.nof_tb_bits, cfg.nof_tb_crc_bits = tbs, 16 if tbs <= 3824 else 24
    cfg.base_graph, cfg.modulation, cfg.nof_segments, cfg.nof_short_segments, cfg.rv = bg, qm, c, c - (nsl % c), rv
    cfg.cw_length_a, cfg.cw_length_b, cfg.lifting_size = lens[0], lens[-1], z
    cfg.Ncb, cfg.Nref = (66 if bg == 1 else 50) * z, nref
    cfg.nof_segment_bits = kp - (24 if c > 1 else 0)
    cfg.nof_filler_bits = nfill
    crc_bytes = [(tb_crc >> 16) & 0xFF, (tb_crc >> 8) & 0xFF, tb_crc & 0xFF]
    if cfg.nof_tb_crc_bits == 16:
        crc_bytes = crc_bytes[1:] + [0]
    for k in range(3):
        cfg.tb_crc[k] = crc_bytes[k]
    cfg.cb_mode = 1 if cb_mode else 0
    return cfg, msg, kp, lens


@pytest.mark.parametrize("case", [(52, 4, 658, 1, 1, 0, 0), (106, 6, 873, 2, 1, 0, 1), (24, 8, 948, 2, 1, 12611, 3),
                                  (25, 2, 120, 1, 2, 0, 2), (273, 2, 193, 2, 2, 0, 0), (1, 2, 120, 1, 2, 0, 1)])
def test_hw_accelerator_call_sequence_cb_and_tb_mode(enc, case):
    """The call sequence of pdsch_encoder_hw_impl::encode (pdsch_encoder_hw_impl.cpp:34-176): configure + enqueue every code
    block, then dequeue them in order (CB mode); one operation for the whole TB (TB mode)."""
    prb, qm, R, nl, bg, nref, rv = case
    rng = np.random.default_rng(517 + prb)
    tbs = synth.tbs_for(prb, qm, R, nl)
    nbits = prb * 156 * qm * nl
    tb = rng.integers(0, 256, tbs // 8, dtype=np.uint8)
    want = _reference_codeword(tb, bg, rv, qm, nref, nl, nbits)
    # CB mode
    hw = pdsch.hw_accelerator_pdsch_enc_cuda(enc, cb_mode=True)
    cfg, msg, kp, lens = _hw_config(tb, bg, rv, qm, nref, nl, nbits, True)
    hw.reserve_queue()
    for cb in range(msg.shape[0]):
        cfg.rm_length = lens[cb]
        hw.configure_operation(cfg, cb)
        assert hw.enqueue_operation(np.packbits(msg[cb, :kp]), None, cb)
    launches = enc.launch_count
    off = 0
    for cb in range(msg.shape[0]):
        bits = np.zeros(lens[cb], np.uint8)
        packed = np.zeros((lens[cb] + 7) // 8, np.uint8)
        assert hw.dequeue_operation(bits, packed, cb)
        assert np.array_equal(bits, want[off:off + lens[cb]]), (case, cb)
        assert np.array_equal(packed, np.packbits(bits)), (case, cb)
        off += lens[cb]
    assert enc.launch_count - launches == 1  # the whole TB was one kernel launch, triggered by the first dequeue
    assert not hw.dequeue_operation(np.zeros(lens[0], np.uint8), None, 0)  # nothing left under that index
    hw.free_queue()
    # TB mode
    hw = pdsch.hw_accelerator_pdsch_enc_cuda(enc, cb_mode=False)
    cfg, msg, kp, lens = _hw_config(tb, bg, rv, qm, nref, nl, nbits, False)
    hw.configure_operation(cfg, 0)
    assert hw.enqueue_operation(tb, None, 0)
    bits = np.zeros(nbits, np.uint8)
    packed = np.zeros((nbits + 7) // 8, np.uint8)
    assert hw.dequeue_operation(bits, packed, 0)
    assert np.array_equal(bits, want), case
    assert np.array_equal(packed, np.packbits(want)), case


def test_code_blocks_all_lifting_sizes(enc):
    """Every lifting size x both base graphs at code-block level (CB mode), random fillers, rv, modulation order, Nref,
    lengths incl. repetition (E > Ncb) - against the compiled reference's ldpc_encoder + the numpy rate matcher."""
    rng = np.random.default_rng(518)
    hw = pdsch.hw_accelerator_pdsch_enc_cuda(enc, cb_mode=True)
    jobs = []
    for bg in (1, 2):
        for z in synth.ALL_Z:
            k, n = synth.kb(bg) * z, synth.ns(bg) * z
            nfill = int(rng.integers(0, max(1, min(k - 2 * z - 1, 3 * z)))) if z > 2 else 0
            msg = np.zeros(k, np.uint8)
            msg[:k - nfill] = rng.integers(0, 2, k - nfill, dtype=np.uint8)
            qm = int(rng.choice([1, 2, 4, 6, 8]))
            e_len = int(rng.integers(max(1, n // (3 * qm)), (2 * n) // qm + 1)) * qm
            rv = int(rng.integers(0, 4))
            nref = int(rng.choice([0, n, (2 * n) // 3, k + 5 * z + 1]))
            jobs.append((bg, z, nfill, msg, qm, e_len, rv, nref))
    for base in range(0, len(jobs), 64):
        chunk = jobs[base:base + 64]
        for i, (bg, z, nfill, msg, qm, e_len, rv, nref) in enumerate(chunk):
            cfg = pdsch.hw_pdsch_encoder_configuration()
            cfg.base_graph, cfg.modulation, cfg.nof_segments, cfg.rv, cfg.lifting_size = bg, qm, 1, rv, z
            cfg.Ncb, cfg.Nref, cfg.nof_filler_bits, cfg.rm_length, cfg.cb_mode = synth.ns(bg) * z, nref, nfill, e_len, 1
            hw.configure_operation(cfg, i)
            assert hw.enqueue_operation(np.packbits(msg[:msg.size - nfill]), None, i)
        for i, (bg, z, nfill, msg, qm, e_len, rv, nref) in enumerate(chunk):
            bits = np.zeros(e_len, np.uint8)
            assert hw.dequeue_operation(bits, None, i)
            cw = ob.ref_ldpc_encode(msg, bg, z) if HAVE_REF else synth.ldpc_encode(msg, bg, z)
            want = synth.rate_match(cw, bg, z, nfill, e_len, rv, qm, nref)
            assert np.array_equal(bits, want), (bg, z, nfill, qm, e_len, rv, nref)


def test_interleaved_enqueue_and_dequeue(enc):
    """Results that are complete but not dequeued yet survive the launch of a later batch (a caller that enqueues more code
    blocks before it has fetched all earlier ones)."""
    rng = np.random.default_rng(520)
    hw = pdsch.hw_accelerator_pdsch_enc_cuda(enc, cb_mode=True)
    ops = []
    for i, (bg, z, qm, e_len) in enumerate([(1, 384, 8, 8960), (2, 208, 2, 5000), (1, 96, 4, 4096), (1, 384, 6, 17100)]):
        k = synth.kb(bg) * z
        msg = rng.integers(0, 2, k, dtype=np.uint8)
        cfg = pdsch.hw_pdsch_encoder_configuration()
        cfg.base_graph, cfg.modulation, cfg.nof_segments, cfg.rv, cfg.lifting_size = bg, qm, 1, i % 4, z
        cfg.Ncb, cfg.Nref, cfg.nof_filler_bits, cfg.rm_length, cfg.cb_mode = synth.ns(bg) * z, 0, 0, e_len, 1
        want = synth.rate_match(synth.ldpc_encode(msg, bg, z), bg, z, 0, e_len, i % 4, qm, 0)
        ops.append((cfg, np.packbits(msg), want))

    def put(i):
        hw.configure_operation(ops[i][0], i)
        assert hw.enqueue_operation(ops[i][1], None, i)

    def get(i):
        bits = np.zeros(ops[i][2].size, np.uint8)
        packed = np.zeros((bits.size + 7) // 8, np.uint8)
        assert hw.dequeue_operation(bits, packed, i)
        assert np.array_equal(bits, ops[i][2]), i
        assert np.array_equal(packed, np.packbits(ops[i][2])), i

    put(0)
    put(1)
    get(0)   # launches 0 and 1
    put(2)   # a new batch while the result of 1 is still waiting
    put(3)
    get(3)   # launches 2 and 3: must not disturb the result of 1
    get(1)
    get(2)
    with pytest.raises(Exception):
        hw.configure_operation(ops[0][0], 0)
        hw.enqueue_operation(ops[0][1], None, 0)
        pdsch.encode_tbs(enc, [pdsch.pdsch_encoder_configuration(1, 0, 2, 0, 1, 156)], [np.zeros(8, np.uint8)])  # hal operations pending
    get(0)


def test_queue_full_and_errors(enc):
    hw = pdsch.hw_accelerator_pdsch_enc_cuda(enc, cb_mode=True)
    cfg = pdsch.hw_pdsch_encoder_configuration()
    cfg.base_graph, cfg.modulation, cfg.nof_segments, cfg.rv, cfg.lifting_size = 1, 2, 1, 0, 8
    cfg.Ncb, cfg.nof_filler_bits, cfg.rm_length, cfg.cb_mode = 66 * 8, 0, 100, 1
    data = np.zeros(22, np.uint8)
    assert not hw.enqueue_operation(data, None, enc.max_ops)  # queue full: the caller dequeues and retries
    from srsran_projectvtlmo_b200 import capi
    with pytest.raises(capi.CudaPuschDecError):
        hw.enqueue_operation(data, None, 3)  # never configured
    hw.configure_operation(cfg, 3)
    with pytest.raises(capi.CudaPuschDecError):
        hw.enqueue_operation(data[:5], None, 3)  # shorter than the configured message
    assert hw.enqueue_operation(data, None, 3)
    with pytest.raises(capi.CudaPuschDecError):
        hw.dequeue_operation(np.zeros(99, np.uint8), None, 3)  # wrong output span
    bits = np.ones(100, np.uint8)
    assert hw.dequeue_operation(bits, None, 3)
    assert not bits.any()  # the all-zero message encodes to the all-zero code word


def test_reference_hw_encoder_over_cuda_accelerator():
    """Drop-in proof on the reference's own classes (oracle/pdsch_hwacc_harness.cpp): the UNMODIFIED pdsch_encoder_hw_impl
    driving the repo's hal::hw_accelerator_pdsch_enc (CB mode, TB mode, TB mode forced into CB mode) produces the code words
    of the reference's software pdsch_encoder_impl."""
    import subprocess

    exe = Path(__file__).resolve().parent.parent / "oracle" / "_ref" / "pdsch_hwacc_parity"
    if not exe.exists():
        pytest.skip("oracle/_ref/pdsch_hwacc_parity not built (needs /root/reference at build time)")
    r = subprocess.run([str(exe), "50"], capture_output=True, text=True, timeout=600)
    print(r.stdout)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " 0 mismatches" in r.stdout
