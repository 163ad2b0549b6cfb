"""Times BASELINE config 1 (592 full-rate BG1 / Z = 384 code blocks, 46 layers, 6 iterations, no CRC) through the unit-level
batch interface for a given decoder variant: kernel span on the device and wall time from host buffers."""
import ctypes as C
import sys
import time

import numpy as np

from srsran_projectvtlmo_b200 import capi, pusch


def main():
    variant = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 592
    acc = pusch.Accelerator(device=0, max_cbs_in_flight=64 * 152, nof_harq_cb_slots=64 * 152)
    acc.set_decoder_variant(variant)
    mt = np.random.RandomState(0)
    llr = ((mt.randint(0, 2 ** 32, (n, 25344), dtype=np.uint64) & 1) * 20 - 10).astype(np.int8)
    bits = np.zeros((n, 1056), np.uint8)
    its = np.zeros(n, np.int32)

    def go():
        st = acc._lib.srsran_cuda_ldpc_decode_batch(acc.h, bits.ctypes.data_as(capi.u8p), llr.ctypes.data_as(capi.i8p), n, 25344,
                                                    1, 384, 0, 0, 6, C.c_float(0.8), its.ctypes.data_as(capi.intp))
        assert st == 0
    for _ in range(9):
        go()
    stage = np.zeros(5)
    t0 = time.perf_counter()
    for _ in range(5):
        go()
        stage += np.array(pusch.last_unit_timing(acc))
    dt = (time.perf_counter() - t0) / 5
    stage /= 5
    print(f"variant {variant} n {n}: stages ms {np.round(stage, 4).tolist()} wall ms {dt * 1e3:.3f} "
          f"edge updates/s (kernel) {n * 6 * 384 * 316 / (stage[2] * 1e-3):.3e} checksum {int(bits.sum())}")
    acc.close()


if __name__ == "__main__":
    main()
