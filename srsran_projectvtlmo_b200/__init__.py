"""B200-native (sm_100a) PUSCH channel-decoding path for srsRAN Project's du_low.

Rate dematching with HARQ soft combining, layered normalized min-sum LDPC decoding (BG1/BG2, 51 lifting sizes) and the
code-block / transport-block CRC check, as hand-written CUDA kernels behind a C ABI (include/srsran_cuda_pusch_dec.h).
The compute path is the shared library libsrsran_cuda_pusch_dec.so; there is no CPU fallback.
"""
from . import capi  # noqa: F401

__all__ = ["capi"]
