import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu():
    if os.environ.get("PUSCH_DEC_FORCE_NO_GPU"):
        return False
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def acc():
    """One accelerator handle (GPU 0) shared by the GPU tests. Fails loudly if the CUDA library is missing."""
    from srsran_projectvtlmo_b200 import pusch

    a = pusch.Accelerator(device=0, max_cbs_in_flight=4096, nof_harq_cb_slots=8192)
    yield a
    a.close()


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device here")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
