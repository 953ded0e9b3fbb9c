import os
import sys

import pytest

# Several virtual ranks of a sharded state share ONE process and GPU in the
# tests (quantum_b200/sharded.py emulated_*): each rank's stream must get its
# own hardware queue, or a kernel of one rank can sit behind another rank's
# spinning peer-wait kernel in a shared queue (false dependency) until the wait
# times out.  Must be set before CUDA initialises; one process per GPU (the
# deployment) never needs it.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
