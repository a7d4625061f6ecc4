import os, sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "gpu_ext: GPU test of an option outside the hot-path scope table (run separately: -m gpu_ext)")


def _has_gpu():
    return os.path.exists("/dev/nvidia0") or os.path.exists("/dev/nvidiactl")


def pytest_collection_modifyitems(config, items):
    # `-m "not gpu"` (the CPU-only run) also selects gpu_ext tests: they need a device just the same, so skip them without one
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="gpu_ext test: no CUDA device here")
    for it in items:
        if "gpu_ext" in it.keywords:
            it.add_marker(skip)
