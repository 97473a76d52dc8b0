import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_device_count() -> int:
    """Number of CUDA devices the driver reports (0 on a CPU-only box); no torch import needed."""
    import ctypes

    for name in ("libcudart.so", "libcudart.so.12", "/usr/local/cuda/lib64/libcudart.so"):
        try:
            rt = ctypes.CDLL(name)
        except OSError:
            continue
        n = ctypes.c_int(0)
        return n.value if rt.cudaGetDeviceCount(ctypes.byref(n)) == 0 else 0
    return 0


def pytest_collection_modifyitems(config, items):
    """A plain `pytest` on a CPU box skips the gpu-marked tests instead of failing in zfvm_create."""
    if not any("gpu" in item.keywords for item in items):
        return
    if _cuda_device_count() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device (the B200 path has no CPU fallback); run with -m gpu on the GPU box")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built_libraries():
    """Build the product library and the oracle once per session (no-ops when up to date)."""
    import __graft_entry__ as ge

    ge.build(quiet=True)
    yield
