"""pytest configuration: the `gpu` marker, and import paths for the package and the oracle.

`-m "not gpu"`: oracle vs the reference's golden vectors, host logic, C-ABI symbol check (no GPU).
`-m gpu`      : parity tests proper -- CUDA path vs oracle, called through the C ABI.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def _have_gpu() -> bool:
    """Is there an NVIDIA device on this box?  Decided WITHOUT the product library, so that a
    missing/broken libqvnt_b200.so on a GPU box makes the gpu tests fail loudly, not skip."""
    if os.environ.get("QVNT_FORCE_GPU_TESTS"):
        return True
    return os.path.exists("/dev/nvidia0") or os.path.exists("/dev/nvidiactl")


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as orc
    orc.build()
    return orc
