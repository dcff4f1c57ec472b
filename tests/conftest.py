import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import oracle as orc
    orc.build()
    return orc


def pytest_collection_modifyitems(config, items):
    """Plain `pytest tests` on a box without a CUDA device (or without the built library) skips the gpu-marked tests instead of
    erroring in them; `-m gpu` on the GPU box is unaffected. The product itself never falls back: it raises."""
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    have_lib = os.path.exists(os.path.join(ROOT, "abcsmc_b200", "libabcsmc_b200.so"))
    if have_gpu and have_lib:
        return
    why = "no CUDA device" if not have_gpu else "abcsmc_b200/libabcsmc_b200.so is not built"
    skip = pytest.mark.skip(reason=f"gpu test: {why}")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
