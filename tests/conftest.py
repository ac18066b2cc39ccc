import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    return pyoracle.Oracle()


@pytest.fixture(scope="session")
def ref():
    from oracle import pyoracle
    if not pyoracle.ref_available():
        pytest.skip("oracle/_ref/libquaffref.so not built here (needs /root/reference)")
    return pyoracle.Ref()


@pytest.fixture(scope="session")
def emu_lib():
    """TEST ONLY: the kernels compiled against the CPU-thread shim (tests/emu)."""
    import build_emu
    return build_emu.build()


@pytest.fixture(scope="session")
def gpu():
    """The product library on a real device.  No silent fallback: a missing library or device fails the test."""
    import torch
    from quaff_b200 import api
    assert torch.cuda.is_available(), "gpu-marked test run without a CUDA device"
    g = api.QuaffGPU(device=0)
    yield g
    g.close()
