import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built():
    import __graft_entry__ as g

    g.build_oracle()
    return g


@pytest.fixture(scope="session")
def emu(built):
    """ctypes handle on the CPU emulation of the kernel source (test infrastructure)."""
    from tests.emu_runner import EmuRunner

    return EmuRunner(built.build_emu())


@pytest.fixture(scope="session")
def gpu_pkg(built):
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    built.build_cuda()
    import multimodal_isic_b200 as pkg

    return pkg
