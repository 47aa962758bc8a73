import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def emu_engine():
    """Engine bound to the host-emulation build of the kernel sources (tests only)."""
    import subprocess
    from metafinishersc_b200 import _lib, engine
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "emu"), "-s"])
    return engine.Engine(_lib.bind(os.path.join(ROOT, "tests", "emu", "_build", "libmfsc_emu.so")))


@pytest.fixture(scope="session")
def gpu_engine():
    """Engine bound to the product library; fails loudly when it is not built or there is no GPU."""
    from metafinishersc_b200 import engine
    E = engine.Engine()
    assert E.device_count() >= 1, "no CUDA device visible"
    return E
