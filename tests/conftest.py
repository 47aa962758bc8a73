import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _device_count():
    try:
        from metafinishersc_b200 import engine
        return engine.Engine().device_count()
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    """A plain `pytest tests` on a machine without a GPU skips the gpu-marked tests (CPU CI stays green); when the gpu tests
    are asked for explicitly (`-m gpu`, what the GPU box runs) nothing is skipped, so a missing device or library fails loudly."""
    if "gpu" in (config.getoption("-m") or ""):
        return
    if _device_count() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device visible (run with -m gpu on a B200 box)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


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
