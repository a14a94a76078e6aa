import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def _have_gpu():
    try:
        from isce3_b200 import _capi
        return _capi.load_library().i3b_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracles():
    """(port, ref-or-None): builds the port on demand; oracle/_ref only exists where
    /root/reference was mounted at build time (it travels as a prebuilt .so)."""
    from oracle import tdbp
    if not tdbp.PORT_LIB.exists() or (Path("/root/reference").exists() and not tdbp.have_ref()):
        tdbp.build()
    return tdbp.port(), (tdbp.ref() if tdbp.have_ref() else None)


@pytest.fixture(scope="session")
def oracle(oracles):
    port, ref = oracles
    return ref if ref is not None else port


@pytest.fixture(scope="session", autouse=True)
def _device_buffer_cache():
    """The GPU tests opt in to the library's device-buffer cache (hundreds of small calls);
    the default -- everything freed before a call returns -- has its own test."""
    if _have_gpu():
        from isce3_b200.focus import keep_device_memory
        keep_device_memory(-1)
    yield
