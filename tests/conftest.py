import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def vp():
    """The product package with libvp_engine.so loaded (built if stale and nvcc is present)."""
    import vocoderproject_b200 as pkg
    from vocoderproject_b200 import build as b
    try:
        b.build()
    except Exception:
        pass
    pkg.load_library()
    return pkg


@pytest.fixture(scope="session")
def oracle():
    import oraclebind
    oraclebind.load()
    return oraclebind
