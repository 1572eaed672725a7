"""pytest configuration: the ``gpu`` marker and shared build fixtures."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def built():
    """Make sure the product library and the CPU checkers exist (cross-compiles without a GPU)."""
    import __graft_entry__ as g
    from webrtc_aecm_b200 import api

    import oracle_lib as ol
    if not (os.path.exists(api.library_path()) and os.path.exists(ol.ORACLE_SO)):
        g.build()
    return True
