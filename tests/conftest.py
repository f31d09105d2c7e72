import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.lib()  # builds on demand
    return oracle


@pytest.fixture(scope="session")
def b2p():
    """The product's C ABI binding; skips (never falls back) when no GPU is present."""
    from paf_baseband2power_b200 import api
    if api.device_count() < 1:
        pytest.skip("no CUDA device")
    return api
