import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.pyoracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reflib():
    from oracle.pyoracle import RefLib
    if not RefLib.available():
        pytest.skip("oracle/_ref/libloam_ref.so not built (reference tree absent)")
    return RefLib()


@pytest.fixture(scope="session")
def ctx():
    """A loamgpu context on cuda:0 — fails loudly (no fallback) when the device or library is missing."""
    from loam_b200 import _capi
    return _capi.Context(0)
