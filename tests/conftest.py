import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import pyoracle
    pyoracle.lib()
    return pyoracle


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference compiled in place (oracle/_ref); tests needing it skip when it was not built."""
    from oracle import refshim
    if not refshim.available():
        pytest.skip("oracle/_ref/libcorto_ref.so not built (needs /root/reference at build time)")
    return refshim
