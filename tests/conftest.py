import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def ref_lib():
    """The unmodified reference (oracle/_ref/libspade_ref.so); prebuilt in the dev container."""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref/libspade_ref.so not built (needs /root/reference)")
    return ref
