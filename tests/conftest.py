import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # a clean checkout has no built library (it is git-ignored): build it once, loudly, before collection
    if not os.path.exists(os.path.join(ROOT, "rebop_b200", "librebop_b200.so")):
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "rebop_b200", "csrc"), "-j8"])


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure), built on demand with gcc."""
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def ffi():
    from rebop_b200 import _ffi
    return _ffi


@pytest.fixture(scope="session")
def gpu(ffi):
    if ffi.device_count() == 0:
        pytest.fail("no CUDA device visible: -m gpu tests must run on the GPU box")
    return 0
