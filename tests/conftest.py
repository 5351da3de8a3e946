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
    from oracle.oracle import Oracle

    path = os.path.join(ROOT, "oracle", "libdvbs2_oracle.so")
    if not os.path.exists(path):
        import subprocess

        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "libdvbs2_oracle.so"])
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    from oracle.oracle import Ref

    if not Ref.available():
        pytest.skip("compiled reference (oracle/_ref) not present on this host")
    r = Ref()
    yield r
    r.close()
