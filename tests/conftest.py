import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built():
    """host + GPU libraries and the oracle are built (idempotent; seconds when up to date)"""
    import subprocess
    subprocess.check_call(["make", "-s", "all", "oracle"], cwd=ROOT, stdout=subprocess.DEVNULL,
                          stderr=subprocess.DEVNULL)
    return True


@pytest.fixture(scope="session")
def gpu(built):
    from imsame_b200 import api
    ctx = api.Imsame(0)  # raises without an sm_100 device: GPU tests must not pass on a fallback
    yield ctx
    ctx.close()
