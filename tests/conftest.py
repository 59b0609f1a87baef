import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the C-ABI library and the oracle exist (both are built by __graft_entry__.build())."""
    so = os.path.join(ROOT, "nettracer_b200", "libnettracer_b200.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "nettracer_b200", "csrc"), "-j8"],
                              stdout=subprocess.DEVNULL)
    from oracle import oracle
    oracle.build()
    yield
