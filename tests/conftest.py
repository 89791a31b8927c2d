import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oraclepy
    oraclepy.build()
    return oraclepy.Oracle()


@pytest.fixture(scope="session")
def oracle64():
    from oracle import oraclepy
    oraclepy.build()
    return oraclepy.Oracle(double=True)


@pytest.fixture(scope="session")
def oracle64r():
    """fp64 linear algebra on the fp32 build's covariance entries: the arbiter of what fp32 pins."""
    from oracle import oraclepy
    oraclepy.build()
    return oraclepy.Oracle(double=True, cov_float=True)


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference built into oracle/_ref (skips when it was never built)."""
    from oracle import oraclepy, refpy
    oraclepy.build()
    if not refpy.available():
        pytest.skip("oracle/_ref/libgpisref.so not built (needs /root/reference at build time)")
    return refpy
