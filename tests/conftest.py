import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def have_ref():
    from oracle import refsw
    return refsw.available()


def have_host_lib():
    from skity_b200 import hostlib
    return os.path.exists(hostlib.LIB_PATH)


requires_ref = pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libskity_ref.so")),
                                  reason="compiled reference (oracle/_ref) not built here")
