import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


def pytest_sessionfinish(session, exitstatus):
    """With a bounds-checking build of the library (DYNENV_B200_LIB=tools/tmp/libbounds.so, tools/gpu_bounds.sh) report
    the shared-memory accesses that left their block's allocation during the whole session; any is a failure."""
    lib_path = os.environ.get("DYNENV_B200_LIB", "")
    if "bounds" not in os.path.basename(lib_path):
        return
    import ctypes as C
    from dynenv_b200 import _lib
    lb = _lib.lib()
    v = (C.c_ulonglong * 4)()
    rc = lb.dynenv_debug_bounds_violations(v)
    print("\n[bounds check] rc=%d violations=%d first word=%d block=%d thread=%d" % (rc, v[0], C.c_longlong(v[1]).value, v[2], v[3]))
    if rc != 0 or v[0] != 0:
        session.exitstatus = 1
    lb.dynenv_debug_bounds_selftest()  # the checker itself: one index behind a 64-byte allocation
    rc = lb.dynenv_debug_bounds_violations(v)
    print("[bounds check] self-test: rc=%d, %d new violation(s), word %d (expected 1, word 1000)" % (rc, v[0], C.c_longlong(v[1]).value))
