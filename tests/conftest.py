import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Compile the native library and the oracle once per session (no-op when up to date)."""
    import __graft_entry__ as entry
    entry.build()


def pytest_sessionfinish(session, exitstatus):
    """Dev run against a -DSMAP_DEBUG_BOUNDS build of the library (tools/bounds_check.sh): every process that loaded it
    reports the index violations its kernels counted."""
    if os.environ.get("SMAP_EXPECT_BOUNDS") != "1":
        return
    from vision_semantic_segmentation_b200 import _native
    if getattr(_native, "_lib", None) is None:
        return   # this process (e.g. the xdist controller) never touched the library
    import ctypes
    out = (ctypes.c_ulonglong * 2)()
    rc = _native._lib.smap_debug_bounds(out)
    worker = os.environ.get("PYTEST_XDIST_WORKER", "main")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "bounds_%s.txt" % worker), "w") as f:
        f.write("rc=%d violations=%d first_code=%d tests_failed=%d\n" % (rc, out[0], out[1], session.testsfailed))
    if rc != 0 or out[0] != 0:
        session.exitstatus = 1
