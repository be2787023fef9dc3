"""Shared fixtures.  `-m "not gpu"` runs here on CPU; `-m gpu` needs a B200 (driver / gpurun).

Only this directory (plus __graft_entry__.smoke() and bench.py's CPU legs) may use anything under
oracle/: the oracle is the checker, never the thing tested for speed or shipped.
"""
import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def xs():
    """The product package (directory name is not an identifier)."""
    mod = importlib.import_module("libxsmm-1_b200")
    from importlib import import_module
    build = import_module("libxsmm-1_b200.build")
    build.build()
    mod.load()
    return mod


@pytest.fixture(scope="session")
def oracle():
    import pyoracle
    pyoracle.build_oracle()
    return pyoracle.Oracle()


@pytest.fixture(scope="session")
def ref():
    """The compiled UNMODIFIED reference (oracle/_ref); tests that need it skip where it is absent."""
    import pyoracle
    if not pyoracle.Ref.available() and os.path.isdir("/root/reference/src"):
        pyoracle.build_ref("avx2")
    if not pyoracle.Ref.available():
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    return pyoracle.Ref()


@pytest.fixture(scope="session")
def gpu(xs):
    if xs.device_count() < 1:
        pytest.fail("this test is marked gpu but no CUDA device is visible (the product has no CPU path)")
    xs.clear_error()
    return xs
