import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
    """gpu-marked tests need a B200: skip them (instead of failing ~200 tests) when the library sees no sm_100 device"""
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if not gpu_items:
        return
    try:
        from ellalgo_rs_b200._lib import lib
        have = lib().ellcuda_device_count() > 0
    except Exception:
        have = False
    if not have:
        skip = pytest.mark.skip(reason="needs a B200 (libellcuda has no CPU fallback)")
        for it in gpu_items:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def ref():
    """the CPU checker (oracle/), built on demand"""
    import ellref_py
    ellref_py.build()
    return ellref_py
