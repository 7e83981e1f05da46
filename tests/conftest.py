import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs the read-only reference tree at /root/reference")


def pytest_sessionstart(session):
    """A fresh checkout has no libvarannealb200.so (it is git-ignored): build it once if nvcc is
    here.  On the GPU box the prebuilt library travels with the snapshot and nothing happens."""
    from varanneal_b200 import _lib, build as _build
    if not os.path.exists(_lib.LIB_PATH) and os.path.exists(_build.NVCC):
        _build.build()


def pytest_collection_modifyitems(config, items):
    from oracle import ref_shim
    if ref_shim.reference_available():
        return
    skip = pytest.mark.skip(reason="reference tree not present on this machine")
    for item in items:
        if "reference" in item.keywords:
            item.add_marker(skip)
