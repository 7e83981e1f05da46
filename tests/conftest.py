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


def pytest_collection_modifyitems(config, items):
    from oracle import ref_shim
    if ref_shim.reference_available():
        return
    skip = pytest.mark.skip(reason="reference tree not present on this machine")
    for item in items:
        if "reference" in item.keywords:
            item.add_marker(skip)
