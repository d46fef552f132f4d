import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")
    config.addinivalue_line("markers", "refsrc: needs /root/reference (only present in the build container)")


def pytest_collection_modifyitems(config, items):
    have_ref = os.path.isdir("/root/reference/marlbase")
    for item in items:
        if "refsrc" in item.keywords and not have_ref:
            item.add_marker(pytest.mark.skip(reason="/root/reference is not present on this box"))
