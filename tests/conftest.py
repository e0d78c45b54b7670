import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs the reference sources at /root/reference")


def pytest_collection_modifyitems(config, items):
    have_ref = os.path.isfile(os.path.join(os.environ.get("SS_REFERENCE", "/root/reference"), "sugarscape.py"))
    skip_ref = pytest.mark.skip(reason="reference sources not present (GPU box)")
    for it in items:
        if "reference" in it.keywords and not have_ref:
            it.add_marker(skip_ref)
