import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # Forced kernel families for the variant tests (tests/test_gpu_variants.py): "name=value,name=value" is applied
    # through the library's explicit option API (cgp_set_option) -- the library itself reads no environment.
    forced = os.environ.get("CGP_TEST_OPTIONS", "")
    if forced:
        from convgp_b200 import _lib
        for item in forced.split(","):
            name, value = item.split("=")
            _lib.set_option(name.strip(), int(value))


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
