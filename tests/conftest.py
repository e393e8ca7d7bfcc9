import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    if have_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def launcher_lib():
    from absinthe_b200.build import build_launcher
    build_launcher()
    from absinthe_b200.launcher import load_library
    return load_library()


@pytest.fixture(scope="session", autouse=True)
def warm_variant_cache(request):
    """On a GPU box with a cold cache, compile the variants of the parity suite in parallel once
    (a cubin takes seconds; the suite uses ~150 of them)."""
    try:
        import torch
        if not torch.cuda.is_available():
            return
    except Exception:
        return
    if not any("gpu" in item.keywords for item in request.session.items):
        return
    cache = os.path.join(ROOT, "build", "variants")
    have = len([f for f in os.listdir(cache) if f.endswith(".cubin")]) if os.path.isdir(cache) else 0
    if have >= 100:
        return
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    try:
        import prebuild_tests
        prebuild_tests.main()
    except Exception as exc:          # the tests build on demand anyway
        print("variant prebuild skipped:", exc)
