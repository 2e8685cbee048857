import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200, sm_100a); run with `-m gpu` on the GPU box")
    # the C-ABI library is a build artefact (git-ignored): (re)build it when it is missing or older than its sources,
    # so a fresh checkout can run the suite directly (nvcc cross-compiles sm_100a without a GPU, ~25 s)
    try:
        from leftrefill_b200 import build as _b
        if _b.is_stale():
            _b.build(force=True, verbose=False)
    except Exception as e:  # noqa: BLE001 - the tests that need the library then fail with its own clear message
        print(f"[conftest] could not build liblr_b200.so: {e}")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
