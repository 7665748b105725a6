import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")
    # a fresh checkout has no built artefacts (they are git-ignored): build the library once, as __graft_entry__.build() does
    lib = os.path.join(ROOT, "raycore.jl_b200", "libraycore_cuda.so")
    if not os.path.exists(lib) and os.environ.get("RAYCORE_CUDA_LIB") is None:
        import subprocess

        subprocess.check_call(["make", "-C", os.path.join(ROOT, "raycore.jl_b200", "csrc"), "-j4", "../libraycore_cuda.so"], stdout=subprocess.DEVNULL)


def _have_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
