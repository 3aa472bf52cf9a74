import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "kmers.jl_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # a fresh checkout has no built artefacts (they are git-ignored): build them once, as __graft_entry__.build() does
    lib = os.path.join(ROOT, "kmers.jl_b200", "libkmerscuda.so")
    oracle = os.path.join(ROOT, "oracle", "libkmers_oracle.so")
    if not (os.path.exists(lib) and os.path.exists(oracle)):
        import subprocess
        jobs = str(os.cpu_count() or 4)
        if not os.path.exists(lib):
            subprocess.run(["make", "-C", os.path.join(ROOT, "kmers.jl_b200", "csrc"), "-j", jobs], check=True)
        if not os.path.exists(oracle):
            subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True)


def _has_gpu() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
