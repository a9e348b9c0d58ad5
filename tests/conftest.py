import json
import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """Without a CUDA device the GPU tier is skipped (not errored), so a plain `pytest tests` on a CPU box shows the CPU
    tier's real result."""
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device (GPU tier: run on the B200 box with -m gpu)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def _ensure_native_built():
    """Build liblzma_b200.so / the oracle in-tree when they are missing or older than their sources (a fresh
    checkout: built artefacts are git-ignored).  nvcc cross-compiles sm_100a without a GPU."""
    import subprocess
    csrc = os.path.join(ROOT, "lzma_rs_b200", "csrc")
    lib = os.path.join(ROOT, "lzma_rs_b200", "liblzma_b200.so")
    srcs = [os.path.join(csrc, f) for f in os.listdir(csrc)] + [os.path.join(ROOT, "include", "lzma_b200.h")]
    if (not os.path.exists(lib)) or any(os.path.getmtime(f) > os.path.getmtime(lib) for f in srcs):
        subprocess.check_call(["make", "-C", csrc, "-s"])
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s"])


_ensure_native_built()


class Golden:
    """The reference's own golden vectors, committed under tests/golden/ (see make_golden.py)."""

    def __init__(self):
        with open(os.path.join(HERE, "golden", "reference_vectors.json")) as f:
            self.meta = json.load(f)["vectors"]
        with open(os.path.join(HERE, "golden", "reference_vectors.bin"), "rb") as f:
            self.blob = f.read()

    def vectors(self, fmt=None, errors=None):
        for v in self.meta:
            if fmt is not None and v["format"] != fmt:
                continue
            if errors is not None and (("error" in v) != errors):
                continue
            yield v

    def compressed(self, v):
        return self.blob[v["offset"]: v["offset"] + v["length"]]


@pytest.fixture(scope="session")
def golden():
    return Golden()
