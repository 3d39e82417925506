import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
    """GPU tests must never silently pass on a box without a GPU."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def toy_inputs():
    """{sample: bytes} of the reference's toy_example inputs (committed xz fixtures)."""
    import lzma
    out = {}
    d = os.path.join(GOLDEN, "fna")
    for f in sorted(os.listdir(d)):
        if f.endswith(".fna.xz"):
            out[f[:-7]] = lzma.decompress(open(os.path.join(d, f), "rb").read())
    return out


@pytest.fixture(scope="session")
def toy_golden_kf():
    """{sample: text} of the reference's committed .kf outputs."""
    import gzip
    out = {}
    d = os.path.join(GOLDEN, "kf")
    for f in sorted(os.listdir(d)):
        if f.endswith(".kf.gz"):
            out[f[:-6]] = gzip.open(os.path.join(d, f), "rb").read().decode()
    return out
