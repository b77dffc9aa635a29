"""pytest configuration: the `gpu` marker and the import roots.

Import roots mirror the reference (its tests run with PYTHONPATH=<repo>:<repo>/src, SURVEY.md
section 4): the package directory plays the role of the reference's ``src/`` so that
``from kernels import get_verify_prefix`` resolves to this repo's drop-in.
"""
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "llm-inference-lab_b200"
for p in (str(PKG), str(ROOT)):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        has_cuda = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_cuda = False
    if has_cuda:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this environment")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    import os

    # SD_GOLDEN_DIR: vectors regenerated from the live reference with another seed (tests/test_oracle_live_reference.py)
    return Path(os.environ.get("SD_GOLDEN_DIR", ROOT / "tests" / "golden"))
