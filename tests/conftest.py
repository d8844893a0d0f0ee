import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("reduced-3dgs_b200", "oracle", "tests", os.path.join("tests", "golden")):
    path = os.path.join(ROOT, p)
    if path not in sys.path:
        sys.path.insert(0, path)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def refC():
    """The compiled, unmodified reference (oracle/_ref/_refC.so); None when it is not available."""
    import torch
    if not torch.cuda.is_available():
        return None
    import build_ref
    return build_ref.load()
