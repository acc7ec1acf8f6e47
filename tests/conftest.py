import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


def pytest_collection_modifyitems(config, items):
    from oracle import ref_loader
    if not ref_loader.available():
        skip = pytest.mark.skip(reason="/root/reference not present (GPU box)")
        for it in items:
            if "reference" in it.keywords:
                it.add_marker(skip)


@pytest.fixture(autouse=True, scope="session")
def _gemm_engine_from_env():
    """EAGCN_GEMM=tcgen05|ffma|tcgen05-nt selects the projection GEMM engine for a GPU test session."""
    eng = os.environ.get("EAGCN_GEMM")
    if eng:
        import torch
        if torch.cuda.is_available():
            from eagcn_b200 import functional as EF
            EF.set_gemm_engine(eng)
    yield
