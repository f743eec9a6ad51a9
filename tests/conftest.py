import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")
    config.addinivalue_line("markers", "ref: needs oracle/_ref/libcgref.so (the compiled reference)")


def _have_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a GPU must fail loudly, not skip: the product has no CPU fallback.
    from oracle import binding as ob

    for item in items:
        if "ref" in item.keywords and not ob.have_ref():
            item.add_marker(pytest.mark.skip(reason="oracle/_ref/libcgref.so not built (no /root/reference here)"))


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import binding as ob

    ob.build()
    return ob


@pytest.fixture(scope="session")
def gpu():
    """The product library; raises if it is missing or no GPU is visible."""
    import cgraytracing_b200 as cg

    cg.load_library()
    return cg
