import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # the product has no CPU fallback: make sure the CUDA library exists before anything imports it
    import importlib.util
    spec = importlib.util.spec_from_file_location("_pydvs_build", os.path.join(ROOT, "pydvs_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.ensure_built()   # digest check: also rebuilds a library that is stale relative to csrc/


def _has_gpu():
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


@pytest.fixture(scope="session")
def oracle_coded():
    from oracle.oracle import Oracle
    return Oracle(uncoded=False)


@pytest.fixture(scope="session")
def oracle_uncoded():
    from oracle.oracle import Oracle
    return Oracle(uncoded=True)


@pytest.fixture(scope="session")
def reference_modules():
    """The compiled, unmodified reference (oracle/_ref) or None when it has not been built."""
    from oracle import ref_loader
    return ref_loader.load()
