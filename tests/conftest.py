import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "le-vins_b200"), os.path.join(ROOT, "oracle"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _have_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def oracle():
    import oracle_py

    oracle_py.build()
    return oracle_py


@pytest.fixture(scope="session")
def built_lib():
    import importlib.util

    spec = importlib.util.spec_from_file_location("levins_build", os.path.join(ROOT, "le-vins_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build()


@pytest.fixture()
def ctx(built_lib):
    """A fresh levins context on cuda:0.  GPU tests must fail loudly, not skip, when the library cannot run."""
    from levins import abi

    c = abi.Context(0)
    yield c
    c.close()
