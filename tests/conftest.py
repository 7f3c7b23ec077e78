import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def ctx():
    if not _have_gpu():
        pytest.skip("no CUDA device")
    import ppo_b200
    c = ppo_b200.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session", autouse=True)
def _build_oracle():
    from oracle import c_oracle
    c_oracle.build()
