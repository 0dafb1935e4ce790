import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
# The B200_TEST_* / tuning variables that force kernel variants and tiny capacities are honoured by the library only
# behind this gate (read once per jnibwa_openIndex: engine_impl.inl EngineConfig).
os.environ.setdefault("B200_TEST_HOOKS", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: larger datasets")


@pytest.fixture(scope="session")
def have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
