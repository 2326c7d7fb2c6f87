import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ctx():
    """One CUDA context for the whole GPU session (2^16-row workspace, 2 slots)."""
    import xfg_stark_b200 as xs
    c = xs.Context(device=0, max_n_log2=16, num_slots=2)
    yield c
    c.close()
