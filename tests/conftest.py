import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# The oracle steps are small (bs <= 64 on the CPU): a handful of threads is as fast as all of them, and the suite stays
# quick when the cores are shared with other jobs (8 OpenMP threads spinning against another process made single
# tests take a minute).
try:
    import torch
    torch.set_num_threads(min(4, os.cpu_count() or 1))
except Exception:  # pragma: no cover
    pass


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
