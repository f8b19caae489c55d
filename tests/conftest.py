import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def readme_obs():
    import oracle
    return oracle.readme_observations()


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def ctx():
    """A libcge context on cuda:0.  No skip-if-missing: GPU tests must fail loudly when the
    CUDA library or the device is unavailable."""
    import cuda_grid_estimation_b200 as cge
    c = cge.Context(0)
    yield c
    c.close()
