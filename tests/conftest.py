import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def paper_objective(point):
    """2-D test function of Hadida et al. 2018 used by the reference's end-to-end test (tests/test_optimisation.py:27-42)."""
    import numpy as np

    x, y = point
    ct, st = np.cos(np.pi / 4), np.sin(np.pi / 4)
    x, y = ct * x + st * y, ct * y - st * x
    return (
        3 * (1 - x) ** 2.0 * np.exp(-(x ** 2) - (y + 1) ** 2)
        - 10 * (x / 5.0 - x ** 3 - y ** 5) * np.exp(-(x ** 2) - y ** 2)
        - 1 / 3 * np.exp(-((x + 1) ** 2) - y ** 2)
    )


@pytest.fixture
def oracle_backend():
    from tests.oracle_backend import OracleBackend

    return OracleBackend()
