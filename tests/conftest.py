import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def gold():
    def load(name):
        return np.load(os.path.join(GOLD, name))
    return load


@pytest.fixture(scope="session")
def scaler2021():
    z = np.load(os.path.join(GOLD, "scaler_DCASE2021.npz"))
    return {"MEL": {k: z[f"MEL_{k}"] for k in ("mean", "std", "max", "min")},
            "IV": {k: z[f"IV_{k}"] for k in ("mean", "std", "max", "min")}}
