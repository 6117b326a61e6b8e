import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def formula_table(n_entries: int, C: int, scale: float) -> np.ndarray:
    """Same closed form as tests/golden/generate_golden.py::formula_table."""
    i = np.arange(n_entries * C, dtype=np.uint64)
    u = (i * np.uint64(2654435761)) & np.uint64(0xFFFFFFFF)
    v = (u.astype(np.float64) / 4294967296.0 - 0.5) * 2.0 * scale
    return v.astype(np.float32).reshape(n_entries, C)


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name))
    return load


@pytest.fixture(scope="session")
def chest_table_unit():
    from oracle.hashgrid import level_offsets
    offs = level_offsets(16, 16, 19, 3)
    return formula_table(int(offs[-1]), 2, 1.0), offs
