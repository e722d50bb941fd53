"""pytest wiring: marker registration and import paths.

`-m "not gpu"` (run in the build container, no GPU) covers the oracle against the golden vectors,
the host logic and the C-ABI symbol table; `-m gpu` (run on a B200) holds the parity tests proper.
"""
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
PKG_PARENT = ROOT / "vision-based-spatio-temporal-analysis_b200"
for p in (str(ROOT), str(PKG_PARENT)):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN_DIR = ROOT / "tests" / "golden"
GOLDEN_CASES = ["rig_small", "seven_views_3x4", "degenerate", "w_guard", "non_finite"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        with np.load(GOLDEN_DIR / f"{name}.npz") as z:
            return {k: z[k] for k in z.files}
    return load
