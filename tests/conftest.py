"""pytest configuration: markers and shared fixture loaders."""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_sample():
    return dict(np.load(GOLDEN / "img2refmap_sample.npz"))


@pytest.fixture(scope="session")
def golden_synth():
    blob = np.load(GOLDEN / "img2refmap_synth.npz")
    cases = {}
    for key in blob.files:
        name, field = key.split("/")
        cases.setdefault(name, {})[field] = blob[key]
    return cases


@pytest.fixture(scope="session")
def golden_mirmap():
    return dict(np.load(GOLDEN / "mirmap_ref.npz"))
