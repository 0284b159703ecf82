import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = ROOT / "tests" / "golden"
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def kats():
    return json.loads((GOLDEN / "reference_kats.json").read_text())


@pytest.fixture(scope="session")
def goldens():
    z = np.load(GOLDEN / "model_goldens.npz")
    cases = json.loads(str(z["cases"]))
    return z, cases


@pytest.fixture(scope="session")
def cohorts():
    from abdpymc_b200.cohort import CohortArrays

    return {name: CohortArrays.load(name) for name in ("cohort", "test_cohort")}
