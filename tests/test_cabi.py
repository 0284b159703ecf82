"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/abd_b200.h declares (no compute call is made here: there is no GPU)."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    from abdpymc_b200 import _lib, build

    build.build()
    return _lib.load()


def declared_symbols():
    text = (ROOT / "include" / "abd_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(abd_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_what_the_binding_binds(lib):
    from abdpymc_b200 import _lib

    names = declared_symbols()
    assert len(names) >= 20
    assert set(names) == set(_lib.SIGNATURES)
    for n in names:
        assert hasattr(lib, n), n


def test_version_and_error_string(lib):
    assert lib.abd_version() >= 100
    assert isinstance(lib.abd_last_error(), bytes)


def test_struct_layout_matches_header():
    from abdpymc_b200._lib import AbdCohort

    # 5 int32 (+4 pad) + 2 ptr + 2 x (int64 + 4 ptr) + 4 int64
    assert ctypes.sizeof(AbdCohort) == 24 + 16 + 2 * 40 + 32
    assert AbdCohort.pcrpos.offset == 24 and AbdCohort.n_rows_s.offset == 40
    assert AbdCohort.n_rows_n.offset == 80 and AbdCohort.total_inds.offset == 120
    assert AbdCohort.ind_offset.offset == 144


def test_no_cpu_fallback(lib, cohorts):
    """Without a CUDA device the product path must fail loudly, not fall back."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from abdpymc_b200._lib import AbdError
    from abdpymc_b200.engine import AbdEngine

    with pytest.raises(AbdError):
        AbdEngine(cohorts["test_cohort"], splits=(14, 20))


def test_product_never_imports_oracle():
    for py in (ROOT / "abdpymc_b200").rglob("*.py"):
        src = py.read_text()
        assert "oracle" not in src.replace("# oracle", ""), py


def test_transforms_roundtrip():
    from abdpymc_b200.engine import backward, forward

    rng = np.random.default_rng(0)
    q = rng.normal(size=(5, 17))
    np.testing.assert_allclose(forward(backward(q)), q, rtol=1e-12, atol=1e-12)
