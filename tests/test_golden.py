"""Golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py):
 - CPU: the oracle and its torch twin still reproduce them (regression pin of the checker);
 - GPU: the CUDA path reproduces values and gradients within BASELINE.json's tolerances."""
import glob
import os

import numpy as np
import pytest
import torch

from tests.golden.make_golden import golden_cases
from tests.util import rel_err

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = golden_cases()


def _load(case):
    g = dict(np.load(os.path.join(GOLDEN_DIR, case.name + ".npz")))
    case.X, case.Z, case.G, case.gd, case.Gu = g["X"], g["Z"], g["G"], g["gd"], g["Gu"]
    case.Wv = g["W"] if case.weighted else None
    return g


def test_every_case_has_a_fixture():
    assert sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))) == \
        sorted(c.name for c in CASES)


@pytest.mark.parametrize("case", CASES, ids=[c.name for c in CASES])
def test_oracle_reproduces_golden(case):
    g = _load(case)
    ok = case.oracle_kernel()
    assert rel_err(ok.Kzx(case.Z, case.X), g["kuf"]) < 1e-12
    assert rel_err(ok.Kdiag(case.X), g["kdiag"]) < 1e-12
    assert rel_err(ok.Kzz(case.Z), g["kuu"]) < 1e-12
    if case.name.startswith("cfg3_mnist_wconv"):  # gradients (autodiff) for one case keeps the CPU suite short
        ref = case.twin_grads()
        assert rel_err(ref["dZ"], g["dZ"]) < 1e-11
        assert rel_err(ref["dW"], g["dW"]) < 1e-11


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[c.name for c in CASES])
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
def test_cuda_reproduces_golden(case, dtype):
    tol = 1e-10 if dtype == torch.float64 else 1e-4
    g = _load(case)
    got = case.package_grads(dtype)
    assert rel_err(got["kuf"], g["kuf"]) < tol
    assert rel_err(got["kdiag"], g["kdiag"]) < tol
    assert rel_err(got["kuu"], g["kuu"]) < tol
    assert rel_err(got["dZ"], g["dZ"]) < tol
    if case.weighted:
        assert rel_err(got["dW"], g["dW"]) < tol
    assert rel_err(np.concatenate([np.ravel(v) for v in got["dvar"]]), g["dvar"]) < 10 * tol
    assert rel_err(np.concatenate([np.ravel(v) for v in got["dls"]]), g["dls"]) < 10 * tol
