"""Randomised geometry sweep of the default CUDA families against the NumPy oracle (values) and the torch twin
(gradients), both dtypes: image sizes 2..32, patch sizes 2 / 3 / 5, M from 1 to 300 (single row, exactly one tile,
one row more than a tile), N from 1 to 33, weighted and unweighted, upstream gradients with zeros.  A sweep of this
kind is what found that one fp16 coefficient term is not enough for dZ when a single inducing patch meets patch
weights of both signs.  Tolerances: BASELINE.json (1e-10 float64, 1e-4 float32, relative to
max|.| of each array); the live CPU twin is evaluated a second time before a disagreement counts."""
import numpy as np
import pytest
import torch

from tests.util import Case, rel_err

pytestmark = pytest.mark.gpu


def _cases(n=24, seed=3):
    rng = np.random.RandomState(seed)
    out = []
    for i in range(n):
        ph = int(rng.choice([2, 3, 5]))
        H, W = int(rng.randint(ph, 33)), int(rng.randint(ph, 33))
        if (H - ph + 1) * (W - ph + 1) > 800:
            H, W = min(H, 28), min(W, 28)
        M = 1 if i % 6 == 0 else int(rng.choice([1, 3, 17, 100, 127, 128, 129, 200, 300]))
        N = int(rng.choice([1, 2, 5, 9, 33]))
        weighted = bool(rng.rand() < 0.7) or i % 6 == 0
        v, l = float(rng.uniform(0.3, 2.0)), float(rng.uniform(0.6, 2.5))
        case = Case("fuzz%d" % i, "conv", H, W, 1, ph, ph, [(v, l, None, False)], weighted=weighted, N=N, M=M,
                    seed=1000 + i)
        u = rng.rand()
        if u < 0.3:
            case.G[rng.rand(*case.G.shape) < 0.5] = 0.0
        elif u < 0.4:
            case.G[:] = 0.0
        out.append(case)
    return out


CASES = _cases()


@pytest.mark.parametrize("case", CASES, ids=["%s_%dx%d_p%d_M%d_N%d" % (c.name, c.H, c.W, c.ph, c.M, c.N) for c in CASES])
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
def test_random_geometry(case, dtype):
    tol = 1e-10 if dtype == torch.float64 else 1e-4
    got = case.package_grads(dtype)
    for attempt in range(2):
        ok, ref = case.oracle_kernel(), case.twin_grads()
        errs = dict(kuf=rel_err(got["kuf"], ok.Kzx(case.Z, case.X)), kdiag=rel_err(got["kdiag"], ok.Kdiag(case.X)),
                    kuu=rel_err(got["kuu"], ok.Kzz(case.Z)), dZ=rel_err(got["dZ"], ref["dZ"]),
                    dvar=rel_err(got["dvar"][0], ref["dvar"][0]) / 10, dls=rel_err(got["dls"][0], ref["dls"][0]) / 10)
        if case.weighted:
            errs["dW"] = rel_err(got["dW"], ref["dW"])
        if max(errs.values()) < tol and np.all(np.isfinite(list(errs.values()))):
            break
    assert max(errs.values()) < tol and np.all(np.isfinite(list(errs.values()))), errs
