"""Development aid: randomised geometry sweep of the default CUDA families against the NumPy oracle / torch twin
(values and gradients, both dtypes).  python tests/tools/fuzz_shapes.py [n_cases] [seed]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests.util import Case, rel_err  # noqa: E402

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.RandomState(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
bad = 0
for i in range(n_cases):
    ph = int(rng.choice([2, 3, 5]))
    H, W = int(rng.randint(ph, 33)), int(rng.randint(ph, 33))
    if (H - ph + 1) * (W - ph + 1) > 800:
        H = min(H, 28)
        W = min(W, 28)
    M = int(rng.choice([1, 3, 17, 100, 127, 128, 129, 200, 300]))
    N = int(rng.choice([1, 2, 5, 9, 33]))
    weighted = bool(rng.rand() < 0.7)
    v, l = float(rng.uniform(0.3, 2.0)), float(rng.uniform(0.6, 2.5))
    case = Case("fuzz%d" % i, "conv", H, W, 1, ph, ph, [(v, l, None, False)], weighted=weighted, N=N, M=M,
                seed=1000 + i)
    if rng.rand() < 0.3:
        case.G[rng.rand(*case.G.shape) < 0.5] = 0.0
    if rng.rand() < 0.15:
        case.G[:] = 0.0
    ok = case.oracle_kernel()
    ref = case.twin_grads()
    for dtype, tol in ((torch.float64, 1e-10), (torch.float32, 1e-4)):
        got = case.package_grads(dtype)
        errs = dict(kuf=rel_err(got["kuf"], ok.Kzx(case.Z, case.X)), kdiag=rel_err(got["kdiag"], ok.Kdiag(case.X)),
                    kuu=rel_err(got["kuu"], ok.Kzz(case.Z)), dZ=rel_err(got["dZ"], ref["dZ"]),
                    dvar=rel_err(got["dvar"][0], ref["dvar"][0]) / 10, dls=rel_err(got["dls"][0], ref["dls"][0]) / 10)
        if weighted:
            errs["dW"] = rel_err(got["dW"], ref["dW"])
        worst = max(errs, key=errs.get)
        flag = "" if errs[worst] < tol and all(np.isfinite(list(errs.values()))) else "   <-- FAIL"
        bad += bool(flag)
        print("%-7s H%2d W%2d p%d M%3d N%2d w%d %s worst %-5s %.2e%s" % (
            case.name, H, W, ph, M, N, weighted, "f64" if dtype == torch.float64 else "f32", worst, errs[worst], flag))
print("failures:", bad)
sys.exit(1 if bad else 0)
