"""Development aid: first-call-in-process check of the float64 dW pieces (fuzz0 geometry)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests.util import Case, rel_err
import oracle.torch_twin as T
from convgp_b200.kernels import as_tensor

case = Case("fuzz0", "conv", 29, 30, 1, 5, 5, [(1.3, 1.7, None, False)], weighted=True, N=1, M=100, seed=1000)
g = case.geom
def twin(which):
    X = torch.tensor(case.X); Z = torch.tensor(case.Z)
    Wt = torch.tensor(case.Wv, requires_grad=True)
    grp = [(torch.tensor(1.3, dtype=torch.float64), torch.tensor(1.7, dtype=torch.float64), None)]
    if which == "kuf":
        loss = (T.Kzx(Z, X, grp, g, Wt) * torch.tensor(case.G)).sum()
    else:
        loss = (T.Kdiag(X, grp, g, Wt) * torch.tensor(case.gd)).sum()
    loss.backward()
    return Wt.grad.numpy()
def cuda(which):
    k = case.package_kernel()
    X, Z = as_tensor(case.X, torch.float64), as_tensor(case.Z, torch.float64)
    if which == "kuf":
        loss = (k.Kzx(Z, X) * as_tensor(case.G, torch.float64)).sum()
    else:
        loss = (k.Kdiag(X) * as_tensor(case.gd, torch.float64)).sum()
    loss.backward()
    return k.W.free.grad.cpu().numpy()
order = sys.argv[1:] or ["kuf", "kdiag"]
for which in order:
    got, ref = cuda(which), twin(which)
    e = rel_err(got, ref)
    print(which, "%.2e" % e, "<-- BAD" if e > 1e-10 else "")
    if e > 1e-10:
        d = np.abs(got - ref).ravel()
        idx = np.argsort(-d)[:8]
        print("   BAD idx", idx.tolist(), "dev", ["%.1e" % v for v in d[idx]], "n>1e-13:", int((d > 1e-13).sum()),
              "of", d.size, "max|ref| %.2e" % np.abs(ref).max())
