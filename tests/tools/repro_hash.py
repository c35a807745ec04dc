"""Development aid: hash the CPU twin's and the CUDA path's float64 dW (Kuf only) in a fresh process."""
import hashlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests.util import Case
import oracle.torch_twin as T
from convgp_b200.kernels import as_tensor
case = Case("fuzz0", "conv", 29, 30, 1, 5, 5, [(1.3, 1.7, None, False)], weighted=True, N=1, M=100, seed=1000)
g = case.geom
X = torch.tensor(case.X); Z = torch.tensor(case.Z)
Wt = torch.tensor(case.Wv, requires_grad=True)
grp = [(torch.tensor(1.3, dtype=torch.float64), torch.tensor(1.7, dtype=torch.float64), None)]
(T.Kzx(Z, X, grp, g, Wt) * torch.tensor(case.G)).sum().backward()
ref = Wt.grad.numpy()
k = case.package_kernel()
Xd, Zd = as_tensor(case.X, torch.float64), as_tensor(case.Z, torch.float64)
kz = k.Kzx(Zd, Xd)
(kz * as_tensor(case.G, torch.float64)).sum().backward()
got = k.W.free.grad.cpu().numpy()
var = k.basekern.variance().detach().cpu().numpy(); ls = k.basekern.lengthscales().detach().cpu().numpy()
print("twin", hashlib.md5(ref.tobytes()).hexdigest()[:8], "cuda", hashlib.md5(np.round(got, 14).tobytes()).hexdigest()[:8],
      "kzx", hashlib.md5(np.round(kz.detach().cpu().numpy(), 14).tobytes()).hexdigest()[:8],
      "var %.17g ls %.17g" % (float(var.ravel()[0]), float(ls.ravel()[0])), "err %.2e" % (np.abs(got - ref).max() / np.abs(ref).max()))
