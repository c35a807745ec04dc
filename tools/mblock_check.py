import sys; sys.path.insert(0, '/root/repo')
import numpy as np, torch
from convgp_b200 import _lib, convkernels as ck, kernels as bk
from convgp_b200.settings import settings
res = {}
for dt in (torch.float64, torch.float32):
    settings.float_type = dt; settings.device = torch.device("cuda", 0)
    k = ck.WeightedConv(bk.RBF(25), [28, 28], [5, 5])
    r = np.random.RandomState(0)
    k.W = r.randn(576)
    X = torch.tensor(r.rand(9, 784), dtype=dt, device="cuda")
    Z = torch.tensor(r.rand(2500, 25), dtype=dt, device="cuda")
    print(dt, [_lib.lib().cgp_kernel_family(k._spec(dt).ref(), 9, 2500, w).decode() for w in range(3)])
    res[dt] = k.Kzx(Z, X).detach().double().cpu().numpy()
a, b = res[torch.float32], res[torch.float64]
print("M=2500 Kuf f32 vs f64:", np.abs(a - b).max() / np.abs(b).max(), np.isfinite(a).all())
