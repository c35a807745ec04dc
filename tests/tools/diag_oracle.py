import numpy as np, torch, sys, os
sys.path.insert(0, os.getcwd())
from tests.util import Case, rel_err
import platform
print(platform.processor(), os.cpu_count(), torch.__config__.parallel_info().split("\n")[0:3])
os.system("lscpu | grep 'Model name'")
case = Case("mnist5x5", "conv", 28, 28, 1, 5, 5, [(0.8, 1.7, None, False)], N=5, M=40)
ok = case.oracle_kernel()
ref = case.twin_grads()
a = ok.Kzx(case.Z, case.X)
print("twin vs numpy", rel_err(ref["kuf"], a))
# long double direct form
Xp = ok._get_patches(case.X).reshape(-1, 25).astype(np.longdouble)
Z = case.Z.astype(np.longdouble)
d = ((Z[:, None, :] - Xp[None, :, :]) ** 2).sum(-1) / np.longdouble(1.7) ** 2
big = np.longdouble(0.8) * np.exp(-d / 2)
K = (big.reshape(40, 5, 576) * case.Wv.astype(np.longdouble)).sum(2) / 576
K = K.astype(np.float64)
print("numpy vs longdouble", rel_err(a, K), " twin vs longdouble", rel_err(ref["kuf"], K))
if torch.cuda.is_available():
    got = case.package_grads(torch.float64)
    print("cuda vs longdouble", rel_err(got["kuf"], K), "cuda vs twin", rel_err(got["kuf"], ref["kuf"]))
