"""Development aid: where does the float32 path differ from the float64 path at full size?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from tests.test_gpu_parity import full_size_case
from tests.util import rel_err
case = full_size_case()
for rep in range(2):
    g32, g64 = case.package_grads(torch.float32), case.package_grads(torch.float64)
    for key in ("kuf", "kdiag", "kuu", "dZ", "dW"):
        a, b = np.asarray(g32[key], np.float64), np.asarray(g64[key], np.float64)
        err = np.abs(a - b) / np.max(np.abs(b))
        bad = np.argwhere(err > 1e-4)
        print(rep, key, "max rel err %.3g" % err.max(), "bad entries:", len(bad), bad[:12].tolist())
    print(rep, "dvar", g32["dvar"][0], g64["dvar"][0], "dls", g32["dls"][0], g64["dls"][0])
    kd32, kd64 = np.asarray(g32["kdiag"], np.float64), np.asarray(g64["kdiag"], np.float64)
    bad = np.argwhere(np.abs(kd32 - kd64) / np.max(np.abs(kd64)) > 1e-4).ravel()
    print("kdiag bad images:", bad.tolist(), "relative deficits * 576:", ((kd64[bad] - kd32[bad]) / kd64[bad] * 576).round(2).tolist())
