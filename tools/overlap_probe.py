import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from convgp_b200 import _lib
from convgp_b200.hotpath import HotPath
dt = torch.float64 if (len(sys.argv) < 2 or sys.argv[1] == "f64") else torch.float32
dev = torch.device("cuda", 0)
inp = bench.make_inputs(bench.HEAD, 0)
t = {k: torch.tensor(v, dtype=dt, device=dev) for k, v in inp.items()}
spec = _lib.Spec(dt, 28, 28, 1, 5, 5, 0, 0, 1, False, False)
for ov in (False, True):
    hp = HotPath(spec, t["X"], t["Z"], t["W"], t["var"], t["ls"], t["G"], t["gd"], t["Gu"], diag_add=1e-3 + 1e-5)
    hp.capture(overlapped=ov)
    for _ in range(10): hp.step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(200): hp.step()
    b.record(); torch.cuda.synchronize()
    print("overlapped" if ov else "serial", a.elapsed_time(b) / 200, "ms/step", float(hp.grads.abs().sum()), float(hp.Kuf.sum()))
