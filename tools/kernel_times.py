"""Per-kernel CUDA-event timings of the bench workload (development aid for tuning knobs)."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from convgp_b200 import _lib
from convgp_b200.hotpath import HotPath

dt = torch.float64 if (len(sys.argv) < 2 or sys.argv[1] == "f64") else torch.float32
dev = torch.device("cuda", 0)
inp = bench.make_inputs(bench.HEAD, 0)
t = {k: torch.tensor(v, dtype=dt, device=dev) for k, v in inp.items()}
spec = _lib.Spec(dt, 28, 28, 1, 5, 5, 0, 0, 1, False, False)
hp = HotPath(spec, t["X"], t["Z"], t["W"], t["var"], t["ls"], t["G"], t["gd"], t["Gu"], diag_add=1e-3 + 1e-5)
L, s, st, p = _lib.lib(), spec, _lib.stream_ptr, _lib.ptr
fns = {
    "kuf_fwd": lambda: L.cgp_kuf_fwd(s.ref(), hp.N, hp.M, p(hp.X), p(hp.Z), p(hp.W), p(hp.variance), p(hp.lengthscales), p(hp.Kuf), None, 0, st()),
    "kdiag_fwd": lambda: L.cgp_kdiag_fwd(s.ref(), hp.N, p(hp.X), p(hp.W), p(hp.variance), p(hp.lengthscales), p(hp.Kdiag), None, 0, st()),
    "kuf_bwd": lambda: L.cgp_kuf_bwd(s.ref(), hp.N, hp.M, p(hp.X), p(hp.Z), p(hp.W), p(hp.variance), p(hp.lengthscales), p(hp.G), p(hp.dZ), p(hp.dW), p(hp.dvar), p(hp.dls), None, 0, st()),
    "kuf_bwd_dz": lambda: L.cgp_kuf_bwd(s.ref(), hp.N, hp.M, p(hp.X), p(hp.Z), p(hp.W), p(hp.variance), p(hp.lengthscales), p(hp.G), p(hp.dZ), None, p(hp.dvar), p(hp.dls), None, 0, st()),
    "kuf_bwd_dw": lambda: L.cgp_kuf_bwd(s.ref(), hp.N, hp.M, p(hp.X), p(hp.Z), p(hp.W), p(hp.variance), p(hp.lengthscales), p(hp.G), None, p(hp.dW), None, None, None, 0, st()),
    "kdiag_bwd": lambda: L.cgp_kdiag_bwd(s.ref(), hp.N, p(hp.X), p(hp.W), p(hp.variance), p(hp.lengthscales), p(hp.gd), p(hp.dW), p(hp.dvar), p(hp.dls), None, 0, st()),
}
out = {}
for name, fn in fns.items():
    ts = []
    for i in range(12):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); _lib.check(fn()); b.record(); torch.cuda.synchronize()
        if i >= 2: ts.append(a.elapsed_time(b))
    out[name] = round(statistics.mean(ts), 4)
print(os.environ.get("TAG", ""), out, "sum", round(sum(out.values()), 4))
