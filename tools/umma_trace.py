"""Development aid: per-unit timeline of CTA 0 of the UMMA kernels (needs a CGP_UMMA_TRACE build:
CGP_NVCC_EXTRA=-DCGP_UMMA_TRACE python convgp_b200/build.py --force)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from convgp_b200 import _lib
from convgp_b200.hotpath import HotPath

dt = torch.float32
dev = torch.device("cuda", 0)
inp = bench.make_inputs(bench.HEAD, 0)
t = {k: torch.tensor(v, dtype=dt, device=dev) for k, v in inp.items()}
spec = _lib.Spec(dt, 28, 28, 1, 5, 5, 0, 0, 1, False, False)
hp = HotPath(spec, t["X"], t["Z"], t["W"], t["var"], t["ls"], t["G"], t["gd"], t["Gu"], diag_add=1e-3 + 1e-5)
L, s, st, p = _lib.lib(), spec, _lib.stream_ptr, _lib.ptr
which = sys.argv[1] if len(sys.argv) > 1 else "kuf_fwd"
fns = {
    "kuf_fwd": lambda: L.cgp_kuf_fwd(s.ref(), hp.N, hp.M, p(hp.X), p(hp.Z), p(hp.W), p(hp.variance), p(hp.lengthscales), p(hp.Kuf), None, 0, st()),
    "kdiag_fwd": lambda: L.cgp_kdiag_fwd(s.ref(), hp.N, p(hp.X), p(hp.W), p(hp.variance), p(hp.lengthscales), p(hp.Kdiag), None, 0, st()),
    "kuf_bwd": lambda: L.cgp_kuf_bwd(s.ref(), hp.N, hp.M, p(hp.X), p(hp.Z), p(hp.W), p(hp.variance), p(hp.lengthscales), p(hp.G), p(hp.dZ), p(hp.dW), p(hp.dvar), p(hp.dls), None, 0, st()),
    "kuf_bwd_dz": lambda: L.cgp_kuf_bwd(s.ref(), hp.N, hp.M, p(hp.X), p(hp.Z), p(hp.W), p(hp.variance), p(hp.lengthscales), p(hp.G), p(hp.dZ), None, p(hp.dvar), p(hp.dls), None, 0, st()),
    "kdiag_bwd": lambda: L.cgp_kdiag_bwd(s.ref(), hp.N, p(hp.X), p(hp.W), p(hp.variance), p(hp.lengthscales), p(hp.gd), p(hp.dW), p(hp.dvar), p(hp.dls), None, 0, st()),
}
for i in range(3):
    _lib.check(fns[which]())
torch.cuda.synchronize()
raw = ctypes.CDLL(_lib.LIB_PATH) if hasattr(_lib, "LIB_PATH") else L
raw.cgp_umma_trace_dump()
