"""Development aid: float32 kernel families of the bench workload against the float64 path of the same library
(values and gradients) and their per-kernel CUDA-event timings.

    python tools/f32_check.py [--rows N] [--opt name=value,...] [--reps K]
"""
import argparse
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from convgp_b200 import _lib  # noqa: E402
from convgp_b200.hotpath import HotPath  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=0)
ap.add_argument("--opt", default="")
ap.add_argument("--reps", type=int, default=12)
ap.add_argument("--cfg", default="cfg3")
ap.add_argument("--images", type=int, default=0, help="minibatch size (default: the config's)")
ap.add_argument("--lib", default="", help="load this copy of the library instead (tools/build_trace.sh)")
ap.add_argument("--trace", default="", help="kuf_fwd | kdiag_fwd | kuf_bwd: dump the umma16 time stamps of one launch of it")
ap.add_argument("--randw", action="store_true", help="patch weights ~ N(0, 1) instead of ones")
a = ap.parse_args()
if a.lib:
    _lib.LIB_PATH = os.path.abspath(a.lib)
for kv in [s for s in a.opt.split(",") if s]:
    k, v = kv.split("=")
    _lib.set_option(k, int(v))
dev = torch.device("cuda", 0)
cfg = bench.CONFIGS[a.cfg]
inp = bench.make_inputs(cfg, 0, a.images or None)
if a.rows:
    inp["X"], inp["G"], inp["gd"] = inp["X"][:a.rows], inp["G"][:, :a.rows], inp["gd"][:a.rows]
if a.randw and inp["W"] is not None:
    import numpy as np
    inp["W"] = np.random.default_rng(3).standard_normal(inp["W"].shape)


def make(dt):
    t = {k: (None if v is None else torch.tensor(v, dtype=dt, device=dev)) for k, v in inp.items()}
    layout = _lib.LAYOUT_PLANAR if cfg.layout == "planar" else _lib.LAYOUT_COLOUR_PATCH
    mask = None
    if cfg.G > 1:
        import numpy as np
        mask = np.zeros((cfg.G, cfg.D), dtype=np.uint8)
        for g in range(cfg.G):
            mask[g, g::cfg.G] = 1
    spec = _lib.Spec(dt, cfg.H, cfg.W, cfg.C, cfg.ph, cfg.pw, layout, _lib.OUT_SUM, cfg.G, False, cfg.layout == "colour", mask)
    return HotPath(spec, t["X"], t["Z"], t["W"], t["var"], t["ls"], t["G"], t["gd"], t["Gu"], diag_add=cfg.diag_add)


h64, h32 = make(torch.float64), make(torch.float32)
print("families f32:", h32.families())
for h in (h64, h32):
    h.step()
torch.cuda.synchronize()


def rel(a32, a64):
    den = float(a64.abs().max())
    return float((a32.double() - a64).abs().max()) / (den if den > 0 else 1.0)


names = ["Kuf", "Kdiag", "Kuu", "dZ", "dW", "dvar", "dls"]
errs = {k: rel(getattr(h32, k), getattr(h64, k)) for k in names if getattr(h32, k) is not None}
print("rel err vs f64:", {k: "%.2e" % v for k, v in errs.items()})
print("nan:", {k: bool(torch.isnan(getattr(h32, k)).any()) for k in names if getattr(h32, k) is not None})

L, s, st, p = _lib.lib(), h32.spec, _lib.stream_ptr, _lib.ptr
hp = h32
fns = {
    "kuf_fwd": lambda: _lib.check(L.cgp_kuf_fwd(s.ref(), hp.N, hp.M, p(hp.X), p(hp.Z), p(hp.W), p(hp.variance), p(hp.lengthscales), p(hp.Kuf), p(hp.ws), hp.ws_bytes, st())),
    "kdiag_fwd": lambda: hp._kdiag_fwd(st()),
    "kuf_bwd": lambda: _lib.check(L.cgp_kuf_bwd(s.ref(), hp.N, hp.M, p(hp.X), p(hp.Z), p(hp.W), p(hp.variance), p(hp.lengthscales), p(hp.G), p(hp.dZ), p(hp.dW), p(hp.dvar), p(hp.dls), p(hp.ws), hp.ws_bytes, st())),
    "kdiag_bwd": lambda: hp._kdiag_bwd(st()),
}
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
out = {}
for name, fn in fns.items():
    ts = []
    for i in range(a.reps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        if i >= 2:
            ts.append(e0.elapsed_time(e1))
    out[name] = "%.1f us (min %.1f)" % (1e3 * statistics.mean(ts), 1e3 * min(ts))
print(out)
if a.trace:
    import ctypes
    fns[a.trace]()
    torch.cuda.synchronize()
    ctypes.CDLL(_lib.LIB_PATH).cgp_umma16_trace_dump()
