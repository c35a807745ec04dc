"""Run a few eager hot-path steps at the bench workload (for ncu: `ncu ... python tools/run_step.py`)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from convgp_b200 import _lib  # noqa: E402
from convgp_b200.hotpath import HotPath  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--dtype", default="f64")
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--rows", type=int, default=0, help="first ROWS images only (one rank's shard of a row-sharded step)")
a = ap.parse_args()
dt = torch.float64 if a.dtype == "f64" else torch.float32
dev = torch.device("cuda", 0)
inp = bench.make_inputs(bench.HEAD, 0)
if a.rows:
    inp["X"], inp["G"], inp["gd"] = inp["X"][:a.rows], inp["G"][:, :a.rows], inp["gd"][:a.rows]
t = {k: torch.tensor(v, dtype=dt, device=dev) for k, v in inp.items()}
spec = _lib.Spec(dt, 28, 28, 1, 5, 5, 0, 0, 1, False, False)
hp = HotPath(spec, t["X"], t["Z"], t["W"], t["var"], t["ls"], t["G"], t["gd"], t["Gu"], diag_add=1e-3 + 1e-5)
for _ in range(a.steps):
    hp.step()
torch.cuda.synchronize()
print("ok", float(hp.Kuf.sum()), float(hp.grads.abs().sum()))
