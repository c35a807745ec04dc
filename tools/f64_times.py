"""Development aid: per-kernel CUDA-event times of the float64 config-3 step (L2 flushed), optionally with another
copy of the library.   python tools/f64_times.py [--lib path] [--rows N]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ap = argparse.ArgumentParser()
ap.add_argument("--lib", default="")
ap.add_argument("--rows", type=int, default=0)
a = ap.parse_args()
from convgp_b200 import _lib  # noqa: E402
if a.lib:
    _lib.LIB_PATH = os.path.abspath(a.lib)
import bench  # noqa: E402


class _Args:
    no_graph = False
    steps = 100
    warmup = 5


os.environ.setdefault("WORLD_SIZE", "1")
ctx = bench.Ctx(_Args())
cfg = bench.HEAD
hp, _, _, _ = bench.build_hotpath(ctx, cfg, "f64", 0, rows=(0, a.rows) if a.rows else None, collective=False)
hp.capture(overlapped=True)
ms, _ = bench.time_steps(ctx, hp, 100, 5)
pk = bench.per_kernel_ms(ctx, hp)
print({"ms_per_step": round(ms / 100, 4), "kernel_us": {k: round(1e3 * v, 1) for k, v in pk.items()}})
