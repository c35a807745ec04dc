"""What one rank of a row-sharded (strong-scaling) step costs: the config-3 hot-path step on ONE GPU with
N = 200, 100, 50, 25 rows (the shard of 1, 2, 4, 8 ranks), graph-replayed with the three-branch fork/join and
without the collective, plus CUDA-event times of each kernel.  Development aid (prints JSON lines)."""
import json
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from convgp_b200 import _lib  # noqa: E402


class _Args:
    no_graph = False
    steps = 100
    warmup = 5


def main():
    opts = [a for a in sys.argv[1:] if "=" in a]
    for o in opts:
        k, v = o.split("=")
        _lib.set_option(k, int(v))
    os.environ.setdefault("WORLD_SIZE", "1")
    ctx = bench.Ctx(_Args())
    for dtype in ("f64", "f32"):
        for cfgkey in ("cfg3", "cfg5"):
            cfg = bench.CONFIGS[cfgkey]
            for div in (1, 2, 4, 8):
                n = max(1, cfg.N // div)
                hp, _, _, _ = bench.build_hotpath(ctx, cfg, dtype, 0, rows=(0, n), gu_scale=1.0 / div, collective=False)
                hp.capture(overlapped=True)
                ms, _ = bench.time_steps(ctx, hp, 100, 5)
                pk = bench.per_kernel_ms(ctx, hp)
                print(json.dumps({"cfg": cfgkey, "dtype": dtype, "rows": n, "ms_per_step": round(ms / 100, 4),
                                  "x_vs_full": None, "kernel_us": {k: round(1e3 * v, 1) for k, v in pk.items()},
                                  "options": opts}))




def timeline(dtype="f64", rows=25, cfgkey="cfg3"):
    """Kernel start / end times inside ONE graph-replayed step (torch profiler, CUPTI): where a sharded step's
    time goes -- kernel bodies, gaps between dependent graph nodes, the fork / join."""
    from torch.profiler import ProfilerActivity, profile
    os.environ.setdefault("WORLD_SIZE", "1")
    ctx = bench.Ctx(_Args())
    cfg = bench.CONFIGS[cfgkey]
    hp, _, _, _ = bench.build_hotpath(ctx, cfg, dtype, 0, rows=(0, rows), gu_scale=1.0, collective=False)
    if os.environ.get("CGP_BRANCH_ORDER"):
        hp.branch_order = os.environ["CGP_BRANCH_ORDER"]
    hp.capture(overlapped=True)
    for _ in range(5):
        hp.step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            ctx.flush_l2()
            hp.step()
        torch.cuda.synchronize()
    evs = sorted((e for e in prof.events() if e.device_type.name == "CUDA"), key=lambda e: e.time_range.start)
    t0 = None
    for e in evs:
        if "fill" in e.name.lower() and e.time_range.end - e.time_range.start > 20:
            t0 = None  # the L2 flush separates steps
            print("---- step", flush=True)
            continue
        if t0 is None:
            t0 = e.time_range.start
        print("%8.1f %8.1f  %s" % (e.time_range.start - t0, e.time_range.end - t0, e.name[:90]))


if __name__ == "__main__":
    tl = [a for a in sys.argv[1:] if a.startswith("timeline")]
    if tl:  # e.g. timeline:f64:25:cfg3
        parts = tl[0].split(":")
        timeline(parts[1] if len(parts) > 1 else "f64", int(parts[2]) if len(parts) > 2 else 25,
                 parts[3] if len(parts) > 3 else "cfg3")
    else:
        main()
