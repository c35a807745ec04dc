"""Time one full SVGP training step (ELBO + gradient w.r.t. every free parameter) at config 3 --
`mnist.py -k wconv -M 750 --minibatch-size 200`, MultiClass(10), whitened full-covariance q -- eagerly and replayed
from ONE CUDA graph.  `measure()` is what bench.py reports as "svgp_step"; run as a script it also prints a
per-kernel breakdown (development aid)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch


def build_model(dtype_name="f64"):
    import bench
    from convgp_b200 import convkernels as ck, kernels as bk, likelihoods as lik, svgp
    from convgp_b200.settings import settings
    dt = torch.float64 if dtype_name == "f64" else torch.float32
    settings.float_type = dt
    cfg = bench.HEAD
    inp = bench.make_inputs(cfg, 0)
    L = 10
    Y = np.random.default_rng(7).integers(0, L, size=(cfg.N, 1)).astype(np.float64)
    k = ck.WeightedConv(bk.RBF(cfg.D), [cfg.H, cfg.W], [cfg.ph, cfg.pw]) + bk.White(1, 1e-3)
    return svgp.SVGP(inp["X"], Y, k, lik.MultiClass(L), inp["Z"], num_latent=L), cfg


def _time(fn, n):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def measure(dtype_name="f64", steps=20):
    m, cfg = build_model(dtype_name)
    params = [p.free for p in m.free_params()]

    def eager():
        for p in params:
            p.grad = None
        f = -m.build_likelihood()
        f.backward()
        return f

    eager_ms = _time(eager, max(5, steps // 2))
    m2, _ = build_model(dtype_name)
    gstep = m2.graphed_objective()
    graph_ms = _time(gstep, steps)
    return {"workload": cfg.workload + ", MultiClass(10), whitened full q_sqrt (750 x 750 x 10): -ELBO and its gradient "
                                       "w.r.t. Z, W, variance, lengthscale, White variance, q_mu, q_sqrt",
            "dtype": dtype_name, "ms_per_step_eager": eager_ms, "ms_per_step": graph_ms,
            "images_per_s": cfg.N / (graph_ms * 1e-3), "mode": "one CUDA graph per step",
            "free_parameters": int(sum(p.numel() for p in params))}


if __name__ == "__main__":
    dtype = "f64" if (len(sys.argv) < 2 or sys.argv[1] == "f64") else "f32"
    print(measure(dtype))
    m, _ = build_model(dtype)
    params = [p.free for p in m.free_params()]

    def step():
        for p in params:
            p.grad = None
        f = -m.build_likelihood()
        f.backward()

    for _ in range(3):
        step()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            step()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
