"""Large-N forward sweep: `predict_y` of 10 000 test images in batches of 1000
(opt_tools/gpflow_tasks.py:70-90) through SVGP.predict_batched -- Kuu, its Cholesky and the projected variational
factors computed ONCE per sweep, every batch one Kuf + Kdiag evaluation and two GEMMs -- at config 3.  Rows are
sharded over the ranks (no collective; each rank returns its shard).  bench.py reports it as "predict"."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

N_TEST = 10000
BATCH = 1000


def measure(ctx, dtype_name="f64", sweeps=3):
    from tools.svgp_step_time import build_model
    m, cfg = build_model(dtype_name)
    rng = np.random.RandomState(5)
    M, L = m.num_inducing, m.num_latent
    m.q_mu = 0.3 * rng.randn(M, L)
    m.q_sqrt = np.tril(0.05 * rng.randn(L, M, M) + np.eye(M)[None]).transpose(1, 2, 0)
    Xtest = np.random.default_rng(11).random((N_TEST, cfg.H * cfg.W))
    dt = torch.float64 if dtype_name == "f64" else torch.float32
    X_host = torch.tensor(Xtest, dtype=dt).pin_memory()
    X_dev = X_host.to(ctx.dev)

    def sweep_host():  # host rows in, host predictions out
        return m.predict_batched(X_host, batch_size=BATCH)

    def sweep_dev():
        return m.predict_batched(X_dev, batch_size=BATCH)

    out = {}
    for name, fn in (("device_resident", sweep_dev), ("host_in_host_out", sweep_host)):
        fn()
        ctx.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(sweeps):
            fn()
        torch.cuda.synchronize()
        dt_s = ctx.max_over_ranks([time.perf_counter() - t0])[0] / sweeps
        out[name] = {"s_per_sweep": dt_s, "value": N_TEST / dt_s, "unit": "images/s"}
    mean, var = sweep_dev()
    out.update({"workload": "predict_y of %d test images, batches of %d, config 3 model (M=750, 10 classes); rows "
                            "sharded over %d rank(s), chol(Kuu) once per sweep" % (N_TEST, BATCH, ctx.world),
                "dtype": dtype_name, "images_total": N_TEST, "scaling": "strong", "n_gpus": ctx.world,
                "rows_this_rank": int(mean.shape[0]), "checksum_rank0": float(np.sum(mean))})
    return out
