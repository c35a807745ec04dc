"""Two data-parallel ranks on the GPU box: the row-sharded CUDA hot path + gradient all-reduce must reproduce the
single-rank gradient, and the sharded SVGP objective the unsharded one.

With >= 2 GPUs each rank owns a GPU, the process group is NCCL and the sum is this library's one-shot peer-memory
all-reduce kernel (cgp_comm_allreduce over CUDA IPC / NVLink).  On a one-GPU box both ranks share the device and
the process group is gloo (NCCL refuses two ranks on one device): the CUDA kernels and the sharding logic are the
same, the sum goes through torch.distributed."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from tests.util import Case, rel_err

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _case():
    return Case("dp2", "conv", 28, 28, 1, 5, 5, [(0.9, 1.3, None, False)], N=9, M=150, seed=5)


def _hotpath(case, dtype, dev, rows, gu_scale, collective):
    from convgp_b200 import _lib
    from convgp_b200.hotpath import HotPath
    lo, hi = rows
    t = lambda a: torch.tensor(np.ascontiguousarray(a), dtype=dtype, device=dev)  # noqa: E731
    spec = _lib.Spec(dtype, case.H, case.W, case.C, case.ph, case.pw, _lib.LAYOUT_PLANAR, _lib.OUT_SUM, 1, False, False)
    return HotPath(spec, t(case.X[lo:hi]), t(case.Z), t(case.Wv), t([0.9]), t([1.3]), t(case.G[:, lo:hi]),
                   t(case.gd[lo:hi]), t(case.Gu * gu_scale), diag_add=1e-3, collective=collective)


def _worker(rank, world, port, ngpu, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = torch.device("cuda", rank % ngpu)
    torch.cuda.set_device(dev)
    if ngpu >= world:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    from convgp_b200 import parallel
    from convgp_b200.settings import settings
    settings.device = dev
    case = _case()
    res = {}
    for dtype, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
        rows = parallel.shard_bounds(case.N, world, rank)
        hp = _hotpath(case, dtype, dev, rows, 1.0 / world, True)
        full = _hotpath(case, dtype, dev, (0, case.N), 1.0, False)
        hp.step()
        full.step()
        hp.capture(overlapped=True)  # and again from the CUDA graph (the peer kernel is captured in it)
        for _ in range(3):
            hp.step()
        torch.cuda.synchronize()
        if hp.reduce is not None:
            hp.reduce.status()
        res[tag + "_sharded"] = hp.grads.double().cpu().numpy()
        res[tag + "_full"] = full.grads.double().cpu().numpy()
        res[tag + "_dZ"] = hp.dZ.double().cpu().numpy().reshape(case.M, -1)
        res[tag + "_dW"] = hp.dW.double().cpu().numpy()
        res[tag + "_name"] = np.array(hp.collective_name)
    # the model-level objective: sharded + all-reduced == unsharded
    from convgp_b200 import convkernels as ck, kernels as bk, likelihoods as lik, svgp
    settings.float_type = torch.float64
    rng = np.random.RandomState(0)
    N, M, L = 11, 9, 3
    X, Y = rng.rand(N, 100), rng.randint(0, L, size=(N, 1)).astype(np.float64)
    k = ck.WeightedConv(bk.RBF(9, variance=0.8, lengthscales=1.4), [10, 10], [3, 3]) + bk.White(1, 1e-3)
    Z = rng.rand(M, 9)
    m = svgp.SVGP(X, Y, k, lik.MultiClass(L), Z, num_latent=L)
    m.q_mu = 0.3 * rng.randn(M, L)
    m.q_sqrt = np.tril(0.2 * rng.randn(L, M, M) + np.eye(M)[None]).transpose(1, 2, 0)
    x0 = m.get_free_state()
    m.data_parallel = False
    f1, g1 = m._objective(x0)
    m.data_parallel = True
    f2, g2 = m._objective(x0)
    res.update(f_full=np.array(f1), g_full=g1, f_dp=np.array(f2), g_dp=g2)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), **res)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_gradients_equal_single_rank(tmp_path):
    ngpu = torch.cuda.device_count()
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), ngpu, str(tmp_path)), nprocs=world, join=True)
    r = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % i)) for i in range(world)]
    ref = _case().twin_grads()
    for tag, tol in (("f64", 1e-10), ("f32", 1e-4)):
        for x in r:
            assert rel_err(x[tag + "_sharded"], x[tag + "_full"]) < tol, (tag, str(x[tag + "_name"]))
            assert rel_err(x[tag + "_dZ"], ref["dZ"]) < tol and rel_err(x[tag + "_dW"], ref["dW"]) < tol
        # every rank holds the bit-identical sum
        assert np.array_equal(r[0][tag + "_sharded"], r[1][tag + "_sharded"])
    if ngpu >= world:
        assert "peer-memory" in str(r[0]["f64_name"]), str(r[0]["f64_name"])
    for x in r:
        assert abs(float(x["f_dp"]) - float(x["f_full"])) < 1e-10 * abs(float(x["f_full"]))
        assert rel_err(x["g_dp"], x["g_full"]) < 1e-10
