"""World-size-2 data-parallel path on CPU (gloo): row-sharded minibatch, replicated parameters, one
packed-gradient all-reduce.  The per-shard arithmetic is done by the torch twin here (no GPU); what
is under test is the host logic of convgp_b200.parallel -- that shards tile the batch and that the
summed packed buffer equals the unsharded gradient."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.util import Case


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from convgp_b200 import parallel
    from oracle import torch_twin as T
    torch.set_num_threads(1)
    case = Case("dp", "conv", 10, 10, 1, 3, 3, [(0.9, 1.3, None, False)], N=5, M=6, seed=3)
    dt = torch.float64
    X = parallel.shard_rows(torch.tensor(case.X, dtype=dt))
    G = parallel.shard_columns(torch.tensor(case.G, dtype=dt))
    gd = parallel.shard_rows(torch.tensor(case.gd, dtype=dt))
    Gu = torch.tensor(case.Gu, dtype=dt) / world  # each rank carries its share of dL/dKuu
    Z = torch.tensor(case.Z, dtype=dt, requires_grad=True)
    W = torch.tensor(case.Wv, dtype=dt, requires_grad=True)
    v = torch.tensor(0.9, dtype=dt, requires_grad=True)
    l = torch.tensor(1.3, dtype=dt, requires_grad=True)
    if X.shape[0] > 0:
        T.hot_path_step(X, Z, [(v, l, None)], case.geom, W, G, gd, Gu, chunk=8)
    flat, shapes = parallel.pack([Z.grad, W.grad, v.grad.reshape(1), l.grad.reshape(1)])
    parallel.allreduce_sum_(flat)
    dZ, dW, dv, dl = parallel.unpack(flat, shapes)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), dZ=dZ.numpy(), dW=dW.numpy(), dv=dv.numpy(), dl=dl.numpy(),
             rows=np.array(parallel.shard_bounds(case.N, world, rank)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    case = Case("dp", "conv", 10, 10, 1, 3, 3, [(0.9, 1.3, None, False)], N=5, M=6, seed=3)
    ref = case.twin_grads()
    r = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % i)) for i in range(world)]
    assert [tuple(x["rows"]) for x in r] == [(0, 3), (3, 5)]
    for x in r:  # every rank ends with the same, full gradient
        np.testing.assert_allclose(x["dZ"], ref["dZ"], rtol=1e-11, atol=1e-13)
        np.testing.assert_allclose(x["dW"], ref["dW"], rtol=1e-11, atol=1e-13)
        np.testing.assert_allclose(x["dv"], ref["dvar"][0], rtol=1e-11)
        np.testing.assert_allclose(x["dl"], ref["dls"][0], rtol=1e-11)
