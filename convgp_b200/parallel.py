"""Data-parallel plumbing for the hot path: one process per GPU, the minibatch rows
of X (hence the columns of Kuf and the entries of Kdiag) are sharded contiguously over ranks,
parameters are replicated, and ONE packed gradient buffer is summed with an all-reduce
(NCCL over NVLink on GPUs; gloo in the CPU tests).  There is no other collective on the path."""
import torch
import torch.distributed as dist


def shard_bounds(n_rows, world_size, rank):
    """Contiguous, balanced row range of `rank`: the first n_rows % world_size ranks get one more
    (cifar's N=30 over 8 ranks -> 4,4,4,4,4,4,3,3)."""
    base, extra = divmod(int(n_rows), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_rows(X, world_size=None, rank=None):
    """This rank's rows of a minibatch tensor (first axis)."""
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    lo, hi = shard_bounds(X.shape[0], world_size, rank)
    return X[lo:hi]


def shard_columns(G, world_size=None, rank=None):
    """This rank's columns of an (M, N[, C]) tensor laid out like Kuf (image index on axis 1)."""
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    lo, hi = shard_bounds(G.shape[1], world_size, rank)
    return G[:, lo:hi]


def allreduce_sum_(buf):
    """In-place sum of the packed gradient buffer over all ranks (no-op without a process group)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    return buf


def pack(tensors):
    """Flatten a list of tensors into one contiguous buffer + the metadata to unpack it."""
    flat = torch.cat([t.reshape(-1) for t in tensors])
    return flat, [tuple(t.shape) for t in tensors]


def unpack(flat, shapes):
    out, o = [], 0
    for shp in shapes:
        n = 1
        for s in shp:
            n *= s
        out.append(flat[o:o + n].view(shp))
        o += n
    return out
