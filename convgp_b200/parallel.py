"""Data-parallel plumbing for the hot path: one process per GPU, the minibatch rows
of X (hence the columns of Kuf and the entries of Kdiag) are sharded contiguously over ranks,
parameters are replicated, and ONE packed gradient buffer is summed across ranks.  On GPUs that sum
is this library's own one-shot peer-memory all-reduce (``GradAllReduce`` -> ``cgp_comm_allreduce``,
csrc/comm.cu: every rank loads every peer's buffer over NVLink inside the step's CUDA graph);
``torch.distributed`` (NCCL / gloo) carries only the set-up exchange of the 64-byte IPC handles, and the
sum itself where peer mapping is unavailable (CPU tests under gloo, GPUs without peer access) or the
buffer is large (the model-level ``dq_sqrt``, bandwidth-bound: NCCL's ring / NVLS is the right tool there)."""
import ctypes

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


class _DeviceBlob:
    """A raw device allocation exposed to torch through __cuda_array_interface__ (zero-copy view)."""

    def __init__(self, ptr, n, dtype, owner):
        self.owner = owner  # keeps the communicator alive as long as a view exists
        typestr = {torch.float64: "<f8", torch.float32: "<f4"}[dtype]
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2}


class GradAllReduce:
    """Sum of one packed buffer of `n` elements over all ranks, in place from the caller's point of view:
    the backward kernels accumulate into ``.local``, ``run()`` enqueues the collective on the current stream, and
    ``.result`` then holds the sum (the same tensor as ``.local`` when there is one rank).

    GPU, world > 1: ``cgp_comm_*`` -- a cudaMalloc'ed segment per rank, mapped into every peer by CUDA IPC, and
    one kernel per step that loads all W copies over NVLink (``capturable``: it runs inside the step's CUDA graph).
    Otherwise ``torch.distributed.all_reduce`` on the buffer (NCCL or gloo), which stays outside CUDA graphs."""

    def __init__(self, n, dtype, device, group=None):
        from . import _lib
        self.n, self.dtype, self.device = int(n), dtype, torch.device(device)
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.group = group
        self._comm = None
        self.capturable = True
        self.name = "none"
        # the peer kernel needs one GPU per rank (its flag polls assume the peers' kernels run concurrently): NCCL
        # groups guarantee that; under gloo several ranks may share a device
        if self.world > 1 and self.device.type == "cuda" and dist.get_backend(group) == "nccl":
            try:
                self._connect(_lib)
                self.name = "one-shot peer-memory all-reduce kernel (cgp_comm_allreduce: NVLink loads of every rank's " \
                            "buffer, rank-ordered sum, in the step's CUDA graph)"
            except _lib.CgpError as exc:
                self._comm = None
                self.name = "NCCL all-reduce (torch.distributed); peer mapping unavailable: %s" % str(exc)[:80]
        if self._comm is None:
            self.local = torch.zeros(self.n, dtype=dtype, device=self.device)
            self.result = self.local
            if self.world > 1:
                self.capturable = False
                if self.name == "none":
                    self.name = "%s all-reduce (torch.distributed)" % dist.get_backend(group)

    def _connect(self, _lib):
        L = _lib.lib()
        es = torch.empty(0, dtype=self.dtype).element_size()
        comm = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(L.cgp_comm_create(self.world, self.rank, self.n * es, ctypes.byref(comm)))
            handle = ctypes.create_string_buffer(64)
            _lib.check(L.cgp_comm_handle(comm, handle))
            # plumbing: the 64-byte IPC handles travel through the process group
            mine = torch.frombuffer(bytearray(handle.raw), dtype=torch.uint8).to(self.device)
            every = [torch.empty(64, dtype=torch.uint8, device=self.device) for _ in range(self.world)]
            dist.all_gather(every, mine, group=self.group)
            blob = b"".join(bytes(t.cpu().numpy().tobytes()) for t in every)
            status = L.cgp_comm_connect(comm, ctypes.c_char_p(blob))
            ok = torch.tensor([1 if status == 0 else 0], device=self.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)  # every rank takes the same path
            if int(ok) == 0:
                msg = (L.cgp_last_error() or b"a peer could not map").decode()
                L.cgp_comm_destroy(comm)
                raise _lib.CgpError(-7, msg)
        self._comm = comm
        self._L = L
        self.local = torch.as_tensor(_DeviceBlob(L.cgp_comm_local(comm), self.n, self.dtype, self), device=self.device)
        self.result = torch.as_tensor(_DeviceBlob(L.cgp_comm_result(comm), self.n, self.dtype, self),
                                      device=self.device)

    def run(self):
        """Enqueue the sum on the current stream."""
        if self.world == 1:
            return self.result
        if self._comm is not None:
            from . import _lib
            _lib.check(self._L.cgp_comm_allreduce(self._comm, _lib.dtype_code(self.dtype), self.n, _lib.stream_ptr()))
        else:
            dist.all_reduce(self.local, op=dist.ReduceOp.SUM, group=self.group)
        return self.result

    def status(self):
        """Raises if a flag poll of the peer kernel ever timed out (a rank never arrived).  Synchronises."""
        if self._comm is not None:
            from . import _lib
            _lib.check(self._L.cgp_comm_status(self._comm))

    def close(self):
        if self._comm is not None:
            self.local = self.result = None
            self._L.cgp_comm_destroy(self._comm)
            self._comm = None
