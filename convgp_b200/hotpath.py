"""The hot path as one unit of work: Kuf + Kdiag + Kuu forward and their backward for one
minibatch, driven straight through the C ABI with pre-allocated buffers (no autograd tape).

This is what a data-parallel rank runs per step: its shard of the minibatch rows, replicated
parameters, and ONE packed gradient buffer [dZ | dW | dvariance | dlengthscales] that is summed
across ranks -- by the one-shot peer-memory all-reduce kernel of this library
(``parallel.GradAllReduce``: NVLink loads from every peer's buffer, inside the step's CUDA graph) or,
where peer mapping is unavailable, by NCCL.  ``step()`` is stream-ordered and allocation-free, so it
is captured in a CUDA graph (``capture()``)."""
import torch

from . import _lib, parallel
from ._lib import check, ptr, stream_ptr


def _pad16(n, es):
    """Element count rounded up so that the next segment of the packed buffer starts 16-byte aligned."""
    q = 16 // es
    return (n + q - 1) // q * q


class HotPath:
    def __init__(self, spec, X, Z, W, variance, lengthscales, G, gd, Gu, diag_add=0.0, collective=False):
        self.spec = spec
        self.X, self.Z, self.W = X.contiguous(), Z.contiguous(), None if W is None else W.contiguous()
        self.variance, self.lengthscales = variance.contiguous(), lengthscales.contiguous()
        self.G, self.gd, self.Gu = G.contiguous(), gd.contiguous(), Gu.contiguous()
        self.diag_add = float(diag_add)
        dev, dt = X.device, X.dtype
        N, M = X.shape[0], Z.shape[0]
        self.N, self.M = N, M
        C = spec.C
        self.Kuf = torch.empty((M, N, C) if spec.per_channel else (M, N), dtype=dt, device=dev)
        self.Kdiag = torch.empty((N, C) if spec.per_channel else (N,), dtype=dt, device=dev)
        self.Kuu = torch.empty((M, M), dtype=dt, device=dev)
        es = X.element_size()
        sizes = [Z.numel(), 0 if W is None else W.numel(), variance.numel(), lengthscales.numel()]
        offs, o = [], 0
        for s in sizes:  # every segment 16-byte aligned (vector loads / stores of the collective and the kernels)
            offs.append(o)
            o += _pad16(s, es)
        total = max(o, 16 // es)
        # the gradient collective owns the buffer the backward kernels accumulate into (peer-mapped when it can be)
        self.reduce = parallel.GradAllReduce(total, dt, dev) if collective else None
        self.acc = self.reduce.local if self.reduce is not None else torch.zeros(total, dtype=dt, device=dev)
        self.grads = self.reduce.result if self.reduce is not None else self.acc
        self._acc_views = [self.acc[o:o + s] if s else None for o, s in zip(offs, sizes)]
        self.dZ, self.dW, self.dvar, self.dls = [self.grads[o:o + s] if s else None for o, s in zip(offs, sizes)]
        self.ws, self.ws_bytes = spec.workspace(N, M, dev)
        # Kdiag: the forward stores the per-image row sums its backward is made of (N * (P + 2) elements) when
        # the configuration is covered, and the backward is then O(N P) instead of a second sweep of the Gram
        n_saved = _lib.lib().cgp_kdiag_saved_elems(spec.ref(), N) if N > 0 else 0
        self.kd_saved = torch.empty(n_saved, dtype=dt, device=dev) if n_saved else None
        self.graph = None
        # Kuf / Kdiag / Kuu forward + backward: one kernel each (+ the all-reduce kernel)
        self.kernels_per_step = 6 + (1 if self.reduce is not None else 0)
        self._side = None

    @property
    def collective_name(self):
        return "none" if self.reduce is None else self.reduce.name

    def families(self):
        """Which CUDA kernel family each launch of this configuration dispatches to (cgp_kernel_family)."""
        L = _lib.lib()
        return {k: L.cgp_kernel_family(self.spec.ref(), self.N, self.M, i).decode()
                for i, k in enumerate(("kuf_fwd", "kuf_bwd", "kdiag_fwd", "kdiag_bwd"))}

    # ------------------------------------------------------------------ launches
    def _kuf_fwd(self, st):
        L, s = _lib.lib(), self.spec
        if self.N > 0:
            check(L.cgp_kuf_fwd(s.ref(), self.N, self.M, ptr(self.X), ptr(self.Z), ptr(self.W), ptr(self.variance),
                                ptr(self.lengthscales), ptr(self.Kuf), ptr(self.ws), self.ws_bytes, st))

    def _kuu_fwd(self, st):
        L, s = _lib.lib(), self.spec
        check(L.cgp_kuu_fwd(s.ref(), self.M, ptr(self.Z), ptr(self.variance), ptr(self.lengthscales), self.diag_add,
                            ptr(self.Kuu), st))

    def _kdiag_fwd(self, st):
        L, s = _lib.lib(), self.spec
        if self.N == 0:
            return
        if self.kd_saved is not None:
            check(L.cgp_kdiag_fwd_saved(s.ref(), self.N, ptr(self.X), ptr(self.W), ptr(self.variance),
                                        ptr(self.lengthscales), ptr(self.Kdiag), ptr(self.kd_saved),
                                        self.kd_saved.numel(), st))
        else:
            check(L.cgp_kdiag_fwd(s.ref(), self.N, ptr(self.X), ptr(self.W), ptr(self.variance),
                                  ptr(self.lengthscales), ptr(self.Kdiag), ptr(self.ws), self.ws_bytes, st))

    def _kuf_bwd(self, st):
        L, s = _lib.lib(), self.spec
        dZ, dW, dvar, dls = self._acc_views
        if self.N > 0:
            check(L.cgp_kuf_bwd(s.ref(), self.N, self.M, ptr(self.X), ptr(self.Z), ptr(self.W), ptr(self.variance),
                                ptr(self.lengthscales), ptr(self.G), ptr(dZ), ptr(dW), ptr(dvar), ptr(dls),
                                ptr(self.ws), self.ws_bytes, st))

    def _kdiag_bwd(self, st):
        L, s = _lib.lib(), self.spec
        _, dW, dvar, dls = self._acc_views
        if self.N == 0:
            return
        if self.kd_saved is not None:
            check(L.cgp_kdiag_bwd_saved(s.ref(), self.N, ptr(self.variance), ptr(self.lengthscales), ptr(self.gd),
                                        ptr(self.kd_saved), self.kd_saved.numel(), ptr(dW), ptr(dvar), ptr(dls), st))
        else:
            check(L.cgp_kdiag_bwd(s.ref(), self.N, ptr(self.X), ptr(self.W), ptr(self.variance),
                                  ptr(self.lengthscales), ptr(self.gd), ptr(dW), ptr(dvar), ptr(dls),
                                  ptr(self.ws), self.ws_bytes, st))

    def _kuu_bwd(self, st):
        L, s = _lib.lib(), self.spec
        dZ, _, dvar, dls = self._acc_views
        check(L.cgp_kuu_bwd(s.ref(), self.M, ptr(self.Z), ptr(self.variance), ptr(self.lengthscales), ptr(self.Gu),
                            ptr(dZ), ptr(dvar), ptr(dls), st))

    # ------------------------------------------------------------------ one step
    def forward(self):
        st = stream_ptr()
        self._kuf_fwd(st)
        self._kdiag_fwd(st)
        self._kuu_fwd(st)

    def backward(self):
        st = stream_ptr()
        self.acc.zero_()
        self._kuf_bwd(st)
        self._kdiag_bwd(st)
        self._kuu_bwd(st)
        if self.reduce is not None:
            self.reduce.run()

    def step_overlapped(self):
        """forward+backward with the three independent kernels of each phase on three streams (fork/join with
        events): every kernel is a persistent grid that fills the GPU, so the gain is the overlap of one kernel's
        ramp-up / tail with another's body.  The three forwards JOIN before any backward starts: in the reference's
        ELBO dL/dKuf, dL/dKdiag and dL/dKuu all depend on all three forward results (svconvgp.py:91-104), so a
        training step cannot overlap a backward kernel with another branch's forward.  Capturable."""
        main = torch.cuda.current_stream()
        if self._side is None:
            self._side = [torch.cuda.Stream(), torch.cuda.Stream()]
        s1, s2 = self._side
        sp = lambda st: _lib.ctypes.c_void_p(st.cuda_stream)  # noqa: E731

        def fork():
            e = torch.cuda.Event()
            e.record(main)
            s1.wait_event(e)
            s2.wait_event(e)

        def join():
            for st in (s1, s2):
                j = torch.cuda.Event()
                j.record(st)
                main.wait_event(j)

        self.acc.zero_()
        fork()
        self._kuf_fwd(sp(main))
        self._kdiag_fwd(sp(s1))
        self._kuu_fwd(sp(s2))
        join()
        fork()
        # backward (all three accumulate atomically into the packed gradient buffer).  Capture order matters: the
        # instantiated graph starts the LAST captured branch first, and the Kuf backward (255 registers x 296 CTAs:
        # the whole register file) must not queue behind the two small kernels (measured at 25 images: 200 vs 206 us)
        self._kuu_bwd(sp(s2))
        self._kdiag_bwd(sp(s1))
        self._kuf_bwd(sp(main))
        join()
        if self.reduce is not None:
            self.reduce.run()

    def step(self):
        if self.graph is not None:
            self.graph.replay()
            if self.reduce is not None and not self.reduce.capturable:
                self.reduce.run()
        else:
            self.forward()
            self.backward()

    def capture(self, overlapped=False):
        """Capture forward+backward (+ the gradient all-reduce when it is this library's own kernel) into a CUDA
        graph: launch latency matters, the whole step is ~1 ms of GPU time at the MNIST shape in float64, 0.2 ms in
        float32, and an eighth of that per GPU when the minibatch is sharded."""
        red = self.reduce
        try:
            if red is not None and not red.capturable:
                self.reduce = None  # NCCL stays outside the graph (launched behind it by step())
            self.forward()
            self.backward()
            if overlapped:
                self.step_overlapped()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                if overlapped:
                    self.step_overlapped()
                else:
                    self.forward()
                    self.backward()
        finally:
            self.reduce = red
        self.graph = g
        return g
