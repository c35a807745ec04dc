"""The hot path as one unit of work: Kuf + Kdiag + Kuu forward and their backward for one
minibatch, driven straight through the C ABI with pre-allocated buffers (no autograd tape).

This is what a data-parallel rank runs per step: its shard of the minibatch rows, replicated
parameters, and ONE packed gradient buffer [dZ | dW | dvariance | dlengthscales] that the NCCL
all-reduce sums across ranks.  ``step()`` is stream-ordered and allocation-free, so
it can be captured in a CUDA graph (``capture()``)."""
import ctypes  # noqa: F401

import torch

from . import _lib
from ._lib import check, ptr, stream_ptr


class HotPath:
    def __init__(self, spec, X, Z, W, variance, lengthscales, G, gd, Gu, diag_add=0.0):
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
        sizes = [Z.numel(), 0 if W is None else W.numel(), variance.numel(), lengthscales.numel()]
        self.grads = torch.zeros(sum(sizes), dtype=dt, device=dev)
        o = 0
        views = []
        for s in sizes:
            views.append(self.grads[o:o + s] if s else None)
            o += s
        self.dZ, self.dW, self.dvar, self.dls = views
        self.ws, self.ws_bytes = spec.workspace(N, M, dev)
        # Kdiag: the forward stores the per-image row sums its backward is made of (N * (P + 2) elements) when
        # the configuration is covered, and the backward is then O(N P) instead of a second sweep of the Gram
        n_saved = _lib.lib().cgp_kdiag_saved_elems(spec.ref(), N) if N > 0 else 0
        self.kd_saved = torch.empty(n_saved, dtype=dt, device=dev) if n_saved else None
        self.graph = None
        # Kuf / Kdiag / Kuu forward + backward: one kernel each
        self.kernels_per_step = 6

    def forward(self):
        L, s, st = _lib.lib(), self.spec, stream_ptr()
        check(L.cgp_kuf_fwd(s.ref(), self.N, self.M, ptr(self.X), ptr(self.Z), ptr(self.W), ptr(self.variance),
                            ptr(self.lengthscales), ptr(self.Kuf), ptr(self.ws), self.ws_bytes, st))
        self._kdiag_fwd(st)
        check(L.cgp_kuu_fwd(s.ref(), self.M, ptr(self.Z), ptr(self.variance), ptr(self.lengthscales), self.diag_add,
                            ptr(self.Kuu), st))

    def _kdiag_fwd(self, st):
        L, s = _lib.lib(), self.spec
        if self.kd_saved is not None:
            check(L.cgp_kdiag_fwd_saved(s.ref(), self.N, ptr(self.X), ptr(self.W), ptr(self.variance),
                                        ptr(self.lengthscales), ptr(self.Kdiag), ptr(self.kd_saved),
                                        self.kd_saved.numel(), st))
        else:
            check(L.cgp_kdiag_fwd(s.ref(), self.N, ptr(self.X), ptr(self.W), ptr(self.variance),
                                  ptr(self.lengthscales), ptr(self.Kdiag), ptr(self.ws), self.ws_bytes, st))

    def _kdiag_bwd(self, st):
        L, s = _lib.lib(), self.spec
        if self.kd_saved is not None:
            check(L.cgp_kdiag_bwd_saved(s.ref(), self.N, ptr(self.variance), ptr(self.lengthscales), ptr(self.gd),
                                        ptr(self.kd_saved), self.kd_saved.numel(), ptr(self.dW), ptr(self.dvar),
                                        ptr(self.dls), st))
        else:
            check(L.cgp_kdiag_bwd(s.ref(), self.N, ptr(self.X), ptr(self.W), ptr(self.variance),
                                  ptr(self.lengthscales), ptr(self.gd), ptr(self.dW), ptr(self.dvar), ptr(self.dls),
                                  ptr(self.ws), self.ws_bytes, st))

    def backward(self):
        L, s, st = _lib.lib(), self.spec, stream_ptr()
        self.grads.zero_()
        check(L.cgp_kuf_bwd(s.ref(), self.N, self.M, ptr(self.X), ptr(self.Z), ptr(self.W), ptr(self.variance),
                            ptr(self.lengthscales), ptr(self.G), ptr(self.dZ), ptr(self.dW), ptr(self.dvar),
                            ptr(self.dls), ptr(self.ws), self.ws_bytes, st))
        self._kdiag_bwd(st)
        check(L.cgp_kuu_bwd(s.ref(), self.M, ptr(self.Z), ptr(self.variance), ptr(self.lengthscales), ptr(self.Gu),
                            ptr(self.dZ), ptr(self.dvar), ptr(self.dls), st))

    def step_overlapped(self):
        """forward+backward with the three independent kernels of each phase on three streams
        (fork/join with events): every kernel is a persistent grid that fills the GPU, so the gain
        is the overlap of one kernel's ramp-up / tail with another's body.  Capturable."""
        L, s = _lib.lib(), self.spec
        main = torch.cuda.current_stream()
        if not hasattr(self, "_side"):
            self._side = [torch.cuda.Stream(), torch.cuda.Stream()]
        s1, s2 = self._side
        sp = lambda st: _lib.ctypes.c_void_p(st.cuda_stream)  # noqa: E731
        self.grads.zero_()
        fork = torch.cuda.Event()
        fork.record(main)
        s1.wait_event(fork)
        s2.wait_event(fork)
        # forward
        check(L.cgp_kuf_fwd(s.ref(), self.N, self.M, ptr(self.X), ptr(self.Z), ptr(self.W), ptr(self.variance),
                            ptr(self.lengthscales), ptr(self.Kuf), ptr(self.ws), self.ws_bytes, sp(main)))
        self._kdiag_fwd(sp(s1))
        check(L.cgp_kuu_fwd(s.ref(), self.M, ptr(self.Z), ptr(self.variance), ptr(self.lengthscales), self.diag_add,
                            ptr(self.Kuu), sp(s2)))
        # backward (all three accumulate atomically into the packed gradient buffer)
        check(L.cgp_kuf_bwd(s.ref(), self.N, self.M, ptr(self.X), ptr(self.Z), ptr(self.W), ptr(self.variance),
                            ptr(self.lengthscales), ptr(self.G), ptr(self.dZ), ptr(self.dW), ptr(self.dvar),
                            ptr(self.dls), ptr(self.ws), self.ws_bytes, sp(main)))
        self._kdiag_bwd(sp(s1))
        check(L.cgp_kuu_bwd(s.ref(), self.M, ptr(self.Z), ptr(self.variance), ptr(self.lengthscales), ptr(self.Gu),
                            ptr(self.dZ), ptr(self.dvar), ptr(self.dls), sp(s2)))
        for st in (s1, s2):
            j = torch.cuda.Event()
            j.record(st)
            main.wait_event(j)

    def step(self):
        if self.graph is not None:
            self.graph.replay()
        else:
            self.forward()
            self.backward()

    def capture(self, overlapped=False, after=None):
        """Capture forward+backward into a CUDA graph (launch-latency matters: the whole step is
        ~1 ms of GPU time at the MNIST shape, 0.2 ms in float32).  overlapped=True captures the three-stream
        variant.  `after`: an optional callable captured behind the step -- the data-parallel all-reduce of
        `self.grads` (NCCL supports capture), so that a multi-GPU step is ONE graph launch."""
        self.forward()
        self.backward()
        if overlapped:
            self.step_overlapped()
        if after is not None:
            after()  # warm the communicator up outside the capture
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            if overlapped:
                self.step_overlapped()
            else:
                self.forward()
                self.backward()
            if after is not None:
                after()
        self.graph = g
        return g
