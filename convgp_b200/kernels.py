"""Base kernels with the GPflow 0.x surface the reference composes its patch kernels from
(``GPflow.kernels.RBF / White / Add``; not in the reference tree -- semantics restated, see
oracle/convgp_oracle.py header).  The patch-level evaluation of RBF members happens inside the
CUDA library; ``RBF.K`` on whole inputs (the ``k2 = RBF(784)`` of sumkern_mnist.py:38-39) is a
plain library GEMM."""
import numpy as np
import torch

from .param import Param, Parameterized, Positive
from .settings import settings


def as_tensor(x, dtype=None):
    """NumPy / torch in -> contiguous tensor of the float type on the compute device."""
    dtype = dtype or settings.float_type
    if isinstance(x, torch.Tensor):
        return x.to(device=settings.device, dtype=dtype)
    return torch.as_tensor(np.ascontiguousarray(np.asarray(x, dtype=np.float64)), dtype=dtype,
                           device=settings.device)


_SIDE_STREAMS = {}


def concurrent(*thunks):
    """Evaluate independent thunks on parallel CUDA streams (fork / join with events) and return their results in
    order.  Kzx, Kdiag and Kzz of one SVGP evaluation do not depend on each other, and each of the library's kernels
    is a persistent grid: on parallel streams the ramp-up / tail of one overlaps the body of another.  Autograd replays
    every backward node on the stream its forward ran on, so the three backward kernels overlap the same way.
    Capturable in a CUDA graph (the side streams join the capture through the fork event).  CPU tensors / a single
    thunk: plain sequential evaluation."""
    if len(thunks) < 2 or not torch.cuda.is_available() or settings.device.type != "cuda":
        return [t() for t in thunks]
    main = torch.cuda.current_stream()
    key = (main.device, len(thunks) - 1)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = [torch.cuda.Stream(device=main.device) for _ in range(len(thunks) - 1)]
        # leaves shared by the branches (Z, W, variance, lengthscales) accumulate their gradient on the stream they were
        # created on; the engine synchronises the branch streams with it -- intended here, not a mistake to warn about
        quiet = getattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch", None)
        if quiet is not None:
            quiet(False)
    side = _SIDE_STREAMS[key]
    fork = torch.cuda.Event()
    fork.record(main)
    outs = [None] * len(thunks)
    for st, i in zip(side, range(1, len(thunks))):
        st.wait_event(fork)
        with torch.cuda.stream(st):
            outs[i] = thunks[i]()
    outs[0] = thunks[0]()
    for st, o in zip(side, outs[1:]):
        main.wait_stream(st)
        for t in (o if isinstance(o, (tuple, list)) else (o,)):
            if isinstance(t, torch.Tensor) and t.is_cuda:
                t.record_stream(main)  # allocated on the side stream, consumed on the main one
    return outs


def svgp_terms(kern, Z, X, jitter=None):
    """(Kdiag(X), Kzx(Z, X), Kzz(Z) + jitter I): the three kernel evaluations of one SVGP prediction
    (misvgp.py:194-201, svsumgp.py:76-88; GPflow's SVGP.build_predict), on parallel streams (``concurrent``)."""
    jitter = settings.jitter_level if jitter is None else jitter
    fused = _fused_terms(kern, Z, X, jitter)
    if fused is not None:
        return fused

    def kzz():
        K = kern.Kzz(Z)
        return K + torch.eye(Z.shape[0], dtype=Z.dtype, device=Z.device) * jitter

    kzx, kdiag, kmm = concurrent(lambda: kern.Kzx(Z, X), lambda: kern.Kdiag(X), kzz)
    return kdiag, kzx, kmm


def _fused_terms(kern, Z, X, jitter):
    """One patch kernel, alone or in a sum with White members (mnist.py:42, rectangles.py:44): one autograd node
    for all three terms (Conv.fused_terms).  White contributes nothing to Kzx (White.K(Z, X) = 0), its variance to
    Kdiag and to the diagonal of Kzz."""
    if not (torch.cuda.is_available() and settings.device.type == "cuda"):
        return None
    members = kern.kern_list if isinstance(kern, Add) else [kern]
    convs = [k for k in members if hasattr(k, "fused_terms")]
    whites = [k for k in members if isinstance(k, White)]
    if len(convs) != 1 or len(convs) + len(whites) != len(members):
        return None
    dp = None
    for w in whites:
        dp = w.variance() if dp is None else dp + w.variance()
    return convs[0].fused_terms(Z, X, jitter, dp)


class Kern(Parameterized):
    def __init__(self, input_dim, active_dims=None):
        self.input_dim = int(input_dim)
        self.active_dims = active_dims

    def __add__(self, other):
        return Add([self, other])

    # inter-domain defaults of the GPflow fork: Kzx = K(Z, X), Kzz = K(Z)
    def Kzx(self, Z, X):
        return self.K(Z, X)

    def Kzz(self, Z):
        return self.K(Z)

    # eager NumPy-in / NumPy-out twins (GPflow AutoFlow wrappers, testing/test_convkerns.py:54,118)
    def compute_K_symm(self, X):
        return self.K(as_tensor(X)).detach().cpu().numpy()

    def compute_K(self, X, X2):
        return self.K(as_tensor(X), as_tensor(X2)).detach().cpu().numpy()

    def compute_Kdiag(self, X):
        return self.Kdiag(as_tensor(X)).detach().cpu().numpy()

    def compute_Kzz(self, Z):
        return self.Kzz(as_tensor(Z)).detach().cpu().numpy()

    def compute_Kzx(self, Z, X):
        return self.Kzx(as_tensor(Z), as_tensor(X)).detach().cpu().numpy()


class RBF(Kern):
    """variance * exp(-|x - x'|^2 / (2 l^2)) on ``active_dims``; ``ARD`` gives one l per dim."""

    def __init__(self, input_dim, variance=1.0, lengthscales=None, active_dims=None, ARD=False):
        Kern.__init__(self, input_dim, active_dims)
        self.ARD = bool(ARD)
        self.variance = Param(np.asarray(variance, dtype=np.float64).reshape(-1)[:1].reshape(()), Positive())
        if lengthscales is None:
            lengthscales = np.ones(self.input_dim) if ARD else 1.0
        ls = np.asarray(lengthscales, dtype=np.float64)
        if ARD and ls.ndim == 0:
            ls = ls * np.ones(self.input_dim)
        self.lengthscales = Param(ls if ARD else ls.reshape(()), Positive())

    def dims(self, D):
        """Indices of the active dims inside a D-vector."""
        ad = self.active_dims
        if ad is None:
            ad = slice(self.input_dim)
        return np.arange(D)[ad]

    def _slice(self, X):
        idx = torch.as_tensor(self.dims(X.shape[1]), device=X.device)
        return X if idx.numel() == X.shape[1] else X[:, idx]

    def K(self, X, X2=None):
        X = self._slice(X) / self.lengthscales()
        Xs = (X * X).sum(1)
        if X2 is None:
            d = -2.0 * (X @ X.T) + Xs[:, None] + Xs[None, :]
        else:
            X2 = self._slice(X2) / self.lengthscales()
            d = -2.0 * (X @ X2.T) + Xs[:, None] + (X2 * X2).sum(1)[None, :]
        return self.variance() * torch.exp(-d / 2.0)

    def Kdiag(self, X):
        return self.variance() * torch.ones(X.shape[0], dtype=X.dtype, device=X.device)


class White(Kern):
    def __init__(self, input_dim, variance=1.0, active_dims=None):
        Kern.__init__(self, input_dim, active_dims)
        self.variance = Param(np.asarray(variance, dtype=np.float64).reshape(()), Positive())

    def K(self, X, X2=None):
        if X2 is None:
            return self.variance() * torch.eye(X.shape[0], dtype=X.dtype, device=X.device)
        return torch.zeros(X.shape[0], X2.shape[0], dtype=X.dtype, device=X.device)

    def Kdiag(self, X):
        return self.variance() * torch.ones(X.shape[0], dtype=X.dtype, device=X.device)


class Add(Kern):
    """Sum of kernels.  Members are reachable as ``kern_list[i]`` (mnist.py:47) and by GPflow's
    auto-names: lower-cased class name, ``_1, _2, ...`` when repeated (sumkern_mnist.py:46-50)."""

    def __init__(self, kern_list):
        flat = []
        for k in kern_list:
            flat.extend(k.kern_list if isinstance(k, Add) else [k])
        Kern.__init__(self, max(k.input_dim for k in flat))
        self.kern_list = flat
        names = [type(k).__name__.lower() for k in flat]
        counts = {}
        for k, nm in zip(flat, names):
            if names.count(nm) > 1:
                counts[nm] = counts.get(nm, 0) + 1
                nm = "%s_%d" % (nm, counts[nm])
            object.__setattr__(self, nm, k)

    def _children(self):
        for i, k in enumerate(self.kern_list):
            yield "kern_list[%d]" % i, k

    def K(self, X, X2=None):
        return sum(k.K(X, X2) for k in self.kern_list)

    def Kdiag(self, X):
        return sum(k.Kdiag(X) for k in self.kern_list)

    def Kzx(self, Z, X):
        out = None
        for k in self.kern_list:
            if isinstance(k, White):
                continue  # White.K(Z, X) == 0
            v = k.Kzx(Z, X)
            out = v if out is None else out + v
        if out is None:
            out = torch.zeros(Z.shape[0], X.shape[0], dtype=X.dtype, device=X.device)
        return out

    def Kzz(self, Z):
        return sum(k.Kzz(Z) for k in self.kern_list)

    @property
    def inducing_outputs(self):
        return max(getattr(k, "inducing_outputs", 1) for k in self.kern_list)
