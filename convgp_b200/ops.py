"""torch.autograd wrappers around the C ABI: the only place PyTorch meets the CUDA library.
PyTorch is the tensor carrier (device memory, streams, autograd tape); all arithmetic of the hot
path happens in libconvgp_b200.so.  Gradients follow SURVEY.md section 8a."""
import torch

from . import _lib
from ._lib import check, ptr, stream_ptr


def _c(t):
    """Contiguous, 16-byte aligned CUDA tensor (None passes through).  Row slices of a float32 tensor with an odd
    row length keep their storage offset under .contiguous() and can start on a 4-byte boundary; the kernels use
    8/16-byte vector loads on X and Z, so such views are copied once here instead of failing in the C ABI."""
    if t is None:
        return None
    if not t.is_cuda:
        raise _lib.CgpError(-3, "tensor is not on a CUDA device (convgp_b200 has no CPU path)")
    t = t.contiguous()
    if t.data_ptr() % 16 != 0 and t.numel() > 0:
        t = t.clone()
    return t


def _numel(t):
    return 0 if t is None else int(t.numel())


def validate(spec, X=None, Z=None, W=None, variance=None, lengthscales=None, X2=None):
    """Shape / dtype checks of everything that crosses the C ABI as a bare pointer (the C side sees only addresses
    and cannot check; the reference's TF graph raises a shape error in all of these cases).  Raises ValueError /
    TypeError -- before any launch, so a wrong argument can never become an out-of-bounds device read."""
    dt = spec.torch_dtype
    HWC = spec.desc.H * spec.desc.W * spec.desc.C
    for name, t in (("X", X), ("X2", X2), ("Z", Z), ("W", W), ("variance", variance), ("lengthscales", lengthscales)):
        if t is None:
            continue
        if t.dtype != dt:
            raise TypeError("%s is %s but the kernel was built for %s (settings.float_type changed?)"
                            % (name, t.dtype, dt))
    for name, t in (("X", X), ("X2", X2)):
        if t is not None and (t.dim() != 2 or t.shape[1] != HWC):
            raise ValueError("%s must be (N, H*W*C) = (N, %d) for a %dx%dx%d image kernel, got %s"
                             % (name, HWC, spec.desc.H, spec.desc.W, spec.desc.C, tuple(t.shape)))
    if Z is not None and (Z.dim() != 2 or Z.shape[1] != spec.D):
        raise ValueError("Z must be (M, patch_len) = (M, %d), got %s" % (spec.D, tuple(Z.shape)))
    if W is not None and _numel(W) != spec.P:
        raise ValueError("W must have num_patches = %d entries, got %s" % (spec.P, tuple(W.shape)))
    if variance is not None and _numel(variance) != spec.n_groups:
        raise ValueError("variance must have n_groups = %d entries, got %d" % (spec.n_groups, _numel(variance)))
    if lengthscales is not None:
        want = spec.n_groups * (spec.D if spec.ard else 1)
        if _numel(lengthscales) != want:
            raise ValueError("lengthscales must have %d entries, got %d" % (want, _numel(lengthscales)))


def _grad_buffers(spec, Z, W, variance, lengthscales):
    """One zeroed, packed gradient buffer [dZ | dW | dvariance | dlengthscales] and its views.
    The C ABI accumulates into it, and it is what the data-parallel all-reduce sums."""
    sizes = [0 if Z is None else Z.numel(), 0 if W is None else W.numel(), variance.numel(), lengthscales.numel()]
    q = 16 // variance.element_size()  # every segment starts 16-byte aligned
    offs, o = [], 0
    for s in sizes:
        offs.append(o)
        o += (s + q - 1) // q * q
    buf = torch.zeros(max(o, q), dtype=variance.dtype, device=variance.device)
    views = [None if ref is None else buf[o:o + s].view(ref.shape)
             for o, s, ref in zip(offs, sizes, (Z, W, variance, lengthscales))]
    return buf, views


class _Kuf(torch.autograd.Function):
    @staticmethod
    def forward(ctx, spec, X, Z, W, variance, lengthscales):
        validate(spec, X=X, Z=Z, W=W, variance=variance, lengthscales=lengthscales)
        X, Z, W, variance, lengthscales = map(_c, (X, Z, W, variance, lengthscales))
        N, M = X.shape[0], Z.shape[0]
        shape = (M, N, spec.C) if spec.per_channel else (M, N)
        out = torch.empty(shape, dtype=X.dtype, device=X.device)
        with torch.cuda.device(X.device):
            ws, nb = spec.workspace(N, M, X.device)
            check(_lib.lib().cgp_kuf_fwd(spec.ref(), N, M, ptr(X), ptr(Z), ptr(W), ptr(variance), ptr(lengthscales),
                                         ptr(out), ptr(ws), nb, stream_ptr()))
        ctx.spec = spec
        ctx.save_for_backward(X, Z, W, variance, lengthscales)
        return out

    @staticmethod
    def backward(ctx, G):
        X, Z, W, variance, lengthscales = ctx.saved_tensors
        spec = ctx.spec
        N, M = X.shape[0], Z.shape[0]
        _, (dZ, dW, dvar, dls) = _grad_buffers(spec, Z, W, variance, lengthscales)
        G = _c(G)
        with torch.cuda.device(X.device):
            ws, nb = spec.workspace(N, M, X.device)
            check(_lib.lib().cgp_kuf_bwd(spec.ref(), N, M, ptr(X), ptr(Z), ptr(W), ptr(variance), ptr(lengthscales),
                                         ptr(G), ptr(dZ), ptr(dW), ptr(dvar), ptr(dls), ptr(ws), nb, stream_ptr()))
        return None, None, dZ, dW, dvar, dls


class _Kdiag(torch.autograd.Function):
    """Kdiag and its gradients.  When a gradient will be asked for and the configuration is covered
    (``cgp_kdiag_saved_elems`` > 0), the forward is the one that also stores the per-image row sums its backward is
    made of (N * (P + 2) elements) and the backward is an O(N P) kernel; otherwise the backward recomputes."""

    @staticmethod
    def forward(ctx, spec, X, W, variance, lengthscales):
        needs_grad = any(ctx.needs_input_grad[2:5])
        validate(spec, X=X, W=W, variance=variance, lengthscales=lengthscales)
        X, W, variance, lengthscales = map(_c, (X, W, variance, lengthscales))
        N = X.shape[0]
        shape = (N, spec.C) if spec.per_channel else (N,)
        out = torch.empty(shape, dtype=X.dtype, device=X.device)
        saved = None
        with torch.cuda.device(X.device):
            L = _lib.lib()
            n_saved = L.cgp_kdiag_saved_elems(spec.ref(), N) if needs_grad and N > 0 else 0
            if n_saved:
                saved = torch.empty(n_saved, dtype=X.dtype, device=X.device)
                check(L.cgp_kdiag_fwd_saved(spec.ref(), N, ptr(X), ptr(W), ptr(variance), ptr(lengthscales), ptr(out),
                                            ptr(saved), n_saved, stream_ptr()))
            else:
                ws, nb = spec.workspace(N, 0, X.device)
                check(L.cgp_kdiag_fwd(spec.ref(), N, ptr(X), ptr(W), ptr(variance), ptr(lengthscales), ptr(out),
                                      ptr(ws), nb, stream_ptr()))
        ctx.spec = spec
        ctx.has_saved = saved is not None
        ctx.save_for_backward(X, W, variance, lengthscales, saved)
        return out

    @staticmethod
    def backward(ctx, g):
        X, W, variance, lengthscales, saved = ctx.saved_tensors
        spec = ctx.spec
        N = X.shape[0]
        _, (_, dW, dvar, dls) = _grad_buffers(spec, None, W, variance, lengthscales)
        g = _c(g)
        with torch.cuda.device(X.device):
            if ctx.has_saved:
                check(_lib.lib().cgp_kdiag_bwd_saved(spec.ref(), N, ptr(variance), ptr(lengthscales), ptr(g),
                                                     ptr(saved), saved.numel(), ptr(dW), ptr(dvar), ptr(dls),
                                                     stream_ptr()))
            else:
                ws, nb = spec.workspace(N, 0, X.device)
                check(_lib.lib().cgp_kdiag_bwd(spec.ref(), N, ptr(X), ptr(W), ptr(variance), ptr(lengthscales),
                                               ptr(g), ptr(dW), ptr(dvar), ptr(dls), ptr(ws), nb, stream_ptr()))
        return None, None, dW, dvar, dls


class _Kuu(torch.autograd.Function):
    @staticmethod
    def forward(ctx, spec, Z, variance, lengthscales, diag_add):
        validate(spec, Z=Z, variance=variance, lengthscales=lengthscales)
        Z, variance, lengthscales = map(_c, (Z, variance, lengthscales))
        M = Z.shape[0]
        out = torch.empty((M, M), dtype=Z.dtype, device=Z.device)
        with torch.cuda.device(Z.device):
            check(_lib.lib().cgp_kuu_fwd(spec.ref(), M, ptr(Z), ptr(variance), ptr(lengthscales), float(diag_add),
                                         ptr(out), stream_ptr()))
        ctx.spec = spec
        ctx.save_for_backward(Z, variance, lengthscales)
        return out

    @staticmethod
    def backward(ctx, Gu):
        Z, variance, lengthscales = ctx.saved_tensors
        spec = ctx.spec
        M = Z.shape[0]
        _, (dZ, _, dvar, dls) = _grad_buffers(spec, Z, None, variance, lengthscales)
        Gu = _c(Gu)
        with torch.cuda.device(Z.device):
            check(_lib.lib().cgp_kuu_bwd(spec.ref(), M, ptr(Z), ptr(variance), ptr(lengthscales), ptr(Gu), ptr(dZ),
                                         ptr(dvar), ptr(dls), stream_ptr()))
        return None, dZ, dvar, dls, None


class _KernelTerms(torch.autograd.Function):
    """Kzx(Z, X), Kdiag(X) and Kzz(Z) + diag of ONE kernel in one autograd node (what an SVGP prediction evaluates:
    misvgp.py:194-201, svsumgp.py:76-88): the three forward launches run on parallel streams, the parameter values are
    read once, and the backward accumulates all three reverse passes into one packed gradient buffer
    [dZ | dW | dvariance | dlengthscales] -- instead of three nodes with their own buffers, accumulations and
    parameter transforms (about forty small torch kernels less per evaluation).  ``diag_const`` (jitter) and
    ``diag_param`` (a 0-dim tensor: the variance of White members, or None) are added to the diagonal of Kzz, and
    ``diag_param`` to Kdiag (White.Kdiag)."""

    @staticmethod
    def forward(ctx, spec, X, Z, W, variance, lengthscales, diag_const, diag_param):
        from .kernels import concurrent
        validate(spec, X=X, Z=Z, W=W, variance=variance, lengthscales=lengthscales)
        X, Z, W, variance, lengthscales = map(_c, (X, Z, W, variance, lengthscales))
        N, M = X.shape[0], Z.shape[0]
        needs_grad = any(ctx.needs_input_grad[2:6])
        L = _lib.lib()
        Kuf = torch.empty((M, N), dtype=X.dtype, device=X.device)
        Kd = torch.empty((N,), dtype=X.dtype, device=X.device)
        Kuu = torch.empty((M, M), dtype=X.dtype, device=X.device)
        with torch.cuda.device(X.device):
            ws, nb = spec.workspace(N, M, X.device)
            n_saved = L.cgp_kdiag_saved_elems(spec.ref(), N) if needs_grad and N > 0 else 0
            saved = torch.empty(n_saved, dtype=X.dtype, device=X.device) if n_saved else None
            ws2 = None if n_saved else spec.workspace(N, 0, X.device)

            def f_kuf():
                if N > 0 and M > 0:
                    check(L.cgp_kuf_fwd(spec.ref(), N, M, ptr(X), ptr(Z), ptr(W), ptr(variance), ptr(lengthscales),
                                        ptr(Kuf), ptr(ws), nb, stream_ptr()))

            def f_kd():
                if N == 0:
                    return
                if n_saved:
                    check(L.cgp_kdiag_fwd_saved(spec.ref(), N, ptr(X), ptr(W), ptr(variance), ptr(lengthscales),
                                                ptr(Kd), ptr(saved), n_saved, stream_ptr()))
                else:
                    check(L.cgp_kdiag_fwd(spec.ref(), N, ptr(X), ptr(W), ptr(variance), ptr(lengthscales), ptr(Kd),
                                          ptr(ws2[0]), ws2[1], stream_ptr()))

            def f_kuu():
                check(L.cgp_kuu_fwd(spec.ref(), M, ptr(Z), ptr(variance), ptr(lengthscales), float(diag_const),
                                    ptr(Kuu), stream_ptr()))

            concurrent(f_kuf, f_kd, f_kuu)
        if diag_param is not None:
            Kuu.diagonal().add_(diag_param)
            Kd.add_(diag_param)
        ctx.spec, ctx.has_diag = spec, diag_param is not None
        ctx.save_for_backward(X, Z, W, variance, lengthscales, saved)
        return Kuf, Kd, Kuu

    @staticmethod
    def backward(ctx, G, gd, Gu):
        from .kernels import concurrent
        X, Z, W, variance, lengthscales, saved = ctx.saved_tensors
        spec = ctx.spec
        N, M = X.shape[0], Z.shape[0]
        _, (dZ, dW, dvar, dls) = _grad_buffers(spec, Z, W, variance, lengthscales)
        G, gd, Gu = _c(G), _c(gd), _c(Gu)
        L = _lib.lib()
        with torch.cuda.device(X.device):
            ws, nb = spec.workspace(N, M, X.device)
            ws2 = None if saved is not None else spec.workspace(N, 0, X.device)

            def b_kuf():
                if N > 0 and M > 0:
                    check(L.cgp_kuf_bwd(spec.ref(), N, M, ptr(X), ptr(Z), ptr(W), ptr(variance), ptr(lengthscales),
                                        ptr(G), ptr(dZ), ptr(dW), ptr(dvar), ptr(dls), ptr(ws), nb, stream_ptr()))

            def b_kd():
                if N == 0:
                    return
                if saved is not None:
                    check(L.cgp_kdiag_bwd_saved(spec.ref(), N, ptr(variance), ptr(lengthscales), ptr(gd), ptr(saved),
                                                saved.numel(), ptr(dW), ptr(dvar), ptr(dls), stream_ptr()))
                else:
                    check(L.cgp_kdiag_bwd(spec.ref(), N, ptr(X), ptr(W), ptr(variance), ptr(lengthscales), ptr(gd),
                                          ptr(dW), ptr(dvar), ptr(dls), ptr(ws2[0]), ws2[1], stream_ptr()))

            def b_kuu():
                check(L.cgp_kuu_bwd(spec.ref(), M, ptr(Z), ptr(variance), ptr(lengthscales), ptr(Gu), ptr(dZ),
                                    ptr(dvar), ptr(dls), stream_ptr()))

            concurrent(b_kuf, b_kd, b_kuu)  # all three accumulate (atomics) into the one packed buffer
        d_diag = (Gu.diagonal().sum() + gd.sum()) if ctx.has_diag else None
        return None, None, dZ, dW, dvar, dls, None, d_diag


def kernel_terms(spec, X, Z, W, variance, lengthscales, diag_const=0.0, diag_param=None):
    """(Kzx(Z, X), Kdiag(X) + diag_param, Kzz(Z) + (diag_const + diag_param) I) in one autograd node."""
    return _KernelTerms.apply(spec, X, Z, W, variance, lengthscales, float(diag_const), diag_param)


def kuf(spec, X, Z, W, variance, lengthscales):
    """Kzx(Z, X): (M, N) or (M, N, C)."""
    return _Kuf.apply(spec, X, Z, W, variance, lengthscales)


def kdiag(spec, X, W, variance, lengthscales):
    """Kdiag(X): (N,) or (N, C)."""
    return _Kdiag.apply(spec, X, W, variance, lengthscales)


def kuu(spec, Z, variance, lengthscales, diag_add=0.0):
    """Kzz(Z) + diag_add * I."""
    return _Kuu.apply(spec, Z, variance, lengthscales, diag_add)


def kfull(spec, X, X2, W, variance, lengthscales):
    """K(X, X2) (forward only: the reference uses it in tests and full-GP baselines, never inside
    the SVGP objective)."""
    validate(spec, X=X, X2=X2, W=W, variance=variance, lengthscales=lengthscales)
    X, X2, W, variance, lengthscales = map(_c, (X, X2, W, variance, lengthscales))
    N = X.shape[0]
    N2 = N if X2 is None else X2.shape[0]
    out = torch.empty((N, N2), dtype=X.dtype, device=X.device)
    with torch.cuda.device(X.device):
        nb = (N + N2) * spec.P * spec.D * X.element_size() + 256  # materialised patches of X and X2
        ws = torch.empty(nb, dtype=torch.uint8, device=X.device)
        check(_lib.lib().cgp_kfull_fwd(spec.ref(), N, N2, ptr(X), ptr(X2), ptr(W), ptr(variance), ptr(lengthscales),
                                       ptr(out), ptr(ws), nb, stream_ptr()))
    return out


def patches(spec, X):
    """_get_patches(X): (N, P, D)."""
    validate(spec, X=X)
    X = _c(X)
    out = torch.empty((X.shape[0], spec.P, spec.D), dtype=X.dtype, device=X.device)
    with torch.cuda.device(X.device):
        check(_lib.lib().cgp_patches(spec.ref(), X.shape[0], ptr(X), ptr(out), stream_ptr()))
    return out
