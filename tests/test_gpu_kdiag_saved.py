"""Kdiag forward-with-residuals / O(N P) backward (cgp_kdiag_fwd_saved, cgp_kdiag_bwd_saved) against the two-pass
pair (cgp_kdiag_fwd, cgp_kdiag_bwd) through the C ABI, at the full MNIST shape and at ragged small shapes, and
against the NumPy oracle / torch twin.  Tolerances: BASELINE.json (1e-10 float64, 1e-4 float32, relative to max|.|)."""
import numpy as np
import pytest
import torch

from tests.util import Case, rel_err

pytestmark = pytest.mark.gpu


def _both_ways(case, dtype):
    from convgp_b200 import _lib
    from convgp_b200.hotpath import HotPath
    from convgp_b200.kernels import as_tensor
    k = case.package_kernel()
    spec = k._spec(dtype) if hasattr(k, "_spec") else None
    if spec is None:
        spec = _lib.Spec(dtype, case.H, case.W, case.C, case.ph, case.pw, _lib.LAYOUT_PLANAR, _lib.OUT_SUM, 1, False,
                         False)
    v, l = case.groups[0][0], case.groups[0][1]
    t = lambda a: as_tensor(np.asarray(a, dtype=np.float64), dtype)  # noqa: E731
    W = t(np.ones(case.wshape()) if case.Wv is None else case.Wv)
    hp = HotPath(spec, t(case.X), t(case.Z), W, t([v]), t([l]), t(case.G), t(case.gd), t(case.Gu))
    assert hp.kd_saved is not None, "configuration should be covered by the saved-residual path"
    L, p, st = _lib.lib(), _lib.ptr, _lib.stream_ptr
    # saved pair
    hp.grads.zero_()
    hp._kdiag_fwd(st())
    hp._kdiag_bwd(st())
    torch.cuda.synchronize()
    a = dict(kd=hp.Kdiag.clone(), dW=hp.dW.clone(), dvar=hp.dvar.clone(), dls=hp.dls.clone())
    # two-pass pair
    hp.grads.zero_()
    hp.Kdiag.zero_()
    _lib.check(L.cgp_kdiag_fwd(spec.ref(), hp.N, p(hp.X), p(hp.W), p(hp.variance), p(hp.lengthscales), p(hp.Kdiag),
                               p(hp.ws), hp.ws_bytes, st()))
    _lib.check(L.cgp_kdiag_bwd(spec.ref(), hp.N, p(hp.X), p(hp.W), p(hp.variance), p(hp.lengthscales), p(hp.gd),
                               p(hp.dW), p(hp.dvar), p(hp.dls), p(hp.ws), hp.ws_bytes, st()))
    torch.cuda.synchronize()
    b = dict(kd=hp.Kdiag.clone(), dW=hp.dW.clone(), dvar=hp.dvar.clone(), dls=hp.dls.clone())
    return {k_: v_.double().cpu().numpy() for k_, v_ in a.items()}, {k_: v_.double().cpu().numpy() for k_, v_ in b.items()}


CASES = [
    Case("saved_mnist_small", "conv", 28, 28, 1, 5, 5, [(0.9, 1.3, None, False)], N=5, M=8, seed=301),
    Case("saved_rect3x3", "conv", 28, 28, 1, 3, 3, [(1.2, 0.8, None, False)], N=3, M=8, seed=302),
    Case("saved_ragged", "conv", 13, 11, 1, 3, 3, [(0.7, 1.1, None, False)], N=7, M=5, seed=303),
    Case("saved_unweighted", "conv", 28, 28, 1, 5, 5, [(1.0, 1.0, None, False)], weighted=False, N=2, M=4, seed=304),
]


@pytest.mark.parametrize("case", CASES, ids=[c.name for c in CASES])
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
def test_saved_pair_matches_two_pass_and_oracle(case, dtype):
    tol = 1e-10 if dtype == torch.float64 else 1e-4
    a, b = _both_ways(case, dtype)
    for key in ("kd", "dW", "dvar", "dls"):
        assert rel_err(a[key], b[key]) < (tol if key in ("kd", "dW") else 10 * tol), key
    ok = case.oracle_kernel()
    assert rel_err(a["kd"], ok.Kdiag(case.X)) < tol
    # gradients of <gd, Kdiag> alone from the torch twin.  The CPU twin on the GPU box's host was seen to deviate
    # by ~1e-9 in about one fresh process out of twelve (bit-identical CUDA results): a
    # disagreement is therefore re-checked once against a second evaluation of the twin before it counts.
    import oracle.torch_twin as T
    g = case.geom

    def twin():
        X = torch.tensor(case.X)
        Wt = torch.tensor(np.ones(case.wshape()) if case.Wv is None else case.Wv, requires_grad=True)
        vt = torch.tensor(float(case.groups[0][0]), dtype=torch.float64, requires_grad=True)
        lt = torch.tensor(float(case.groups[0][1]), dtype=torch.float64, requires_grad=True)
        kd = T.Kdiag(X, [(vt, lt, None)], g, Wt)
        (kd * torch.tensor(case.gd)).sum().backward()
        return Wt.grad.numpy(), vt.grad.numpy(), lt.grad.numpy()

    for attempt in range(2):
        dW, dv, dl = twin()
        errs = (rel_err(a["dW"], dW) / tol, rel_err(a["dvar"], dv) / (10 * tol), rel_err(a["dls"], dl) / (10 * tol))
        if max(errs) < 1.0:
            break
    assert max(errs) < 1.0, errs


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
def test_saved_pair_full_mnist_shape(dtype):
    """BASELINE config 3 at full size (N=200, P=576): the two routes agree; repeated to catch ordering races."""
    tol = 1e-10 if dtype == torch.float64 else 1e-4
    case = Case("saved_full", "conv", 28, 28, 1, 5, 5, [(1.0, 1.0, None, False)], N=200, M=16, seed=305)
    for _ in range(3):
        a, b = _both_ways(case, dtype)
        for key in ("kd", "dW", "dvar", "dls"):
            assert rel_err(a[key], b[key]) < (tol if key in ("kd", "dW") else 10 * tol), key
