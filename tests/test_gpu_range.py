"""The documented range of the float32 default family (umma16.cuh: two fp16 terms per operand): c2 |patch|^2 must stay
below fp16's 65504.  Outside it the result is NaN -- never a silently wrong finite number -- and the 3xTF32 family
(options *_f32 = 2) computes the same inputs correctly."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _terms(ls, fam):
    from convgp_b200 import _lib, convkernels as ck, kernels as bk
    from convgp_b200.settings import settings
    names = ("kuf_fwd_f32", "kuf_bwd_f32", "kdiag_fwd_f32", "kdiag_bwd_f32")
    for n in names:
        _lib.set_option(n, fam)
    try:
        out = {}
        for dt in (torch.float32, torch.float64):
            settings.float_type = dt
            settings.device = torch.device("cuda", 0)
            k = ck.WeightedConv(bk.RBF(25, lengthscales=ls), [28, 28], [5, 5])
            r = np.random.RandomState(0)
            X = r.rand(6, 784)
            k.W = r.randn(576)
            Z = ck.Conv(bk.RBF(25), [28, 28], [5, 5]).compute_patches(X).reshape(-1, 25)[::131][:20] + 1e-4 * r.rand(20, 25)
            Xt = torch.tensor(X, dtype=dt, device="cuda")
            Zt = torch.tensor(Z, dtype=dt, device="cuda", requires_grad=True)
            kuf = k.Kzx(Zt, Xt)
            kuf.sum().backward()
            out[dt] = (kuf.detach().double().cpu().numpy(), Zt.grad.double().cpu().numpy())
        return out
    finally:
        for n in names:
            _lib.set_option(n, -1)


def test_inside_the_range_both_families_agree_with_float64():
    for fam in (3, 2):
        o = _terms(0.3, fam)  # c2 |patch|^2 up to ~ 200
        for (a, b), tol in zip(zip(o[torch.float32], o[torch.float64]), (1e-4, 2e-3)):  # Kuf, dZ (cancels at small l)
            assert np.isfinite(a).all()
            assert np.abs(a - b).max() <= tol * np.abs(b).max()


def test_outside_the_range_is_nan_not_a_wrong_number():
    """l = 0.004 on [0, 1] pixels: c2 |patch|^2 reaches ~ 1e6 > 65504.  (At such exponents float32 itself resolves the
    expanded distance |x|^2 - 2 x.z + |z|^2 only to ~ 0.1, for the reference's float32 TensorFlow graph as for the 3xTF32
    family: the comparison with float64 is correspondingly loose.)"""
    o16 = _terms(0.004, 3)
    kuf16, kuf64 = o16[torch.float32][0], o16[torch.float64][0]
    assert np.isfinite(kuf64).all() and np.abs(kuf64).max() > 0.0
    wrong = np.isfinite(kuf16) & (np.abs(kuf16 - kuf64) > 0.5 * np.abs(kuf64).max())
    assert not wrong.any()          # no finite entry is far off ...
    assert np.isnan(kuf16).any()    # ... and the overflow is visible
    o32 = _terms(0.004, 2)   # the 3xTF32 family has float32's range
    kuf32 = o32[torch.float32][0]
    assert np.isfinite(kuf32).all() and np.isfinite(o32[torch.float32][1]).all()
    assert np.abs(kuf32 - kuf64).max() <= 0.5 * np.abs(kuf64).max()
