"""kernels.svgp_terms: the fused autograd node (ops.kernel_terms: Kzx, Kdiag, Kzz of one kernel, one packed gradient
buffer) against the three separate nodes it replaces -- values and every parameter gradient, both dtypes, with and
without White members and patch weights, planar and colour-patch kernels."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _build(kind, dtype):
    from convgp_b200 import convkernels as ck, kernels as bk
    from convgp_b200.settings import settings
    settings.float_type = dtype
    settings.device = torch.device("cuda", 0)
    if kind == "wconv+white":
        k = ck.WeightedConv(bk.RBF(25), [12, 12], [5, 5]) + bk.White(1, 1e-3)
        C = 1
    elif kind == "conv":
        k = ck.Conv(bk.RBF(9), [10, 11], [3, 3])
        C = 1
    elif kind == "conv_c3+2white":
        k = ck.Conv(bk.RBF(25), [12, 12], [5, 5], colour_channels=3) + bk.White(1, 1e-3) + bk.White(1, 2e-3)
        C = 3
    else:  # the additive colour-patch kernel of cifar.py:33-41
        base = bk.Add([bk.RBF(25, variance=1.0 + 0.1 * c, lengthscales=0.5 + 0.2 * c, active_dims=np.s_[c::3])
                       for c in range(3)])
        k = ck.WeightedColourPatchConv(base, [12, 12], [5, 5], colour_channels=3) + bk.White(1, 1e-3)
        C = 3
    conv = k.kern_list[0] if isinstance(k, bk.Add) else k
    r = np.random.RandomState(3)
    if hasattr(conv, "W"):
        conv.W = r.randn(conv.num_patches)
    H, W = conv.img_size
    X = torch.tensor(r.rand(5, H * W * C), dtype=dtype, device="cuda")
    Z = torch.tensor(r.rand(7, conv.patch_len), dtype=dtype, device="cuda", requires_grad=True)
    return k, X, Z


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-12), (torch.float32, 2e-5)])
@pytest.mark.parametrize("kind", ["wconv+white", "conv", "conv_c3+2white", "addwconv+white"])
def test_fused_terms_match_separate_nodes(kind, dtype, tol):
    from convgp_b200 import kernels as bk
    k, X, Z = _build(kind, dtype)
    free = [p.free for p in k.free_params()] + [Z]
    r = np.random.RandomState(5)
    G = torch.tensor(r.randn(7, 5), dtype=dtype, device="cuda")
    gd = torch.tensor(r.randn(5), dtype=dtype, device="cuda")
    Gu = torch.tensor(r.randn(7, 7), dtype=dtype, device="cuda")

    def run(fused):
        for t in free:
            t.grad = None
        if fused:
            kd, kzx, kmm = bk.svgp_terms(k, Z, X, jitter=1e-5)
            assert bk._fused_terms(k, Z, X, 1e-5) is not None
        else:
            kzx, kd = k.Kzx(Z, X), k.Kdiag(X)
            kmm = k.Kzz(Z) + torch.eye(7, dtype=dtype, device="cuda") * 1e-5
        ((kzx * G).sum() + (kd * gd).sum() + (kmm * Gu).sum()).backward()
        return [kzx.detach(), kd.detach(), kmm.detach()] + [t.grad.clone() for t in free]

    a, b = run(True), run(False)
    assert len(a) == len(b)
    for x, y in zip(a, b):
        den = float(y.abs().max()) or 1.0
        assert float((x - y).abs().max()) / den < tol


def test_weighted_conv_rbf_is_not_fused():
    """WeightedConvRBF weights only Kzx (convkernels.py:228-248): its three terms take different arguments."""
    from convgp_b200 import convkernels as ck, kernels as bk
    from convgp_b200.settings import settings
    settings.float_type = torch.float64
    settings.device = torch.device("cuda", 0)
    k = ck.WeightedConvRBF([10, 10], [3, 3])
    X = torch.rand(3, 100, dtype=torch.float64, device="cuda")
    Z = torch.rand(4, 9, dtype=torch.float64, device="cuda")
    assert bk._fused_terms(k, Z, X, 0.0) is None
    kd, kzx, kmm = bk.svgp_terms(k, Z, X, jitter=0.0)
    assert kzx.shape == (4, 3) and kd.shape == (3,) and kmm.shape == (4, 4)
