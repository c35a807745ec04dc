"""Parity against golden vectors produced by the REFERENCE'S OWN SOURCE (tests/golden/ref/*.npz, made by
tests/golden/make_ref_golden.py from /root/reference/convgp/*.py executed on the stand-ins in oracle/ref_shim/).

 - CPU (`-m "not gpu"`): the NumPy oracle and its torch twin reproduce the reference's outputs -- this is what
   pins the oracle;
 - GPU (`-m gpu`): the CUDA path, called through the public classes over the C-ABI, reproduces them within
   BASELINE.json's tolerances (1e-10 float64, 1e-4 float32, relative to max|.| of each array)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import convgp_oracle as O
from oracle import svgp_oracle as SO
from tests.golden.make_ref_golden import kernel_cases
from tests.util import rel_err

REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref")
CASES = kernel_cases()
EXTRA = ["ref_full_K", "ref_convrbf", "ref_conditional", "ref_models"]


def _load(case):
    g = dict(np.load(os.path.join(REF_DIR, case.name + ".npz")))
    if "X" in g:
        case.X, case.Z, case.G, case.gd, case.Gu = g["X"], g["Z"], g["G"], g["gd"], g["Gu"]
        case.Wv = g["W"] if case.weighted else None
    else:  # full-M fixtures store outputs only: the Case regenerates the same seeded inputs
        assert np.allclose([case.X.sum(), case.Z.sum(), case.G.sum(), case.Gu.sum()], g["x_checksum"], rtol=1e-14)
    return g


def _kuu_err(kuu, g):
    return rel_err(kuu[::7, ::5], g["kuu_sub"]) if "kuu_sub" in g else rel_err(kuu, g["kuu"])


def _extra(name):
    return dict(np.load(os.path.join(REF_DIR, name + ".npz")))


def test_every_case_has_a_fixture():
    assert sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(REF_DIR, "*.npz"))) == \
        sorted([c.name for c in CASES] + EXTRA)


# ------------------------------------------------------------------------------------------- CPU: oracle
@pytest.mark.parametrize("case", CASES, ids=[c.name for c in CASES])
def test_oracle_matches_reference_source(case):
    g = _load(case)
    ok = case.oracle_kernel()
    assert rel_err(ok.Kzx(case.Z, case.X), g["kuf"]) < 1e-12
    assert rel_err(ok.Kdiag(case.X), g["kdiag"]) < 1e-12
    assert _kuu_err(ok.Kzz(case.Z), g) < 1e-12
    assert np.array_equal(ok._get_patches(case.X[:1]), g["patches"])  # layout and values, bitwise
    ref = case.twin_grads()  # autodiff through OUR restatement vs autodiff through the reference's code
    assert rel_err(ref["dZ"], g["dZ"]) < 1e-11
    if case.weighted:
        assert rel_err(ref["dW"], g["dW"]) < 1e-11
    assert rel_err(np.concatenate([np.ravel(v) for v in ref["dvar"]]), g["dvar"]) < 1e-11
    assert rel_err(np.concatenate([np.ravel(v) for v in ref["dls"]]), g["dls"]) < 1e-11


def test_oracle_full_K_matches_reference_source():
    g = _extra("ref_full_K")
    wk = O.WeightedConv(O.RBF(9, 0.8, 1.7), [8, 8], [3, 3])
    wk.W = g["W"]
    assert rel_err(wk.K(g["X"]), g["K_wconv"]) < 1e-12
    assert rel_err(wk.K(g["X"], g["X2"]), g["K_wconv_x2"]) < 1e-12
    assert rel_err(wk.Kdiag(g["X"]), g["Kdiag_wconv"]) < 1e-12
    assert rel_err(O.Conv(O.RBF(9, 0.8, 1.7), [8, 8], [3, 3]).K(g["X"]), g["K_conv"]) < 1e-12
    mk = O.WeightedMultiChannelConvGP(O.RBF(4, 1.2, 0.9), [6, 6], [2, 2], 2)
    mk.W = g["Wc"]
    assert rel_err(mk.K(g["Xc"]), g["K_multi"]) < 1e-12
    assert rel_err(mk.Kdiag(g["Xc"]), g["Kdiag_multi"]) < 1e-12


def test_oracle_convrbf_matches_reference_source():
    # the reference evaluates xTx / xTz in float32 (convkernels.py:208-219): agreement is float32-level
    g = _extra("ref_convrbf")
    kr, wkr = O.ConvRBF([28, 28], [5, 5]), O.WeightedConvRBF([28, 28], [5, 5])
    for k in (kr, wkr):
        k.basekern.variance, k.basekern.lengthscales = 0.6, 1.9
    wkr.W = g["W"]
    assert rel_err(kr.Kzx(g["Z"], g["X"]), g["kzx_convrbf"]) < 1e-5
    assert rel_err(wkr.Kzx(g["Z"], g["X"]), g["kzx_wconvrbf"]) < 1e-5


def test_oracle_conditional_matches_reference_source():
    g = _extra("ref_conditional")
    for whiten in (True, False):
        for name, q in (("q3", g["q_sqrt3"]), ("q2", g["q_sqrt2"]), ("none", None)):
            mu, var = O.conditional(g["Kdiag"], g["Kmn"], g["Kmm"], g["f"], q, whiten)
            assert rel_err(mu, g["mu_%s_w%d" % (name, whiten)]) < 1e-12
            assert rel_err(var, g["var_%s_w%d" % (name, whiten)]) < 1e-12


def test_oracle_models_match_reference_source():
    from scipy.linalg import solve_triangular
    g = _extra("ref_models")
    # MultiOutputInducingSVGP (misvgp.py:187-210)
    k = O.WeightedMultiChannelConvGP(O.RBF(9, 0.9, 1.2), [7, 7], [3, 3], 3)
    k.W = g["mo_W"]
    mu, var = SO.multioutput_predict(k, g["mo_Z"], g["mo_X"], g["mo_q_mu"], g["mo_q_sqrt"], 2)
    assert rel_err(mu, g["mo_mu"]) < 1e-11 and rel_err(var, g["mo_var"]) < 1e-11
    # FullSVSumGP (svsumgp.py:74-98)
    X, M = g["sum_X"], g["sum_Z1"].shape[0]
    k1 = O.WeightedConv(O.RBF(9, 1.0, 1.1), [8, 8], [3, 3])
    k1.W = g["sum_W"]
    k2 = O.Add([O.RBF(64, 1.1, 2.0), O.White(64, 1e-2)])
    A = np.concatenate([solve_triangular(np.linalg.cholesky(kk.Kzz(Zk) + 1e-5 * np.eye(M)), kk.Kzx(Zk, X), lower=True)
                        for kk, Zk in ((k1, g["sum_Z1"]), (k2, g["sum_Z2"]))], 0)
    fvar = k1.Kdiag(X) + k2.Kdiag(X) - (A ** 2).sum(0)
    Lq = np.tril(np.transpose(g["full_q_sqrt"], (2, 0, 1)))
    LTA = np.matmul(np.transpose(Lq, (0, 2, 1)), A[None])
    var = (fvar[None] + (LTA ** 2).sum(1)).T
    mu = A.T @ g["full_q_mu"]
    assert rel_err(mu, g["full_mu"]) < 1e-11 and rel_err(var, g["full_var"]) < 1e-11
    elbo = SO.varexp_gaussian(mu, var, g["sum_Y"], 0.3).sum() - O.gauss_kl_white(g["full_q_mu"], g["full_q_sqrt"])
    assert abs(elbo - float(g["full_elbo"])) < 1e-10 * abs(elbo)
    # SVConvGP (svconvgp.py:106-122)
    mu, var = SO.svconvgp_predict_bruteforce(O.RBF(25, 0.7, 1.3), g["svc_Z"], g["svc_X"], g["svc_q_mu"],
                                             g["svc_q_sqrt"], [28, 28], 5)
    assert rel_err(mu, g["svc_mu"]) < 1e-11 and rel_err(var, g["svc_var"]) < 1e-10
    assert np.array_equal(O.Conv(O.RBF(25), [28, 28], [5, 5])._get_patches(g["svc_X"])[:, 0, :], g["svc_Zinit"])
    # SVColourConvGP (svconvgp.py:47-63)
    k = O.WeightedConv(O.RBF(9, 0.8, 1.2), [8, 8], [3, 3])
    k.W = g["scc_W"]
    mu, var = SO.svcolourconvgp_predict(k, g["scc_Z"], g["scc_X"], g["scc_q_mu"], g["scc_q_sqrt"], g["scc_wc"], 2, 3)
    assert rel_err(mu, g["scc_mu"]) < 1e-11 and rel_err(var, g["scc_var"]) < 1e-11


# ------------------------------------------------------------------------------------------- GPU: CUDA path
@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[c.name for c in CASES])
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32], ids=["f64", "f32"])
def test_cuda_matches_reference_source(case, dtype):
    tol = 1e-10 if dtype == torch.float64 else 1e-4
    g = _load(case)
    got = case.package_grads(dtype)
    assert rel_err(got["kuf"], g["kuf"]) < tol
    assert rel_err(got["kdiag"], g["kdiag"]) < tol
    assert _kuu_err(got["kuu"], g) < tol
    assert rel_err(got["dZ"], g["dZ"]) < tol
    if case.weighted:
        assert rel_err(got["dW"], g["dW"]) < tol
    assert rel_err(np.concatenate([np.ravel(v) for v in got["dvar"]]), g["dvar"]) < 10 * tol
    assert rel_err(np.concatenate([np.ravel(v) for v in got["dls"]]), g["dls"]) < 10 * tol
    k = case.package_kernel()
    assert np.array_equal(k.compute_patches(case.X[:1]), g["patches"])
    assert rel_err(k.compute_Kzx(case.Z, case.X), g["kuf"]) < 1e-10


def _pkg():
    from convgp_b200 import convkernels as ck, kernels as bk, likelihoods as lik, misvgp, svgp, svsumgp, svconvgp
    return ck, bk, lik, misvgp, svgp, svsumgp, svconvgp


@pytest.mark.gpu
def test_cuda_full_K_and_convrbf_match_reference_source():
    ck, bk, _, misvgp, _, _, _ = _pkg()
    g = _extra("ref_full_K")
    wk = ck.WeightedConv(bk.RBF(9, variance=0.8, lengthscales=1.7), [8, 8], [3, 3])
    wk.W = g["W"]
    assert rel_err(wk.compute_K_symm(g["X"]), g["K_wconv"]) < 1e-10
    assert rel_err(wk.compute_K(g["X"], g["X2"]), g["K_wconv_x2"]) < 1e-10
    assert rel_err(wk.compute_Kdiag(g["X"]), g["Kdiag_wconv"]) < 1e-10
    k = ck.Conv(bk.RBF(9, variance=0.8, lengthscales=1.7), [8, 8], [3, 3])
    assert rel_err(k.compute_K_symm(g["X"]), g["K_conv"]) < 1e-10
    mk = misvgp.WeightedMultiChannelConvGP(bk.RBF(4, variance=1.2, lengthscales=0.9), [6, 6], [2, 2], 2)
    mk.W = g["Wc"]
    assert rel_err(mk.compute_K_symm(g["Xc"]), g["K_multi"]) < 1e-10
    assert rel_err(mk.compute_Kdiag(g["Xc"]), g["Kdiag_multi"]) < 1e-10
    g = _extra("ref_convrbf")
    kr, wkr = ck.ConvRBF([28, 28], [5, 5]), ck.WeightedConvRBF([28, 28], [5, 5])
    for k in (kr, wkr):
        k.basekern.variance, k.basekern.lengthscales = 0.6, 1.9
    wkr.W = g["W"]
    # the reference's own criterion for this pair is np.allclose (testing/test_convkerns.py:33-36): its conv2d is float32
    assert rel_err(kr.compute_Kzx(g["Z"], g["X"]), g["kzx_convrbf"]) < 1e-5
    assert rel_err(wkr.compute_Kzx(g["Z"], g["X"]), g["kzx_wconvrbf"]) < 1e-5


@pytest.mark.gpu
def test_cuda_conditional_matches_reference_source():
    from convgp_b200.misvgp import conditional
    from convgp_b200.kernels import as_tensor
    g = _extra("ref_conditional")
    t = lambda a: as_tensor(a, torch.float64)  # noqa: E731
    for whiten in (True, False):
        for name, q in (("q3", g["q_sqrt3"]), ("q2", g["q_sqrt2"]), ("none", None)):
            mu, var = conditional(None, None, t(g["Kdiag"]), t(g["Kmn"]), t(g["Kmm"]), t(g["f"]), full_cov=False,
                                  q_sqrt=None if q is None else t(q), whiten=whiten)
            assert rel_err(mu.cpu().numpy(), g["mu_%s_w%d" % (name, whiten)]) < 1e-10
            assert rel_err(var.cpu().numpy(), g["var_%s_w%d" % (name, whiten)]) < 1e-10


@pytest.mark.gpu
def test_cuda_models_match_reference_source():
    ck, bk, lik, misvgp, _, svsumgp, svconvgp = _pkg()
    g = _extra("ref_models")
    # cfg 4 structure: FullSVSumGP / MeanFieldSVSumGP
    def kerns():
        k1 = ck.WeightedConv(bk.RBF(9, variance=1.0, lengthscales=1.1), [8, 8], [3, 3])
        k1.W = g["sum_W"]
        return k1, bk.RBF(64, variance=1.1, lengthscales=2.0) + bk.White(64, 1e-2)

    k1, k2 = kerns()
    full = svsumgp.FullSVSumGP(g["sum_X"], g["sum_Y"], k1, k2, lik.Gaussian(0.3), g["sum_Z1"], g["sum_Z2"],
                               num_latent=2)
    full.q_mu, full.q_sqrt = g["full_q_mu"], g["full_q_sqrt"]
    mu, var = full.predict_f(g["sum_X"])
    assert rel_err(mu, g["full_mu"]) < 1e-9 and rel_err(var, g["full_var"]) < 1e-9
    assert abs(full.compute_log_likelihood() - float(g["full_elbo"])) < 1e-9 * abs(float(g["full_elbo"]))
    k1, k2 = kerns()
    mf = svsumgp.MeanFieldSVSumGP(g["sum_X"], g["sum_Y"], k1, k2, lik.Gaussian(0.3), g["sum_Z1"], g["sum_Z2"],
                                  num_latent=2)
    mf.q_mu, mf.q_sqrt = g["mf_q_mu"], g["mf_q_sqrt"]
    mu, var = mf.predict_f(g["sum_X"])
    assert rel_err(mu, g["mf_mu"]) < 1e-9 and rel_err(var, g["mf_var"]) < 1e-9
    assert abs(mf.compute_log_likelihood() - float(g["mf_elbo"])) < 1e-9 * abs(float(g["mf_elbo"]))
    # cifar 'multi'
    k = misvgp.WeightedMultiChannelConvGP(bk.RBF(9, variance=0.9, lengthscales=1.2), [7, 7], [3, 3], 3)
    k.W = g["mo_W"]
    m = misvgp.MultiOutputInducingSVGP(g["mo_X"], g["mo_Y"], k, lik.Gaussian(1.0), g["mo_Z"], num_latent=2)
    m.q_mu, m.q_sqrt = g["mo_q_mu"], g["mo_q_sqrt"]
    mu, var = m.predict_f(g["mo_X"])
    assert rel_err(mu, g["mo_mu"]) < 1e-9 and rel_err(var, g["mo_var"]) < 1e-9
    # SVConvGP
    m = svconvgp.SVConvGP(g["svc_X"], g["svc_Y"], bk.RBF(25, variance=0.7, lengthscales=1.3), lik.Gaussian(0.1),
                          g["svc_Z"], 5)
    m.q_mu, m.q_sqrt = g["svc_q_mu"], g["svc_q_sqrt"]
    mu, var = m.predict_f(g["svc_X"])
    assert rel_err(mu, g["svc_mu"]) < 1e-9 and rel_err(var, g["svc_var"]) < 1e-8
    assert abs(m.compute_log_likelihood() - float(g["svc_elbo"])) < 1e-8 * abs(float(g["svc_elbo"]))
    assert rel_err(m.compute_patch_predictions(g["svc_X"]), g["svc_patch_mu"]) < 1e-9
    # SVColourConvGP (svconvgp.py:25-63)
    k = ck.WeightedConv(bk.RBF(9, variance=0.8, lengthscales=1.2), [8, 8], [3, 3])
    k.W = g["scc_W"]
    m = svconvgp.SVColourConvGP(g["scc_X"], g["scc_Y"], k, lik.Gaussian(1.0), g["scc_Z"], 3, num_latent=2)
    m.wc = g["scc_wc"]
    m.q_mu, m.q_sqrt = g["scc_q_mu"], g["scc_q_sqrt"]
    mu, var = m.predict_f(g["scc_X"])
    assert rel_err(mu, g["scc_mu"]) < 1e-9 and rel_err(var, g["scc_var"]) < 1e-9
