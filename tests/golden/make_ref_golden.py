"""Golden vectors produced by the REFERENCE'S OWN SOURCE (tests/golden/ref/*.npz).

Run in the build container, from the repo root:   python tests/golden/make_ref_golden.py

/root/reference/convgp/{convkernels,misvgp,svsumgp,svconvgp}.py are imported unmodified (oracle/run_reference.py)
on top of the TensorFlow-1.x / GPflow-0.x stand-ins in oracle/ref_shim/ (the real stacks do not install on this
image).  Values come out of the reference's forward code; gradients are torch autodiff THROUGH that
code (the reference has no gradient code of its own -- TF autodiff plays that role upstream).  What remains
restated rather than executed is GPflow's share (RBF / White / Add / conditional / KL / Gaussian likelihood): see
oracle/ref_shim/GPflow/__init__.py.

The fixtures travel to the GPU box; /root/reference does not.  tests/test_ref_golden.py checks the NumPy oracle
(CPU) and the CUDA path (GPU) against them.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.util import Case  # noqa: E402

S = np.s_
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref")


# ------------------------------------------------------------------------------------ kernel cases
def kernel_cases():
    return [
        # cfg 1  rectangles.py:41-44   Conv(RBF(9)) 28x28, 3x3 (X = real rectangles images, see main())
        Case("ref_cfg1_rectangles_conv", "conv", 28, 28, 1, 3, 3, [(1.0, 1.0, None, False)], weighted=False, N=3,
             M=16, seed=201),
        # rectangles.py:43 with --kernel-ard
        Case("ref_rectangles_wconv_ard", "conv", 28, 28, 1, 3, 3,
             [(0.7, np.linspace(0.6, 1.8, 9), None, True)], N=2, M=10, seed=202),
        # rectangles.py:48-56  wconv-add: RBF(9) + six 3-dim RBFs on rows / columns of the patch
        Case("ref_rectangles_wconv_add", "conv", 28, 28, 1, 3, 3,
             [(1.0, 1.0, None, False), (0.9, 0.8, S[0:3], False), (1.1, 1.2, S[3:6], False),
              (0.8, 0.9, S[6:9], False), (1.2, 1.1, S[0:7:3], False), (0.7, 1.3, S[1:8:3], False),
              (1.05, 0.95, S[2:9:3], False)], N=2, M=10, seed=203),
        # cfg 2/3  mnist01.py:45-47, mnist.py:41-42   WeightedConv(RBF(25)) 28x28, 5x5
        Case("ref_cfg3_mnist_wconv", "conv", 28, 28, 1, 5, 5, [(1.0, 1.0, None, False)], N=3, M=24, seed=204),
        # the hyper-parameters the reference's own test uses (testing/test_convkerns.py:28-31), randn inputs
        Case("ref_cfg3_mnist_wconv_l2p34", "conv", 28, 28, 1, 5, 5, [(0.001, 2.34, None, False)], N=2, M=12,
             seed=205, xdist="randn"),
        Case("ref_mnist_conv_unweighted", "conv", 28, 28, 1, 5, 5, [(1.3, 0.8, None, False)], weighted=False, N=2,
             M=12, seed=206),
        # cfg 5  cifar.py:33-41   WeightedColourPatchConv(sum_c RBF(25, active_dims=c::3, l=0.5)) 32x32x3
        Case("ref_cfg5_cifar_addwconv", "colour", 32, 32, 3, 5, 5,
             [(1.004, 0.5, S[0::3], False), (0.991, 0.5, S[1::3], False), (1.013, 0.5, S[2::3], False)],
             N=2, M=16, seed=207, xdist="uint8"),
        # cifar.py:30-31  wconv / cpwconv ; cifar.py:42-44 multi
        Case("ref_cifar_wconv_c3", "conv", 32, 32, 3, 5, 5, [(1.0, 1.0, None, False)], N=2, M=12, seed=208,
             xdist="uint8"),
        Case("ref_cifar_cpwconv", "colour", 32, 32, 3, 5, 5, [(1.1, 2.0, None, False)], N=2, M=12, seed=209,
             xdist="uint8"),
        Case("ref_cifar_multi_c3", "multi", 32, 32, 3, 5, 5, [(1.0, 1.0, None, False)], N=2, M=12, seed=210,
             xdist="uint8"),
        # BASELINE's full inducing sets (M = 750 / 1000: every row tile incl. the ragged last one) on a slice of the
        # minibatch small enough for the reference's materialised (M x N P) tensor to fit in host memory
        Case("ref_cfg3_mnist_wconv_M750", "conv", 28, 28, 1, 5, 5, [(1.0, 1.0, None, False)], N=24, M=750, seed=213),
        Case("ref_cfg5_cifar_addwconv_M1000", "colour", 32, 32, 3, 5, 5,
             [(1.004, 0.5, S[0::3], False), (0.991, 0.5, S[1::3], False), (1.013, 0.5, S[2::3], False)],
             N=4, M=1000, seed=214, xdist="uint8"),
        # small geometries of the reference's tests (testing/test_convkerns.py:44-49, 104-110)
        Case("ref_tiny_2x2_c2", "conv", 3, 3, 2, 2, 2, [(1.0, 1.0, None, False)], N=3, M=8, seed=211, xdist="randn"),
        Case("ref_tiny_2x2_colour_c2", "colour", 3, 3, 2, 2, 2, [(1.0, 1.0, None, False)], N=3, M=8, seed=212,
             xdist="randn"),
    ]


def ref_kernel(case, GP, ck, mi):
    """The reference's kernel object for a Case, built the way its scripts build it."""
    D = case.geom.D
    members = []
    for (v, l, ad, ard) in case.groups:
        dim = D if ad is None else len(np.arange(D)[ad])
        members.append(GP.kernels.RBF(dim, variance=v, lengthscales=l, active_dims=ad, ARD=ard))
    base = members[0]
    for m in members[1:]:
        base = base + m
    img, pat = [case.H, case.W], [case.ph, case.pw]
    if case.kind == "multi":
        k = mi.WeightedMultiChannelConvGP(base, img, pat, case.C)
    elif case.kind == "colour":
        k = (ck.WeightedColourPatchConv if case.weighted else ck.ColourPatchConv)(base, img, pat, case.C)
    else:
        k = (ck.WeightedConv if case.weighted else ck.Conv)(base, img, pat, case.C)
    if case.weighted:
        k.W = case.Wv
    return k, members


def run_kernel_case(case, GP, ck, mi):
    import torch
    from GPflow.param import tf_mode
    k, members = ref_kernel(case, GP, ck, mi)
    X = torch.tensor(case.X, dtype=torch.float64)
    Z = torch.tensor(case.Z, dtype=torch.float64, requires_grad=True)
    with tf_mode():
        kuf, kd, kuu = k.Kzx(Z, X), k.Kdiag(X), k.Kzz(Z)
        loss = (kuf * torch.tensor(case.G)).sum() + (kd * torch.tensor(case.gd)).sum() + \
            (kuu * torch.tensor(case.Gu)).sum()
        loss.backward()
    out = dict(
        X=case.X, Z=case.Z, W=np.ones(case.wshape()) if case.Wv is None else case.Wv, G=case.G, gd=case.gd,
        Gu=case.Gu, kuf=kuf.detach().numpy(), kdiag=kd.detach().numpy(), kuu=kuu.detach().numpy(),
        dZ=Z.grad.numpy(), dW=k.__dict__["W"].grad() if case.weighted else np.zeros(0),
        dvar=np.concatenate([np.ravel(m.__dict__["variance"].grad()) for m in members]),
        dls=np.concatenate([np.ravel(m.__dict__["lengthscales"].grad()) for m in members]))
    # the eager NumPy entry points the reference's tests use
    assert np.array_equal(k.compute_Kzx(case.Z, case.X), out["kuf"])
    out["patches"] = k.compute_patches(case.X)[:1]  # first image: pins the patch layout
    if case.M > 100:
        # full inducing sets: keep the fixture small -- the inputs are regenerated from the Case's seed at test time
        # (NumPy's legacy RandomState stream is stable), Kzz is kept on a sub-grid
        for key in ("X", "Z", "W", "G", "gd", "Gu"):
            del out[key]
        out["kuu_sub"] = out.pop("kuu")[::7, ::5].copy()
        out["x_checksum"] = np.array([case.X.sum(), case.Z.sum(), case.G.sum(), case.Gu.sum()])
    return out


# ------------------------------------------------------------------------------------ other fixtures
def run_full_K(GP, ck, mi):
    """K(X), K(X, X2): convkernels.py:64-71,157-171; misvgp.py:33-49 (symmetric only)."""
    rng = np.random.RandomState(220)
    out = {}
    X, X2 = rng.randn(4, 64), rng.randn(3, 64)
    wk = ck.WeightedConv(GP.kernels.RBF(9, variance=0.8, lengthscales=1.7), [8, 8], [3, 3])
    wk.W = rng.randn(wk.num_patches)
    out.update(X=X, X2=X2, W=wk.W.value, K_wconv=wk.compute_K_symm(X), K_wconv_x2=wk.compute_K(X, X2),
               Kdiag_wconv=wk.compute_Kdiag(X))
    k = ck.Conv(GP.kernels.RBF(9, variance=0.8, lengthscales=1.7), [8, 8], [3, 3])
    out.update(K_conv=k.compute_K_symm(X))
    Xc = rng.randn(3, 36 * 2)
    mk = mi.WeightedMultiChannelConvGP(GP.kernels.RBF(4, variance=1.2, lengthscales=0.9), [6, 6], [2, 2], 2)
    mk.W = rng.randn(2, 25)
    out.update(Xc=Xc, Wc=mk.W.value, K_multi=mk.compute_K_symm(Xc), Kdiag_multi=mk.compute_Kdiag(Xc))
    return out


def run_convrbf(GP, ck):
    """ConvRBF / WeightedConvRBF.Kzx (convkernels.py:204-248): the reference's conv2d formulation, float32 inside."""
    rng = np.random.RandomState(221)
    X, Z, W = rng.rand(4, 784), rng.rand(7, 25), rng.randn(576)
    kr, wkr = ck.ConvRBF([28, 28], [5, 5]), ck.WeightedConvRBF([28, 28], [5, 5])
    for k in (kr, wkr):
        k.basekern.variance = 0.6
        k.basekern.lengthscales = 1.9
    wkr.W = W
    return dict(X=X, Z=Z, W=W, variance=0.6, lengthscales=1.9, kzx_convrbf=kr.compute_Kzx(Z, X),
                kzx_wconvrbf=wkr.compute_Kzx(Z, X))


def run_conditional(mi):
    """misvgp.py:83-165 with pre-evaluated blocks: all four (whiten, q_sqrt rank) combinations."""
    import torch
    rng = np.random.RandomState(222)
    M, N, L = 9, 6, 3
    B = rng.randn(M, M)
    Kmm = B @ B.T + M * np.eye(M)
    Kmn = rng.randn(M, N)
    Kdiag = 5.0 + rng.rand(N) + (np.linalg.solve(np.linalg.cholesky(Kmm), Kmn) ** 2).sum(0)
    f = rng.randn(M, L)
    q3 = np.tril(0.3 * rng.randn(L, M, M) + np.eye(M)[None]).transpose(1, 2, 0)
    q2 = 0.5 + rng.rand(M, L)
    out = dict(Kmm=Kmm, Kmn=Kmn, Kdiag=Kdiag, f=f, q_sqrt3=q3, q_sqrt2=q2)
    t = lambda a: torch.tensor(a, dtype=torch.float64)  # noqa: E731
    for whiten in (True, False):
        for name, q in (("q3", q3), ("q2", q2), ("none", None)):
            mu, var = mi.conditional(None, t(np.zeros((M, 1))), t(Kdiag), t(Kmn), t(Kmm), t(f), full_cov=False,
                                     q_sqrt=None if q is None else t(q), whiten=whiten)
            out["mu_%s_w%d" % (name, whiten)] = mu.numpy()
            out["var_%s_w%d" % (name, whiten)] = var.numpy()
    return out


def _variational(rng, M, L):
    q_mu = 0.3 * rng.randn(M, L)
    q_sqrt = np.tril(0.2 * rng.randn(L, M, M) + np.eye(M)[None]).transpose(1, 2, 0)
    return q_mu, q_sqrt


def run_models(GP, ck, mi, ss, sc):
    """build_predict / bound of the reference's model classes (svsumgp.py, misvgp.py:168-210, svconvgp.py)."""
    out = {}
    # -- cfg 4 structure (sumkern_mnist.py:31-55): k1 = WeightedConv(RBF), k2 = RBF(all pixels) + White, Gaussian lik
    rng = np.random.RandomState(230)
    N, M, L = 8, 6, 2
    X, Y = rng.rand(N, 64), rng.randn(N, L)
    W = rng.rand(36) + 0.5
    Z1 = X.reshape(N, 8, 8)[:M, :3, :3].reshape(M, 9) + 0.01 * rng.rand(M, 9)
    Z2 = X[:M].copy()

    def kerns():
        k1 = ck.WeightedConv(GP.kernels.RBF(9, variance=1.0, lengthscales=1.1), [8, 8], [3, 3])
        k1.W = W
        k2 = GP.kernels.RBF(64, variance=1.1, lengthscales=2.0) + GP.kernels.White(64, 1e-2)
        return k1, k2

    k1, k2 = kerns()
    full = ss.FullSVSumGP(X, Y, k1, k2, GP.likelihoods.Gaussian(), Z1.copy(), Z2.copy(), num_latent=L)
    full.likelihood.variance = 0.3
    fq_mu, fq_sqrt = _variational(rng, 2 * M, L)
    full.q_mu, full.q_sqrt = fq_mu, fq_sqrt
    mu, var = full.predict_f(X)
    out.update(sum_X=X, sum_Y=Y, sum_W=W, sum_Z1=Z1, sum_Z2=Z2, full_q_mu=fq_mu, full_q_sqrt=fq_sqrt, full_mu=mu,
               full_var=var, full_elbo=full.compute_log_likelihood(), sum_lik_var=0.3)
    k1, k2 = kerns()
    mf = ss.MeanFieldSVSumGP(X, Y, k1, k2, GP.likelihoods.Gaussian(), Z1.copy(), Z2.copy(), num_latent=L)
    mf.likelihood.variance = 0.3
    mq_mu, mq_sqrt = _variational(rng, M, 2 * L)
    mf.q_mu, mf.q_sqrt = mq_mu, mq_sqrt
    mu, var = mf.predict_f(X)
    out.update(mf_q_mu=mq_mu, mf_q_sqrt=mq_sqrt, mf_mu=mu, mf_var=var, mf_elbo=mf.compute_log_likelihood())

    # -- cifar.py 'multi': WeightedMultiChannelConvGP + MultiOutputInducingSVGP
    rng = np.random.RandomState(231)
    N, M, L, C = 6, 7, 2, 3
    X, Y = rng.rand(N, 49 * C), rng.randn(N, L)
    k = mi.WeightedMultiChannelConvGP(GP.kernels.RBF(9, variance=0.9, lengthscales=1.2), [7, 7], [3, 3], C)
    Wm = rng.rand(C, 25) + 0.5
    k.W = Wm
    Z = rng.rand(M, 9)
    m = mi.MultiOutputInducingSVGP(X, Y, k, GP.likelihoods.Gaussian(), Z, num_latent=L)
    q_mu, q_sqrt = _variational(rng, M, L * C)
    m.q_mu, m.q_sqrt = q_mu, q_sqrt
    mu, var = m.predict_f(X)
    out.update(mo_X=X, mo_Y=Y, mo_W=Wm, mo_Z=Z, mo_q_mu=q_mu, mo_q_sqrt=q_sqrt, mo_mu=mu, mo_var=var)

    # -- SVConvGP (svconvgp.py:66-144; 28x28 hard-coded at :125)
    rng = np.random.RandomState(232)
    N, M, L = 3, 5, 1
    X, Y, Z = rng.rand(N, 784), rng.randn(N, 1), rng.rand(M, 25)
    # the array-Z branch of the reference's constructor cannot run (np.zeros((Z, D)), svconvgp.py:83): build with
    # the int form (Z := first patch of the first M images, :85-86), pin that initialisation, then assign Z.
    m = sc.SVConvGP(X, Y, GP.kernels.RBF(25, variance=0.7, lengthscales=1.3), GP.likelihoods.Gaussian(), M, 5)
    out["svc_Zinit"] = m.Z.value
    m.Z = Z
    m.likelihood.variance = 0.1
    q_mu, q_sqrt = _variational(rng, M, L)
    m.q_mu, m.q_sqrt = q_mu, q_sqrt
    mu, var = m.predict_f(X)
    out.update(svc_X=X, svc_Y=Y, svc_Z=Z, svc_q_mu=q_mu, svc_q_sqrt=q_sqrt, svc_mu=mu, svc_var=var,
               svc_elbo=m.compute_log_likelihood(), svc_patch_mu=m.compute_patch_predictions(X))

    # -- SVColourConvGP (svconvgp.py:25-63): one conv kernel applied per colour channel, outputs mixed by wc
    rng = np.random.RandomState(233)
    N, M, L, C = 5, 6, 2, 3
    X, Y, Z = rng.rand(N, 64 * C), rng.randn(N, L), rng.rand(M, 9)
    k = ck.WeightedConv(GP.kernels.RBF(9, variance=0.8, lengthscales=1.2), [8, 8], [3, 3])
    Wk = rng.rand(36) + 0.5
    k.W = Wk
    m = sc.SVColourConvGP(X, Y, k, GP.likelihoods.Gaussian(), Z, C, num_latent=L)
    wc = rng.rand(C) + 0.5
    m.wc = wc
    q_mu, q_sqrt = _variational(rng, M, L * C)
    m.q_mu, m.q_sqrt = q_mu, q_sqrt
    mu, var = m.predict_f(X)
    out.update(scc_X=X, scc_Y=Y, scc_Z=Z, scc_W=Wk, scc_wc=wc, scc_q_mu=q_mu, scc_q_sqrt=q_sqrt, scc_mu=mu, scc_var=var)
    return out


def rectangles_rows(n):
    import zipfile
    z = zipfile.ZipFile("/root/reference/datasets/rectangles.zip")
    rows = []
    with z.open("rectangles_train.amat") as f:
        for _ in range(n):
            rows.append([float(v) for v in f.readline().split()][:784])
    return np.array(rows)


def prepare(case):
    """Input overrides applied identically at generation and at test time (tests load X/Z from the fixture)."""
    if case.name == "ref_cfg1_rectangles_conv":
        case.X = rectangles_rows(case.N)
        rng = np.random.RandomState(case.seed)
        pat = case.oracle_kernel(np.ones(case.wshape()))._get_patches(case.X).reshape(-1, case.geom.D)
        case.Z = pat[rng.permutation(len(pat))[:case.M]] + 0.001 * rng.rand(case.M, case.geom.D)
    return case


def main(only=()):
    """Regenerate every fixture, or only the named ones (`python tests/golden/make_ref_golden.py ref_models`)."""
    from oracle import run_reference as R
    GP, ck, mi, ss, sc = R.load()
    os.makedirs(OUT, exist_ok=True)
    for case in kernel_cases():
        if only and case.name not in only:
            continue
        np.savez_compressed(os.path.join(OUT, case.name + ".npz"), **run_kernel_case(prepare(case), GP, ck, mi))
        print("wrote", case.name)
    for name, fn in (("ref_full_K", lambda: run_full_K(GP, ck, mi)), ("ref_convrbf", lambda: run_convrbf(GP, ck)),
                     ("ref_conditional", lambda: run_conditional(mi)),
                     ("ref_models", lambda: run_models(GP, ck, mi, ss, sc))):
        if only and name not in only:
            continue
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **fn())
        print("wrote", name)


if __name__ == "__main__":
    main(tuple(sys.argv[1:]))
