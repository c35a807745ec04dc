"""The hand-written linear algebra of the SVGP consumers (csrc/linalg.cu through the C ABI) against library results
(cuBLAS / cuSOLVER through torch -- used here ONLY as the cross-check SURVEY.md hard-part 6 allows) and against torch
autograd through the textbook formulation of the conditional (misvgp.py:128-163)."""
import numpy as np
import pytest
import torch

from tests.util import rel_err

pytestmark = pytest.mark.gpu

DT = [(torch.float64, 1e-12, 1e-10), (torch.float32, 2e-5, 1e-4)]


def _dev():
    return torch.device("cuda", torch.cuda.current_device())


def _rand(rng, *shape, dtype):
    return torch.tensor(rng.standard_normal(shape), dtype=dtype, device=_dev())


def _spd(rng, M, dtype, jitter=1e-2):
    Z = rng.random((M, 9))
    d = ((Z[:, None, :] - Z[None, :, :]) ** 2).sum(-1)
    K = np.exp(-0.5 * d / 1.5 ** 2) + jitter * np.eye(M)
    return torch.tensor(K, dtype=dtype, device=_dev())


@pytest.mark.parametrize("dtype,tol,_", DT)
@pytest.mark.parametrize("m,n,k", [(64, 64, 64), (750, 200, 750), (130, 67, 45), (1, 1, 1), (33, 200, 7), (200, 750, 129)])
def test_gemm_all_layouts(dtype, tol, _, m, n, k):
    from convgp_b200 import linalg
    rng = np.random.default_rng(m * 1000 + n * 10 + k)
    for ta in (False, True):
        for tb in (False, True):
            A = _rand(rng, *((k, m) if ta else (m, k)), dtype=dtype)
            B = _rand(rng, *((n, k) if tb else (k, n)), dtype=dtype)
            ref = (A.T if ta else A).double() @ (B.T if tb else B).double()
            got = linalg.gemm(A, B, trans_a=ta, trans_b=tb, alpha=0.5)
            assert rel_err(got.cpu().numpy(), 0.5 * ref.cpu().numpy()) < tol, (ta, tb)


@pytest.mark.parametrize("dtype,tol,_", DT)
def test_gemm_triangular_symmetric_batched(dtype, tol, _):
    from convgp_b200 import linalg
    rng = np.random.default_rng(3)
    M, N, L = 150, 70, 3
    Lo = torch.tril(_rand(rng, M, M, dtype=dtype))
    garbage = Lo + torch.triu(torch.full_like(Lo, float("nan")), 1)  # entries above the diagonal must never be read
    B = _rand(rng, M, N, dtype=dtype)
    for ta in (False, True):
        ref = (Lo.T if ta else Lo).double() @ B.double()
        got = linalg.gemm(garbage, B, trans_a=ta, a_tri=True)
        assert rel_err(got.cpu().numpy(), ref.cpu().numpy()) < tol
    X = _rand(rng, N, M, dtype=dtype)
    for tb in (False, True):  # B operand triangular: X (N x M) times Lo / Lo^T
        ref = X.double() @ (Lo.T if tb else Lo).double()
        got = linalg.gemm(X, garbage, trans_b=tb, b_tri=True)
        assert rel_err(got.cpu().numpy(), ref.cpu().numpy()) < tol
    # lower-only and mirrored outputs
    P = B.double() @ B.double().T
    low = linalg.gemm(B, B, trans_b=True, c_mode=1)
    assert rel_err(low.cpu().numpy(), torch.tril(P).cpu().numpy()) < tol
    sym = linalg.gemm(B, B, trans_b=True, c_mode=2)
    assert rel_err(sym.cpu().numpy(), P.cpu().numpy()) < tol
    # batched (shared A) and batched operands
    Bb = _rand(rng, L, M, N, dtype=dtype)
    got = linalg.gemm(garbage, Bb, a_tri=True)
    assert rel_err(got.cpu().numpy(), (Lo.double()[None] @ Bb.double()).cpu().numpy()) < tol
    Ab = torch.tril(_rand(rng, L, M, M, dtype=dtype))
    got = linalg.gemm(Ab, Bb, trans_a=True, a_tri=True)
    assert rel_err(got.cpu().numpy(), (Ab.double().transpose(1, 2) @ Bb.double()).cpu().numpy()) < tol


@pytest.mark.parametrize("dtype,tol,_", DT)
@pytest.mark.parametrize("M", [1, 31, 32, 33, 64, 100, 257, 750, 1000])
def test_potrf_inv_against_cusolver(dtype, tol, _, M):
    from convgp_b200 import linalg
    rng = np.random.default_rng(M)
    K = _spd(rng, M, dtype)
    Lm, Li, info = linalg.potrf_inv(K)
    assert int(info) == 0
    ref = torch.linalg.cholesky(K.double())
    scale = 100 if dtype == torch.float32 else 1000  # conditioning of K (1e2 .. 1e4) amplifies the rounding
    assert rel_err(Lm.cpu().numpy(), ref.cpu().numpy()) < scale * tol
    assert float(torch.triu(Lm, 1).abs().max()) == 0.0 if M > 1 else True
    eye = torch.eye(M, dtype=torch.float64, device=K.device)
    assert rel_err((Li.double() @ Lm.double()).cpu().numpy(), eye.cpu().numpy()) < scale * tol
    assert float(torch.triu(Li, 1).abs().max()) == 0.0 if M > 1 else True
    assert rel_err((Lm.double() @ Lm.double().T).cpu().numpy(), K.double().cpu().numpy()) < 10 * tol


def test_potrf_reports_the_failing_minor():
    from convgp_b200 import _lib, linalg
    rng = np.random.default_rng(0)
    K = _spd(rng, 100, torch.float64)
    K[70, 70] = -1.0
    with pytest.raises(_lib.CgpError) as e:
        linalg.potrf_inv(K)
    assert e.value.status == -6 and "order 71" in str(e.value)
    _, _, info = linalg.potrf_inv(K, check_pd=False)  # asynchronous use: the flag stays on the device
    assert int(info) == 71
    with pytest.raises(_lib.CgpError):
        linalg.potrf_inv(-torch.eye(4, dtype=torch.float32, device=_dev()))


def _ref_conditional(Kdiag, Kmn, Kmm, f, q_sqrt, whiten):
    """misvgp.py:128-165 in plain torch (library Cholesky / triangular solves; autograd for the gradients)."""
    Lm = torch.linalg.cholesky(Kmm)
    A = torch.linalg.solve_triangular(Lm, Kmn, upper=False)
    fvar = (Kdiag - (A * A).sum(0))[None, :].expand(f.shape[1], -1)
    if not whiten:
        A = torch.linalg.solve_triangular(Lm.T, A, upper=True)
    fmean = A.T @ f
    if q_sqrt is not None:
        if q_sqrt.dim() == 2:
            LTA = A[None] * q_sqrt.T[:, :, None]
        else:
            LTA = torch.tril(q_sqrt.permute(2, 0, 1)).transpose(1, 2) @ A[None]
        fvar = fvar + (LTA * LTA).sum(1)
    return fmean, fvar.T


@pytest.mark.parametrize("dtype,_,tol", DT)
@pytest.mark.parametrize("whiten", [True, False])
@pytest.mark.parametrize("qrank", [0, 2, 3])
@pytest.mark.parametrize("M,N,L", [(70, 45, 3), (200, 130, 10)])
def test_conditional_values_and_gradients(dtype, _, tol, whiten, qrank, M, N, L):
    from convgp_b200.conditional import conditional
    rng = np.random.default_rng(M + N + qrank)
    Kmm = _spd(rng, M, dtype, jitter=3e-2)
    Kmn = 0.3 * _rand(rng, M, N, dtype=dtype)
    Kdiag = torch.tensor(rng.random(N) + 1.0, dtype=dtype, device=_dev())
    f = _rand(rng, M, L, dtype=dtype)
    if qrank == 0:
        q = None
    elif qrank == 2:
        q = torch.tensor(rng.random((M, L)) + 0.5, dtype=dtype, device=_dev())
    else:
        q = torch.tril(0.2 * _rand(rng, L, M, M, dtype=dtype) + torch.eye(M, dtype=dtype, device=_dev())).permute(1, 2, 0)
        q = q + torch.triu(torch.ones(M, M, dtype=dtype, device=_dev()), 1)[:, :, None]  # upper part must be ignored
    wm, wv = _rand(rng, N, L, dtype=dtype), _rand(rng, N, L, dtype=dtype)

    def run(fn, dt):
        leaves = [t.detach().to(dt).clone().requires_grad_(True) if t is not None else None for t in (Kdiag, Kmn, Kmm, f, q)]
        mu, var = fn(leaves[0], leaves[1], leaves[2], leaves[3], leaves[4], whiten)
        (mu * wm.to(dt)).sum().add((var * wv.to(dt)).sum()).backward()
        return [mu, var] + [t.grad for t in leaves if t is not None]

    got = run(lambda a, b, c, d, e, w: conditional(a, b, c, d, q_sqrt=e, whiten=w), dtype)
    ref = run(_ref_conditional, torch.float64)
    names = ["fmean", "fvar", "dKdiag", "dKmn", "dKmm", "df"] + (["dq"] if q is not None else [])
    amp = 30 if dtype == torch.float32 else 100  # the ill-conditioned solve amplifies rounding (cond ~ 1e2)
    for nm, g, r in zip(names, got, ref):
        if nm == "dq" and qrank == 3:
            g, r = torch.tril(g.permute(2, 0, 1)), torch.tril(r.permute(2, 0, 1))
        if nm == "dKmm":  # a symmetric-matrix gradient is defined up to its antisymmetric part: compare sym()
            g, r = 0.5 * (g + g.T), 0.5 * (r + r.T)
        assert rel_err(g.detach().cpu().numpy(), r.detach().cpu().numpy()) < amp * tol, (nm, whiten, qrank)


@pytest.mark.parametrize("dtype,_,tol", DT)
def test_shared_factorisation_and_logdet(dtype, _, tol):
    """C right-hand sides against one Cholesky (multi-output models) and the log-determinant output (gauss_kl)."""
    from convgp_b200 import linalg
    rng = np.random.default_rng(5)
    M, N, C = 90, 40, 3
    Kmm = _spd(rng, M, dtype, jitter=3e-2).requires_grad_(True)
    Kmn = (0.3 * _rand(rng, C, M, N, dtype=dtype)).requires_grad_(True)
    w = _rand(rng, C, M, N, dtype=dtype)
    A, _, logdet = linalg.chol_solve(Kmm, Kmn)
    ((A * w).sum() + 1.7 * logdet).backward()
    K64 = Kmm.detach().double().requires_grad_(True)
    R64 = Kmn.detach().double().requires_grad_(True)
    Lm = torch.linalg.cholesky(K64)
    A64 = torch.linalg.solve_triangular(Lm[None].expand(C, -1, -1), R64, upper=False)
    ((A64 * w.double()).sum() + 1.7 * torch.log(torch.diagonal(Lm)).sum()).backward()
    amp = 30 if dtype == torch.float32 else 100
    assert rel_err(A.detach().cpu().numpy(), A64.detach().cpu().numpy()) < amp * tol
    assert abs(float(logdet) - float(torch.log(torch.diagonal(Lm)).sum())) < amp * tol * abs(float(logdet)) + 1e-6
    assert rel_err(Kmn.grad.cpu().numpy(), R64.grad.cpu().numpy()) < amp * tol
    g, r = Kmm.grad, K64.grad
    assert rel_err((0.5 * (g + g.T)).cpu().numpy(), (0.5 * (r + r.T)).cpu().numpy()) < amp * tol


def test_full_size_config3_and_config4_shapes():
    """M = 750 (config 3) and the stacked 1500 (config 4, svsumgp.py:86) with L = 10 at the full minibatch."""
    from convgp_b200 import linalg
    rng = np.random.default_rng(9)
    for Mt, N, L in ((750, 200, 10), (1500, 200, 10)):
        A = 0.05 * _rand(rng, Mt, N, dtype=torch.float64)
        q_mu = _rand(rng, Mt, L, dtype=torch.float64)
        Lq = torch.tril(0.02 * _rand(rng, L, Mt, Mt, dtype=torch.float64) + torch.eye(Mt, dtype=torch.float64, device=_dev()))
        Kd = torch.tensor(rng.random(N) + 1.0, dtype=torch.float64, device=_dev())
        A.requires_grad_(True)
        q_mu.requires_grad_(True)
        q = Lq.permute(1, 2, 0).detach().requires_grad_(True)
        fm, fv = linalg.project(A, q_mu, q, Kd)
        wm, wv = _rand(rng, N, L, dtype=torch.float64), _rand(rng, N, L, dtype=torch.float64)
        ((fm * wm).sum() + (fv * wv).sum()).backward()
        A2 = A.detach().clone().requires_grad_(True)
        m2 = q_mu.detach().clone().requires_grad_(True)
        q2 = Lq.detach().clone().requires_grad_(True)
        LTA = q2.transpose(1, 2) @ A2[None]
        rm = A2.T @ m2
        rv = ((Kd - (A2 * A2).sum(0))[None] + (LTA * LTA).sum(1)).T
        ((rm * wm).sum() + (rv * wv).sum()).backward()
        assert rel_err(fm.detach().cpu().numpy(), rm.detach().cpu().numpy()) < 1e-11
        assert rel_err(fv.detach().cpu().numpy(), rv.detach().cpu().numpy()) < 1e-11
        assert rel_err(A.grad.cpu().numpy(), A2.grad.cpu().numpy()) < 1e-11
        assert rel_err(q_mu.grad.cpu().numpy(), m2.grad.cpu().numpy()) < 1e-11
        assert rel_err(torch.tril(q.grad.permute(2, 0, 1)).cpu().numpy(), torch.tril(q2.grad).cpu().numpy()) < 1e-11
