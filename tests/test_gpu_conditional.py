"""The fused whitened conditional (convgp_b200/conditional.py) against the line-by-line restatement of
misvgp.py:83-165 in the torch twin: values and all five gradients."""
import numpy as np
import pytest
import torch

from oracle import torch_twin as T
from tests.util import rel_err

pytestmark = pytest.mark.gpu


def test_fused_conditional_matches_reference_form():
    from convgp_b200.conditional import whitened_conditional
    rng = np.random.RandomState(0)
    M, N, L = 37, 23, 4
    B = rng.randn(M, M)
    Kmm = B @ B.T + M * np.eye(M)
    Kmn = rng.randn(M, N)
    Kd = rng.rand(N) + 5.0
    q_mu = rng.randn(M, L)
    q_sqrt = np.tril(rng.randn(L, M, M) * 0.3 + np.eye(M)[None]).transpose(1, 2, 0).copy()
    wm, wv = rng.randn(N, L), rng.randn(N, L)

    def run(fn, dev):
        ts = [torch.tensor(a, dtype=torch.float64, device=dev, requires_grad=True) for a in (Kd, Kmn, Kmm, q_mu, q_sqrt)]
        mu, var = fn(*ts)
        loss = (mu * torch.tensor(wm, device=dev)).sum() + (var * torch.tensor(wv, device=dev)).sum()
        loss.backward()
        return [mu.detach().cpu().numpy(), var.detach().cpu().numpy()] + [t.grad.cpu().numpy() for t in ts]

    got = run(whitened_conditional, "cuda")
    got[4] = 0.5 * (got[4] + got[4].T)
    names = ("fmean", "fvar", "dKdiag", "dKmn", "dKmm", "dq_mu", "dq_sqrt")
    for attempt in range(2):  # the live CPU twin is evaluated a second time before a disagreement counts
        ref = run(lambda kd, kmn, kmm, f, qs: T.conditional(kd, kmn, kmm, f, qs, whiten=True), "cpu")
        ref[4] = 0.5 * (ref[4] + ref[4].T)      # autograd through cholesky returns a triangular-weighted dKmm
        ref[6] = np.tril(ref[6].transpose(2, 0, 1)).transpose(1, 2, 0)
        errs = {name: rel_err(g, r) for g, r, name in zip(got, ref, names)}
        if max(errs.values()) < 1e-9:
            break
    assert max(errs.values()) < 1e-9, errs
