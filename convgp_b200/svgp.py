"""Sparse variational GP with inter-domain inducing patches: the model the reference scripts build
(``GPflow.svgp.SVGP`` of the GPflow-inter-domain fork, mirrored in-tree by svconvgp.py:91-104 for
the ELBO and misvgp.py:83-165 for the conditional).  Same constructor, parameters (Z, q_mu, q_sqrt)
and methods (build_prior_KL, build_predict, build_likelihood, predict_y, compute_log_likelihood,
_objective, get_free_state / set_state).  Kernel matrices come from the CUDA hot path; Cholesky, the
triangular inverse, every GEMM of the conditional and of its reverse mode, and the variational projection are
this library's own kernels (linalg.py -> csrc/linalg.cu); torch carries tensors and the autograd tape."""
import numpy as np
import torch

from . import parallel
from .kernels import as_tensor, svgp_terms
from .misvgp import conditional
from .param import LowerTriangular, Param, Parameterized, Positive
from .settings import settings


def gauss_kl_white(q_mu, q_sqrt):
    """KL[N(q_mu, q_sqrt q_sqrt^T) || N(0, I)] summed over latents (GPflow 0.x gauss_kl_white)."""
    KL = 0.5 * (q_mu ** 2).sum() - 0.5 * q_mu.numel()
    L = torch.tril(q_sqrt.permute(2, 0, 1))
    KL = KL - 0.5 * torch.log(torch.diagonal(L, dim1=1, dim2=2) ** 2).sum()
    return KL + 0.5 * (L ** 2).sum()


def gauss_kl_white_diag(q_mu, q_sqrt):
    KL = 0.5 * (q_mu ** 2).sum() - 0.5 * q_mu.numel()
    return KL - 0.5 * torch.log(q_sqrt ** 2).sum() + 0.5 * (q_sqrt ** 2).sum()


def gauss_kl(q_mu, q_sqrt, K):
    """KL[N(q_mu, S) || N(0, K)] summed over latents (GPflow 0.x gauss_kl / gauss_kl_diag).  chol(K), the solves
    against it and log|K| come from one ``linalg.chol_solve`` node (hand-written Cholesky / inverse / GEMM)."""
    from . import linalg
    M, L_ = q_mu.shape
    if q_sqrt.dim() == 2:
        rhs = torch.cat([q_mu, torch.eye(M, dtype=K.dtype, device=K.device)], 1)
    else:
        Lq = torch.tril(q_sqrt.permute(2, 0, 1))                              # (L, M, M)
        rhs = torch.cat([q_mu, Lq.permute(1, 0, 2).reshape(M, L_ * M)], 1)    # [q_mu | Lq_1 | ... | Lq_L]
    sol, _, logdet = linalg.chol_solve(K, rhs)                                # Lk^-1 rhs, sum log diag(Lk)
    alpha, rest = sol[:, :L_], sol[:, L_:]
    KL = 0.5 * (alpha ** 2).sum() + L_ * logdet - 0.5 * M * L_
    if q_sqrt.dim() == 2:
        KL = KL - 0.5 * torch.log(q_sqrt ** 2).sum()
        kinv_diag = (rest ** 2).sum(0)                                        # diag(K^-1) = colsum(Lk^-1 ** 2)
        return KL + 0.5 * (kinv_diag[:, None] * q_sqrt ** 2).sum()
    KL = KL - 0.5 * torch.log(torch.diagonal(Lq, dim1=1, dim2=2) ** 2).sum()
    return KL + 0.5 * (rest ** 2).sum()


class Model(Parameterized):
    """Objective plumbing shared by the SVGP family."""

    def build_likelihood(self):
        raise NotImplementedError

    def build_prior(self):
        return 0.0

    def compute_log_likelihood(self):
        with torch.no_grad():
            return float(self.build_likelihood())

    def _objective(self, x):
        """(-ELBO, gradient) at free state x, float64 NumPy like GPflow's (opt_tools/helpers.py:108-117).
        With an initialised process group the data term is this rank's shard and the packed
        [f | g] buffer is summed across ranks."""
        self.set_state(x)
        ps = self.free_params()
        for p in ps:
            p.free.grad = None
        f = -(self.build_likelihood() + self.build_prior())
        f.backward()
        flat = torch.cat([f.detach().reshape(1)] + [
            (p.free.grad if p.free.grad is not None else torch.zeros_like(p.free)).reshape(-1) for p in ps])
        if getattr(self, "data_parallel", True):
            parallel.allreduce_sum_(flat)
        out = flat.double().cpu().numpy()
        return out[0], out[1:]

    def graphed_objective(self, warmup=3):
        """Capture -ELBO and its gradient (w.r.t. the free parameters) for a static minibatch shape into
        ONE CUDA graph and return ``step() -> flat`` with ``flat = [f | g]`` (a static device tensor,
        summed across ranks when a process group is initialised).  At the MNIST shape the whole step
        is a few milliseconds of GPU time spread over ~150 launches, so replaying a graph removes
        the Python/launch overhead that otherwise dominates the wall clock.  Every call to
        ``step()`` draws a fresh minibatch (the indices are copied into a static device buffer)."""
        ps = self.free_params()
        free = [p.free for p in ps]
        static_idx = self._begin_static_minibatch()  # owned by the returned step, not left on the model

        def body():
            f = -(self.build_likelihood() + self.build_prior())
            grads = torch.autograd.grad(f, free, allow_unused=True)
            return torch.cat([f.detach().reshape(1)] + [(g if g is not None else torch.zeros_like(t)).reshape(-1)
                                                for g, t in zip(grads, free)])

        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(warmup):
                    body()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                flat = body()
        finally:
            self._mb_idx = None  # eager _objective / optimize / compute_log_likelihood sample fresh minibatches again

        def step():
            self._next_static_minibatch(static_idx)
            graph.replay()
            if getattr(self, "data_parallel", True):
                parallel.allreduce_sum_(flat)
            return flat

        step.graph = graph
        return step

    def _begin_static_minibatch(self):
        return None

    def _next_static_minibatch(self, static_idx):
        pass

    def optimize(self, method="L-BFGS-B", maxiter=1000, callback=None, lr=1e-3, **kw):
        """Drive the objective with scipy (method string) or a torch optimiser class / 'adam'."""
        if isinstance(method, str) and method.lower() != "adam":
            from scipy.optimize import minimize
            res = minimize(self._objective, self.get_free_state(), jac=True, method=method, callback=callback,
                           options=dict(maxiter=maxiter, **kw))
            self.set_state(res.x)
            return res
        opt_cls = torch.optim.Adam if isinstance(method, str) else method
        opt = opt_cls([p.free for p in self.free_params()], lr=lr)
        for _ in range(maxiter):
            opt.zero_grad()
            f = -(self.build_likelihood() + self.build_prior())
            f.backward()
            opt.step()
            if callback is not None:
                callback(self.get_free_state())
        return None


class SVGP(Model):
    def __init__(self, X, Y, kern, likelihood, Z, mean_function=None, num_latent=None, q_diag=False, whiten=True,
                 minibatch_size=None):
        self.X = as_tensor(X)
        self.Y = as_tensor(Y)
        self.kern, self.likelihood = kern, likelihood
        self.mean_function = mean_function
        self.q_diag, self.whiten = q_diag, whiten
        self.num_data = self.X.shape[0]
        self.minibatch_size = minibatch_size
        self._rng = np.random.RandomState(0)
        self.num_latent = num_latent or self.Y.shape[1]
        if Z is not None:
            self.Z = Param(Z)
            self.num_inducing = self.Z.shape[0]
            self._init_variational(self.num_inducing, self.num_latent)

    def _init_variational(self, M, L):
        self.q_mu = Param(np.zeros((M, L)))
        if self.q_diag:
            self.q_sqrt = Param(np.ones((M, L)), Positive())
        else:
            self.q_sqrt = Param(np.array([np.eye(M) for _ in range(L)]).swapaxes(0, 2), LowerTriangular(M, L))

    # ------------------------------------------------------------------ minibatching
    def _minibatch(self):
        if self.minibatch_size is None or self.minibatch_size >= self.num_data:
            return self.X, self.Y
        if getattr(self, "_mb_idx", None) is not None:  # static index buffer (CUDA-graph mode)
            return torch.index_select(self.X, 0, self._mb_idx), torch.index_select(self.Y, 0, self._mb_idx)
        idx = torch.as_tensor(self._rng.permutation(self.num_data)[:self.minibatch_size], device=self.X.device)
        return self.X[idx], self.Y[idx]

    def _begin_static_minibatch(self):
        """Static device index buffer for CUDA-graph mode.  ``self._mb_idx`` points at it only while the graph is
        being captured (``graphed_objective`` clears it afterwards); the returned state belongs to the graphed step.
        Two pinned host buffers alternate, each guarded by an event recorded after its H2D copy, so rewriting a
        host buffer can never race a copy that is still queued."""
        if self.minibatch_size is None or self.minibatch_size >= self.num_data:
            return None
        st = {"host": [torch.empty(self.minibatch_size, dtype=torch.long).pin_memory() for _ in range(2)],
              "event": [None, None], "turn": 0,
              "dev": torch.zeros(self.minibatch_size, dtype=torch.long, device=self.X.device)}
        self._mb_idx = st["dev"]
        self._next_static_minibatch(st)
        return st

    def _next_static_minibatch(self, st):
        if st is None:
            return
        i = st["turn"]
        st["turn"] = 1 - i
        if st["event"][i] is not None:
            st["event"][i].synchronize()
        st["host"][i].copy_(torch.from_numpy(self._rng.permutation(self.num_data)[:self.minibatch_size]))
        st["dev"].copy_(st["host"][i], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        st["event"][i] = ev

    # ------------------------------------------------------------------ model terms
    def build_prior_KL(self):
        q_mu, q_sqrt = self.q_mu(), self.q_sqrt()
        if self.whiten:
            return gauss_kl_white_diag(q_mu, q_sqrt) if self.q_diag else gauss_kl_white(q_mu, q_sqrt)
        Z = self.Z()
        K = self.kern.Kzz(Z) + torch.eye(Z.shape[0], dtype=Z.dtype, device=Z.device) * settings.jitter_level
        return gauss_kl(q_mu, q_sqrt, K)

    def build_predict(self, Xnew, full_cov=False):
        Z = self.Z()
        Kdiag, Kmn, Kmm = svgp_terms(self.kern, Z, Xnew)  # three independent kernel evaluations, parallel streams
        mu, var = conditional(Xnew, Z, Kdiag, Kmn, Kmm, self.q_mu(),
                              full_cov=full_cov, q_sqrt=self.q_sqrt(), whiten=self.whiten)
        if self.mean_function is not None:
            mu = mu + self.mean_function(Xnew)
        return mu, var

    def build_likelihood(self):  # svconvgp.py:91-104
        Xb, Yb = self._minibatch()
        dp = torch.distributed.is_initialized() and getattr(self, "data_parallel", True)
        world = torch.distributed.get_world_size() if dp else 1
        rank = torch.distributed.get_rank() if dp else 0
        if world > 1:  # data parallel: this rank's rows; the KL term is counted once (rank 0)
            Xb, Yb = parallel.shard_rows(Xb, world, rank), parallel.shard_rows(Yb, world, rank)
        n_batch = self.minibatch_size if (self.minibatch_size and self.minibatch_size < self.num_data) \
            else self.num_data
        fmean, fvar = self.build_predict(Xb)
        var_exp = self.likelihood.variational_expectations(fmean, fvar, Yb)
        scale = float(self.num_data) / float(n_batch)
        ll = var_exp.sum() * scale
        if rank == 0:
            ll = ll - self.build_prior_KL()
        return ll

    # ------------------------------------------------------------------ prediction
    def predict_f(self, Xnew):
        with torch.no_grad():
            mu, var = self.build_predict(as_tensor(Xnew))
        return mu.cpu().numpy(), var.cpu().numpy()

    def predict_batched(self, Xnew, batch_size=1000, density_Y=None):
        """Large-N forward sweep (the reference predicts 10 000 test images in batches of 1000 / 500 / 200,
        opt_tools/gpflow_tasks.py:70-90, cifar.py:44, and re-evaluates Kzz and its Cholesky for every
        batch): here ``Kuu``, its Cholesky, ``Li = Lm^-1`` and ``Lq^T Li`` are computed ONCE and every
        batch costs one Kuf + Kdiag evaluation, one triangular GEMM and the fused projection.  Returns ``(mean, var)`` of y, or the log
        predictive density when ``density_Y`` is given.  Rows of ``Xnew`` are sharded over ranks when a
        process group is initialised (each rank returns its shard)."""
        if type(self).build_predict is not SVGP.build_predict:
            raise NotImplementedError("predict_batched re-uses the factors of SVGP.build_predict; %s overrides "
                                      "build_predict -- use predict_y / predict_density in batches"
                                      % type(self).__name__)
        if not (self.whiten and not self.q_diag):
            raise NotImplementedError("predict_batched covers the whitened full-covariance model")
        from . import linalg
        with torch.no_grad():
            Xnew = parallel.shard_rows(as_tensor(Xnew))
            Yd = None if density_Y is None else parallel.shard_rows(as_tensor(density_Y))
            Z = self.Z()
            M = Z.shape[0]
            Kmm = self.kern.Kzz(Z) + torch.eye(M, dtype=Z.dtype, device=Z.device) * settings.jitter_level
            _, Li, _ = linalg.potrf_inv(Kmm)                       # once per sweep
            q_mu, q_sqrt = self.q_mu(), self.q_sqrt()
            outs = []
            for s in range(0, Xnew.shape[0], batch_size):
                xb = Xnew[s:s + batch_size]
                A = linalg.gemm(Li, self.kern.Kzx(Z, xb), a_tri=True)
                fmean, fvar = linalg.project(A, q_mu, q_sqrt, self.kern.Kdiag(xb))  # LTA never leaves the kernel
                if self.mean_function is not None:
                    fmean = fmean + self.mean_function(xb)
                if Yd is None:
                    outs.append(self.likelihood.predict_mean_and_var(fmean, fvar))
                else:
                    outs.append((self.likelihood.predict_density(fmean, fvar, Yd[s:s + batch_size]),))
            if not outs:
                return None
            res = tuple(torch.cat([o[i] for o in outs]).cpu().numpy() for i in range(len(outs[0])))
        return res if Yd is None else res[0]

    def predict_y(self, Xnew):
        with torch.no_grad():
            mu, var = self.build_predict(as_tensor(Xnew))
            m, v = self.likelihood.predict_mean_and_var(mu, var)
        return m.cpu().numpy(), v.cpu().numpy()

    def predict_density(self, Xnew, Ynew):
        with torch.no_grad():
            mu, var = self.build_predict(as_tensor(Xnew))
            return self.likelihood.predict_density(mu, var, as_tensor(Ynew)).cpu().numpy()
