#!/usr/bin/env python
"""Benchmark of the convgp patch-kernel hot path (BASELINE.json metric):

    conv/wconv Kuf + Kdiag (+ Kuu) forward + backward, images/s, 28x28, 5x5 patches, M=750,
    minibatch 200 (config 3: mnist.py -k wconv -M 750 --minibatch-size 200), float64 AND float32.

One "step" = one pass of the hot path over one minibatch: Kuf, Kdiag, Kuu forward, then -- after a join, because
dL/dKuf depends on all three forwards in the reference's ELBO -- their backward w.r.t. Z, W, variance, lengthscales
from given upstream gradients.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--dtype both|f64|f32] [--impl reference]
                  [--secondary / --no-secondary] [--workload hotpath|predict]

Prints ONE JSON line (rank 0): the float64 headline, an "f32" sub-object with the tcgen05 path, the secondary
configs under "configs", the SVGP training step under "svgp_step", and for N > 1 "strong" (the same 200 images
split by rows, all-reduce inside the timed region) and "grad_check" (all-reduced sharded gradient vs rank 0's
unsharded gradient) next to the weak-scaling headline.  See DESIGN.md section 6 for how each field is obtained.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


# ----------------------------------------------------------------------------- workloads
class Cfg:
    """One BASELINE.json configuration of the hot path.  groups: [(variance, lengthscale)] of the additive RBF base
    kernel; layout 'planar' (Conv._get_patches) or 'colour' (ColourPatchConv._get_patches)."""

    def __init__(self, key, workload, H, W, C, ph, pw, layout, M, N, weighted, groups, white, xdist="rand"):
        self.key, self.workload = key, workload
        self.H, self.W, self.C, self.ph, self.pw, self.layout = H, W, C, ph, pw, layout
        self.M, self.N, self.weighted, self.groups, self.white, self.xdist = M, N, weighted, groups, white, xdist
        self.Pc = (H - ph + 1) * (W - pw + 1)
        if layout == "planar":
            self.P, self.D = self.Pc * C, ph * pw
        else:
            self.P, self.D = self.Pc, ph * pw * C
        self.G = len(groups)
        self.diag_add = white + 1e-5  # White variance (mnist.py:42) + jitter (gpflowrc:11)


def _cifar_groups():
    r = np.random.default_rng(2)
    return [(1.0 + 0.01 * r.standard_normal(), 0.5) for _ in range(3)]  # cifar.py:36-37


CONFIGS = {
    "cfg1": Cfg("cfg1", "rectangles.py -k conv -M 16 --minibatch-size 100 (28x28x1, 3x3 patches, P=676, D=9)",
                28, 28, 1, 3, 3, "planar", 16, 100, False, [(1.0, 1.0)], 1e-3),
    "cfg2": Cfg("cfg2", "mnist01.py -k wconv -M 50 (28x28x1, 5x5 patches, P=576, D=25, minibatch 100)",
                28, 28, 1, 5, 5, "planar", 50, 100, True, [(1.0, 1.0)], 1e-3),
    "cfg3": Cfg("cfg3", "mnist.py -k wconv -M 750 --minibatch-size 200 (28x28x1, 5x5 patches, P=576, D=25)",
                28, 28, 1, 5, 5, "planar", 750, 200, True, [(1.0, 1.0)], 1e-3),
    # the patch-kernel half of sumkern_mnist.py (k1 = WeightedConv(RBF(25)), no White: sumkern_mnist.py:32); the
    # RBF(784) half and the 1500 x 1500 variational covariance are in "svgp_step" (model level)
    "cfg4": Cfg("cfg4", "sumkern_mnist.py -k1 wconv -k2 rbf -M 750 --vardist full --minibatch-size 200: k1 patch kernels",
                28, 28, 1, 5, 5, "planar", 750, 200, True, [(1.0, 1.0)], 0.0),
    "cfg5": Cfg("cfg5", "cifar.py -k addwconv -M 1000 --minibatch-size 30 (32x32x3 colour patches, 5x5, P=784, "
                        "D=75 = 3 RBF groups x 25)",
                32, 32, 3, 5, 5, "colour", 1000, 30, True, _cifar_groups(), 0.0, xdist="uint8"),
}
HEAD = CONFIGS["cfg3"]
METRIC = "conv/wconv Kuf+Kdiag fwd+bwd images/sec (28x28, M=750)"


def im2col(X, cfg):
    """NumPy patches (N, P, D) in the reference's two layouts (convkernels.py:38-62, :128-141)."""
    N = X.shape[0]
    Ho, Wo = cfg.H - cfg.ph + 1, cfg.W - cfg.pw + 1
    if cfg.layout == "planar":
        img = X.reshape(N, cfg.H * cfg.W, cfg.C).transpose(0, 2, 1).reshape(N * cfg.C, cfg.H, cfg.W)
        out = np.empty((N * cfg.C, Ho, Wo, cfg.ph, cfg.pw))
        for i in range(cfg.ph):
            for j in range(cfg.pw):
                out[:, :, :, i, j] = img[:, i:i + Ho, j:j + Wo]
        return out.reshape(N, cfg.C * Ho * Wo, cfg.ph * cfg.pw)
    img = X.astype(np.float32).astype(np.float64).reshape(N, cfg.H, cfg.W, cfg.C)
    out = np.empty((N, Ho, Wo, cfg.ph, cfg.pw, cfg.C))
    for i in range(cfg.ph):
        for j in range(cfg.pw):
            out[:, :, :, i, j, :] = img[:, i:i + Ho, j:j + Wo, :]
    return out.reshape(N, Ho * Wo, cfg.ph * cfg.pw * cfg.C)


def _draw_x(cfg, seed, n):
    r = np.random.default_rng(seed)
    if cfg.xdist == "uint8":
        return r.integers(0, 256, size=(n, cfg.H * cfg.W * cfg.C)) / 255.0
    return r.random((n, cfg.H * cfg.W * cfg.C))


def make_inputs(cfg, rank=0, n_images=None):
    """Seeded synthetic inputs: X ~ U[0,1) (seed 0+rank; uint8/255 for config 5), Z = M patches of
    the seed-0 minibatch + 1e-3 U (seed 1), W = 1, upstream gradients ~ N(0,1) (seeds 4,5,6)."""
    n = cfg.N if n_images is None else n_images
    X = _draw_x(cfg, 0 + rank, n)
    r1 = np.random.default_rng(1)
    pat = im2col(_draw_x(cfg, 0, min(cfg.N, 64)), cfg).reshape(-1, cfg.D)
    Z = pat[r1.permutation(len(pat))[:cfg.M]] + 0.001 * r1.random((cfg.M, cfg.D))
    G = np.random.default_rng(4).standard_normal((cfg.M, n))
    gd = np.random.default_rng(5).standard_normal(n)
    Gu = np.random.default_rng(6).standard_normal((cfg.M, cfg.M))
    return dict(X=X, Z=Z, W=np.ones(cfg.P) if cfg.weighted else None, var=np.array([g[0] for g in cfg.groups]),
                ls=np.array([g[1] for g in cfg.groups]), G=G, gd=gd, Gu=Gu)


def algorithmic_work(cfg, n_images):
    """FMA-pipe flops and exps of one step: per unique pair with G_k RBF groups over D dims, forward
    2D+2G_k flops + G_k exps; backward 4D+8G_k (Kuf) / 2D+8G_k (Kdiag) flops + G_k exps (recompute)."""
    D, G, P, M = cfg.D, cfg.G, cfg.P, cfg.M
    kuf = n_images * P * M
    kdiag = n_images * P * (P + 1) // 2
    kuu = M * M
    w = {
        "kuf_fwd": (kuf * (2 * D + 2 * G), kuf * G), "kdiag_fwd": (kdiag * (2 * D + 2 * G), kdiag * G),
        "kuu_fwd": (kuu * (2 * D + 2 * G), kuu * G), "kuf_bwd": (kuf * (4 * D + 8 * G), kuf * G),
        "kdiag_bwd": (kdiag * (2 * D + 8 * G), kdiag * G), "kuu_bwd": (kuu * (4 * D + 8 * G), kuu * G),
    }
    w["step"] = (sum(v[0] for v in w.values()), sum(v[1] for v in w.values()))
    return w


def algorithmic_bytes(cfg, n_images, es):
    """HBM bytes one fused step has to move: X read forward + backward, Z, Kuf out + G in, Kuu + Gu, small vectors."""
    return es * (2 * n_images * cfg.H * cfg.W * cfg.C + 4 * cfg.M * cfg.D + 2 * cfg.M * n_images
                 + 2 * cfg.M * cfg.M + 4 * cfg.P + 2 * n_images)


# ----------------------------------------------------------------------------- CPU reference arm
def cpu_port_images_per_s(cfg, dtype_name, n_images, reps, threads=None):
    """The reference's algorithm on host cores: oracle/torch_twin (materialised patches -> GEMM -> exp -> patch
    sums, per-image Kdiag loop, autodiff backward).  The only place bench.py executes oracle/."""
    import torch
    from oracle import torch_twin as T
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    dt = torch.float64 if dtype_name == "f64" else torch.float32
    inp = make_inputs(cfg, 0, n_images)
    g = T.Geom(cfg.H, cfg.W, cfg.C, cfg.ph, cfg.pw, "conv" if cfg.layout == "planar" else "colour")
    X = torch.tensor(inp["X"], dtype=dt)
    G, gd, Gu = (torch.tensor(inp[k], dtype=dt) for k in ("G", "gd", "Gu"))
    times = []
    for rep in range(reps + 1):
        Z = torch.tensor(inp["Z"], dtype=dt, requires_grad=True)
        Wt = None if inp["W"] is None else torch.tensor(inp["W"], dtype=dt, requires_grad=True)
        groups = []
        for gi, (v, l) in enumerate(cfg.groups):
            ad = None if cfg.G == 1 else torch.arange(gi, cfg.D, cfg.G)
            groups.append((torch.tensor(float(v), dtype=dt, requires_grad=True),
                           torch.tensor(float(l), dtype=dt, requires_grad=True), ad))
        t0 = time.perf_counter()
        T.hot_path_step(X, Z, groups, g, Wt, G, gd, Gu, chunk=8)
        dt_s = time.perf_counter() - t0
        if rep > 0 or reps == 0:  # first pass is warm-up (unless it is the only one asked for)
            times.append(dt_s)
    return n_images / statistics.median(times), threads, times


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle/torch_twin port; the TF1/GPflow stack cannot be
    installed) on ALL host cores, on the full minibatch of the headline configuration, W warm-up + K timed steps."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = HEAD
    dtype = "f64" if args.dtype == "both" else args.dtype
    total = args.steps + args.warmup
    v, threads, times = cpu_port_images_per_s(cfg, dtype, cfg.N, max(0, total - 1))
    times = times[-args.steps:] if len(times) >= args.steps else times
    ms = 1e3 * statistics.mean(times)
    value = cfg.N / (ms * 1e-3)
    line = {
        "impl": "reference", "metric": METRIC, "value": value,
        "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
        "config": {"workload": cfg.workload, "images_per_gpu_per_step": cfg.N, "M": cfg.M,
                   "note": "reference TF1/GPflow stack not installable; oracle/torch_twin (same algorithm) on host cores, "
                           "the full %d-image minibatch per step" % cfg.N},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": threads, "kind": "port",
                         "sample": "the full %d-image minibatch per step, M=%d, fwd+bwd" % (cfg.N, cfg.M)},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ----------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:  # pragma: no cover
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if mask & bit:
                        self.reasons.add(nm)
            except Exception:  # pragma: no cover
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------- GPU arm
class Ctx:
    """Per-process state of the GPU arm: device, ranks, flush buffer, probed pipe peaks, the gradient collective."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.args = args
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback "
                             "(use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
                os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout: ONE JSON line
            dist.init_process_group("nccl", device_id=self.dev)
        self.flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=self.dev)  # > 126 MB L2
        self.peaks = {}
        self.launches = 0

    def flush_l2(self):
        self.flush_buf.fill_(self.rank & 0xFF)

    def probe_peaks(self):
        """Pipe rates of this GPU (MEASURED_PEAKS.json has HBM / bf16 only, neither bounds this path)."""
        import ctypes
        from convgp_b200 import _lib
        if self.peaks or self.rank != 0:
            return self.peaks
        for kind, name in ((0, "fp64_fma"), (1, "fp32_fma"), (2, "mufu_ex2"), (3, "f64_exp"), (4, "fp64_mma"),
                           (16, "tf32_mma"), (17, "tcgen05_tf32")):
            v = ctypes.c_double()
            _lib.check(_lib.lib().cgp_peak_probe(kind, ctypes.byref(v)))
            self.peaks[name] = v.value
        return self.peaks

    def max_over_ranks(self, values):
        import torch
        import torch.distributed as dist
        if self.world == 1:
            return list(values)
        t = torch.tensor(list(values), dtype=torch.float64, device=self.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist()

    def barrier(self):
        import torch.distributed as dist
        if self.world > 1:
            dist.barrier()


def build_hotpath(ctx, cfg, dtype_name, rank_seed, rows=None, gu_scale=1.0, collective=True):
    """HotPath of `cfg` on this rank.  rows = (lo, hi): this rank's rows of the SAME minibatch (strong scaling);
    None: the rank's own minibatch (weak scaling, seed 0+rank).  gu_scale: share of dL/dKuu this rank carries (Kuu is
    replicated, so under a sharded step every rank back-propagates 1/world of it and the all-reduce restores the sum)."""
    import torch
    from convgp_b200 import _lib
    from convgp_b200.hotpath import HotPath
    dt = torch.float64 if dtype_name == "f64" else torch.float32
    inp = make_inputs(cfg, rank_seed)
    if rows is not None:
        lo, hi = rows
        inp["X"], inp["G"], inp["gd"] = inp["X"][lo:hi], inp["G"][:, lo:hi], inp["gd"][lo:hi]
    inp["Gu"] = inp["Gu"] * gu_scale
    tens = {k: (None if v is None else torch.tensor(np.ascontiguousarray(v), dtype=dt, device=ctx.dev))
            for k, v in inp.items()}
    layout = _lib.LAYOUT_PLANAR if cfg.layout == "planar" else _lib.LAYOUT_COLOUR_PATCH
    mask = None
    if cfg.G > 1:  # one RBF per colour channel of the colour patch: dims c::C (cifar.py:33-41)
        mask = np.zeros((cfg.G, cfg.D), dtype=np.uint8)
        for g in range(cfg.G):
            mask[g, g::cfg.G] = 1
    spec = _lib.Spec(dt, cfg.H, cfg.W, cfg.C, cfg.ph, cfg.pw, layout, _lib.OUT_SUM, cfg.G, False,
                     cfg.layout == "colour", mask)
    hp = HotPath(spec, tens["X"], tens["Z"], tens["W"], tens["var"], tens["ls"], tens["G"], tens["gd"], tens["Gu"],
                 diag_add=cfg.diag_add, collective=collective and ctx.world > 1)
    return hp, inp, tens, spec


def time_steps(ctx, hp, steps, warmup, sample_clocks=False):
    """W untimed steps, then K steps each bracketed by CUDA events on the launching stream with the L2 flushed in
    between; barrier + synchronize on both sides; returns (total ms max over ranks, clocks)."""
    import torch
    for _ in range(max(warmup, 3)):
        hp.step()
    torch.cuda.synchronize()
    sampler = ClockSampler(ctx.local_rank) if sample_clocks else None
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    ctx.barrier()
    torch.cuda.synchronize()
    if sampler:
        sampler.start()
    for a, b in ev:
        ctx.flush_l2()
        a.record()
        hp.step()
        b.record()
    torch.cuda.synchronize()
    ctx.barrier()
    clocks = sampler.stop() if sampler else None
    total_ms = sum(a.elapsed_time(b) for a, b in ev)
    ctx.launches += hp.kernels_per_step * steps
    return ctx.max_over_ranks([total_ms])[0], clocks


def per_kernel_ms(ctx, hp):
    """CUDA events around each of the six launches, eager, L2 flushed (rank 0)."""
    import torch
    from convgp_b200 import _lib
    L, s, st, p = _lib.lib(), hp.spec, _lib.stream_ptr, _lib.ptr

    def k_kuf_fwd():
        _lib.check(L.cgp_kuf_fwd(s.ref(), hp.N, hp.M, p(hp.X), p(hp.Z), p(hp.W), p(hp.variance), p(hp.lengthscales),
                                 p(hp.Kuf), p(hp.ws), hp.ws_bytes, st()))

    def k_kuu_fwd():
        _lib.check(L.cgp_kuu_fwd(s.ref(), hp.M, p(hp.Z), p(hp.variance), p(hp.lengthscales), hp.diag_add,
                                 p(hp.Kuu), st()))

    def k_kuf_bwd():
        _lib.check(L.cgp_kuf_bwd(s.ref(), hp.N, hp.M, p(hp.X), p(hp.Z), p(hp.W), p(hp.variance), p(hp.lengthscales),
                                 p(hp.G), p(hp.dZ), p(hp.dW), p(hp.dvar), p(hp.dls), p(hp.ws), hp.ws_bytes, st()))

    def k_kuu_bwd():
        _lib.check(L.cgp_kuu_bwd(s.ref(), hp.M, p(hp.Z), p(hp.variance), p(hp.lengthscales), p(hp.Gu), p(hp.dZ),
                                 p(hp.dvar), p(hp.dls), st()))

    out = {}
    for name, fn in (("kuf_fwd", k_kuf_fwd), ("kdiag_fwd", lambda: hp._kdiag_fwd(st())), ("kuu_fwd", k_kuu_fwd),
                     ("kuf_bwd", k_kuf_bwd), ("kdiag_bwd", lambda: hp._kdiag_bwd(st())), ("kuu_bwd", k_kuu_bwd)):
        ts = []
        for i in range(13):
            ctx.flush_l2()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            if i >= 3:
                ts.append(a.elapsed_time(b))
        out[name] = statistics.mean(ts)
        ctx.launches += 13
    return out


def bound_fn(cfg, dtype_name, peaks):
    """T_bound(flops, exps) of the binding pipe(s) and a description."""
    D = cfg.D
    if dtype_name == "f64":
        fma_rate = max(peaks["fp64_fma"], peaks["fp64_mma"])  # DMMA and DFMA share the pipe; DMMA is the faster feed

        def t_bound(fl, ex):
            return fl / (2 * fma_rate) + ex / peaks["f64_exp"]
        return t_bound, "fp64", ("FP64 pipe: flops/(2*max(DFMA, DMMA) rate) + exps/(library exp rate), all probed on "
                                 "this GPU by cgp_peak_probe (MEASURED_PEAKS.json has no FP64 figure)")

    # float32 runs on three pipes at once: the contraction FMAs (2D of the 2D+2G / 2D+8G flops of a pair, 4D of the
    # 4D+8G of a Kuf backward pair) on tcgen05 tensor cores at three kind::f16 MMAs per float32 product (two fp16 terms
    # per operand, umma16.cuh; kind::f16 runs at twice the probed kind::tf32 rate), the exps on MUFU, everything else on
    # the FP32 FMA pipe.  The slowest of the three bounds the kernel.
    def t_bound(fl, ex):
        pairs = ex / cfg.G
        per_pair = fl / pairs
        contraction = pairs * (4 * D if per_pair > 3 * D + 4 * cfg.G else 2 * D)
        return max(3 * (contraction / 2) / (2 * peaks["tcgen05_tf32"]), ex / peaks["mufu_ex2"],
                   (fl - contraction) / (2 * peaks["fp32_fma"]))
    return t_bound, "mufu", ("max(3 * contraction FMAs / tcgen05 kind::f16 rate (= 2 x the probed kind::tf32 rate), exps / "
                             "MUFU ex2 rate, remaining flops / (2 * FFMA rate)): the three pipes run concurrently; MUFU "
                             "binds at this shape.  All rates probed on this GPU by cgp_peak_probe")


def step_bound(cfg, dtype_name, peaks, n_images, executed_only=False, kd_saved=True):
    work = algorithmic_work(cfg, n_images)
    t_bound, _, _ = bound_fn(cfg, dtype_name, peaks)
    names = [k for k in work if k != "step"]
    if executed_only and kd_saved:
        # The Kdiag backward consumes row sums its forward saved, so the library EXECUTES one sweep of the P x P
        # Gram per step, not the two that SURVEY.md 8d's algorithmic count assumes.
        names = [k for k in names if k != "kdiag_bwd"]
    if dtype_name == "f64":
        return t_bound(sum(work[k][0] for k in names), sum(work[k][1] for k in names))
    return sum(t_bound(*work[k]) for k in names)  # each kernel has its own binding pipe


class Pipeline:
    """Double-buffered end-to-end loop: while step i computes, a copy stream moves the minibatch of step i+1 from
    pinned host memory into the other device buffer and the result of step i-1 back to pinned host memory.  Every
    step still gets its own H2D copy of X and its own D2H read of the result; the host waits for the result of step
    i-1 before it enqueues step i+1 (it is never more than one step ahead)."""

    def __init__(self, h2d, compute, d2h):
        import torch
        self.h2d, self.compute, self.d2h = h2d, compute, d2h
        self.copy = torch.cuda.Stream()

    def run(self, n):
        import torch
        comp, copy = torch.cuda.current_stream(), self.copy
        ev_h2d, ev_comp, ev_d2h = [None, None], [None, None], [None, None]
        copy.wait_stream(comp)
        with torch.cuda.stream(copy):
            self.h2d(0)
            ev_h2d[0] = torch.cuda.Event()
            ev_h2d[0].record(copy)
        for i in range(n):
            b, nb = i % 2, (i + 1) % 2
            if i + 1 < n:  # prefetch the next minibatch into the buffer step i-1 has finished with
                with torch.cuda.stream(copy):
                    if ev_comp[nb] is not None:
                        copy.wait_event(ev_comp[nb])
                    self.h2d(nb)
                    ev_h2d[nb] = torch.cuda.Event()
                    ev_h2d[nb].record(copy)
            comp.wait_event(ev_h2d[b])
            if ev_d2h[b] is not None:
                comp.wait_event(ev_d2h[b])  # the result buffer of step i-2 has been read back
            self.compute(b)
            ev_comp[b] = torch.cuda.Event()
            ev_comp[b].record(comp)
            with torch.cuda.stream(copy):
                copy.wait_event(ev_comp[b])
                self.d2h(b)
                ev_d2h[b] = torch.cuda.Event()
                ev_d2h[b].record(copy)
            if i >= 1:
                ev_d2h[nb].synchronize()  # the host consumes the result of step i-1
        if n:
            ev_d2h[(n - 1) % 2].synchronize()
        comp.wait_stream(copy)


def measure_e2e(ctx, cfg, dtype_name, hp, inp, tens, steps):
    """The same step through the public kernel classes + torch autograd (host minibatch in, packed gradients + loss
    out, every step; CUDA-graph replay when capturable) and straight over the C ABI (HotPath).  Headline numbers are
    double-buffered (Pipeline); the strictly serial copy -> compute -> copy -> wait loop is reported next to them."""
    import torch
    import torch.distributed as dist
    from convgp_b200 import convkernels as ck, kernels as bk
    from convgp_b200.settings import settings
    world, dev = ctx.world, ctx.dev
    dt = tens["X"].dtype
    es = 8 if dtype_name == "f64" else 4
    settings.float_type = dt
    settings.device = dev
    kern = ck.WeightedConv(bk.RBF(cfg.D), [cfg.H, cfg.W], [cfg.ph, cfg.pw]) + bk.White(1, 1e-3)
    conv = kern.kern_list[0]
    Zp = tens["Z"].clone().requires_grad_(True)
    free = [conv.W.free, conv.basekern.variance.free, conv.basekern.lengthscales.free, kern.white.variance.free, Zp]
    X_host = torch.tensor(inp["X"], dtype=dt).pin_memory()
    n_out = sum(t.numel() for t in free) + 1
    out_host = [torch.empty(n_out, dtype=dt).pin_memory() for _ in range(2)]
    X_dev = [torch.empty_like(tens["X"]) for _ in range(2)]

    def e2e_body(b):
        for t in free:
            t.grad = None
        # the package's own evaluation of one SVGP prediction's kernel terms (kernels.svgp_terms: what SVGP.build_predict
        # calls -- one autograd node, the three launches on parallel streams)
        kd, kuf, kuu = bk.svgp_terms(kern, Zp, X_dev[b], jitter=0.0)
        loss = (kuf * tens["G"]).sum() + (kd * tens["gd"]).sum() + (kuu * tens["Gu"]).sum()
        loss.backward()
        return torch.cat([t.grad.reshape(-1) for t in free] + [loss.detach().reshape(1)])

    def e2e_eager_step():
        X_dev[0].copy_(X_host, non_blocking=True)  # H2D of the minibatch (the reference feeds it per step)
        packed = e2e_body(0)
        if world > 1:
            dist.all_reduce(packed)
        out_host[0].copy_(packed, non_blocking=True)  # D2H of gradients + loss
        torch.cuda.current_stream().synchronize()

    def timed(fn, n):
        ctx.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    def timed_pipeline(pipe, n):
        pipe.run(3)
        ctx.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pipe.run(n)
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    for _ in range(3):
        e2e_eager_step()
    e2e_steps = max(10, min(steps, 100))
    e2e_eager_s = timed(e2e_eager_step, e2e_steps)
    ctx.launches += 6 * (e2e_steps + 3)

    e2e_s, e2e_serial_s, e2e_mode = e2e_eager_s, e2e_eager_s, "eager"
    if not ctx.args.no_graph:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    e2e_body(0)
                    e2e_body(1)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graphs, packed_static = [], []
            for b in range(2):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    packed_static.append(e2e_body(b))
                graphs.append(g)

            def compute(b):
                graphs[b].replay()
                if world > 1:
                    dist.all_reduce(packed_static[b])

            def e2e_graph_step():  # strictly serial: copy in, compute, copy out, wait
                X_dev[0].copy_(X_host, non_blocking=True)
                compute(0)
                out_host[0].copy_(packed_static[0], non_blocking=True)
                torch.cuda.current_stream().synchronize()

            e2e_graph_step()
            ref_out = out_host[0].clone()
            e2e_eager_step()
            tol = dict(rtol=1e-6 if dtype_name == "f64" else 1e-3, atol=1e-6)
            if not torch.allclose(ref_out, out_host[0], **tol):
                raise RuntimeError("graph-replayed step disagrees with the eager step")
            e2e_serial_s = timed(e2e_graph_step, e2e_steps)
            pipe = Pipeline(lambda b: X_dev[b].copy_(X_host, non_blocking=True), compute,
                            lambda b: out_host[b].copy_(packed_static[b], non_blocking=True))
            e2e_s = timed_pipeline(pipe, e2e_steps)
            for b in range(2):  # both buffers of the pipelined loop hold the same step's result as the serial loop
                if not torch.allclose(ref_out, out_host[b], **tol):
                    raise RuntimeError("pipelined step disagrees with the serial step")
            e2e_mode = "cuda_graph, double-buffered"
            ctx.launches += 6 * (2 * e2e_steps + 3)
        except Exception as exc:  # keep the eager number, say why
            e2e_s, e2e_mode = e2e_serial_s, "eager (graph capture failed: %s)" % str(exc)[:120]

    # ---- straight over the C ABI: HotPath.step() on pre-allocated buffers
    ng = hp.grads.numel()
    out_c = [torch.empty(ng + hp.Kdiag.numel(), dtype=dt).pin_memory() for _ in range(2)]
    res_dev = [torch.empty(ng + hp.Kdiag.numel(), dtype=dt, device=dev) for _ in range(2)]
    X_stage = [torch.empty_like(hp.X) for _ in range(2)]

    def cabi_compute(b):
        hp.X.copy_(X_stage[b], non_blocking=True)  # device-to-device: the step's graph reads the one static buffer
        hp.step()
        res_dev[b][:ng].copy_(hp.grads, non_blocking=True)
        res_dev[b][ng:].copy_(hp.Kdiag.reshape(-1), non_blocking=True)

    def e2e_cabi_step():
        X_stage[0].copy_(X_host, non_blocking=True)
        cabi_compute(0)
        out_c[0].copy_(res_dev[0], non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(3):
        e2e_cabi_step()
    e2e_cabi_serial_s = timed(e2e_cabi_step, e2e_steps)
    pipe_c = Pipeline(lambda b: X_stage[b].copy_(X_host, non_blocking=True), cabi_compute,
                      lambda b: out_c[b].copy_(res_dev[b], non_blocking=True))
    e2e_cabi_s = timed_pipeline(pipe_c, e2e_steps)
    ctx.launches += hp.kernels_per_step * (2 * e2e_steps + 6)
    e2e_s, e2e_serial_s, e2e_eager_s, e2e_cabi_s, e2e_cabi_serial_s = ctx.max_over_ranks(
        [e2e_s, e2e_serial_s, e2e_eager_s, e2e_cabi_s, e2e_cabi_serial_s])
    images = cfg.N * world
    h2d = cfg.N * cfg.H * cfg.W * cfg.C * es
    return {"value": images * e2e_steps / e2e_s, "unit": "images/s", "h2d_bytes_per_step": h2d,
            "d2h_bytes_per_step": n_out * es, "steps": e2e_steps, "mode": e2e_mode,
            "serial_value": images * e2e_steps / e2e_serial_s,
            "eager_value": images * e2e_steps / e2e_eager_s,
            "api": "kernel classes (Kzx / Kdiag / Kzz) + torch autograd; X from pinned host memory, packed "
                   "gradients + loss back to host memory, every step.  value: double-buffered (the copy of step i+1's "
                   "minibatch and the read-back of step i-1's result overlap step i's kernels; the host waits for "
                   "every result, one step behind); serial_value: copy in -> compute -> copy out -> wait, one step at "
                   "a time; eager_value: the same without CUDA-graph replay",
            "c_abi": {"value": images * e2e_steps / e2e_cabi_s, "unit": "images/s", "h2d_bytes_per_step": h2d,
                      "serial_value": images * e2e_steps / e2e_cabi_serial_s,
                      "d2h_bytes_per_step": int(out_c[0].numel()) * es,
                      "api": "HotPath.step() over the C ABI (pre-allocated buffers, one CUDA graph incl. the gradient "
                             "all-reduce), same host-to-device and device-to-host copies every step, double-buffered "
                             "like value; serial_value: one step at a time"}}


def _traffic(dtype_name, kernel):
    """DRAM bytes per launch of the dominant kernel, from the committed ncu --set full capture (latest first)."""
    try:
        for name in ("r2z_traffic.json", "r1z_traffic.json", "r1_traffic.json"):
            path = os.path.join(ROOT, "profiles", name)
            if os.path.exists(path):
                v = json.load(open(path)).get(dtype_name, {}).get(kernel)
                if v is not None:
                    return v
    except Exception:
        pass
    return None


def measure_headline(ctx, dtype_name, with_cpu=True):
    """Config 3 in one dtype: weak-scaling value, per-kernel roofline, e2e.  Returns the dict of the JSON line's
    dtype-specific fields (rank 0; other ranks get the collective parts only)."""
    import torch  # noqa: F401
    args, cfg = ctx.args, HEAD
    es = 8 if dtype_name == "f64" else 4
    hp, inp, tens, spec = build_hotpath(ctx, cfg, dtype_name, ctx.rank)
    if not args.no_graph:
        hp.capture(overlapped=True)
    total_ms, clocks = time_steps(ctx, hp, args.steps, args.warmup, sample_clocks=True)
    e2e = measure_e2e(ctx, cfg, dtype_name, hp, inp, tens, args.steps)
    out = {"total_ms": total_ms}
    if ctx.rank != 0:
        return out
    peaks = ctx.probe_peaks()
    pk = per_kernel_ms(ctx, hp)
    work = algorithmic_work(cfg, cfg.N)
    t_bound, pipe, peak_note = bound_fn(cfg, dtype_name, peaks)
    dom = max(pk, key=pk.get)
    fl, ex = work[dom]
    dur = pk[dom] * 1e-3
    achieved = fl / dur / 1e12
    peak = fl / t_bound(fl, ex) / 1e12
    ms_per_step = total_ms / args.steps
    kd_saved = hp.kd_saved is not None
    step_frac = step_bound(cfg, dtype_name, peaks, cfg.N) / (ms_per_step * 1e-3)
    step_frac_exec = step_bound(cfg, dtype_name, peaks, cfg.N, True, kd_saved) / (ms_per_step * 1e-3)
    hbm_peak = None
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    images = cfg.N * ctx.world
    out.update({
        "value": images * args.steps / (total_ms * 1e-3), "ms_per_step": ms_per_step, "dtype": dtype_name,
        "roofline": {
            "bound": pipe, "kernel": dom, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
            "frac": achieved / peak, "traffic": _traffic(dtype_name, dom), "peak_source": peak_note,
            "kernel_ms": pk, "kernel_frac": {k: t_bound(*work[k]) / (pk[k] * 1e-3) for k in pk},
            "step_frac": step_frac, "step_frac_executed_work": step_frac_exec, "probes": peaks,
            "hbm": {"algorithmic_bytes_per_step": algorithmic_bytes(cfg, cfg.N, es),
                    "achieved_gbs": algorithmic_bytes(cfg, cfg.N, es) / (ms_per_step * 1e-3) / 1e9,
                    "peak_gbs_measured": hbm_peak},
        },
        "e2e": e2e, "clocks": clocks, "collective": hp.collective_name,
    })
    if with_cpu:
        cpu_val, cores, _ = cpu_port_images_per_s(cfg, dtype_name, cfg.N, 2)
        out["cpu_baseline"] = {"value": cpu_val, "unit": "images/s", "cores": cores, "kind": "port",
                               "sample": "the full %d-image minibatch, M=%d, fwd+bwd, median of 2 after one warm-up"
                                         % (cfg.N, cfg.M)}
    return out


def measure_strong(ctx, dtype_name):
    """N > 1: the SAME 200-image minibatch split by rows over the ranks (parallel.shard_bounds), replicated
    parameters and Kuu, the gradient all-reduce inside the step; next to it the unsharded step on one GPU (every rank
    runs it, rank 0's gradient is the reference of grad_check)."""
    import torch
    from convgp_b200 import parallel
    args, cfg = ctx.args, HEAD
    lo, hi = parallel.shard_bounds(cfg.N, ctx.world, ctx.rank)
    hp, _, _, _ = build_hotpath(ctx, cfg, dtype_name, 0, rows=(lo, hi), gu_scale=1.0 / ctx.world)
    full, _, _, _ = build_hotpath(ctx, cfg, dtype_name, 0, collective=False)
    if not args.no_graph:
        hp.capture(overlapped=True)
        full.capture(overlapped=True)
    sharded_ms, _ = time_steps(ctx, hp, args.steps, args.warmup)
    # single-GPU time of the same 200 images: rank 0's own clock (no collective, no other rank involved)
    for _ in range(3):
        full.step()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in ev:
        ctx.flush_l2()
        a.record()
        full.step()
        b.record()
    torch.cuda.synchronize()
    full_ms = sum(a.elapsed_time(b) for a, b in ev)
    ctx.launches += full.kernels_per_step * args.steps
    # gradient check: all-reduced sharded gradient against the unsharded one, per segment, relative to max |.|
    hp.step()
    full.step()
    torch.cuda.synchronize()
    worst = 0.0
    for a, b in zip((hp.dZ, hp.dW, hp.dvar, hp.dls), (full.dZ, full.dW, full.dvar, full.dls)):
        if a is None:
            continue
        den = float(b.abs().max())
        worst = max(worst, float((a - b).abs().max()) / (den if den > 0 else 1.0))
    worst = ctx.max_over_ranks([worst])[0]
    ctx.barrier()
    if ctx.rank != 0:
        return None
    ms, ms1 = sharded_ms / args.steps, full_ms / args.steps
    return {"images_total": cfg.N, "rows_per_rank": [parallel.shard_bounds(cfg.N, ctx.world, r)[1]
                                                      - parallel.shard_bounds(cfg.N, ctx.world, r)[0]
                                                      for r in range(ctx.world)],
            "ms_per_step": ms, "value": cfg.N / (ms * 1e-3), "unit": "images/s",
            "ms_per_step_1gpu": ms1, "speedup_vs_1gpu": ms1 / ms, "collective": hp.collective_name,
            "grad_check": worst, "grad_check_tol": 1e-10 if dtype_name == "f64" else 1e-4,
            "note": "grad_check = max over [dZ|dW|dvar|dls] of max|all-reduced sharded gradient - rank 0's unsharded "
                    "gradient| / max|unsharded|, worst rank"}


def measure_secondary(ctx, key, dtype_name):
    """A secondary BASELINE config: device-resident hot-path step, images/s and the step's roofline fraction.  Under
    N > 1 the minibatch rows are sharded over the ranks (config 5: 30 rows -> 4,4,4,4,4,4,3,3 at 8)."""
    from convgp_b200 import parallel
    args, cfg = ctx.args, CONFIGS[key]
    rows, gu = None, 1.0
    if ctx.world > 1:
        rows, gu = parallel.shard_bounds(cfg.N, ctx.world, ctx.rank), 1.0 / ctx.world
    hp, _, _, _ = build_hotpath(ctx, cfg, dtype_name, 0, rows=rows, gu_scale=gu)
    if not args.no_graph:
        hp.capture(overlapped=True)
    steps = max(20, min(args.steps, 100))
    total_ms, _ = time_steps(ctx, hp, steps, min(args.warmup, 5))
    if ctx.rank != 0:
        return None
    peaks = ctx.probe_peaks()
    ms = total_ms / steps
    pk = per_kernel_ms(ctx, hp) if ctx.world == 1 else None
    kd_saved = hp.kd_saved is not None
    out = {"workload": cfg.workload, "dtype": dtype_name, "images_per_step": cfg.N, "M": cfg.M, "ms_per_step": ms,
           "value": cfg.N / (ms * 1e-3), "unit": "images/s",
           "step_frac": step_bound(cfg, dtype_name, peaks, cfg.N) / (ms * 1e-3),
           "step_frac_executed_work": step_bound(cfg, dtype_name, peaks, cfg.N, True, kd_saved) / (ms * 1e-3),
           "families": hp.families()}
    if pk:
        work = algorithmic_work(cfg, cfg.N)
        t_bound, pipe, _ = bound_fn(cfg, dtype_name, peaks)
        out["kernel_ms"] = pk
        out["kernel_frac"] = {k: t_bound(*work[k]) / (pk[k] * 1e-3) for k in pk}
        out["bound"] = pipe
    if ctx.world > 1:
        out["rows_per_rank"] = [parallel.shard_bounds(cfg.N, ctx.world, r)[1] - parallel.shard_bounds(cfg.N, ctx.world, r)[0]
                                for r in range(ctx.world)]
        out["scaling"] = "strong"
    return out


def run_gpu(args):
    import torch.distributed as dist
    ctx = Ctx(args)
    dtypes = ["f64", "f32"] if args.dtype == "both" else [args.dtype]
    head = measure_headline(ctx, dtypes[0])
    sub = {d: measure_headline(ctx, d, with_cpu=False) for d in dtypes[1:]}
    strong = {}
    if ctx.world > 1:
        for d in dtypes:
            strong[d] = measure_strong(ctx, d)
    configs = {}
    if args.secondary:
        for key in ("cfg1", "cfg2", "cfg4", "cfg5"):
            for d in dtypes:
                try:
                    configs.setdefault(key, {})[d] = measure_secondary(ctx, key, d)
                except Exception as exc:  # a secondary line must not take the headline down; say what happened
                    configs.setdefault(key, {})[d] = {"error": str(exc)[:300]}
    extras = {}
    if args.svgp_step and ctx.world == 1:
        try:
            from tools import svgp_step_time
            extras["svgp_step"] = {d: svgp_step_time.measure(d, steps=20) for d in dtypes}
        except Exception as exc:
            extras["svgp_step"] = {"error": str(exc)[:300]}
    if args.predict:
        try:
            from tools import predict_bench
            extras["predict"] = predict_bench.measure(ctx, dtypes[0])
        except Exception as exc:
            extras["predict"] = {"error": str(exc)[:300]}

    if ctx.rank != 0:
        if ctx.world > 1:
            dist.destroy_process_group()
        return
    cfg = HEAD
    line = {
        "metric": METRIC, "value": head["value"], "unit": "images/s", "n_gpus": ctx.world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": head["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": head["dtype"], "data": "synthetic",
        "config": {"workload": cfg.workload, "images_per_gpu_per_step": cfg.N, "M": cfg.M,
                   "l2": "256 MB buffer written between timed steps (inputs are ~5 MB, far below the 126 MB L2)",
                   "cuda_graph": not args.no_graph,
                   "graph_branches": "Kuf | Kdiag | Kuu forward on parallel branches, JOIN, then the three backward "
                                     "kernels on parallel branches (dL/dKuf depends on all three forwards in the ELBO)",
                   "collective": head["collective"] if ctx.world > 1 else "none"},
        "roofline": head["roofline"], "cpu_baseline": head["cpu_baseline"], "e2e": head["e2e"],
        "gpu_launches": ctx.launches, "clocks": head["clocks"],
    }
    for d, s in sub.items():
        line[d] = {k: s[k] for k in ("value", "ms_per_step", "dtype", "roofline", "e2e", "clocks")}
    if strong:
        line["strong"] = strong[dtypes[0]]
        line["grad_check"] = strong[dtypes[0]]["grad_check"]
        for d in dtypes[1:]:
            line[d]["strong"] = strong[d]
            line[d]["grad_check"] = strong[d]["grad_check"]
    if configs:
        line["configs"] = configs
    line.update(extras)
    emit(line)
    if ctx.world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def _claim_stdout():
    """ONE JSON line on stdout: anything a library prints to file descriptor 1 (NCCL's version banner does, whatever
    NCCL_DEBUG says) is sent to stderr instead; emit() writes the line to the original descriptor."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--dtype", choices=["both", "f64", "f32"], default="both",
                    help="both (default): float64 headline + an 'f32' sub-object with the tcgen05 path")
    ap.add_argument("--impl", choices=["native", "reference"], default="native")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--secondary", action=argparse.BooleanOptionalAction, default=True,
                    help="also time configs 1, 2, 4, 5 ('configs' sub-object)")
    ap.add_argument("--svgp-step", action=argparse.BooleanOptionalAction, default=True,
                    help="also time the full SVGP training step of config 3 ('svgp_step', one GPU)")
    ap.add_argument("--predict", action=argparse.BooleanOptionalAction, default=True,
                    help="also time the 10 000-image forward sweep ('predict'; rows sharded over the ranks)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
