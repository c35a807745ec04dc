#!/usr/bin/env python
"""Benchmark of the convgp patch-kernel hot path (BASELINE.json metric):

    conv/wconv Kuf + Kdiag (+ Kuu) forward + backward, images/s, 28x28, 5x5 patches, M=750,
    minibatch 200 per GPU (config 3: mnist.py -k wconv -M 750 --minibatch-size 200), float64 (default)
    or float32.

One "step" = one pass of the hot path over one minibatch: Kuf, Kdiag, Kuu forward, then their
backward w.r.t. Z, W, variance, lengthscales from given upstream gradients.
N>1: one process per GPU (torchrun), each rank runs the step on its own 200 images (weak scaling)
and the packed gradient buffer is all-reduced over NCCL every step inside the timed region.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--dtype f64|f32] [--impl reference]

Prints ONE JSON line (rank 0).  See DESIGN.md section 6 for how each field is obtained.
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

H = W_IMG = 28
PH = PW = 5
M_IND = 750
N_BATCH = 200
P = (H - PH + 1) * (W_IMG - PW + 1)  # 576
D = PH * PW  # 25
WORKLOAD = "mnist.py -k wconv -M 750 --minibatch-size 200 (28x28x1, 5x5 patches, P=576, D=25)"


# ----------------------------------------------------------------------------- workload
def im2col(X):
    N = X.shape[0]
    img = X.reshape(N, H, W_IMG)
    out = np.empty((N, H - PH + 1, W_IMG - PW + 1, PH, PW))
    for i in range(PH):
        for j in range(PW):
            out[:, :, :, i, j] = img[:, i:i + H - PH + 1, j:j + W_IMG - PW + 1]
    return out.reshape(N, P, D)


def make_inputs(rank, n_images=N_BATCH):
    """Seeded synthetic inputs of config 3: X ~ U[0,1) (seed 0+rank), Z = M patches of X + 1e-3 U
    (seed 1), l = 1, var = 1, W = 1, upstream gradients ~ N(0,1) (seeds 4,5,6)."""
    X = np.random.default_rng(0 + rank).random((n_images, H * W_IMG))
    r1 = np.random.default_rng(1)
    pat = im2col(np.random.default_rng(0).random((N_BATCH, H * W_IMG))).reshape(-1, D)
    Z = pat[r1.permutation(len(pat))[:M_IND]] + 0.001 * r1.random((M_IND, D))
    G = np.random.default_rng(4).standard_normal((M_IND, n_images))
    gd = np.random.default_rng(5).standard_normal(n_images)
    Gu = np.random.default_rng(6).standard_normal((M_IND, M_IND))
    return dict(X=X, Z=Z, W=np.ones(P), var=np.array([1.0]), ls=np.array([1.0]), G=G, gd=gd, Gu=Gu)


def algorithmic_work(n_images):
    """FMA-pipe flops and exps of one step: per unique pair, forward 2D+2 flops + 1 exp;
    backward 4D+8 (Kuf) / 2D+8 (Kdiag) flops + 1 exp (recompute)."""
    kuf = n_images * P * M_IND
    kdiag = n_images * P * (P + 1) // 2
    kuu = M_IND * M_IND
    w = {
        "kuf_fwd": (kuf * (2 * D + 2), kuf), "kdiag_fwd": (kdiag * (2 * D + 2), kdiag),
        "kuu_fwd": (kuu * (2 * D + 2), kuu), "kuf_bwd": (kuf * (4 * D + 8), kuf),
        "kdiag_bwd": (kdiag * (2 * D + 8), kdiag), "kuu_bwd": (kuu * (4 * D + 8), kuu),
    }
    w["step"] = (sum(v[0] for v in w.values()), sum(v[1] for v in w.values()))
    # algorithmic HBM bytes/step: X read fwd+bwd, Z, Kuf out + G in, small vectors
    return w


def algorithmic_bytes(n_images, es):
    return es * (2 * n_images * H * W_IMG + 4 * M_IND * D + 2 * M_IND * n_images + 2 * M_IND * M_IND + 4 * P
                 + 2 * n_images)


# ----------------------------------------------------------------------------- CPU reference arm
def cpu_port_images_per_s(dtype_name, n_images, reps, threads=None):
    """The reference's algorithm on host cores: oracle/torch_twin (materialised patches -> GEMM ->
    exp -> patch sums, per-image Kdiag loop, autodiff backward)."""
    import torch
    from oracle import torch_twin as T
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    dt = torch.float64 if dtype_name == "f64" else torch.float32
    inp = make_inputs(0, n_images)
    g = T.Geom(H, W_IMG, 1, PH, PW, "conv")
    X = torch.tensor(inp["X"], dtype=dt)
    G, gd, Gu = (torch.tensor(inp[k], dtype=dt) for k in ("G", "gd", "Gu"))
    times = []
    for rep in range(reps + 1):
        Z = torch.tensor(inp["Z"], dtype=dt, requires_grad=True)
        Wt = torch.tensor(inp["W"], dtype=dt, requires_grad=True)
        v = torch.tensor(1.0, dtype=dt, requires_grad=True)
        l = torch.tensor(1.0, dtype=dt, requires_grad=True)
        t0 = time.perf_counter()
        T.hot_path_step(X, Z, [(v, l, None)], g, Wt, G, gd, Gu, chunk=8)
        dt_s = time.perf_counter() - t0
        if rep > 0:  # first pass is warm-up
            times.append(dt_s)
    return n_images / statistics.median(times), threads, times


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = 40
    per_step = []
    total = args.steps + args.warmup
    v, threads, times = cpu_port_images_per_s(args.dtype, sample, max(1, total - 1))
    times = times[-args.steps:] if len(times) >= args.steps else times
    ms = 1e3 * statistics.mean(times)
    value = sample / (ms * 1e-3)
    line = {
        "impl": "reference", "metric": "conv/wconv Kuf+Kdiag fwd+bwd images/sec (28x28, M=750)", "value": value,
        "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": WORKLOAD, "images_per_step": sample,
                   "note": "reference TF1/GPflow stack not installable; oracle/torch_twin (same algorithm) on host cores"},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": threads, "kind": "port",
                         "sample": "%d images per step of the %d-image minibatch, M=750, fwd+bwd" % (sample, N_BATCH)},
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
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from convgp_b200 import _lib
    from convgp_b200.hotpath import HotPath

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout: ONE JSON line
        dist.init_process_group("nccl", device_id=dev)
    dt = torch.float64 if args.dtype == "f64" else torch.float32
    es = 8 if args.dtype == "f64" else 4

    inp = make_inputs(rank)
    tens = {k: torch.tensor(v, dtype=dt, device=dev) for k, v in inp.items()}
    spec = _lib.Spec(dt, H, W_IMG, 1, PH, PW, _lib.LAYOUT_PLANAR, _lib.OUT_SUM, 1, False, False)
    hp = HotPath(spec, tens["X"], tens["Z"], tens["W"], tens["var"], tens["ls"], tens["G"], tens["gd"], tens["Gu"],
                 diag_add=1e-3 + 1e-5)  # White(1, 1e-3) + jitter (mnist.py:42, gpflowrc:11)
    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def flush_l2():
        flush_buf.fill_(rank & 0xFF)

    def allreduce():
        if world > 1:
            dist.all_reduce(hp.grads)

    # ---- peaks of the binding pipes, measured on this box (MEASURED_PEAKS.json has HBM / bf16 only)
    import ctypes
    peaks = {}
    if rank == 0:
        for kind, name in ((0, "fp64_fma"), (1, "fp32_fma"), (2, "mufu_ex2"), (3, "f64_exp"), (4, "fp64_mma"), (16, "tf32_mma"),
                           (17, "tcgen05_tf32")):
            v = ctypes.c_double()
            _lib.check(_lib.lib().cgp_peak_probe(kind, ctypes.byref(v)))
            peaks[name] = v.value

    collective_in_graph = False
    if not args.no_graph:
        # Kuf / Kdiag / Kuu of each phase on parallel graph branches; with several ranks the NCCL all-reduce of the
        # packed gradients is captured behind them, so a data-parallel step is one graph launch
        if world > 1 and args.graph_collective:
            try:
                hp.capture(overlapped=True, after=lambda: dist.all_reduce(hp.grads))
                collective_in_graph = True
            except Exception as exc:  # keep the collective outside the graph, say why
                sys.stderr.write("all-reduce not captured (%s); launching it after the graph\n" % str(exc)[:200])
                hp.graph = None
        if hp.graph is None:
            hp.capture(overlapped=True)
    if world > 1:  # every rank must take the same path
        flag = torch.tensor([1.0 if collective_in_graph else 0.0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if collective_in_graph and flag.item() < 1.0:
            collective_in_graph = False
            hp.capture(overlapped=True)
    if collective_in_graph:
        def allreduce():  # noqa: F811  (already inside hp.step())
            pass
    # ---- warm-up
    for _ in range(max(args.warmup, 3)):
        hp.step()
        allreduce()
    torch.cuda.synchronize()

    # ---- timed region: K steps, device-resident inputs, CUDA events per step, L2 flushed between steps
    sampler = ClockSampler(local_rank)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    for a, b in ev:
        flush_l2()
        a.record()
        hp.step()
        allreduce()
        b.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop()
    total_ms = sum(a.elapsed_time(b) for a, b in ev)

    # ---- end-to-end through the public kernel-class API, host buffers in, gradients + loss out
    from convgp_b200 import convkernels as ck, kernels as bk
    from convgp_b200.settings import settings
    settings.float_type = dt
    settings.device = dev
    kern = ck.WeightedConv(bk.RBF(D), [H, W_IMG], [PH, PW]) + bk.White(1, 1e-3)
    conv = kern.kern_list[0]
    Zp = tens["Z"].clone().requires_grad_(True)
    free = [conv.W.free, conv.basekern.variance.free, conv.basekern.lengthscales.free, kern.white.variance.free, Zp]
    X_host = torch.tensor(inp["X"], dtype=dt).pin_memory()
    n_out = sum(t.numel() for t in free) + 1
    out_host = torch.empty(n_out, dtype=dt).pin_memory()
    X_dev = torch.empty_like(tens["X"])

    def e2e_body():
        X_dev.copy_(X_host, non_blocking=True)  # H2D of the minibatch (the reference feeds it per step)
        for t in free:
            t.grad = None
        kuf, kd, kuu = kern.Kzx(Zp, X_dev), kern.Kdiag(X_dev), kern.Kzz(Zp)
        loss = (kuf * tens["G"]).sum() + (kd * tens["gd"]).sum() + (kuu * tens["Gu"]).sum()
        loss.backward()
        return torch.cat([t.grad.reshape(-1) for t in free] + [loss.detach().reshape(1)])

    def e2e_eager_step():
        packed = e2e_body()
        if world > 1:
            dist.all_reduce(packed)
        out_host.copy_(packed, non_blocking=True)  # D2H of gradients + loss
        torch.cuda.current_stream().synchronize()

    def timed(fn, n):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    for _ in range(3):
        e2e_eager_step()
    e2e_steps = max(10, min(args.steps, 100))
    e2e_eager_s = timed(e2e_eager_step, e2e_steps)

    # The same public-API step (kernel classes + autograd, H2D in, D2H out) captured once in a CUDA
    # graph and replayed: the deployment pattern for a ~1 ms step, where Python dispatch of the
    # small parameter-transform ops otherwise costs as much as the kernels.  The NCCL all-reduce
    # stays outside the graph.
    e2e_s, e2e_mode = e2e_eager_s, "eager"
    if not args.no_graph:
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    e2e_body()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g_e2e = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_e2e):
                packed_static = e2e_body()
                if world == 1:
                    out_host.copy_(packed_static, non_blocking=True)

            def e2e_graph_step():
                g_e2e.replay()
                if world > 1:
                    dist.all_reduce(packed_static)
                    out_host.copy_(packed_static, non_blocking=True)
                torch.cuda.current_stream().synchronize()

            e2e_graph_step()
            ref_out = out_host.clone()
            e2e_eager_step()
            if not torch.allclose(ref_out, out_host, rtol=1e-6 if args.dtype == "f64" else 1e-3, atol=1e-6):
                raise RuntimeError("graph-replayed step disagrees with the eager step")
            e2e_s, e2e_mode = timed(e2e_graph_step, e2e_steps), "cuda_graph"
        except Exception as exc:  # keep the eager number, say why
            e2e_mode = "eager (graph capture failed: %s)" % str(exc)[:120]

    # ---- the same step driven straight through the C ABI (HotPath: pre-allocated buffers, the three-branch CUDA
    # graph), still with the minibatch coming from pinned host memory and the packed gradients + Kdiag going back
    # to host memory every step.  Reported next to the kernel-class number: the difference is torch's autograd
    # and parameter-transform dispatch, not the library.
    out_c = torch.empty(hp.grads.numel() + hp.Kdiag.numel(), dtype=dt).pin_memory()
    ng = hp.grads.numel()

    def e2e_cabi_step():
        hp.X.copy_(X_host, non_blocking=True)
        hp.step()
        allreduce()
        out_c[:ng].copy_(hp.grads, non_blocking=True)
        out_c[ng:].copy_(hp.Kdiag.reshape(-1), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(3):
        e2e_cabi_step()
    e2e_cabi_s = timed(e2e_cabi_step, e2e_steps)

    # ---- max over ranks
    if world > 1:
        tt = torch.tensor([total_ms, e2e_s, e2e_eager_s, e2e_cabi_s], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms, e2e_s, e2e_eager_s, e2e_cabi_s = tt.tolist()

    # ---- per-kernel durations (rank 0): CUDA events around each launch, eager, L2 flushed
    per_kernel = {}
    if rank == 0:
        L, s, st = _lib.lib(), spec, _lib.stream_ptr
        p = _lib.ptr

        def k_kuf_fwd():
            _lib.check(L.cgp_kuf_fwd(s.ref(), hp.N, hp.M, p(hp.X), p(hp.Z), p(hp.W), p(hp.variance), p(hp.lengthscales),
                                     p(hp.Kuf), None, 0, st()))

        def k_kdiag_fwd():  # the forward that stores the per-image row sums its backward is made of, when covered
            hp._kdiag_fwd(st())

        def k_kuu_fwd():
            _lib.check(L.cgp_kuu_fwd(s.ref(), hp.M, p(hp.Z), p(hp.variance), p(hp.lengthscales), hp.diag_add,
                                     p(hp.Kuu), st()))

        def k_kuf_bwd():
            _lib.check(L.cgp_kuf_bwd(s.ref(), hp.N, hp.M, p(hp.X), p(hp.Z), p(hp.W), p(hp.variance), p(hp.lengthscales),
                                     p(hp.G), p(hp.dZ), p(hp.dW), p(hp.dvar), p(hp.dls), None, 0, st()))

        def k_kdiag_bwd():
            hp._kdiag_bwd(st())

        def k_kuu_bwd():
            _lib.check(L.cgp_kuu_bwd(s.ref(), hp.M, p(hp.Z), p(hp.variance), p(hp.lengthscales), p(hp.Gu), p(hp.dZ),
                                     p(hp.dvar), p(hp.dls), st()))

        for name, fn in (("kuf_fwd", k_kuf_fwd), ("kdiag_fwd", k_kdiag_fwd), ("kuu_fwd", k_kuu_fwd),
                         ("kuf_bwd", k_kuf_bwd), ("kdiag_bwd", k_kdiag_bwd), ("kuu_bwd", k_kuu_bwd)):
            ts = []
            for i in range(13):
                flush_l2()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                b.record()
                torch.cuda.synchronize()
                if i >= 3:
                    ts.append(a.elapsed_time(b))
            per_kernel[name] = statistics.mean(ts)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline
    work = algorithmic_work(N_BATCH)
    if args.dtype == "f64":
        fma_rate = max(peaks["fp64_fma"], peaks["fp64_mma"])  # DMMA and DFMA share the pipe; DMMA is the faster feed

        def t_bound(fl, ex):
            return fl / (2 * fma_rate) + ex / peaks["f64_exp"]
        pipe, peak_note = "fp64", ("FP64 pipe: flops/(2*max(DFMA, DMMA) rate) + exps/(library exp rate), all probed on "
                                   "this GPU by cgp_peak_probe (MEASURED_PEAKS.json has no FP64 figure)")
    else:
        # float32 runs on three pipes at once: the contraction FMAs (2D of the 2D+2 / 2D+8 flops of a pair, 4D of
        # the 4D+8 of a Kuf backward pair) on tcgen05 tensor cores at three TF32 MMAs per float32 product, the exps on
        # MUFU, everything else on the FP32 FMA pipe.  The slowest of the three bounds the kernel.
        def t_bound(fl, ex):
            per_pair = fl / ex
            contraction = ex * (4 * D if per_pair > 3 * D else 2 * D)
            return max(3 * (contraction / 2) / peaks["tcgen05_tf32"], ex / peaks["mufu_ex2"],
                       (fl - contraction) / (2 * peaks["fp32_fma"]))
        pipe, peak_note = "mufu", ("max(3 * contraction FMAs / tcgen05 kind::tf32 rate, exps / MUFU ex2 rate, remaining flops / "
                                   "(2 * FFMA rate)): the three pipes run concurrently; MUFU binds at this shape.  All rates "
                                   "probed on this GPU by cgp_peak_probe")
    dom = max(per_kernel, key=per_kernel.get)
    fl, ex = work[dom]
    if args.dtype != "f64":  # the step bound is the sum of the per-kernel bounds (each kernel has its own binding pipe)
        step_bound = sum(t_bound(*work[k]) for k in work if k != "step")
    dur = per_kernel[dom] * 1e-3
    achieved = fl / dur / 1e12
    peak = fl / t_bound(fl, ex) / 1e12
    sfl, sex = work["step"]
    ms_per_step = total_ms / args.steps
    step_frac = (t_bound(sfl, sex) if args.dtype == "f64" else step_bound) / (ms_per_step * 1e-3)
    # The Kdiag backward consumes row sums its forward saved, so the library EXECUTES one sweep of the P x P Gram
    # per step, not the two that SURVEY.md 8d's algorithmic count assumes: the same fraction on the executed work
    # (the forward's 2D+2 flops + 1 exp per Kdiag pair taken out) is reported next to it.
    kfl, kex = work["kdiag_fwd"]
    step_frac_executed = ((t_bound(sfl - kfl, sex - kex) if args.dtype == "f64" else step_bound - t_bound(kfl, kex))
                          / (ms_per_step * 1e-3)) if hp.kd_saved is not None else step_frac
    hbm_peak = None
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    traffic = None
    try:  # DRAM bytes per launch of the dominant kernel, from the committed ncu --set full capture
        for name in ("r1z_traffic.json", "r1_traffic.json"):  # latest capture first
            path = os.path.join(ROOT, "profiles", name)
            if os.path.exists(path):
                traffic = json.load(open(path)).get(args.dtype, {}).get(dom)
                if traffic is not None:
                    break
    except Exception:
        pass
    images = N_BATCH * world
    value = images * args.steps / (total_ms * 1e-3)
    e2e_value = images * e2e_steps / e2e_s

    cpu_val, cores, _ = cpu_port_images_per_s(args.dtype, 40, 3)

    line = {
        "metric": "conv/wconv Kuf+Kdiag fwd+bwd images/sec (28x28, M=750)",
        "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": WORKLOAD, "images_per_gpu_per_step": N_BATCH, "M": M_IND,
                   "l2": "256 MB buffer written between timed steps (inputs are ~5 MB, far below the 126 MB L2)",
                   "cuda_graph": not args.no_graph, "graph_branches": "Kuf | Kdiag | Kuu of each phase run on parallel branches",
                   "collective": ("NCCL all-reduce of the packed gradient buffer per step"
                                  + (", captured in the step's CUDA graph" if collective_in_graph else ""))
                   if world > 1 else "none"},
        "roofline": {
            "bound": pipe, "kernel": dom, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
            "frac": achieved / peak, "traffic": traffic, "peak_source": peak_note,
            "kernel_ms": per_kernel, "step_frac": step_frac, "step_frac_executed_work": step_frac_executed,
            "probes": peaks,
            "hbm": {"algorithmic_bytes_per_step": algorithmic_bytes(N_BATCH, es),
                    "achieved_gbs": algorithmic_bytes(N_BATCH, es) / (ms_per_step * 1e-3) / 1e9,
                    "peak_gbs_measured": hbm_peak},
        },
        "cpu_baseline": {"value": cpu_val, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": "40 images of the 200-image minibatch, M=750, fwd+bwd, median of 3"},
        "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": N_BATCH * H * W_IMG * es,
                "d2h_bytes_per_step": n_out * es, "steps": e2e_steps, "mode": e2e_mode,
                "eager_value": images * e2e_steps / e2e_eager_s,
                "api": "kernel classes (Kzx / Kdiag / Kzz) + torch autograd; X from pinned host memory, packed "
                       "gradients + loss back to host memory, every step",
                "c_abi": {"value": images * e2e_steps / e2e_cabi_s, "unit": "images/s",
                          "h2d_bytes_per_step": N_BATCH * H * W_IMG * es,
                          "d2h_bytes_per_step": int(out_c.numel()) * es,
                          "api": "HotPath.step() over the C ABI (pre-allocated buffers, three-branch CUDA graph), "
                                 "same host-to-device and device-to-host copies every step"}},
        "gpu_launches": hp.kernels_per_step * args.steps,
        "clocks": clocks,
    }
    emit(line)
    if world > 1:
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
    ap.add_argument("--dtype", choices=["f64", "f32"], default="f64")
    ap.add_argument("--impl", choices=["native", "reference"], default="native")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--graph-collective", action="store_true",
                    help="experimental: capture the NCCL all-reduce in the step's CUDA graph instead of launching it "
                         "after the graph (worked at 2 GPUs in float64, hung once in float32: off by default)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
