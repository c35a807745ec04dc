"""Development aid: kernels launched by one end-to-end step through the public kernel classes (what bench.py's e2e
graph replays), from the torch profiler.   python tools/e2e_profile.py [f64|f32]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from convgp_b200 import convkernels as ck, kernels as bk  # noqa: E402
from convgp_b200.settings import settings  # noqa: E402

dtn = sys.argv[1] if len(sys.argv) > 1 else "f64"
dt = torch.float64 if dtn == "f64" else torch.float32
dev = torch.device("cuda", 0)
cfg = bench.HEAD
inp = bench.make_inputs(cfg, 0)
tens = {k: (None if v is None else torch.tensor(v, dtype=dt, device=dev)) for k, v in inp.items()}
settings.float_type = dt
settings.device = dev
kern = ck.WeightedConv(bk.RBF(cfg.D), [cfg.H, cfg.W], [cfg.ph, cfg.pw]) + bk.White(1, 1e-3)
conv = kern.kern_list[0]
Zp = tens["Z"].clone().requires_grad_(True)
free = [conv.W.free, conv.basekern.variance.free, conv.basekern.lengthscales.free, kern.white.variance.free, Zp]


def body():
    for t in free:
        t.grad = None
    kd, kuf, kuu = bk.svgp_terms(kern, Zp, tens["X"], jitter=0.0)
    loss = (kuf * tens["G"]).sum() + (kd * tens["gd"]).sum() + (kuu * tens["Gu"]).sum()
    loss.backward()
    return torch.cat([t.grad.reshape(-1) for t in free] + [loss.detach().reshape(1)])


for _ in range(3):
    body()
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        body()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=90))
