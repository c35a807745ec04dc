"""Time one full SVGP training step (ELBO + gradient) at config 3 and break it down (development aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from convgp_b200 import convkernels as ck, kernels as bk, likelihoods as lik, svgp
from convgp_b200.settings import settings

dt = torch.float64 if (len(sys.argv) < 2 or sys.argv[1] == "f64") else torch.float32
settings.float_type = dt
inp = bench.make_inputs(0)
N, M, L = 200, 750, 10
X = inp["X"]; Y = np.random.default_rng(7).integers(0, L, size=(N, 1)).astype(np.float64)
k = ck.WeightedConv(bk.RBF(25), [28, 28], [5, 5]) + bk.White(1, 1e-3)
m = svgp.SVGP(X, Y, k, lik.MultiClass(L), inp["Z"], num_latent=L)
params = [p.free for p in m.free_params()]

def step():
    for p in params: p.grad = None
    f = -m.build_likelihood()
    f.backward()
    return f

for _ in range(3): step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): step()
b.record(); torch.cuda.synchronize()
print("full SVGP step ms:", a.elapsed_time(b) / 10)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=60))

# ---- the same step replayed from a CUDA graph
m2 = svgp.SVGP(X, Y, k, lik.MultiClass(L), inp["Z"], num_latent=L)
gstep = m2.graphed_objective()
for _ in range(3): gstep()
torch.cuda.synchronize()
a.record()
for _ in range(20): gstep()
b.record(); torch.cuda.synchronize()
print("graphed SVGP step ms:", a.elapsed_time(b) / 20)
