"""Runs one diagonal log-pdf kernel (csrc/logpdf.cu) a few times at 2^25 elements (for ncu): bernoulli | uniform | cauchy."""
import os, sys
import numpy as np, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import fastad_b200 as F
F.load(); L = F.lib(); F.check(L.fadb_init(0)); F.use_torch_stream()
which = sys.argv[1] if len(sys.argv) > 1 else "bernoulli"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
n = 1 << 25
g = torch.Generator(device="cuda").manual_seed(1)
u = lambda lo, hi: torch.rand(n, dtype=torch.float64, device="cuda", generator=g) * (hi - lo) + lo
if which == "bernoulli":
    x = (u(0, 1) < 0.5).double(); b = u(0.05, 0.95); c = None
    adj = dict(b_adj=torch.zeros_like(b))
elif which == "uniform":
    b = u(-2, -1); c = u(1, 2); x = u(-1, 1)
    adj = dict(b_adj=torch.zeros_like(b), c_adj=torch.zeros_like(c))
else:
    x = u(-1, 1); b = u(-1, 1); c = u(0.5, 2)
    adj = dict(x_adj=torch.zeros_like(x), b_adj=torch.zeros_like(b), c_adj=torch.zeros_like(c))
step = lambda: F.logpdf(which, x, b, c, seed=1.0, **adj)
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps): step()
e1.record(); torch.cuda.synchronize()
print(which, "ms/step", e0.elapsed_time(e1) / reps)
