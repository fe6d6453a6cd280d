"""Runs the two generic-path bench expressions a few times (for ncu): the 3-deep sum chain at 2^24 and
the Black-Scholes call expression at 2^20, on the specialised stage kernels."""
import os, sys
import numpy as np, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import fastad_b200 as F, synth, exprs
F.load(); L = F.lib(); F.check(L.fadb_init(0)); F.use_torch_stream()
which = sys.argv[1] if len(sys.argv) > 1 else "chain"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
if which == "chain":
    n2 = 1 << 24
    e = F.Expr()
    x = e.var(synth.sum_inputs(n2)); w0, w1 = e.var(2.0), e.var(1.3)
    cc = e.const(synth.sum_inputs(n2, seed=5))
    z = e.binary("div", e.binary("sub", x, e.binary("mul", w0, cc)), w1)
    e.bind(e.sum(e.binary("mul", e.const(-0.5), e.pow(2, z))))
    step = lambda: F.check(L.fadb_expr_autodiff(e.h, 1.0, None, None))
else:
    nb = 1 << 20
    exprs.RNG = np.random.default_rng(1)
    e = F.Expr()
    root, leaves, _ = exprs.black_scholes_vec(nb)(e)
    e.bind(root)
    ones = torch.ones(nb, dtype=torch.float64, device="cuda")
    step = lambda: F.check(L.fadb_expr_autodiff(e.h, 0.0, F._ptr(ones), None))
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps): step()
e1.record(); torch.cuda.synchronize()
print(which, "ms/step", e0.elapsed_time(e1) / reps, F.jit_stats())
if F.jit_stats()["failed"]:
    print("JIT diagnostics:", L.fadb_jit_diagnostics().decode()[:3000])
