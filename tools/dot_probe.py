"""Least squares through the generic path: sum(pow<2>(y - dot(X, w))), X rows x cols constant (README.md:606-662
batched).  Prints ms per autodiff and the describe() of the plan."""
import os, sys
import numpy as np, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import fastad_b200 as F, synth
F.load(); L = F.lib(); F.check(L.fadb_init(0)); F.use_torch_stream()
rows = 1 << int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
cols = int(sys.argv[2]) if len(sys.argv) > 2 else 64
reps = 20
Xh, yh, _ = synth.linreg_inputs(rows, cols)
e = F.Expr()
X = e.const(Xh); y = e.const(yh); w = e.var(0.01 * np.arange(cols, dtype=np.float64).reshape(cols, 1))
inner = len(sys.argv) > 3
res = e.binary("sub", y, e.dot(X, w))
if inner:
    res = e.binary("mul", res, e.const(1.0 + 0 * yh))      # defeats the least-squares pattern: generic GEMV kernels
root = e.sum(e.pow(2, res))
e.bind(root)
print(e.describe().split("\n")[0], [l for l in e.describe().split("\n") if l.startswith("stage")])
step = lambda: F.check(L.fadb_expr_autodiff(e.h, 1.0, None, None))
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps): step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print("rows", rows, "cols", cols, "ms/step %.4f" % ms, "X passes at roofline (6550 GB/s): %.2f" % (ms * 1e-3 * 6550e9 / (rows * cols * 8)))
v = e.autodiff(1.0)
wh = 0.01 * np.arange(cols)
r = yh - Xh @ wh
print("value rel err", abs(v[0] - r @ r) / (r @ r), "adj rel err", np.max(np.abs(e.adj(w).ravel() / (reps + 4) - (-2 * Xh.T @ r))) / np.max(np.abs(2 * Xh.T @ r)))

# ---- wide case: a 3x2 weight matrix times a 2 x 2^20 batch (the batched layer of the README network)
B = 1 << 20
rng = np.random.default_rng(5)
e = F.Expr()
W = e.var(rng.normal(size=(3, 2))); Xb = e.const(rng.uniform(-1, 1, size=(2, B))); T = e.const(rng.uniform(0, 1, size=(3, B)))
e.bind(e.sum(e.pow(2, e.binary("sub", e.unary("sigmoid", e.dot(W, Xb)), T))))
print([l for l in e.describe().split("\n") if l.startswith("stage")])
for _ in range(3): step2 = F.check(L.fadb_expr_autodiff(e.h, 1.0, None, None))
torch.cuda.synchronize()
e0.record()
for _ in range(reps): F.check(L.fadb_expr_autodiff(e.h, 1.0, None, None))
e1.record(); torch.cuda.synchronize()
print("wide 3x2 * 2x2^20: ms/step %.4f" % (e0.elapsed_time(e1) / reps))

# ---- matrix x matrix: 1024 x 1024 x 1024, both variable
rng = np.random.default_rng(6)
e = F.Expr()
A = e.var(rng.normal(size=(1024, 1024))); Bm = e.var(rng.normal(size=(1024, 1024)))
e.bind(e.sum(e.pow(2, e.dot(A, Bm))))
for _ in range(3): F.check(L.fadb_expr_autodiff(e.h, 1.0, None, None))
torch.cuda.synchronize()
e0.record()
for _ in range(reps): F.check(L.fadb_expr_autodiff(e.h, 1.0, None, None))
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print("matmul 1024^3 fwd+bwd: ms/step %.4f  (%.1f TFLOP/s FP64 on 3 x 2 n^3 flop)" % (ms, 6 * 1024**3 / (ms * 1e-3) / 1e12))
