"""One-pass dot -> map -> sum route (fast=6) at 2^20 x 64 for maps of growing cost, against the same plan with the route off:
least squares through a cheap map, logistic log-likelihood with intercept (exp + log per row), a heavy map."""
import os, sys, ctypes as C
import numpy as np, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import fastad_b200 as F, synth
F.load(); L = F.lib(); F.check(L.fadb_init(0)); F.use_torch_stream()
rows, cols = 1 << 20, 64
X, y, wt = synth.linreg_inputs(rows, cols)
yb = (y > 0).astype(np.float64)


def build(kind):
    e = F.Expr()
    Xc = e.const(X)
    w, b = e.var((0.01 * wt).reshape(cols, 1)), e.var(0.1)
    eta = e.binary("add", e.dot(Xc, w), b)
    if kind == "logistic":
        body = e.binary("sub", e.binary("mul", e.const(yb), eta), e.unary("log", e.binary("add", e.const(1.0), e.unary("exp", eta))))
    elif kind == "heavy":
        t = e.unary("tanh", eta)
        body = e.binary("mul", e.unary("erf", t), e.binary("add", e.unary("sin", eta), e.unary("log", e.binary("add", e.const(2.0), e.unary("cos", eta)))))
    else:
        r = e.binary("sub", e.const(y), eta)
        body = e.binary("mul", r, r)
    e.bind(e.sum(body))
    return e


e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for kind in ("square", "logistic", "heavy"):
    for on in (1, 0):
        prev = L.fadb_glm_enable(on)
        e = build(kind)
        for _ in range(3): F.check(L.fadb_expr_autodiff(e.h, 1.0, None, None))
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20): F.check(L.fadb_expr_autodiff(e.h, 1.0, None, None))
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print("%-9s one-pass route %s: %.4f ms  (%.2f of one pass over X)" % (kind, "on " if on else "off", ms, 8 * rows * (cols + 1) / (ms * 1e-3) / 6550.1e9), flush=True)
        L.fadb_glm_enable(prev)
        del e
