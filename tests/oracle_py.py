"""ctypes front-end to oracle/ (TEST INFRASTRUCTURE — never imported by the product).

`OracleExpr` exposes the node-builder vocabulary shared with `fastad_b200.Expr` so a parity
test can build the same expression on both and compare.
"""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCL, VEC, MAT = 0, 1, 2
UNARY = dict(neg=0, sin=1, cos=2, tan=3, asin=4, acos=5, atan=6, exp=7, log=8, sqrt=9, erf=10,
             sigmoid=11, sinh=12, cosh=13, tanh=14, mock2x=100)
BINARY = dict(add=0, sub=1, mul=2, div=3, lt=4, le=5, gt=6, ge=7, eq=8, ne=9, land=10, lor=11, mockb=100)

DIST = dict(cauchy=0, uniform=1, bernoulli=2)
DET_POLICY = dict(fullpivlu=0, ldlt=1, llt=2)
_dp = C.POINTER(C.c_double)
_lib = None


def lib(fast=False):
    global _lib
    name = "libfastad_oracle_fast.so" if fast else "libfastad_oracle.so"
    if not fast and _lib is not None:
        return _lib
    L = C.CDLL(os.path.join(ROOT, "oracle", "_build", name))
    vp, sz, i, d = C.c_void_p, C.c_size_t, C.c_int, C.c_double
    sig = {
        "orc_new": (vp, []), "orc_free": (None, [vp]),
        "orc_var": (i, [vp, i, sz, sz, vp, vp]),
        "orc_const_scalar": (i, [vp, d]),
        "orc_const_view": (i, [vp, i, sz, sz, vp]),
        "orc_rebind": (None, [vp, i, vp, vp, sz, sz]),
        "orc_unary": (i, [vp, i, i]), "orc_binary": (i, [vp, i, i, i]),
        "orc_pow": (i, [vp, C.c_long, i]), "orc_sum": (i, [vp, i]),
        "orc_sum_iter": (i, [vp, i, C.POINTER(i)]), "orc_prod": (i, [vp, i]),
        "orc_prod_iter": (i, [vp, i, C.POINTER(i)]), "orc_dot": (i, [vp, i, i]),
        "orc_normal": (i, [vp, i, i, i]), "orc_eq": (i, [vp, i, i]), "orc_glue": (i, [vp, i, i]),
        "orc_norm": (i, [vp, i]), "orc_transpose": (i, [vp, i]), "orc_if_else": (i, [vp, i, i, i]),
        "orc_opeq": (i, [vp, i, i, i]), "orc_logpdf": (i, [vp, i, i, i, i]),
        "orc_det": (i, [vp, i, i, i]), "orc_wishart": (i, [vp, i, i, d]),
        "orc_bind": (None, [vp, i, C.POINTER(sz)]),
        "orc_feval": (_dp, [vp, i]), "orc_beval": (None, [vp, i, d, vp]),
        "orc_size": (sz, [vp, i]), "orc_is_const": (i, [vp, i]),
        "orc_black_scholes": (None, [sz, vp, vp, vp, vp, vp, C.POINTER(vp), i]),
        "orc_sum_unary": (d, [i, sz, vp, vp, d, i]),
        "orc_normal_lpdf": (d, [i, i, sz, vp, vp, vp, vp, vp, vp, d, i]),
        "orc_linreg": (d, [sz, sz, vp, vp, vp, vp, vp, vp, d, i]),
        "orc_mlp3": (d, [sz, vp, vp, vp, vp, i]),
    }
    for k, (r, a) in sig.items():
        f = getattr(L, k)
        f.restype, f.argtypes = r, a
    if not fast:
        _lib = L
    return L


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


class OracleExpr:
    """Run-time expression on the CPU oracle.  Leaves keep numpy arrays alive."""

    def __init__(self):
        self.L = lib()
        self.g = self.L.orc_new()
        self.keep = []
        self.vars = {}

    def __del__(self):
        try:
            self.L.orc_free(self.g)
        except Exception:
            pass

    # ---- leaves ----
    def var(self, value, shape=None, name=None):
        v = np.array(value, dtype=np.float64, order="F")
        if shape is None:
            shape = (SCL, VEC, MAT)[v.ndim]
        rows = v.shape[0] if v.ndim >= 1 else 1
        cols = v.shape[1] if v.ndim == 2 else 1
        v = np.asfortranarray(v).reshape(-1, order="F").copy()
        a = np.zeros_like(v)
        self.keep += [v, a]
        nid = self.L.orc_var(self.g, shape, rows, cols, _p(v), _p(a))
        self.vars[nid] = (v, a, rows, cols)
        return nid

    def const(self, value):
        v = np.array(value, dtype=np.float64, order="F")
        if v.ndim == 0:
            return self.L.orc_const_scalar(self.g, float(v))
        shape = (SCL, VEC, MAT)[v.ndim]
        rows = v.shape[0]
        cols = v.shape[1] if v.ndim == 2 else 1
        v = np.asfortranarray(v).reshape(-1, order="F").copy()
        self.keep.append(v)
        return self.L.orc_const_view(self.g, shape, rows, cols, _p(v))

    # ---- nodes ----
    def unary(self, op, c):
        return self.L.orc_unary(self.g, UNARY[op], c)

    def binary(self, op, l, r):
        return self.L.orc_binary(self.g, BINARY[op], l, r)

    def pow(self, n, c):
        return self.L.orc_pow(self.g, n, c)

    def sum(self, c):
        return self.L.orc_sum(self.g, c)

    def prod(self, c):
        return self.L.orc_prod(self.g, c)

    def sum_iter(self, cs):
        arr = (C.c_int * len(cs))(*cs)
        return self.L.orc_sum_iter(self.g, len(cs), arr)

    def prod_iter(self, cs):
        arr = (C.c_int * len(cs))(*cs)
        return self.L.orc_prod_iter(self.g, len(cs), arr)

    def dot(self, l, r):
        return self.L.orc_dot(self.g, l, r)

    def normal(self, x, mu, sigma):
        return self.L.orc_normal(self.g, x, mu, sigma)

    def eq(self, w, c):
        return self.L.orc_eq(self.g, w, c)

    def glue(self, *ids):
        out = ids[0]
        for nxt in ids[1:]:
            out = self.L.orc_glue(self.g, out, nxt)
        return out

    def norm(self, c):
        return self.L.orc_norm(self.g, c)

    def transpose(self, c):
        return self.L.orc_transpose(self.g, c)

    def if_else(self, c, a, b):
        return self.L.orc_if_else(self.g, c, a, b)

    def opeq(self, op, w, c):
        return self.L.orc_opeq(self.g, BINARY[op], w, c)

    def logpdf(self, dist, a, b, c=-1):
        return self.L.orc_logpdf(self.g, DIST[dist], a, b, c)

    def det(self, c, policy="fullpivlu", log=False):
        return self.L.orc_det(self.g, DET_POLICY[policy], 1 if log else 0, c)

    def wishart(self, x, v, df):
        return self.L.orc_wishart(self.g, x, v, float(df))

    # ---- driver ----
    def bind(self, root):
        out = (C.c_size_t * 2)()
        self.L.orc_bind(self.g, root, out)
        self.root = root
        return int(out[0]), int(out[1])

    def feval(self, root=None):
        root = self.root if root is None else root
        n = self.L.orc_size(self.g, root)
        p = self.L.orc_feval(self.g, root)
        return np.ctypeslib.as_array(p, shape=(n,)).copy()

    def beval(self, seed=1.0, root=None):
        root = self.root if root is None else root
        if np.ndim(seed) == 0:
            self.L.orc_beval(self.g, root, float(seed), None)
        else:
            s = f64(np.asfortranarray(seed).reshape(-1, order="F"))
            self.L.orc_beval(self.g, root, 0.0, _p(s))

    def autodiff(self, seed=1.0, root=None):
        v = self.feval(root)
        self.beval(seed, root)
        return v

    def value(self, nid):
        return self.vars[nid][0]

    def adj(self, nid):
        return self.vars[nid][1]

    def reset_adj(self):
        for v, a, _, _ in self.vars.values():
            a[:] = 0


# ---- whole-workload entry points (bench cpu_baseline + parity) ----

def black_scholes(S, K, sigma, tau, r, nthreads=1, fast=False):
    L = lib(fast)
    n = len(S)
    outs = [np.zeros(n) for _ in range(8)]
    arr = (C.c_void_p * 8)(*[o.ctypes.data for o in outs])
    L.orc_black_scholes(n, _p(f64(S)), _p(f64(K)), _p(f64(sigma)), _p(f64(tau)), _p(f64(r)), arr,
                        nthreads)
    return outs


def sum_unary(op, x, x_adj, seed=1.0, nthreads=1, fast=False):
    code = UNARY[op] if isinstance(op, str) else 1000 + op[1]
    return lib(fast).orc_sum_unary(code, len(x), _p(x), _p(x_adj), seed, nthreads)


def normal_lpdf(x, mu, sigma, x_adj=None, mu_adj=None, sigma_adj=None, seed=1.0, nthreads=1,
                fast=False):
    shape = (1 if np.ndim(mu) and len(mu) > 1 else 0) | (2 if np.ndim(sigma) and len(sigma) > 1 else 0)
    mask = (1 if x_adj is None else 0) | (2 if mu_adj is None else 0) | (4 if sigma_adj is None else 0)
    mu = np.atleast_1d(f64(mu))
    sigma = np.atleast_1d(f64(sigma))
    return lib(fast).orc_normal_lpdf(shape, mask, len(x), _p(x), _p(mu), _p(sigma), _p(x_adj),
                                     _p(mu_adj), _p(sigma_adj), seed, nthreads)


def linreg(X, y, w, w_adj, sigma, sigma_adj, seed=1.0, nthreads=1, fast=False):
    rows, cols = X.shape
    assert X.flags.f_contiguous
    return lib(fast).orc_linreg(rows, cols, _p(X), _p(y), _p(w), _p(w_adj), _p(sigma), _p(sigma_adj),
                                seed, nthreads)


def mlp3(X, y, parm, grad, nthreads=1, fast=False):
    assert X.flags.f_contiguous
    return lib(fast).orc_mlp3(X.shape[0], _p(X), _p(y), _p(parm), _p(grad), nthreads)
