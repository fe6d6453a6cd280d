"""fastad_b200 — ctypes front-end to libfastad_b200.so (the C ABI in include/fastad_b200.h).

The product is the CUDA library; this module only binds it for the Python test and bench
harness.  PyTorch supplies device memory, streams and torch.distributed.  There is NO CPU
fallback: if the shared library is missing, or there is no CUDA device, calls fail loudly.
"""
import ctypes as C
import os

__all__ = ["lib", "load", "FadbError", "OVERWRITE_ADJ", "TRUST_SIGMA", "STRICT_SIGMA", "UNARY", "BINARY",
           "black_scholes", "sum_unary", "prod", "normal_lpdf", "linreg_nll", "mlp3"]

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libfastad_b200.so")

OVERWRITE_ADJ = 1
TRUST_SIGMA = 2
STRICT_SIGMA = 4
SCL, VEC, MAT = 0, 1, 2
UNARY = dict(neg=0, sin=1, cos=2, tan=3, asin=4, acos=5, atan=6, exp=7, log=8, sqrt=9, erf=10,
             sigmoid=11, sinh=12, cosh=13, tanh=14, mock2x=100)
BINARY = dict(add=0, sub=1, mul=2, div=3, lt=4, le=5, gt=6, ge=7, eq=8, ne=9, land=10, lor=11, mockb=100)


class FadbError(RuntimeError):
    pass


_vp, _sz, _i, _d, _u = C.c_void_p, C.c_size_t, C.c_int, C.c_double, C.c_uint
_dp = C.POINTER(C.c_double)

# name -> (restype, argtypes); mirrors include/fastad_b200.h one to one
SIGNATURES = {
    "fadb_init": (_i, [_i]),
    "fadb_shutdown": (None, []),
    "fadb_device_count": (_i, []),
    "fadb_get_device": (_i, []),
    "fadb_set_stream": (_i, [_vp]),
    "fadb_get_stream": (_i, [C.POINTER(_vp)]),
    "fadb_sync": (_i, []),
    "fadb_last_error": (C.c_char_p, []),
    "fadb_launch_count": (C.c_ulonglong, []),
    "fadb_version": (C.c_char_p, []),
    "fadb_malloc": (_i, [_sz, C.POINTER(_vp)]),
    "fadb_free": (_i, [_vp]),
    "fadb_h2d": (_i, [_vp, _vp, _sz]),
    "fadb_d2h": (_i, [_vp, _vp, _sz]),
    "fadb_memset_zero": (_i, [_vp, _sz]),
    "fadb_host_alloc_pinned": (_i, [_sz, C.POINTER(_vp)]),
    "fadb_host_free_pinned": (_i, [_vp]),
    "fadb_host_register": (_i, [_vp, _sz]),
    "fadb_host_unregister": (_i, [_vp]),
    "fadb_black_scholes": (_i, [_sz, _vp, _vp, _vp, _vp, _vp, C.POINTER(_vp), _u]),
    "fadb_black_scholes_host": (_i, [_sz, _vp, _vp, _vp, _vp, _vp, C.POINTER(_vp)]),
    "fadb_black_scholes_host_mode": (_i, [_dp]),
    "fadb_black_scholes_batches": (_i, [_sz, _sz, _sz, _sz, C.POINTER(_vp), C.POINTER(_vp), _u]),
    "fadb_sum_unary": (_i, [_i, _sz, _vp, _vp, _d, _u, _dp]),
    "fadb_sum_unary_async": (_i, [_i, _sz, _vp, _vp, _d, _u, _vp]),
    "fadb_prod": (_i, [_sz, _vp, _vp, _d, _u, _dp]),
    "fadb_normal_lpdf": (_i, [_i, _sz, _vp, _vp, _vp, _vp, _vp, _vp, _d, _u, _dp]),
    "fadb_sigma_check": (_i, [_sz, _vp, _vp]),
    "fadb_normal_lpdf_partial": (_i, [_i, _sz, _vp, _vp, _vp, _vp, _vp, _vp, _d, _u, _vp, _vp]),
    "fadb_normal_lpdf_finish": (_i, [_i, _sz, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _d, _u, _dp]),
    "fadb_glm_enable": (_i, [_i]),
    "fadb_additive_root_enable": (_i, [_i]),
    "fadb_jit_ring_enable": (_i, [_i]),
    "fadb_normal_lpdf_undo": (_i, [_i, _sz, _vp, _vp, _vp, _vp, _vp, _vp, _d, _u, _vp]),
    "fadb_normal_dense": (_i, [_sz, _vp, _vp, _i, _vp, _vp, _vp, _vp, _d, _u, _dp]),
    "fadb_normal_dense_async": (_i, [_sz, _vp, _vp, _i, _vp, _vp, _vp, _vp, _d, _u, _i, _vp]),
    "fadb_det": (_i, [_sz, _vp, _vp, _d, _u, _i, _i, _dp]),
    "fadb_det_async": (_i, [_sz, _vp, _vp, _d, _u, _i, _i, _i, _vp]),
    "fadb_wishart": (_i, [_sz, _vp, _vp, _d, _vp, _vp, _d, _u, _dp]),
    "fadb_wishart_async": (_i, [_sz, _vp, _vp, _d, _vp, _vp, _d, _u, _i, _vp]),
    "fadb_logpdf": (_i, [_i, _i, _sz, _vp, _vp, _vp, _vp, _vp, _vp, _d, _u, _dp]),
    "fadb_logpdf_async": (_i, [_i, _i, _sz, _vp, _vp, _vp, _vp, _vp, _vp, _d, _u, _vp]),
    "fadb_linreg_partial": (_i, [_sz, _sz, _vp, _vp, _vp, _vp]),
    "fadb_linreg_finish": (_i, [_sz, _sz, _vp, _vp, _vp, _vp, _d, _u, _dp]),
    "fadb_linreg_nll": (_i, [_sz, _sz, _vp, _vp, _vp, _vp, _vp, _vp, _d, _u, _dp]),
    "fadb_lsq": (_i, [_sz, _sz, _vp, _vp, _vp, _vp, _d, _u, _dp]),
    "fadb_lsq_finish": (_i, [_sz, _vp, _vp, _d, _u, _vp]),
    "fadb_mlp3": (_i, [_sz, _vp, _vp, _vp, _vp, _u, _dp]),
    "fadb_mlp3_async": (_i, [_sz, _vp, _vp, _vp, _vp, _u, _vp]),
    "fadb_expr_create": (_i, [C.POINTER(_vp)]),
    "fadb_expr_destroy": (None, [_vp]),
    "fadb_expr_var": (_i, [_vp, _i, _sz, _sz, _vp, _vp]),
    "fadb_expr_const_scalar": (_i, [_vp, _d]),
    "fadb_expr_const_view": (_i, [_vp, _i, _sz, _sz, _vp]),
    "fadb_expr_rebind": (_i, [_vp, _i, _vp, _vp]),
    "fadb_expr_unary": (_i, [_vp, _i, _i]),
    "fadb_expr_binary": (_i, [_vp, _i, _i, _i]),
    "fadb_expr_pow": (_i, [_vp, C.c_long, _i]),
    "fadb_expr_sum": (_i, [_vp, _i]),
    "fadb_expr_sum_iter": (_i, [_vp, _i, C.POINTER(_i)]),
    "fadb_expr_prod": (_i, [_vp, _i]),
    "fadb_expr_prod_iter": (_i, [_vp, _i, C.POINTER(_i)]),
    "fadb_expr_dot": (_i, [_vp, _i, _i]),
    "fadb_expr_normal": (_i, [_vp, _i, _i, _i]),
    "fadb_expr_eq": (_i, [_vp, _i, _i]),
    "fadb_expr_glue": (_i, [_vp, _i, _i]),
    "fadb_expr_logpdf": (_i, [_vp, _i, _i, _i, _i]),
    "fadb_expr_norm": (_i, [_vp, _i]),
    "fadb_expr_det": (_i, [_vp, _i, _i, _i]),
    "fadb_expr_wishart": (_i, [_vp, _i, _i, _d]),
    "fadb_expr_transpose": (_i, [_vp, _i]),
    "fadb_expr_if_else": (_i, [_vp, _i, _i, _i]),
    "fadb_expr_opeq": (_i, [_vp, _i, _i, _i]),
    "fadb_expr_node_size": (_i, [_vp, _i, C.POINTER(_sz), C.POINTER(_sz)]),
    "fadb_expr_is_const": (_i, [_vp, _i]),
    "fadb_expr_plan": (_i, [_vp, _i, C.POINTER(_sz)]),
    "fadb_expr_describe": (_i, [_vp, C.c_char_p, _sz]),
    "fadb_expr_bind": (_i, [_vp, _i, _u, C.POINTER(_sz)]),
    "fadb_expr_feval": (_i, [_vp, _dp]),
    "fadb_expr_beval": (_i, [_vp, _d, _vp]),
    "fadb_expr_autodiff": (_i, [_vp, _d, _vp, _dp]),
    "fadb_expr_value_ptr": (_i, [_vp, C.POINTER(_vp), C.POINTER(_sz)]),
    "fadb_p2p_create": (_i, [_i, _i, _vp]),
    "fadb_p2p_connect": (_i, [_vp]),
    "fadb_p2p_create_local": (_i, [_i, C.POINTER(_i)]),
    "fadb_p2p_allreduce": (_i, [_vp, _i]),
    "fadb_p2p_status": (_i, [C.POINTER(C.c_ulonglong)]),
    "fadb_p2p_destroy": (_i, []),
    "fadb_normal_lpdf_partial_push": (_i, [_i, _sz, _vp, _vp, _vp, _vp, _vp, _vp, _d, _u, _vp, _vp]),
    "fadb_normal_lpdf_finish_pull": (_i, [_i, _sz, _vp, _vp, _vp, _vp, _vp, _vp, _d, _u, _dp]),
    "fadb_linreg_partial_push": (_i, [_sz, _sz, _vp, _vp, _vp, _vp]),
    "fadb_linreg_finish_pull": (_i, [_sz, _sz, _vp, _vp, _vp, _d, _u, _dp]),
    "fadb_sum_unary_allreduce_async": (_i, [_i, _sz, _vp, _vp, _d, _u, _vp]),
    "fadb_mlp3_allreduce_async": (_i, [_sz, _vp, _vp, _vp, _vp, _u, _vp]),
    "fadb_jit_enable": (_i, [_i]),
    "fadb_jit_stats": (_i, [C.POINTER(C.c_ulonglong)]),
    "fadb_jit_diagnostics": (C.c_char_p, []),
    "fadb_expr_jit_program": (_i, [_vp, _i, _i, C.c_char_p, _sz]),
    "fadb_jit_compile_only": (_i, [C.c_char_p, C.POINTER(_sz), C.c_char_p, _sz]),
}

_lib = None


def load(path=None):
    """dlopen the shared library and attach signatures.  Raises if it is not built."""
    global _lib
    path = path or os.environ.get("FASTAD_B200_LIB") or LIB_PATH      # the override serves A/B runs of two builds
    if not os.path.exists(path):
        raise FadbError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  fastad_b200 has no CPU fallback.")
    L = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)          # AttributeError if the symbol is not exported
        fn.restype, fn.argtypes = res, args
    _lib = L
    return L


def lib():
    return _lib if _lib is not None else load()


def check(rc):
    if rc != 0:
        raise FadbError(f"fastad_b200 error {rc}: {lib().fadb_last_error().decode()}")


def jit_enable(on):
    """Specialised (NVRTC) stage kernels on/off; returns the previous setting."""
    return bool(lib().fadb_jit_enable(1 if on else 0))


def jit_stats():
    out = (C.c_ulonglong * 4)()
    check(lib().fadb_jit_stats(out))
    return {"compiled": out[0], "hits": out[1], "failed": out[2], "launched": out[3]}


def jit_compile_only(program):
    """NVRTC-compiles a stage program text to an sm_100a cubin (no device needed); returns (bytes, log)."""
    n = _sz(0)
    log = C.create_string_buffer(1 << 16)
    check(lib().fadb_jit_compile_only(program.encode(), C.byref(n), log, len(log)))
    return n.value, log.value.decode()


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def use_torch_stream():
    """Route the library's launches to torch's current CUDA stream on the current device."""
    import torch
    h = torch.cuda.current_stream().cuda_stream
    # torch's default stream is the legacy NULL stream: name it explicitly (cudaStreamLegacy == 1),
    # because NULL means "the library's own stream" to fadb_set_stream.
    check(lib().fadb_set_stream(C.c_void_p(h if h else 1)))


def _chk_f64(*ts):
    import torch
    for t in ts:
        if t is None:
            continue
        if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()):
            raise FadbError("fastad_b200 expects contiguous float64 CUDA tensors")


# ------------------------------------------------------------------------------------------
# Thin helpers over torch CUDA tensors (float64, contiguous).  Values come back as floats.
# ------------------------------------------------------------------------------------------

def black_scholes(S, K, sigma, tau, r, out=None, flags=OVERWRITE_ADJ):
    import torch
    _chk_f64(S, K, sigma, tau, r)
    n = S.numel()
    if out is None:
        out = [torch.zeros(n, dtype=torch.float64, device=S.device) for _ in range(8)]
    _chk_f64(*out)
    arr = (C.c_void_p * 8)(*[o.data_ptr() for o in out])
    check(lib().fadb_black_scholes(n, _ptr(S), _ptr(K), _ptr(sigma), _ptr(tau), _ptr(r), arr, flags))
    return out


class BlackScholesBatches:
    """Operand tables for fadb_black_scholes_batches: `sets` = [(inputs[5], outputs[8]), ...] of float64 CUDA tensors of
    one length; run(count, first) launches `count` batches back to back from ONE host call (no Python between launches)."""

    def __init__(self, sets, flags=OVERWRITE_ADJ):
        for ins, outs in sets:
            _chk_f64(*ins, *outs)
        self.keep = sets
        self.n = sets[0][0][0].numel()
        self.nsets = len(sets)
        self.flags = flags
        self.tin = (C.c_void_p * (5 * self.nsets))(*[t.data_ptr() for ins, _ in sets for t in ins])
        self.tout = (C.c_void_p * (8 * self.nsets))(*[t.data_ptr() for _, outs in sets for t in outs])
        self.fn = lib().fadb_black_scholes_batches

    def run(self, count, first=0):
        check(self.fn(count, self.n, self.nsets, first, self.tin, self.tout, self.flags))


def black_scholes_host(S, K, sigma, tau, r, out):
    """numpy (host) arrays in, numpy arrays out — the call a user of the reference makes."""
    n = len(S)
    arr = (C.c_void_p * 8)(*[o.ctypes.data for o in out])
    p = lambda a: C.c_void_p(a.ctypes.data)
    check(lib().fadb_black_scholes_host(n, p(S), p(K), p(sigma), p(tau), p(r), arr))
    return out


def _op_code(op):
    if isinstance(op, str):
        return UNARY[op]
    kind, n = op
    assert kind == "pow"
    return 1000 + int(n)


def sum_unary(op, x, x_adj, seed=1.0, flags=0):
    _chk_f64(x, x_adj)
    v = C.c_double()
    check(lib().fadb_sum_unary(_op_code(op), x.numel(), _ptr(x), _ptr(x_adj), seed, flags, C.byref(v)))
    return v.value


def prod(x, x_adj, seed=1.0, flags=0):
    _chk_f64(x, x_adj)
    v = C.c_double()
    check(lib().fadb_prod(x.numel(), _ptr(x), _ptr(x_adj), seed, flags, C.byref(v)))
    return v.value


def normal_lpdf(x, mu, sigma, x_adj=None, mu_adj=None, sigma_adj=None, seed=1.0, flags=0):
    """mu / sigma: 1-element tensors are scalars (ad::scl), n-element tensors vectors."""
    _chk_f64(x, mu, sigma, x_adj, mu_adj, sigma_adj)
    n = x.numel()
    shape = (1 if (mu.numel() == n and n > 1) else 0) | (2 if (sigma.numel() == n and n > 1) else 0)
    v = C.c_double()
    check(lib().fadb_normal_lpdf(shape, n, _ptr(x), _ptr(mu), _ptr(sigma), _ptr(x_adj), _ptr(mu_adj),
                                 _ptr(sigma_adj), seed, flags, C.byref(v)))
    return v.value


def normal_dense(x, mu, Sigma, x_adj=None, mu_adj=None, Sigma_adj=None, seed=1.0, flags=0):
    """Sigma: (n, n) torch tensor holding the column-major matrix (symmetric: layout-agnostic)."""
    _chk_f64(x, mu, Sigma, x_adj, mu_adj, Sigma_adj)
    v = C.c_double()
    check(lib().fadb_normal_dense(x.numel(), _ptr(x), _ptr(mu), 1 if mu.numel() > 1 else 0, _ptr(Sigma), _ptr(x_adj),
                                  _ptr(mu_adj), _ptr(Sigma_adj), seed, flags, C.byref(v)))
    return v.value


DET_POLICY = dict(fullpivlu=0, ldlt=1, llt=2)


def det(A, A_adj=None, seed=1.0, policy="fullpivlu", log=False, flags=0):
    """det / log_det of the (n, n) column-major matrix A; A_adj += seed * [det] * A^-T."""
    _chk_f64(A, A_adj)
    n = int(round(A.numel() ** 0.5))
    v = C.c_double()
    check(lib().fadb_det(n, _ptr(A), _ptr(A_adj), seed, flags, DET_POLICY[policy], 1 if log else 0, C.byref(v)))
    return v.value


def wishart(X, V, df, X_adj=None, V_adj=None, seed=1.0, flags=0):
    """wishart_adj_log_pdf(X, V, df) for (p, p) column-major X, V."""
    _chk_f64(X, V, X_adj, V_adj)
    p = int(round(X.numel() ** 0.5))
    v = C.c_double()
    check(lib().fadb_wishart(p, _ptr(X), _ptr(V), float(df), _ptr(X_adj), _ptr(V_adj), seed, flags, C.byref(v)))
    return v.value


def logpdf(dist, x, b, c=None, x_adj=None, b_adj=None, c_adj=None, seed=1.0, flags=0):
    """cauchy(x, loc, scale) / uniform(x, min, max) / bernoulli(x, p); 1-element tensors are scalars."""
    _chk_f64(x, b, c, x_adj, b_adj, c_adj)
    n = x.numel()
    shape = (1 if (b.numel() == n and n > 1) else 0) | (2 if (c is not None and c.numel() == n and n > 1) else 0)
    v = C.c_double()
    check(lib().fadb_logpdf(dict(cauchy=0, uniform=1, bernoulli=2)[dist], shape, n, _ptr(x), _ptr(b), _ptr(c),
                            _ptr(x_adj), _ptr(b_adj), _ptr(c_adj), seed, flags, C.byref(v)))
    return v.value


def linreg_nll(X, y, w, w_adj, sigma, sigma_adj, seed=1.0, flags=0):
    """X: (cols, rows) contiguous torch tensor == rows x cols column-major."""
    _chk_f64(X, y, w, w_adj, sigma, sigma_adj)
    cols, rows = X.shape
    v = C.c_double()
    check(lib().fadb_linreg_nll(rows, cols, _ptr(X), _ptr(y), _ptr(w), _ptr(w_adj), _ptr(sigma),
                                _ptr(sigma_adj), seed, flags, C.byref(v)))
    return v.value


def mlp3(X, y, parm, grad, flags=0):
    """X: (2, batch) contiguous torch tensor == batch x 2 column-major."""
    _chk_f64(X, y, parm, grad)
    v = C.c_double()
    check(lib().fadb_mlp3(y.numel(), _ptr(X), _ptr(y), _ptr(parm), _ptr(grad), flags, C.byref(v)))
    return v.value


# ------------------------------------------------------------------------------------------
# Expr: the run-time expression builder of the C ABI (what include/fastad lowers ad:: expression
# templates into at bind() time), with the same vocabulary as tests/oracle_py.OracleExpr.
# device=False builds and plans on the host only (no CUDA call): used by the CPU tests.
# ------------------------------------------------------------------------------------------
class Expr:
    def __init__(self, device=True):
        self.L = lib()
        h = C.c_void_p()
        check(self.L.fadb_expr_create(C.byref(h)))
        self.h = h
        self.device = device
        self.vars = {}
        self.keep = []
        self.root = None

    def __del__(self):
        try:
            self.L.fadb_expr_destroy(self.h)
        except Exception:
            pass

    def _id(self, rc):
        if rc < 0:
            check(rc)
        return rc

    def _buf(self, flat):
        import numpy as np
        if not self.device:
            a = np.ascontiguousarray(flat, dtype=np.float64)
            self.keep.append(a)
            return a, a.ctypes.data
        import torch
        t = torch.from_numpy(np.ascontiguousarray(flat, dtype=np.float64)).cuda()
        self.keep.append(t)
        return t, t.data_ptr()

    @staticmethod
    def _shape_of(v):
        import numpy as np
        v = np.array(v, dtype=np.float64, order="F")
        shape = (SCL, VEC, MAT)[v.ndim]
        rows = v.shape[0] if v.ndim >= 1 else 1
        cols = v.shape[1] if v.ndim == 2 else 1
        return np.asfortranarray(v).reshape(-1, order="F").copy(), shape, rows, cols

    def var(self, value):
        import numpy as np
        flat, shape, rows, cols = self._shape_of(value)
        tv, pv = self._buf(flat)
        ta, pa = self._buf(np.zeros_like(flat))
        nid = self._id(self.L.fadb_expr_var(self.h, shape, rows, cols, C.c_void_p(pv), C.c_void_p(pa)))
        self.vars[nid] = (tv, ta)
        return nid

    def const(self, value):
        import numpy as np
        if np.ndim(value) == 0:
            return self._id(self.L.fadb_expr_const_scalar(self.h, float(value)))
        flat, shape, rows, cols = self._shape_of(value)
        tv, pv = self._buf(flat)
        return self._id(self.L.fadb_expr_const_view(self.h, shape, rows, cols, C.c_void_p(pv)))

    def unary(self, op, c):
        return self._id(self.L.fadb_expr_unary(self.h, UNARY[op], c))

    def binary(self, op, l, r):
        return self._id(self.L.fadb_expr_binary(self.h, BINARY[op], l, r))

    def pow(self, n, c):
        return self._id(self.L.fadb_expr_pow(self.h, n, c))

    def sum(self, c):
        return self._id(self.L.fadb_expr_sum(self.h, c))

    def prod(self, c):
        return self._id(self.L.fadb_expr_prod(self.h, c))

    def sum_iter(self, cs):
        return self._id(self.L.fadb_expr_sum_iter(self.h, len(cs), (C.c_int * len(cs))(*cs)))

    def prod_iter(self, cs):
        return self._id(self.L.fadb_expr_prod_iter(self.h, len(cs), (C.c_int * len(cs))(*cs)))

    def dot(self, l, r):
        return self._id(self.L.fadb_expr_dot(self.h, l, r))

    def normal(self, x, mu, sigma):
        return self._id(self.L.fadb_expr_normal(self.h, x, mu, sigma))

    def eq(self, w, c):
        return self._id(self.L.fadb_expr_eq(self.h, w, c))

    def glue(self, *ids):
        out = ids[0]
        for nxt in ids[1:]:
            out = self._id(self.L.fadb_expr_glue(self.h, out, nxt))
        return out

    def logpdf(self, dist, a, b, c=-1):
        return self._id(self.L.fadb_expr_logpdf(self.h, dict(cauchy=0, uniform=1, bernoulli=2)[dist], a, b, c))

    def norm(self, c):
        return self._id(self.L.fadb_expr_norm(self.h, c))

    def det(self, c, policy="fullpivlu", log=False):
        return self._id(self.L.fadb_expr_det(self.h, DET_POLICY[policy], 1 if log else 0, c))

    def wishart(self, x, v, df):
        return self._id(self.L.fadb_expr_wishart(self.h, x, v, float(df)))

    def transpose(self, c):
        return self._id(self.L.fadb_expr_transpose(self.h, c))

    def if_else(self, c, a, b):
        return self._id(self.L.fadb_expr_if_else(self.h, c, a, b))

    def opeq(self, op, w, c):
        return self._id(self.L.fadb_expr_opeq(self.h, BINARY[op], w, c))

    def plan(self, root):
        out = (C.c_size_t * 2)()
        check(self.L.fadb_expr_plan(self.h, root, out))
        self.root = root
        return int(out[0]), int(out[1])

    def jit_program(self, stage, mode):
        """Constants that specialise planned stage `stage` for mode 1 (forward), 2 (backward), 3 (fused)."""
        buf = C.create_string_buffer(1 << 16)
        check(self.L.fadb_expr_jit_program(self.h, stage, mode, buf, len(buf)))
        return buf.value.decode()

    def describe(self):
        buf = C.create_string_buffer(1 << 16)
        self.L.fadb_expr_describe(self.h, buf, len(buf))
        return buf.value.decode()

    def bind(self, root, flags=0):
        out = (C.c_size_t * 2)()
        check(self.L.fadb_expr_bind(self.h, root, flags, out))
        self.root = root
        rows, cols = C.c_size_t(), C.c_size_t()
        check(self.L.fadb_expr_node_size(self.h, root, C.byref(rows), C.byref(cols)))
        self.root_size = rows.value * cols.value
        return int(out[0]), int(out[1])

    def _seed(self, seed):
        import numpy as np
        if np.ndim(seed) == 0:
            return float(seed), None, None
        import torch
        flat = np.asfortranarray(np.asarray(seed, dtype=np.float64)).reshape(-1, order="F")
        t = torch.from_numpy(np.ascontiguousarray(flat)).cuda()
        return 0.0, t, C.c_void_p(t.data_ptr())

    def feval(self):
        import numpy as np
        out = np.zeros(self.root_size)
        check(self.L.fadb_expr_feval(self.h, out.ctypes.data_as(_dp)))
        return out

    def beval(self, seed=1.0):
        s, keep, p = self._seed(seed)
        check(self.L.fadb_expr_beval(self.h, s, p))
        check(self.L.fadb_sync())

    def autodiff(self, seed=1.0):
        import numpy as np
        s, keep, p = self._seed(seed)
        out = np.zeros(self.root_size)
        check(self.L.fadb_expr_autodiff(self.h, s, p, out.ctypes.data_as(_dp)))
        return out

    def value(self, nid):
        check(self.L.fadb_sync())
        return self.vars[nid][0].cpu().numpy()

    def adj(self, nid):
        check(self.L.fadb_sync())
        return self.vars[nid][1].cpu().numpy()

    def reset_adj(self):
        for v, a in self.vars.values():
            a.zero_()
