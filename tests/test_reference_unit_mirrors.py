"""Mirrors of the reference's per-node unit tests (test/reverse/core/*_unittest.cpp), same fixture values
(test/testutil/base_fixture.hpp:96-127: scl 2.31, vec {3.1,-2.3,1.3,0,5.1}, mat 2x3), same seeds, expected
values restated in numpy from the functor definitions.  Every case runs on the CPU oracle (not gpu) and,
marked gpu, through the C ABI on the device (specialised kernels by default)."""
import numpy as np
import pytest

import oracle_py as O

SCL = 2.31
VEC = np.array([3.1, -2.3, 1.3, 0.0, 5.1])
MAT = np.array([[3.1, -2.3, 1.3], [0.2, 5.1, -0.9]])
SEED = 3.14
VSEED = np.array([2.13, 0.3231, 4.231, 0.2, 3.21])          # binary_unittest.cpp:78
MSEED = np.array([[2.0, 3.0, 4.0], [1.0, 0.5, 0.25]])       # same shape as binary_unittest.cpp:79-80 (values free)
ULP4 = dict(rtol=2e-15, atol=1e-300)


def _gpu_backend():
    import torch
    assert torch.cuda.is_available()
    import fastad_b200 as F
    F.load()
    F.check(F.lib().fadb_init(0))
    F.use_torch_stream()
    return F.Expr


@pytest.fixture(params=["oracle", pytest.param("gpu", marks=pytest.mark.gpu)])
def E(request):
    return O.OracleExpr if request.param == "oracle" else _gpu_backend()


def close(a, b, **kw):
    kw = kw or ULP4
    np.testing.assert_allclose(np.asarray(a, dtype=float).reshape(-1, order="F"),
                               np.asarray(b, dtype=float).reshape(-1, order="F"), **kw)


# ---------------------------------------------------------------- unary_unittest.cpp:169-333
UNARY = {   # name: (f, f')  -- unary.hpp:190-296
    "neg": (lambda x: -x, lambda x: -np.ones_like(x)),
    "sin": (np.sin, np.cos),
    "cos": (np.cos, lambda x: -np.sin(x)),
    "tan": (np.tan, lambda x: 1 / np.cos(x) ** 2),
    "asin": (np.arcsin, lambda x: 1 / np.sqrt(1 - x * x)),
    "acos": (np.arccos, lambda x: -1 / np.sqrt(1 - x * x)),
    "atan": (np.arctan, lambda x: 1 / (1 + x * x)),
    "exp": (np.exp, np.exp),
    "log": (np.log, lambda x: 1 / x),
    "sqrt": (np.sqrt, lambda x: 0.5 / np.sqrt(x)),
    "sigmoid": (lambda x: 1 / (1 + np.exp(-x)), lambda x: np.exp(-x) / (1 + np.exp(-x)) ** 2),
    "sinh": (np.sinh, np.cosh),
    "cosh": (np.cosh, np.sinh),
    "tanh": (np.tanh, lambda x: 1 - np.tanh(x) ** 2),
}


@pytest.mark.parametrize("name", sorted(UNARY))
def test_unary_scl_and_vec(E, name):
    f, df = UNARY[name]
    v3 = np.array([0.3, 0.71, 0.12]) if name in ("asin", "acos", "log", "sqrt") else np.array([3.1, -2.3, 1.3])
    vseed = np.array([2.13, 0.3231, 4.231])                 # unary_unittest.cpp:44
    s = 0.431 if name in ("asin", "acos") else SCL
    e = E(); x = e.var(s); e.bind(e.unary(name, x))
    close(e.autodiff(SEED), f(np.array(s)), rtol=4e-15)
    close(e.adj(x), SEED * df(np.array(s)), rtol=1e-14)
    e = E(); v = e.var(v3); e.bind(e.unary(name, v))
    close(e.autodiff(vseed), f(v3), rtol=4e-15)
    close(e.adj(v), vseed * df(v3), rtol=1e-14)


def test_unary_vec_seed_with_zeros(E):
    # unary_unittest.cpp:120-133: only the seeded entries receive an adjoint (MockUnary f = 2x)
    e = E(); v = e.var(VEC); e.bind(e.unary("mock2x", v))
    vs = np.zeros(5); vs[1] = vs[3] = SEED
    close(e.autodiff(vs), 2 * VEC)
    close(e.adj(v), 2 * vs)


# ---------------------------------------------------------------- binary_unittest.cpp:181-384
def _mockb(x, y):
    return x - 2 * y


def test_binary_mock_all_shape_pairs(E):
    # scl-scl: both operands are the SAME variable: adj = blmap + brmap = seed - 2 seed   (:189-196)
    e = E(); s = e.var(SCL); e.bind(e.binary("mockb", s, s))
    close(e.autodiff(SEED), _mockb(SCL, SCL)); close(e.adj(s), SEED - 2 * SEED)
    # scl-vec (:204-213): scalar adjoint is the sum of the seed
    e = E(); s, v = e.var(SCL), e.var(VEC); e.bind(e.binary("mockb", s, v))
    close(e.autodiff(VSEED), _mockb(SCL, VEC)); close(e.adj(s), VSEED.sum()); close(e.adj(v), -2 * VSEED)
    # vec-scl
    e = E(); s, v = e.var(SCL), e.var(VEC); e.bind(e.binary("mockb", v, s))
    close(e.autodiff(VSEED), _mockb(VEC, SCL)); close(e.adj(v), VSEED); close(e.adj(s), -2 * VSEED.sum())
    # vec-vec, same variable (:223-232)
    e = E(); v = e.var(VEC); e.bind(e.binary("mockb", v, v))
    close(e.autodiff(VSEED), _mockb(VEC, VEC)); close(e.adj(v), VSEED - 2 * VSEED)
    # mat-mat, same variable (:242-251)
    e = E(); m = e.var(MAT); e.bind(e.binary("mockb", m, m))
    close(e.autodiff(MSEED), _mockb(MAT, MAT)); close(e.adj(m), MSEED - 2 * MSEED)
    # scl-(scl-scl) (:253-264): f(s, f(s,s)); adjoint 3 seed
    e = E(); s = e.var(SCL); e.bind(e.binary("mockb", s, e.binary("mockb", s, s)))
    close(e.autodiff(SEED), _mockb(SCL, _mockb(SCL, SCL))); close(e.adj(s), 3 * SEED)


@pytest.mark.parametrize("op", ["add", "sub", "mul", "div"])
def test_binary_operators_scl_vec_mat(E, op):
    # binary_unittest.cpp:270-384: operator results and adjoints for every operand shape pair
    f = dict(add=lambda x, y: x + y, sub=lambda x, y: x - y, mul=lambda x, y: x * y, div=lambda x, y: x / y)[op]
    dl = dict(add=lambda x, y: 1 + 0 * x * y, sub=lambda x, y: 1 + 0 * x * y, mul=lambda x, y: y + 0 * x, div=lambda x, y: 1 / y + 0 * x)[op]
    dr = dict(add=lambda x, y: 1 + 0 * x * y, sub=lambda x, y: -1 + 0 * x * y, mul=lambda x, y: x + 0 * y, div=lambda x, y: -x / y ** 2)[op]
    u = np.array([0.7, 1.9, -2.4, 3.3, 0.5])
    w = np.array([[1.5, -0.4, 2.2], [0.9, 3.1, -1.7]])
    for a, b, seed in ((SCL, 0.83, SEED), (SCL, u, VSEED), (u, SCL, VSEED), (VEC + 0.05, u, VSEED), (MAT, w, MSEED), (SCL, w, MSEED)):
        e = E(); x, y = e.var(a), e.var(b); e.bind(e.binary(op, x, y))
        close(e.autodiff(seed), f(np.asarray(a), np.asarray(b)) + 0 * np.asarray(seed), rtol=4e-15)
        gl, gr = seed * dl(np.asarray(a), np.asarray(b)), seed * dr(np.asarray(a), np.asarray(b))
        close(e.adj(x), gl.sum() if np.ndim(a) == 0 else gl, rtol=1e-14)
        close(e.adj(y), gr.sum() if np.ndim(b) == 0 else gr, rtol=1e-14)


# ---------------------------------------------------------------- sum_unittest.cpp:71-233, prod_unittest.cpp:95-273
def test_sum_iter_over_scalars_vectors_matrices(E):
    size = 5
    vseed = np.array([0.23, 0.1, 3.231, 4.23, 5.0])         # sum_unittest.cpp:38
    for base, seed in ((SCL, SEED), (VEC, vseed), (MAT, MSEED)):
        e = E(); xs = [e.var(base) for _ in range(size)]
        e.bind(e.sum_iter([e.unary("mock2x", x) for x in xs]))
        close(e.autodiff(seed), size * 2 * np.asarray(base))
        for x in xs:
            close(e.adj(x), 2 * np.asarray(seed) + 0 * np.asarray(base))


def test_sum_elem_of_vec_and_mat(E):
    # sum_unittest.cpp:150-233: sum(expr) of all elements, seed broadcast to every element
    for base in (VEC, MAT):
        e = E(); x = e.var(base); e.bind(e.sum(e.unary("mock2x", x)))
        close(e.autodiff(SEED), 2 * base.sum())
        close(e.adj(x), 2 * SEED + 0 * base)


def test_prod_iter_and_prod_elem(E):
    size = 5
    vseed = np.array([0.23, 0.1, 3.231, 4.23, 5.0])         # prod_unittest.cpp:37
    e = E(); xs = [e.var(SCL) for _ in range(size)]
    e.bind(e.prod_iter([e.unary("mock2x", x) for x in xs]))
    p = (2 * SCL) ** size
    close(e.autodiff(SEED), p, rtol=4e-15)
    for x in xs:
        close(e.adj(x), SEED * 2 * p / (2 * SCL), rtol=1e-14)
    # vectors, elementwise product of the list; fixture vec has an exact zero at index 3 (base_fixture.hpp:113)
    e = E(); xs = [e.var(VEC) for _ in range(size)]
    e.bind(e.prod_iter([e.unary("mock2x", x) for x in xs]))
    close(e.autodiff(vseed), (2 * VEC) ** size, rtol=4e-15)
    with np.errstate(divide="ignore", invalid="ignore"):
        ref = np.where(VEC == 0, 0.0, vseed * 2 * (2 * VEC) ** size / (2 * VEC))
    for x in xs:
        close(e.adj(x), ref, rtol=1e-14)
    # prod(expr) over the elements of one vector with one exact zero (prod_unittest.cpp:200-273)
    e = E(); x = e.var(VEC); e.bind(e.prod(e.unary("mock2x", x)))
    close(e.autodiff(SEED), 0.0)
    others = np.prod(2 * VEC[VEC != 0])
    close(e.adj(x), np.where(VEC == 0, SEED * 2 * others, 0.0), rtol=1e-14)


# ---------------------------------------------------------------- pow_unittest.cpp:72-183
def test_pow_nodes(E):
    seed, vseed = 0.32188, np.array([2.13, 0.3231, 4.231, 2.41, 3.13])      # pow_unittest.cpp:30, 48
    e = E(); s = e.var(SCL); e.bind(e.pow(2, e.unary("mock2x", s)))
    close(e.autodiff(seed), (2 * SCL) ** 2); close(e.adj(s), seed * 8 * SCL, rtol=1e-14)
    e = E(); s = e.var(SCL); e.bind(e.pow(-2, e.unary("mock2x", s)))
    close(e.autodiff(seed), 1 / (2 * SCL) ** 2, rtol=4e-15); close(e.adj(s), seed * 2 * -2 / (2 * SCL) ** 3, rtol=1e-14)
    e = E(); s = e.var(SCL); e.bind(e.pow(0, e.unary("mock2x", s)))
    close(e.autodiff(seed), 1.0); close(e.adj(s), 0.0)
    e = E(); v = e.var(VEC); e.bind(e.pow(2, e.unary("mock2x", v)))          # :139-147, zero element included
    close(e.autodiff(vseed), (2 * VEC) ** 2); close(e.adj(v), vseed * 8 * VEC, rtol=1e-14)
    vv = np.array([3.1, -2.3, 1.3, 0.7, 5.1])
    e = E(); v = e.var(vv); e.bind(e.pow(-1, e.unary("mock2x", v)))          # :157-165
    close(e.autodiff(vseed), 1 / (2 * vv), rtol=4e-15); close(e.adj(v), vseed * 2 * -1 / (2 * vv) ** 2, rtol=1e-14)


# ---------------------------------------------------------------- dot_unittest.cpp:84-179
DOT_VEC = np.array([0.312332, -3.213, 9.135])
DOT_MT = np.array([[2.3, 0.0], [0.523, -2.3], [4.2, 95.2]])
DOT_VSEED = np.array([2.3, 1.32])
DOT_MSEED = np.array([[2.34, -3.2], [3.4, 1.23455]])


def test_dot_vars_unary_and_constant(E):
    e = E(); m, v = e.var(MAT), e.var(DOT_VEC); e.bind(e.dot(m, v))
    close(e.autodiff(DOT_VSEED), MAT @ DOT_VEC, rtol=1e-14)
    close(e.adj(m), np.outer(DOT_VSEED, DOT_VEC), rtol=1e-14); close(e.adj(v), MAT.T @ DOT_VSEED, rtol=1e-14)
    # operands through MockUnary (f = 2x): :103-126
    e = E(); m, v = e.var(MAT), e.var(DOT_VEC); e.bind(e.dot(e.unary("mock2x", m), e.unary("mock2x", v)))
    close(e.autodiff(DOT_VSEED), (2 * MAT) @ (2 * DOT_VEC), rtol=1e-14)
    close(e.adj(m), 2 * np.outer(DOT_VSEED, 2 * DOT_VEC), rtol=1e-14); close(e.adj(v), 2 * (2 * MAT).T @ DOT_VSEED, rtol=1e-14)
    # matrix x matrix: :128-153
    e = E(); m, t = e.var(MAT), e.var(DOT_MT); e.bind(e.dot(e.unary("mock2x", m), e.unary("mock2x", t)))
    close(e.autodiff(DOT_MSEED), (2 * MAT) @ (2 * DOT_MT), rtol=1e-14)
    close(e.adj(m), 2 * DOT_MSEED @ (2 * DOT_MT).T, rtol=1e-14); close(e.adj(t), 2 * (2 * MAT).T @ DOT_MSEED, rtol=1e-14)
    # constant matrix: :155-179
    e = E(); v = e.var(DOT_VEC); e.bind(e.dot(e.const(MAT), v))
    close(e.autodiff(DOT_VSEED), MAT @ DOT_VEC, rtol=1e-14); close(e.adj(v), MAT.T @ DOT_VSEED, rtol=1e-14)


# ---------------------------------------------------------------- glue_unittest.cpp:65-108, eq_unittest.cpp:63-137
def test_glue_and_placeholders(E):
    vseed = np.array([2.3, 1.4, -2.3, 0.3, 1.3])            # glue_unittest.cpp:47
    # (place = 2 x, 2 place): value 4 x, place_adj 2 seed, x_adj 4 seed  (:84-93)
    for base, seed in ((SCL, SEED), (VEC, vseed), (MAT, MSEED)):
        e = E(); x, place = e.var(base), e.var(0 * np.asarray(base))
        e.bind(e.glue(e.eq(place, e.unary("mock2x", x)), e.unary("mock2x", place)))
        close(e.autodiff(seed), 4 * np.asarray(base))
        close(e.value(place), 2 * np.asarray(base))
        close(e.adj(place), 2 * np.asarray(seed) + 0 * np.asarray(base))
        close(e.adj(x), 4 * np.asarray(seed) + 0 * np.asarray(base))


def test_eq_node_value_and_adjoint(E):
    # eq_unittest.cpp:63-137: w = f(x): value stored in w, adjoint flows through w into x
    for base, seed in ((SCL, SEED), (VEC, VSEED), (MAT, MSEED)):
        e = E(); x, w = e.var(base), e.var(0 * np.asarray(base))
        e.bind(e.eq(w, e.unary("mock2x", x)))
        close(e.autodiff(seed), 2 * np.asarray(base))
        close(e.value(w), 2 * np.asarray(base))
        close(e.adj(w), np.asarray(seed) + 0 * np.asarray(base))
        close(e.adj(x), 2 * np.asarray(seed) + 0 * np.asarray(base))


# ---------------------------------------------------------------- eval_unittest.cpp:32-48, var_view_unittest.cpp
def test_autodiff_accumulates_and_reset(E):
    e = E(); x = e.var(VEC); e.bind(e.sum(e.binary("mul", x, x)))
    e.autodiff(1.0); e.autodiff(1.0)
    close(e.adj(x), 4 * VEC)                                 # adjoints accumulate (var_view.hpp:91-94)
    e.reset_adj(); e.autodiff(0.5)
    close(e.adj(x), VEC)
