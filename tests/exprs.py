"""Expression cases shared by the oracle/product parity tests.  Each case builds the SAME
expression through the common builder vocabulary (OracleExpr / fastad_b200.Expr) and returns
(root, [variable node ids], seed)."""
import numpy as np

import kats

RNG = np.random.default_rng(20261019)


def _vec(n, lo=0.5, hi=2.0):
    return RNG.uniform(lo, hi, n)


def scalar_chain(e):   # node_inttest.cpp:129-159
    l1, l2, l3 = e.var(1.5928), e.var(-0.291), e.var(5.1023)
    root = e.binary("sub",
                    e.binary("add", e.binary("mul", l1, l3),
                             e.binary("mul", e.unary("sin", e.unary("cos", e.binary("add", l1, l2))), l2)),
                    e.binary("div", l1, e.unary("exp", l3)))
    return root, [l1, l2, l3], 1.0


def placeholders(e):   # bind_unittest.cpp:13-35
    w1, w2, w3, w4 = e.var(1.0), e.var(2.0), e.var(0.0), e.var(0.0)
    root = e.glue(e.eq(w3, e.binary("mul", w1, w2)), e.eq(w4, e.binary("mul", w3, w3)))
    return root, [w1, w2, w3, w4], 1.0


def sumnode_benchmark(e):   # benchmark/sum_benchmark.cpp:46-61
    vs = [e.var(float(i)) for i in range(10)]
    w4, w5 = e.var(0.0), e.var(0.0)
    s = e.sum_iter([e.binary("mul", v, v) for v in vs])
    root = e.glue(e.eq(w4, s), e.eq(w5, e.binary("add", e.binary("mul", w4, w4), e.unary("sin", w4))))
    return root, vs + [w4, w5], 1.0


def prodnode_benchmark(e):   # benchmark/prod_benchmark.cpp:46-64 (values shifted off zero)
    vs = [e.var(0.5 + 0.1 * i) for i in range(10)]
    w4, w5 = e.var(0.0), e.var(0.0)
    p = e.prod_iter([e.binary("mul", v, v) for v in vs])
    root = e.glue(e.eq(w4, p), e.eq(w5, e.binary("add", e.binary("mul", w4, w4), e.unary("cos", w4))))
    return root, vs + [w4, w5], 1.0


def vec_sum_chain(n):
    def build(e):   # cf. sum_benchmark.cpp:84-85: sum(-0.5 * pow<2>((x - c*w0) / w1))
        x = e.var(_vec(n))
        w0, w1 = e.var(2.0), e.var(1.3)
        c = e.const(_vec(n))
        z = e.binary("div", e.binary("sub", x, e.binary("mul", w0, c)), w1)
        root = e.sum(e.binary("mul", e.const(-0.5), e.pow(2, z)))
        return root, [x, w0, w1], 0.75
    return build


def vec_nested_sum(n):
    def build(e):   # a reduction nested under scalar nodes: w4 = sum(exp(x)), w5 = w4*w4 + sin(w4)
        x = e.var(_vec(n, 0.1, 0.9))
        w4, w5 = e.var(0.0), e.var(0.0)
        root = e.glue(e.eq(w4, e.sum(e.unary("exp", x))),
                      e.eq(w5, e.binary("add", e.binary("mul", w4, w4), e.unary("sin", w4))))
        return root, [x, w4, w5], 1.0
    return build


def vec_scalar_times_sum(n):
    def build(e):   # scalar broadcast of a reduced value back into a vector map
        x, y = e.var(_vec(n)), e.var(_vec(n))
        root = e.sum(e.binary("mul", x, e.sum(e.unary("log", y))))
        return root, [x, y], 1.0
    return build


def vec_prod(n, zeros=()):
    def build(e):
        v = _vec(n, 0.9, 1.1)
        for z in zeros:
            v[z] = 0.0
        x = e.var(v)
        root = e.prod(e.binary("mul", x, e.const(1.0 + 0 * v)))
        return root, [x], 2.0
    return build


def vec_root_with_vector_seed(n):
    def build(e):   # vector-valued root, vector seed (eval.hpp:68-73)
        x, y = e.var(_vec(n)), e.var(_vec(n))
        s = e.var(0.7)
        root = e.binary("add", e.binary("mul", e.unary("tanh", x), y), e.binary("div", s, x))
        return root, [x, y, s], _vec(n, -1, 1)
    return build


def vec_placeholder_chain(n):
    def build(e):   # vector placeholders + glue, as the batched Black-Scholes expression uses them
        x = e.var(_vec(n))
        a, b = e.var(np.zeros(n)), e.var(np.zeros(n))
        root = e.glue(e.eq(a, e.unary("log", x)),
                      e.eq(b, e.binary("mul", a, a)),
                      e.binary("sub", e.unary("exp", b), e.binary("mul", a, x)))
        return root, [x, a, b], np.ones(n)
    return build


def placeholder_sum(n):
    def build(e):   # a reduction over a stage that assigns a placeholder (sum terminal + scalar channel + placeholder store)
        x, a, s = e.var(_vec(n)), e.var(np.zeros(n)), e.var(0.8)
        root = e.glue(e.eq(a, e.unary("log", x)),
                      e.sum(e.binary("add", e.binary("mul", a, a), e.binary("mul", s, x))))
        return root, [x, a, s], 1.0
    return build


def placeholder_reassigned(n):
    def build(e):   # the same placeholder assigned twice, read between and after the assignments
        x = e.var(_vec(n))
        a, b = e.var(np.zeros(n)), e.var(np.zeros(n))
        root = e.glue(e.eq(a, e.unary("log", x)),
                      e.eq(b, e.binary("mul", a, a)),
                      e.eq(a, e.binary("add", b, x)),
                      e.binary("sub", e.unary("sin", a), e.binary("mul", b, x)))
        return root, [x, a, b], _vec(n, -1, 1)
    return build


def placeholder_normal(n):
    def build(e):   # log-pdf terminal fed by a placeholder: normal(y, a = w x, sigma)
        x, y = e.const(_vec(n)), e.const(_vec(n))
        w, sg, a = e.var(0.9), e.var(1.4), e.var(np.zeros(n))
        return e.glue(e.eq(a, e.binary("mul", w, x)), e.normal(y, a, sg)), [w, sg, a], 1.0
    return build


# Not compared with the oracle: the reference's EqNode::beval reads the placeholder's STORAGE (eq.hpp:101-107), so a placeholder
# assigned twice is differentiated with its last value throughout -- not a derivative of anything; the stage kernels keep
# one value per assignment.  The case stays for the ring-vs-plain bit-identity test (two assignments of one ring slot).
NO_ORACLE = {"placeholder_reassigned_ring"}


def normal_with_expr_mean(n):
    def build(e):   # univariate regression: normal(y | alpha + beta * x, sigma)
        y, x = e.const(_vec(n, -1, 1)), e.const(_vec(n))
        alpha, beta, sigma = e.var(0.3), e.var(-0.2), e.var(1.7)
        mu = e.binary("add", alpha, e.binary("mul", beta, x))
        root = e.normal(y, mu, sigma)
        return root, [alpha, beta, sigma], 1.0
    return build


def normal_vec_sigma_expr(n):
    def build(e):   # vector sigma produced by an expression: normal(x | mu, exp(ls))
        x, mu, ls = e.var(_vec(n, -1, 1)), e.var(_vec(n, -1, 1)), e.var(_vec(n, -0.5, 0.5))
        root = e.normal(x, mu, e.unary("exp", ls))
        return root, [x, mu, ls], 1.0
    return build


def sum_of_normals(n):
    def build(e):   # log-posterior shape: normal(x|mu,sigma) + normal(mu|0,10)
        x = e.const(_vec(n, -1, 1))
        mu, sigma = e.var(0.1), e.var(1.3)
        root = e.binary("add", e.normal(x, mu, sigma), e.normal(mu, e.const(0.0), e.const(10.0)))
        return root, [mu, sigma], 1.0
    return build


def dot_norm_sum(e):   # node_inttest.cpp:402-442 shape: sum(dot(M, v) elementwise-squared)
    M = e.var(np.array([[1.0, 2.0], [-1.0, 3.0], [2.0, -3.0]]))
    v = e.var(np.array([[0.5], [-1.5]]))
    d = e.dot(M, v)
    root = e.sum(e.binary("mul", d, d))
    return root, [M, v], 1.0


def quadratic_form(e):   # README.md:582-607
    x = e.var(np.array([[0.5], [0.6]]))
    xt = e.var(np.array([[0.5, 0.6]]))
    S = e.const(np.array([[2.0, 3.0], [3.0, 6.0]]))
    return e.dot(e.dot(xt, S), x), [x, xt], np.ones((1, 1))


def mlp3_row(e):   # README.md:676-726 for one data row
    p = kats.NN_PARM
    A1, b1 = e.var(p[0:6].reshape(3, 2, order="F")), e.var(p[6:9].reshape(3, 1))
    A2, b2 = e.var(p[9:18].reshape(3, 3, order="F")), e.var(p[18:21].reshape(3, 1))
    A3, b3 = e.var(p[21:24].reshape(1, 3)), e.var(p[24:25].reshape(1, 1))
    xi, yi = e.const(kats.NN_X[2].reshape(2, 1)), e.const(kats.NN_Y[2].reshape(1, 1))
    x1a = e.binary("add", e.dot(A1, xi), b1)
    x1b = e.binary("add", e.dot(A1, xi), b1)
    y1 = e.unary("sigmoid", x1a)
    y2 = e.binary("add", e.unary("sigmoid", e.binary("add", e.dot(A2, y1), b2)), x1b)
    y3 = e.binary("add", e.dot(A3, y2), b3)
    return e.pow(2, e.binary("sub", yi, y3)), [A1, b1, A2, b2, A3, b3], np.ones((1, 1))


def linreg_expr(rows, cols):
    def build(e):   # the C4 expression: normal(y, dot(X, w), sigma)
        import synth
        X, y, w_true = synth.linreg_inputs(rows, cols, seed=7)
        Xc, yc = e.const(X), e.const(y)
        w, sigma = e.var(w_true + 0.01), e.var(1.3)
        return e.normal(yc, e.dot(Xc, w), sigma), [w, sigma], 1.0
    return build


def linreg_with_intercept(rows, cols, sigma=1.3, link=False, const_sigma=False):
    def build(e):   # normal(y, dot(X, w) + b, sigma): C4 with an intercept (and optionally a link function on the mean)
        import synth
        X, y, w_true = synth.linreg_inputs(rows, cols, seed=13)
        Xc, yc = e.const(X), e.const(y + 0.4)
        w, b = e.var(w_true + 0.01), e.var(0.35)
        sg = e.const(sigma) if const_sigma else e.var(sigma)
        mean = e.binary("add", e.dot(Xc, w), b)
        if link:
            mean = e.binary("mul", e.const(3.0), e.unary("tanh", e.binary("mul", e.const(0.3), mean)))
        return e.normal(yc, mean, sg), ([w, b] if const_sigma else [w, b, sg]), 0.8
    return build


def posterior_linreg(rows, cols):
    def build(e):   # a posterior: log-likelihood + log-priors.  normal(y | X w + b, s) + normal(w | 0, 2) + normal(b | 0.5, 10) - 0.5 * sum(s^2)
        import synth
        X, y, w_true = synth.linreg_inputs(rows, cols, seed=21)
        w, b, sg = e.var(w_true + 0.02), e.var(0.3), e.var(1.2)
        lik = e.normal(e.const(y + 0.3), e.binary("add", e.dot(e.const(X), w), b), sg)
        pw = e.normal(w, e.const(0.0), e.const(2.0))
        pb = e.normal(b, e.const(0.5), e.const(10.0))
        ps = e.binary("mul", e.const(-0.5), e.binary("mul", sg, sg))
        root = e.binary("add", e.binary("add", e.binary("add", lik, pw), pb), ps)
        return root, [w, b, sg], 0.9
    return build


def least_squares(rows, cols, flipped=False, x_var=False, inner=False):
    def build(e):   # README.md:640-662 batched: sum(pow<2>(y - dot(X, w))); X constant (or a variable: outer-product adjoint)
        import synth
        X, y, w_true = synth.linreg_inputs(rows, cols, seed=9)
        Xn = e.var(X) if x_var else e.const(X)
        yc = e.const(y)
        w = e.var(w_true + 0.05)
        d = e.dot(Xn, w)
        r = e.binary("sub", d, yc) if flipped else e.binary("sub", yc, d)
        if inner:          # not the recognised shape: the residual goes through another map first
            r = e.binary("mul", r, e.unary("exp", e.binary("mul", e.const(-0.01), r)))
        return e.sum(e.pow(2, r)), ([Xn, w] if x_var else [w]), 0.75
    return build


def glm_logistic(rows, cols, weights=False):
    def build(e):   # logistic-regression log-likelihood: sum(w_i (y_i eta_i - log(1 + exp(eta_i)))), eta = dot(X, beta) + b
        X = np.asfortranarray(RNG.normal(size=(rows, cols)) / np.sqrt(cols))
        y = RNG.integers(0, 2, rows).astype(np.float64)
        beta, b = e.var(0.3 * RNG.normal(size=cols)), e.var(0.2)
        eta = e.binary("add", e.dot(e.const(X), beta), b)
        ll = e.binary("sub", e.binary("mul", e.const(y), eta), e.unary("log", e.binary("add", e.const(1.0), e.unary("exp", eta))))
        if weights:
            ll = e.binary("mul", e.const(RNG.uniform(0.5, 1.5, rows)), ll)
        return e.sum(ll), [beta, b], 0.75
    return build


def many_scalar_params(n, nparams):
    def build(e):   # a vector model with more scalar parameters than the interpreter has reduction channels
        x = e.var(RNG.uniform(0.2, 1.5, n))
        ps = [e.var(float(v)) for v in RNG.uniform(0.5, 1.5, nparams)]
        acc = e.binary("mul", ps[0], x)
        for k, p in enumerate(ps[1:], 1):
            t = e.binary("mul", p, e.unary(("sin", "cos", "exp", "tanh")[k % 4], e.binary("mul", e.const(0.3 * k), x)))
            acc = e.binary("add", acc, t)
        return e.sum(acc), [x] + ps, 0.5
    return build


def long_chain(n, depth):
    def build(e):   # more nodes than the interpreter's 64 slots, one stage
        x = e.var(RNG.uniform(0.5, 1.5, n)); c = e.const(RNG.uniform(0.5, 1.5, n))
        v = x
        for k in range(depth):
            v = e.binary("add", e.binary("mul", e.unary(("sin", "cos", "atan")[k % 3], v), c), e.binary("div", x, e.const(1.0 + k)))
        return e.sum(v), [x], 1.0
    return build


def many_vector_leaves(n, nl):
    def build(e):   # more distinct per-element operands than the interpreter's leaf table holds
        xs = [e.var(RNG.uniform(0.5, 1.5, n)) for _ in range(nl)]
        acc = xs[0]
        for k, x in enumerate(xs[1:], 1):
            acc = e.binary("add", acc, e.binary("mul", e.const(1.0 / k), e.pow(2, x)))
        return e.sum(acc), xs, 2.0
    return build


JIT_ONLY = ("many_scalar_params", "long_chain_200_nodes", "many_vector_leaves")


def dot_tall_multi_rhs(rows, cols, outs, x_var=False):
    def build(e):   # multi-output regression: X (rows x cols) times W (cols x outs); sum of squared residuals
        X = RNG.normal(size=(rows, cols)); W = RNG.normal(size=(cols, outs)); Y = X @ W + 0.1 * RNG.normal(size=(rows, outs))
        Xn = e.var(X) if x_var else e.const(X)
        Wn = e.var(W + 0.05)
        r = e.binary("sub", e.const(Y), e.dot(Xn, Wn))
        return e.sum(e.pow(2, r)), ([Xn, Wn] if x_var else [Wn]), 0.5
    return build


def dot_wide_batch(hidden, inputs, batch):
    def build(e):   # a small weight matrix times a batch of columns: dot(W, X), X inputs x batch (the batched layer of a network)
        W = e.var(RNG.normal(size=(hidden, inputs))); b = e.var(RNG.normal(size=(hidden, 1)))
        X = e.const(RNG.uniform(-1, 1, size=(inputs, batch)))
        T = e.const(RNG.uniform(0, 1, size=(hidden, batch)))
        h = e.unary("sigmoid", e.dot(W, X))
        return e.sum(e.pow(2, e.binary("sub", h, T))), [W], 1.0
    return build


def dot_matrix_product(m, k, p):
    def build(e):   # matrix x matrix with both operands variable, through a map on each side (dot_unittest.cpp:128-153 shape, large)
        A = e.var(RNG.normal(size=(m, k))); B = e.var(RNG.normal(size=(k, p)))
        prod = e.dot(e.unary("tanh", A), B)
        return e.sum(e.pow(2, prod)), [A, B], 0.5
    return build


def black_scholes_vec(n, put=False):
    def build(e):   # example/black_scholes.cpp:16-39 over vec leaves
        import synth
        inp = synth.black_scholes_inputs(n, seed=11)
        S, sigma, r = e.var(inp["S"]), e.var(inp["sigma"]), e.var(inp["r"])
        K, tau, sq = e.const(inp["K"]), e.const(inp["tau"]), e.const(np.sqrt(inp["tau"]))
        c0, c1, c2 = e.var(np.zeros(n)), e.var(np.zeros(n)), e.var(np.zeros(n))

        def phi(x):
            er = e.unary("erf", e.binary("div", x, e.const(np.sqrt(2.0))))
            return e.binary("mul", e.const(0.5), e.binary("add", er, e.const(1.0)))
        e0 = e.eq(c0, e.unary("log", e.binary("div", S, K)))
        drift = e.binary("mul", e.binary("add", r, e.binary("div", e.binary("mul", sigma, sigma), e.const(2.0))), tau)
        e1 = e.eq(c1, e.binary("div", e.binary("add", c0, drift), e.binary("mul", sigma, sq)))
        e2 = e.eq(c2, e.binary("sub", c1, e.binary("mul", sigma, sq)))
        pv = e.binary("mul", K, e.unary("exp", e.binary("mul", e.unary("neg", r), tau)))
        if not put:
            price = e.binary("sub", e.binary("mul", phi(c1), S), e.binary("mul", phi(c2), pv))
        else:
            price = e.binary("sub", e.binary("mul", phi(e.unary("neg", c2)), pv),
                             e.binary("mul", phi(e.unary("neg", c1)), S))
        return e.glue(e0, e1, e2, price), [S, sigma, r, c0, c1, c2], np.ones(n)
    return build


# ---- SURVEY.md 8f "next" rows -------------------------------------------------------------
def norm_vec(e):   # norm_unittest.cpp:31-49: NormNode over MockUnary(vec), seed 9.2313
    v = e.var(kats.FIX_VEC)
    return e.norm(e.unary("mock2x", v)), [v], 9.2313


def norm_mat_chain(n):
    def build(e):   # node_inttest.cpp:402-442 shape: norm of an elementwise expression of a matrix
        M = e.var(_vec(n * 3).reshape(n, 3))
        c = e.var(0.7)
        return e.norm(e.binary("mul", e.unary("sin", M), c)), [M, c], 1.0
    return build


def transpose_quadratic(e):   # README.md:582-607 with the real transpose node
    x = e.var(np.array([[0.5], [0.6]]))
    S = e.const(np.array([[2.0, 3.0], [3.0, 6.0]]))
    return e.dot(e.dot(e.transpose(x), S), x), [x], np.ones((1, 1))


def transpose_of_expr(e):
    M = e.var(np.array([[1.0, 2.0, -1.0], [0.5, -3.0, 2.0]]))
    w = e.var(np.array([[0.3], [-0.2]]))
    t = e.transpose(e.unary("exp", M))                # 3x2
    return e.sum(e.dot(t, w)), [M, w], 1.0


def if_else_scalar(taken):
    def build(e):   # if_else_unittest: scalar condition from a comparison of variables
        a, b = e.var(1.5 if taken else -0.5), e.var(0.25)
        x = e.var(np.array([0.5, 1.5, 2.5]))
        cond = e.binary("lt", b, a)
        root = e.sum(e.if_else(cond, e.binary("mul", x, a), e.unary("exp", e.binary("mul", x, b))))
        return root, [a, b, x], 1.0
    return build


def if_else_nested(n, c1, c2):
    def build(e):   # nested IfElseNode over vector branches, large enough for the 4-elements-per-thread specialised form
        a, b = e.var(c1), e.var(c2)
        x, y = e.var(_vec(n, 0.2, 1.8)), e.var(_vec(n, 0.5, 1.5))
        inner = e.if_else(e.binary("gt", b, e.const(0.0)), e.binary("mul", e.unary("sin", x), y), e.binary("div", x, y))
        outer = e.if_else(e.binary("lt", a, e.const(1.0)), inner, e.binary("mul", e.unary("exp", e.binary("mul", a, x)), b))
        return e.sum(e.binary("add", outer, e.binary("mul", a, y))), [a, b, x, y], 0.6
    return build


def if_else_logical(e):
    a, b, c = e.var(0.3), e.var(-1.2), e.var(2.0)
    cond = e.binary("land", e.binary("ge", a, e.const(0.0)), e.binary("lor", e.binary("gt", b, e.const(0.0)), e.binary("ne", c, a)))
    root = e.if_else(cond, e.binary("mul", a, c), e.binary("sub", b, c))
    return root, [a, b, c], 1.0


def opeq_scl_alias(e):   # eq_unittest.cpp:143-160: w *= w
    w = e.var(kats.FIX_SCL)
    return e.opeq("mul", w, w), [w], 3.2
def opeq_scl_noalias(e):   # eq_unittest.cpp:162-178: w /= MockUnary(w)
    w = e.var(kats.FIX_SCL)
    return e.opeq("div", w, e.unary("mock2x", w)), [w], 3.2
def opeq_vec_scl(e):   # eq_unittest.cpp:180-199: v *= s
    v, s_ = e.var(kats.FIX_VEC), e.var(kats.FIX_SCL)
    return e.opeq("mul", v, s_), [v, s_], np.array([2.1, -0.3, 1.0, 0.5, 4.0])
def opeq_vec_noalias(e):   # eq_unittest.cpp:220-240: v *= MockUnary(v)
    v = e.var(kats.FIX_VEC)
    return e.opeq("mul", v, e.unary("mock2x", v)), [v], np.array([2.1, -0.3, 1.0, 0.5, 4.0])
def opeq_chain(n):
    def build(e):   # accumulate-style program: w = x ; w += sin(x) ; w -= c*y ; sum(w*w)
        x, y, c = e.var(_vec(n)), e.var(_vec(n)), e.var(0.4)
        w = e.var(np.zeros(n))
        prog = e.glue(e.eq(w, e.unary("exp", x)), e.opeq("add", w, e.unary("sin", x)), e.opeq("sub", w, e.binary("mul", c, y)),
                      e.opeq("div", w, e.binary("add", y, e.const(1.0))))
        return e.glue(prog, e.sum(e.binary("mul", w, w))), [x, y, c, w], 1.0
    return build


def opeq_mul_then_sum(n):
    def build(e):   # (w *= x, sum(w)) -- the op= statement is not the root
        x, w = e.var(_vec(n)), e.var(_vec(n, -1.5, 1.5))
        return e.glue(e.opeq("mul", w, x), e.sum(w)), [x, w], 1.0
    return build
def opeq_div_self_then_sum(n):
    def build(e):   # (w /= exp(w), sum(w)): the right-hand side reads the PREVIOUS value of w (eq.hpp:240-271)
        w = e.var(_vec(n, -1.0, 1.0))
        return e.glue(e.opeq("div", w, e.unary("exp", w)), e.sum(w)), [w], 0.7
    return build
def opeq_then_dot(e):   # the op= stage is an inner stage: its output w is read by a dot stage through memory
    A = e.const(np.array([[1.0, 2.0, -0.5], [-1.0, 3.0, 0.25], [2.0, -3.0, 1.5], [0.5, 0.5, -2.0]]))
    w, x = e.var(np.array([[0.5], [-1.5], [2.0]])), e.var(np.array([[1.25], [0.75], [-0.5]]))
    d = e.dot(A, w)
    return e.glue(e.opeq("mul", w, x), e.sum(e.binary("mul", d, d))), [w, x], 1.0


# ---- other diagonal log-pdfs ------------------------------------------------
C_VX, C_VL, C_VS = np.array([0.5, -1.3, -3.2414999]), np.array([0.4, -2.30000001, -10.32]), np.array([0.51, 0.01, 3.4])


def cauchy_case(kind, seed=1.0, n=None):
    def build(e):   # cauchy_unittest.cpp:59-76 fixture values, or n random elements
        if n is None:
            xv, lv, sv = C_VX, C_VL, C_VS
        else:
            xv, lv, sv = _vec(n, -3, 3), _vec(n, -3, 3), _vec(n, 0.2, 3)
        x = e.var(xv if kind[0] == "v" else 0.421)
        l = e.var(lv if kind[1] == "v" else 0.341)
        s_ = e.var(sv if kind[2] == "v" else 2.132)
        return e.logpdf("cauchy", x, l, s_), [x, l, s_], seed
    return build


def uniform_case(kind, n=None, bad=False):
    def build(e):   # uniform_unittest.cpp:62-76 fixture values
        if n is None:
            xv, lo, hi = np.array([0.5, -2.3, -3.2414999]), np.array([0.4, -2.30000001, -10.32]), np.array([0.51, 0.0, 3.4])
        else:
            xv = _vec(n, -1, 1); lo = xv - _vec(n, 0.1, 2); hi = xv + _vec(n, 0.1, 2)
        if kind == "vss":
            xv = np.array([0.5, -2.3, -3.2414999]) if n is None else _vec(n, -3, 0.5)
        if bad:
            xv = xv.copy(); xv[0] = -10000.0
        x = e.var(xv if kind[0] == "v" else 0.45)
        a = e.var(lo if kind[1] == "v" else -3.2415)
        b = e.var(hi if kind[2] == "v" else 0.5231)
        return e.logpdf("uniform", x, a, b), [x, a, b], 0.7
    return build


def bernoulli_case(kind, xs=(1.0, 0.0, 1.0), ps=(0.3, 0.42, 0.98), p_scalar=0.0001):
    def build(e):   # bernoulli_unittest.cpp:48-57 fixture values
        x = e.const(np.array(xs)) if kind[0] == "v" else e.const(np.array([xs[0]]))
        p = e.var(np.array(ps)) if kind[1] == "v" else e.var(p_scalar)
        return e.logpdf("bernoulli", x, p), [p], 1.3
    return build


def cauchy_in_model(n):
    def build(e):   # log-posterior shape with an expression location: cauchy(y | a + b*x, exp(ls))
        y, x = e.const(_vec(n, -2, 2)), e.const(_vec(n))
        a, b, ls = e.var(0.1), e.var(0.7), e.var(-0.2)
        loc = e.binary("add", a, e.binary("mul", b, x))
        return e.binary("add", e.logpdf("cauchy", y, loc, e.unary("exp", ls)), e.logpdf("uniform", a, e.const(-5.0), e.const(5.0))), [a, b, ls], 1.0
    return build


# ---- dense-covariance normal -------------------------------------------------
SIG3 = np.array([[1.0, 0.0, 0.0], [0.3, 2.0, 0.0], [0.2, -0.3, 3.0]])    # normal_unittest.cpp:94-102 (lower only)


def dense_normal(mu_vec, n=None, const_sigma=False):
    def build(e):
        if n is None:
            xv, mv, Sig = kats.N_VX, (kats.N_VMU if mu_vec else kats.N_MU), SIG3
        else:
            Lo = np.tril(_vec(n * n, -1, 1).reshape(n, n), -1) / np.sqrt(n) + 1.2 * np.eye(n)
            Sig = Lo @ Lo.T
            xv, mv = _vec(n, -1, 1), (_vec(n, -1, 1) if mu_vec else 0.3)
        x, mu = e.var(xv), e.var(mv)
        S = e.const(Sig) if const_sigma else e.var(Sig)
        return e.normal(x, mu, S), [x, mu] + ([] if const_sigma else [S]), 1.0
    return build


def dense_normal_in_model(n):
    def build(e):   # multivariate normal with an expression mean plus a prior: N(y | a + b*t, Sigma) + N(a | 0, 10)
        Lo = np.tril(_vec(n * n, -1, 1).reshape(n, n), -1) / np.sqrt(n) + 1.5 * np.eye(n)
        y, t = e.const(_vec(n, -1, 1)), e.const(_vec(n))
        a, b = e.var(0.2), e.var(-0.4)
        S = e.var(Lo @ Lo.T)
        mean = e.binary("add", a, e.binary("mul", b, t))
        return e.binary("add", e.normal(y, mean, S), e.normal(a, e.const(0.0), e.const(10.0))), [a, b, S], 1.0
    return build


# ---- det / log_det / wishart ----------------------------------------------------
def _spd(n, diag=1.2):
    Lo = np.tril(_vec(n * n, -1, 1).reshape(n, n), -1) / np.sqrt(n) + diag * np.eye(n)
    return Lo @ Lo.T


def det_case(policy, log, n=None):
    def build(e):
        if n is None:
            A = kats.DET_FPLU if policy == "fullpivlu" else kats.DET_LLT
        else:
            A = (_vec(n * n, -1, 1).reshape(n, n) + 0.5 * np.eye(n)) if policy == "fullpivlu" else _spd(n)
        a = e.var(A)
        return e.det(a, policy, log), [a], kats.DET_SEED
    return build


def log_det_of_expression(n):
    def build(e):   # log det of a computed matrix, inside a larger scalar expression: 0.5*log_det(s*A + B) - sum(A*A)
        A, B, s = e.var(_spd(n)), e.const(_spd(n, 0.8)), e.var(1.3)
        M = e.binary("add", e.binary("mul", s, A), B)
        root = e.binary("sub", e.binary("mul", e.const(0.5), e.det(M, "llt", True)), e.sum(e.binary("mul", A, A)))
        return root, [A, s], 0.9
    return build


def wishart_case(p=None, const_v=False):
    def build(e):
        X, V, df = (kats.WISHART_X, kats.WISHART_V, kats.WISHART_N) if p is None else (_spd(p), _spd(p), p + 2.5)
        x = e.var(X)
        v = e.const(V) if const_v else e.var(V)
        return e.wishart(x, v, df), [x] + ([] if const_v else [v]), 1.0
    return build


def wishart_plus_det(p):
    def build(e):   # two dense stages sharing the device workspace, combined by a scalar stage
        x, v = e.var(_spd(p)), e.var(_spd(p))
        root = e.binary("add", e.wishart(x, v, p + 4), e.binary("mul", e.const(0.25), e.det(v, "fullpivlu", True)))
        return root, [x, v], 1.1
    return build


CASES = {
    "det_fplu": det_case("fullpivlu", False), "det_ldlt": det_case("ldlt", False), "det_llt": det_case("llt", False),
    "log_det_fplu": det_case("fullpivlu", True), "log_det_ldlt": det_case("ldlt", True), "log_det_llt": det_case("llt", True),
    "det_fplu_37": det_case("fullpivlu", False, 37), "log_det_llt_70": det_case("llt", True, 70),
    "log_det_of_expression": log_det_of_expression(45),
    "wishart": wishart_case(), "wishart_40": wishart_case(40), "wishart_const_v_33": wishart_case(33, True),
    "wishart_plus_det": wishart_plus_det(36),
    "dense_normal_vsm": dense_normal(False), "dense_normal_vvm": dense_normal(True),
    "dense_normal_vvm_100": dense_normal(True, n=100), "dense_normal_vsm_65_const_sigma": dense_normal(False, n=65, const_sigma=True),
    "dense_normal_in_model": dense_normal_in_model(48),
    "cauchy_sss": cauchy_case("sss"), "cauchy_vss": cauchy_case("vss"), "cauchy_vsv": cauchy_case("vsv"),
    "cauchy_vvs": cauchy_case("vvs"), "cauchy_vvv": cauchy_case("vvv"),
    "cauchy_vvv_big_seed": cauchy_case("vvv", seed=0.6, n=4099), "cauchy_vss_big": cauchy_case("vss", n=3001),
    "uniform_sss": uniform_case("sss"), "uniform_vss": uniform_case("vss"), "uniform_vsv": uniform_case("vsv"),
    "uniform_vvs": uniform_case("vvs"), "uniform_vvv": uniform_case("vvv"), "uniform_vvv_big": uniform_case("vvv", n=5003),
    "uniform_out_of_range": uniform_case("vvv", bad=True),
    "bernoulli_ss": bernoulli_case("ss"), "bernoulli_vs": bernoulli_case("vs"), "bernoulli_vv": bernoulli_case("vv"),
    "bernoulli_p_out_of_range": bernoulli_case("vs", p_scalar=1.341), "bernoulli_x_not_binary": bernoulli_case("vv", xs=(2.0, 0.0, 1.0)),
    "bernoulli_p_zero_all_x_zero": bernoulli_case("vs", xs=(0.0, 0.0, 0.0), p_scalar=0.0),
    "cauchy_in_model": cauchy_in_model(2500),
    "norm_vec": norm_vec,
    "norm_mat_chain": norm_mat_chain(500),
    "transpose_quadratic": transpose_quadratic,
    "transpose_of_expr": transpose_of_expr,
    "if_else_taken": if_else_scalar(True),
    "if_else_not_taken": if_else_scalar(False),
    "if_else_logical": if_else_logical,
    "if_else_nested_tt": if_else_nested(5003, 0.5, 0.7),
    "if_else_nested_tf": if_else_nested(5003, 0.5, -0.7),
    "if_else_nested_f": if_else_nested(4100, 1.5, 0.7),
    "opeq_scl_alias": opeq_scl_alias,
    "opeq_scl_noalias": opeq_scl_noalias,
    "opeq_vec_scl": opeq_vec_scl,
    "opeq_vec_noalias": opeq_vec_noalias,
    "opeq_chain": opeq_chain(777),
    "opeq_mul_then_sum": opeq_mul_then_sum(1001),
    "opeq_div_self_then_sum": opeq_div_self_then_sum(333),
    "opeq_then_dot": opeq_then_dot,
    "scalar_chain": scalar_chain,
    "placeholders": placeholders,
    "sumnode_benchmark": sumnode_benchmark,
    "prodnode_benchmark": prodnode_benchmark,
    "vec_sum_chain_5": vec_sum_chain(5),
    "vec_sum_chain_100k": vec_sum_chain(100003),
    "vec_nested_sum": vec_nested_sum(4099),
    "vec_scalar_times_sum": vec_scalar_times_sum(3001),
    "vec_prod": vec_prod(300),
    "vec_prod_one_zero": vec_prod(300, zeros=(17,)),
    "vec_prod_two_zeros": vec_prod(300, zeros=(17, 200)),
    "vec_prod_4099": vec_prod(4099),                               # 4 elements per thread in the specialised kernel
    "vec_prod_4099_one_zero": vec_prod(4099, zeros=(4097,)),
    "vec_root_vector_seed_4100": vec_root_with_vector_seed(4100),
    "vec_root_vector_seed": vec_root_with_vector_seed(2049),
    "vec_placeholder_chain": vec_placeholder_chain(1025),
    # >= 4096 elements with placeholder stores: the cp.async operand ring of the specialised kernel (several trips per thread)
    "placeholder_chain_ring": vec_placeholder_chain(333337),
    "placeholder_sum_ring": placeholder_sum(200003),
    "placeholder_reassigned_ring": placeholder_reassigned(170001),
    "placeholder_normal_ring": placeholder_normal(160001),
    "black_scholes_call_ring": black_scholes_vec(400003),
    "black_scholes_put_ring_4099": black_scholes_vec(4099, put=True),
    # several trips per thread of the four-element form (no placeholder stores), full trips + a partial last one
    "vec_sum_chain_1300003": vec_sum_chain(1300003),
    "vec_root_vector_seed_1250001": vec_root_with_vector_seed(1250001),
    "normal_expr_mean": normal_with_expr_mean(5001),
    "normal_vec_sigma_expr": normal_vec_sigma_expr(3000),
    "sum_of_normals": sum_of_normals(2000),
    "dot_norm_sum": dot_norm_sum,
    "quadratic_form": quadratic_form,
    "mlp3_row": mlp3_row,
    "linreg_expr": linreg_expr(2000, 64),
    "linreg_expr_small": linreg_expr(130, 5),
    "many_scalar_params": many_scalar_params(5003, 14),
    "long_chain_200_nodes": long_chain(4100, 40),
    "many_vector_leaves": many_vector_leaves(4097, 30),
    "dot_matrix_product_tiled": dot_matrix_product(150, 70, 97),
    "dot_matrix_product_small": dot_matrix_product(40, 70, 97),
    "dot_tall_3_rhs": dot_tall_multi_rhs(4500, 6, 3),
    "dot_tall_2_rhs_x_variable": dot_tall_multi_rhs(4100, 3, 2, x_var=True),
    "dot_wide_batch": dot_wide_batch(3, 2, 6000),
    "dot_wide_batch_16x8": dot_wide_batch(16, 8, 4099),
    "least_squares": least_squares(5000, 64),
    "least_squares_flipped_odd": least_squares(4099, 7, flipped=True),
    "least_squares_small": least_squares(100, 3),
    "least_squares_tall_generic": least_squares(6001, 19, inner=True),          # odd row count: gemv -> stage -> gemv^T
    "glm_least_squares_inner": least_squares(6000, 19, inner=True),            # even: the one-pass kernel (fast=6)
    "glm_logistic": glm_logistic(4160, 64),
    "posterior_linreg": posterior_linreg(4224, 16),
    "posterior_linreg_odd_rows": posterior_linreg(4099, 5),
    "glm_linreg_intercept": linreg_with_intercept(4224, 64),
    "glm_linreg_intercept_link_const_sigma": linreg_with_intercept(2050, 7, link=True, const_sigma=True),
    "glm_linreg_intercept_invalid_sigma": linreg_with_intercept(1024, 5, sigma=-0.7),
    "glm_logistic_weighted_ragged": glm_logistic(4290, 5, weights=True),        # last tile holds 2 of 64 rows
    "glm_logistic_tiny": glm_logistic(66, 3),
    "least_squares_x_variable": least_squares(4500, 5, x_var=True),
    "black_scholes_call": black_scholes_vec(2001),
    "black_scholes_put": black_scholes_vec(2001, put=True),
}
