"""Pins oracle/ (the CPU restatement) against the reference's own golden vectors.

CPU only.  Every expectation cites the reference test it is transcribed from.
"""
import math

import numpy as np
import pytest

import kats
import oracle_py as O
from kats import assert_double_eq


def _normal(kind, const_x=False, const_sigma=False):
    e = O.OracleExpr()
    xv = kats.N_VX if kind[0] == "v" else kats.N_X
    mv = kats.N_VMU if kind[1] == "v" else kats.N_MU
    sv = kats.N_VSIGMA if kind[2] == "v" else kats.N_SIGMA
    x = e.const(xv) if const_x else e.var(xv)
    mu = e.var(mv)
    s = e.const(sv) if const_sigma else e.var(sv)
    root = e.normal(x, mu, s)
    e.bind(root)
    return e, x, mu, s


@pytest.mark.parametrize("kind", ["sss", "vss", "vvs", "vsv", "vvv"])
def test_normal_goldens(kind):
    # test/reverse/stat/normal_unittest.cpp:105-297
    val, xa, ma, sa = kats.NORMAL_KAT[kind]
    e, x, mu, s = _normal(kind)
    assert_double_eq(e.feval(), [val], what=kind + " feval")
    e.beval(1.0)
    assert_double_eq(e.adj(x), xa, what=kind + " x_adj")
    assert_double_eq(e.adj(mu), ma, what=kind + " mu_adj")
    assert_double_eq(e.adj(s), sa, what=kind + " sigma_adj")


def test_normal_constant_operand_paths():
    # normal_unittest.cpp:151-172 (vss, x constant) and :237-257 (vsv, x and sigma constant)
    val, _, ma, sa = kats.NORMAL_KAT["vss"]
    e, x, mu, s = _normal("vss", const_x=True)
    assert_double_eq(e.feval(), [val])
    e.beval(1.0)
    assert_double_eq(e.adj(mu), ma)
    assert_double_eq(e.adj(s), sa)
    val, _, ma, _ = kats.NORMAL_KAT["vsv"]
    e, x, mu, s = _normal("vsv", const_x=True, const_sigma=True)
    assert_double_eq(e.feval(), [val])
    e.beval(1.0)
    assert_double_eq(e.adj(mu), ma)


@pytest.mark.parametrize("kind,bad", [("vsv", -1.0), ("vvv", 0.0)])
def test_normal_not_pos_def(kind, bad):
    # normal_unittest.cpp:207-213, 266-272
    e = O.OracleExpr()
    sig = kats.N_VSIGMA.copy()
    sig[0] = bad
    x = e.var(kats.N_VX)
    mu = e.var(kats.N_VMU if kind[1] == "v" else kats.N_MU)
    s = e.var(sig)
    e.bind(e.normal(x, mu, s))
    assert e.feval()[0] == -math.inf
    e.beval(1.0)
    assert np.all(e.adj(x) == 0)


def test_erf_kat():
    # test/reverse/core/unary_unittest.cpp:320-324
    x0, f, df = kats.ERF_KAT
    e = O.OracleExpr()
    x = e.var(x0)
    e.bind(e.unary("erf", x))
    assert_double_eq(e.feval(), [f])
    e.beval(3.14)
    assert_double_eq(e.adj(x), [3.14 * df])


def test_eval_mock_unary_accumulates():
    # test/reverse/core/eval_unittest.cpp:32-48
    e = O.OracleExpr()
    x = e.var(kats.FIX_SCL)
    e.bind(e.unary("mock2x", x))
    assert_double_eq(e.autodiff(), [4.62])
    assert_double_eq(e.adj(x), [2.0])
    e.beval(1.0)
    assert_double_eq(e.adj(x), [4.0])


def test_bind_placeholders():
    # test/reverse/core/bind_unittest.cpp:13-35
    e = O.OracleExpr()
    w1, w2, w3, w4 = e.var(1.0), e.var(2.0), e.var(0.0), e.var(0.0)
    root = e.glue(e.eq(w3, e.binary("mul", w1, w2)), e.eq(w4, e.binary("mul", w3, w3)))
    sizes = e.bind(root)
    assert sizes == (0, 0)   # both roots alias their placeholders (eq.hpp:141-145)
    for _ in range(2):
        assert_double_eq(e.autodiff(), [4.0])
        assert_double_eq(e.adj(w4), [1.0])
        assert_double_eq(e.adj(w3), [4.0])
        assert_double_eq(e.adj(w2), [4.0])
        assert_double_eq(e.adj(w1), [8.0])
        e.reset_adj()


def test_integration_chains():
    # test/integration/node_inttest.cpp:27-159
    e = O.OracleExpr()
    x = e.var(3.1)
    e.bind(e.unary("sin", e.unary("log", x)))
    assert_double_eq(e.autodiff(), [math.sin(math.log(3.1))])
    assert_double_eq(e.adj(x), [math.cos(math.log(3.1)) / 3.1])

    x1, x2, x3 = 1.5928, -0.291, 5.1023
    e = O.OracleExpr()
    l1, l2, l3 = e.var(x1), e.var(x2), e.var(x3)
    root = e.binary("sub",
                    e.binary("add", e.binary("mul", l1, l3),
                             e.binary("mul", e.unary("sin", e.unary("cos", e.binary("add", l1, l2))), l2)),
                    e.binary("div", l1, e.unary("exp", l3)))
    e.bind(root)
    assert_double_eq(e.feval(), [x1 * x3 + math.sin(math.cos(x1 + x2)) * x2 - x1 / math.exp(x3)])
    e.beval(1.0)
    assert_double_eq(e.adj(l1), [x3 - x2 * math.cos(math.cos(x1 + x2)) * math.sin(x1 + x2) - math.exp(-x3)])
    assert_double_eq(e.adj(l2), [math.sin(math.cos(x1 + x2)) - x2 * math.cos(math.cos(x1 + x2)) * math.sin(x1 + x2)])
    assert_double_eq(e.adj(l3), [x1 + x1 * math.exp(-x3)])


def test_sum_iter_reevaluate():
    # node_inttest.cpp:299-341: sum over 3 vars of cos(sin(v)*v), evaluated twice
    vals = [0.203104, 1.4231, -1.231]
    e = O.OracleExpr()
    vs = [e.var(v) for v in vals]
    root = e.sum_iter([e.unary("cos", e.binary("mul", e.unary("sin", v), v)) for v in vs])
    e.bind(root)
    want = sum(math.cos(math.sin(v) * v) for v in vals)
    for _ in range(2):
        assert_double_eq(e.autodiff(), [want])
        for v, nid in zip(vals, vs):
            d = -math.sin(math.sin(v) * v) * (math.cos(v) * v + math.sin(v))
            assert_double_eq(e.adj(nid), [d])
        e.reset_adj()


def test_mat_sum_of_squares():
    # node_inttest.cpp:380-400: sum(v*v) == 29 over a 3x2 matrix
    m = np.array([[1.0, 2.0], [-1.0, 3.0], [2.0, -3.0]])
    e = O.OracleExpr()
    v = e.var(m)
    e.bind(e.sum(e.binary("mul", v, v)))
    assert_double_eq(e.autodiff(), [float((m * m).sum())])
    assert_double_eq(e.adj(v), (2 * m).reshape(-1, order="F"))


def test_prod_with_zero_element():
    # test/reverse/core/prod_unittest.cpp (fixture vec has an exact zero, base_fixture.hpp:113)
    e = O.OracleExpr()
    v = e.var(kats.FIX_VEC)
    e.bind(e.prod(v))
    assert_double_eq(e.autodiff(2.0), [0.0])
    others = np.prod(np.delete(kats.FIX_VEC, 3))
    assert_double_eq(e.adj(v), [0, 0, 0, 2.0 * others, 0])


def test_pow_zero_guards():
    # pow.hpp:125-160 via test/reverse/core/pow_unittest.cpp
    e = O.OracleExpr()
    v = e.var(kats.FIX_VEC)
    e.bind(e.sum(e.pow(2, v)))
    assert_double_eq(e.autodiff(), [float((kats.FIX_VEC ** 2).sum())])
    assert_double_eq(e.adj(v), 2 * kats.FIX_VEC)
    e = O.OracleExpr()
    x = e.var(0.0)
    e.bind(e.pow(-2, x))
    assert e.autodiff()[0] == math.inf
    assert e.adj(x)[0] == -math.inf


def test_quadratic_form_readme():
    # README.md:582-607: x^T Sigma x, x=[0.5,0.6], Sigma=[[2,3],[3,6]] -> 4.46, adj [5.6,10.2]
    e = O.OracleExpr()
    x = e.var(np.array([[0.5], [0.6]]))      # 2x1 mat
    xt = e.var(np.array([[0.5, 0.6]]))       # 1x2 mat standing for transpose(x)
    S = e.const(np.array([[2.0, 3.0], [3.0, 6.0]]))
    e.bind(e.dot(e.dot(xt, S), x))
    assert_double_eq(e.autodiff(np.ones((1, 1))), [4.46])
    assert_double_eq(e.adj(x) + e.adj(xt), [5.6, 10.2])


def test_linreg_readme():
    # README.md:623-662, GenTestData.nb:216-221: loss 6655, adj [-1210, -12100]
    e = O.OracleExpr()
    theta = e.var(kats.LINREG_THETA.reshape(2, 1))
    loss = 0.0
    for i in range(5):
        g = O.OracleExpr()
        th = g.var(kats.LINREG_THETA.reshape(2, 1))
        xi = g.const(kats.LINREG_X[i].reshape(1, 2))
        yi = g.const(kats.LINREG_Y[i].reshape(1, 1))
        g.bind(g.pow(2, g.binary("sub", yi, g.dot(xi, th))))
        loss += g.autodiff(np.ones((1, 1)))[0]
        e.adj(theta)[:] += g.adj(th)
    assert_double_eq([loss], [kats.LINREG_LOSS])
    assert_double_eq(e.adj(theta), kats.LINREG_ADJ)


def test_linreg_workload_matches_row_loop():
    # the C4 workload (normal_adj_log_pdf(y, dot(X,w), sigma)) equals -0.5*loss/sigma^2 - n log sigma
    X = np.asfortranarray(kats.LINREG_X)
    w = kats.LINREG_THETA.copy()
    wa = np.zeros(2)
    sg = np.array([1.0])
    sa = np.zeros(1)
    v = O.linreg(X, kats.LINREG_Y.copy(), w, wa, sg, sa)
    assert_double_eq([v], [-0.5 * kats.LINREG_LOSS])
    assert_double_eq(wa, -0.5 * kats.LINREG_ADJ)
    assert_double_eq(sa, [kats.LINREG_LOSS - 5])


def test_mlp3_mathematica_goldens():
    # GenTestData.nb:691-700 (README.md:670-745 network)
    X = np.asfortranarray(kats.NN_X)
    grad = np.zeros(25)
    loss = O.mlp3(X, kats.NN_Y.copy(), kats.NN_PARM.copy(), grad)
    assert abs(loss - 92030.04305391687) <= 1e-12 * 92030.0
    assert abs(grad[0] - kats.NN_GRAD_A1_11) <= 1e-12 * abs(kats.NN_GRAD_A1_11)
    assert abs(grad[3] - kats.NN_GRAD_A1_12) <= 1e-12 * abs(kats.NN_GRAD_A1_12)
    g2 = np.zeros(25)
    loss2 = O.mlp3(X, kats.NN_Y.copy(), kats.NN_PARM.copy(), g2, nthreads=2)
    np.testing.assert_allclose(g2, grad, rtol=1e-13)
    assert abs(loss2 - loss) <= 1e-12 * abs(loss)


def test_black_scholes_example_point():
    # example/black_scholes.cpp:43-70; closed-form values from SURVEY.md §6
    p = kats.BS_POINT
    arr = lambda v: np.array([v, v])
    out = O.black_scholes(arr(p["S"]), arr(p["K"]), arr(p["sigma"]), arr(p["tau"]), arr(p["r"]))
    k = kats.BS_KAT
    want = [k["call"], k["put"], k["delta_call"], k["vega"], k["rho_call"], k["delta_put"],
            k["vega"], k["rho_put"]]
    for got, w in zip(out, want):
        assert abs(got[0] - w) <= 2e-14 * max(1.0, abs(w)) * 8, (got[0], w)
        assert got[0] == got[1]


def test_threaded_workloads_agree_with_single_thread():
    rng = np.random.default_rng(7)
    n = 1000
    x = rng.normal(size=n); mu = rng.normal(size=n); s = rng.uniform(0.5, 2, size=n)
    xa, ma, sa = np.zeros(n), np.zeros(n), np.zeros(n)
    v1 = O.normal_lpdf(x, mu, s, xa, ma, sa)
    xb, mb, sb = np.zeros(n), np.zeros(n), np.zeros(n)
    v2 = O.normal_lpdf(x, mu, s, xb, mb, sb, nthreads=3)
    assert abs(v1 - v2) <= 1e-12 * abs(v1)
    np.testing.assert_array_equal(xa, xb)
    # scipy cross-check of the value (constant -n/2 log 2pi dropped, normal.hpp:14-30)
    from scipy.stats import norm
    ref = norm.logpdf(x, mu, s).sum() + 0.5 * n * math.log(2 * math.pi)
    assert abs(v1 - ref) <= 1e-12 * abs(ref)
    xs = rng.uniform(0.5, 2, size=n)
    a1, a2 = np.zeros(n), np.zeros(n)
    s1 = O.sum_unary("exp", xs, a1)
    s2 = O.sum_unary("exp", xs, a2, nthreads=4)
    assert abs(s1 - s2) <= 1e-13 * abs(s1)
    np.testing.assert_array_equal(a1, a2)
    assert_double_eq(a1, np.exp(xs), ulps=1)


# ------------------------------------------------------------------ SURVEY.md 8f "next" rows
def test_norm_goldens():
    # test/reverse/core/norm_unittest.cpp:31-49: squaredNorm of MockUnary(vec), seed 9.2313
    e = O.OracleExpr()
    v = e.var(kats.FIX_VEC)
    e.bind(e.norm(e.unary("mock2x", v)))
    assert_double_eq(e.feval(), [float(((2 * kats.FIX_VEC) ** 2).sum())])
    e.beval(9.2313)
    assert_double_eq(e.adj(v), 2 * (9.2313 * 2 * (2 * kats.FIX_VEC)))


def test_opeq_goldens():
    # test/reverse/core/eq_unittest.cpp:143-240
    seed, x0 = 3.2, kats.FIX_SCL
    e = O.OracleExpr(); w = e.var(x0)
    e.bind(e.opeq("mul", w, w))
    assert_double_eq(e.feval(), [x0 * x0])
    e.beval(seed)
    assert_double_eq(e.adj(w), [seed * 2 * x0])
    assert_double_eq(e.value(w), [x0])                       # the previous value is restored (eq.hpp:255)
    e = O.OracleExpr(); w = e.var(x0)
    e.bind(e.opeq("div", w, e.unary("mock2x", w)))
    assert_double_eq(e.feval(), [x0 / (2 * x0)])
    e.beval(seed)
    assert abs(e.adj(w)[0]) <= 1e-16
    vseed = np.array([2.1, -0.3, 1.0, 0.5, 4.0])
    e = O.OracleExpr(); v = e.var(kats.FIX_VEC); s = e.var(x0)
    e.bind(e.opeq("mul", v, s))
    assert_double_eq(e.feval(), kats.FIX_VEC * x0)
    e.beval(vseed)
    assert_double_eq(e.adj(v), vseed * x0)
    assert_double_eq(e.adj(s), [float((vseed * kats.FIX_VEC).sum())])
    e = O.OracleExpr(); v = e.var(kats.FIX_VEC)
    e.bind(e.opeq("mul", v, e.unary("mock2x", v)))
    e.feval(); e.beval(vseed)
    assert_double_eq(e.adj(v), vseed * (2 * kats.FIX_VEC + 2 * kats.FIX_VEC))


def test_for_each_goldens():
    # test/reverse/core/for_each_unittest.cpp:33-46: for_each over {x += 1, x += 1} (ForEachIterNode, for_each.hpp:19-111,
    # is the glue of its sub-expressions): value x + 2, stored back into x; beval(seed) leaves x_adj = seed
    seed, x0 = 3.14, kats.FIX_SCL
    e = O.OracleExpr(); x = e.var(x0)
    e.bind(e.glue(e.opeq("add", x, e.const(1.0)), e.opeq("add", x, e.const(1.0))))
    res = e.feval()
    assert_double_eq(res, [x0 + 2.0])
    assert_double_eq(e.value(x), res)
    e.beval(seed)
    assert_double_eq(e.adj(x), [seed])


def test_if_else_and_comparisons():
    # test/reverse/core/if_else_unittest.cpp: only the taken branch receives the seed;
    # comparison nodes have no adjoint (binary.hpp:124)
    for a0, taken in ((1.5, True), (-0.5, False)):
        e = O.OracleExpr()
        a, b = e.var(a0), e.var(0.25)
        x = e.var(np.array([0.5, 1.5, 2.5]))
        root = e.sum(e.if_else(e.binary("lt", b, a), e.binary("mul", x, a), e.unary("exp", e.binary("mul", x, b))))
        e.bind(root)
        xv = np.array([0.5, 1.5, 2.5])
        want = (xv * a0).sum() if taken else np.exp(xv * 0.25).sum()
        assert_double_eq(e.autodiff(), [want])
        if taken:
            assert_double_eq(e.adj(a), [xv.sum()]); assert_double_eq(e.adj(b), [0.0]); assert_double_eq(e.adj(x), [a0] * 3)
        else:
            assert_double_eq(e.adj(a), [0.0]); assert_double_eq(e.adj(x), 0.25 * np.exp(xv * 0.25))


def test_transpose_quadratic_form_readme():
    # README.md:582-607 with ad::transpose: x^T Sigma x = 4.46, adj [5.6, 10.2]
    e = O.OracleExpr()
    x = e.var(np.array([[0.5], [0.6]]))
    S = e.const(np.array([[2.0, 3.0], [3.0, 6.0]]))
    e.bind(e.dot(e.dot(e.transpose(x), S), x))
    assert_double_eq(e.autodiff(np.ones((1, 1))), [4.46])
    assert_double_eq(e.adj(x), [5.6, 10.2])


def test_other_logpdf_goldens():
    # test/reverse/stat/cauchy_unittest.cpp:80-262 (values generated by ref/cauchy_ref.py there)
    import exprs

    def run(build):
        e = O.OracleExpr()
        root, vs, seed = build(e)
        e.bind(root)
        return e.autodiff(seed)[0], [e.adj(v).copy() for v in vs]
    v, (xa, la, sa) = run(exprs.cauchy_case("sss"))
    assert_double_eq([v], [-0.7584675254495730]); assert_double_eq(xa, [-0.0351507439654960])
    assert_double_eq(la, [0.0351507439654960]); assert_double_eq(sa, [-0.4677241747104879])
    v, (xa, la, sa) = run(exprs.cauchy_case("vss"))
    assert_double_eq([v], [-4.0831775617990589])
    assert_double_eq(xa, [-0.0695735121824751, 0.4534210702643782, 0.4122618701395337])
    assert_double_eq(la, [-0.7961094282214367]); assert_double_eq(sa, [-0.3601996841981470])
    v, (xa, la, sa) = run(exprs.cauchy_case("vvv"))
    assert_double_eq([v], [-6.867595467569049816])
    assert_double_eq(xa, [-0.740466493891151267, -1.999800000003999267, -0.229578571732138997])
    assert_double_eq(la, -np.array(xa))
    assert_double_eq(sa, [-1.815594805119381983, 99.980002000199945655, 0.183844689107000858])
    v, _ = run(exprs.cauchy_case("vvs")); assert_double_eq([v], [-4.959070146586642025])
    v, _ = run(exprs.cauchy_case("vsv")); assert_double_eq([v], [-6.9858076420069022])
    # test/reverse/stat/uniform_unittest.cpp:80-230
    v, (xa, lo, hi) = run(exprs.uniform_case("sss"))
    assert_double_eq([v], [-1.3256416139079406])
    assert_double_eq(lo, [0.7 / (0.5231 + 3.2415)]); assert_double_eq(hi, [-0.7 / (0.5231 + 3.2415)]); assert xa[0] == 0
    for kind, want in (("vss", -3.9769248417238217), ("vsv", -4.3915297872269807), ("vvs", -1.3266062626903801),
                       ("vvv", -1.2444888363909485)):
        v, _ = run(exprs.uniform_case(kind))
        assert_double_eq([v], [want], what="uniform " + kind)
    v, (xa, lo, hi) = run(exprs.uniform_case("vvv", bad=True))
    assert v == -np.inf and np.all(lo == 0) and np.all(hi == 0)
    # test/reverse/stat/bernoulli_unittest.cpp:60-200
    v, _ = run(exprs.bernoulli_case("vs")); assert_double_eq([v], [-18.42078074895269779176])
    v, (pa,) = run(exprs.bernoulli_case("vs", p_scalar=1.341)); assert v == -np.inf and pa[0] == 0
    v, (pa,) = run(exprs.bernoulli_case("vs", xs=(0.0, 0.0, 0.0), p_scalar=0.0)); assert v == 0 and pa[0] == 0
    v, (pa,) = run(exprs.bernoulli_case("vv", xs=(2.0, 0.0, 1.0))); assert v == -np.inf and np.all(pa == 0)
    v, (pa,) = run(exprs.bernoulli_case("vv"))
    assert_double_eq([v], [np.log(0.3) + np.log(1 - 0.42) + np.log(0.98)])
    assert_double_eq(pa, [1.3 / 0.3, -1.3 / (1 - 0.42), 1.3 / 0.98])


def test_dense_covariance_normal_goldens():
    # test/reverse/stat/normal_unittest.cpp:298-395 (vsm / vvm); only the LOWER triangle of Sigma is
    # set in the fixture (:94-102) and read by Eigen::LLT<Lower>
    Sig = np.array([[1.0, 0.0, 0.0], [0.3, 2.0, 0.0], [0.2, -0.3, 3.0]])
    e = O.OracleExpr()
    x, mu, S = e.var(kats.N_VX), e.var(kats.N_MU), e.var(Sig)
    e.bind(e.normal(x, mu, S))
    assert_double_eq(e.autodiff(), [-8.8105250497069019])
    assert_double_eq(e.adj(x), [-3.7624909485879803, 1.6010137581462704, -0.0890658942795076])
    assert_double_eq(e.adj(mu), [2.2505430847212176])
    sa = e.adj(S).reshape(3, 3, order="F")
    assert_double_eq([sa[0, 0], sa[1, 0], sa[1, 1], sa[2, 1], sa[2, 2]],
                     [6.5432306187049774, -2.9250063314004424, 1.0137007310866775, -0.1038829443345370, -0.1689156028253514])
    assert abs(sa[2, 0] - 0.2119067294266190) <= 1e-15
    e = O.OracleExpr()
    x, mu, S = e.var(kats.N_VX), e.var(kats.N_VMU), e.var(Sig)
    e.bind(e.normal(x, mu, S))
    assert_double_eq(e.autodiff(), [-7.3649692930088602])
    assert_double_eq(e.adj(x), [-3.4158218682114407, 0.4279507603186097, -0.5628167994207096])
    assert_double_eq(e.adj(mu), -e.adj(x))
    sa = e.adj(S).reshape(3, 3, order="F")
    assert_double_eq([sa[0, 0], sa[1, 0], sa[2, 0], sa[0, 1], sa[1, 1], sa[2, 1]],
                     [5.2989810672774862, -0.6440082274123684, 1.0055928845283644, -0.6440082274123684,
                      -0.1763508691715067, -0.1530140218890801])
    e = O.OracleExpr()
    x, mu, S = e.var(kats.N_VX), e.var(kats.N_VMU), e.var(np.zeros((3, 3)))
    e.bind(e.normal(x, mu, S))
    assert e.autodiff()[0] == -np.inf and np.all(e.adj(x) == 0)       # normal_unittest.cpp:304-310


@pytest.mark.parametrize("log", [False, True])
@pytest.mark.parametrize("policy", ["fullpivlu", "ldlt", "llt"])
def test_det_log_det_goldens(policy, log):
    # test/reverse/core/det_unittest.cpp:57-155 and log_det_unittest.cpp:57-152: value == determinant()
    # (EXPECT_DOUBLE_EQ; the fixtures are integer matrices, exact determinants 153 and 65), adjoint ==
    # seed * det * inverse().transpose() (resp. seed * inverse().transpose()) within 3e-13
    A = kats.DET_FPLU if policy == "fullpivlu" else kats.DET_LLT
    exact = 153.0 if policy == "fullpivlu" else 65.0
    e = O.OracleExpr()
    a = e.var(A)
    sizes = e.bind(e.det(a, policy, log))
    assert sizes == (1, 0)                                           # det.hpp:75-85
    v = e.autodiff(kats.DET_SEED)[0]
    assert_double_eq([v], [math.log(exact) if log else exact])
    want = kats.DET_SEED * (1.0 if log else exact) * np.linalg.inv(A).T
    assert np.abs(e.adj(a).reshape(4, 4, order="F") - want).max() <= 3e-13
    # constant operand folds to a constant (det.hpp:222-227)
    e = O.OracleExpr()
    c = e.det(e.const(A), policy, log)
    assert e.L.orc_is_const(e.g, c)
    e.bind(c)
    assert_double_eq(e.feval(), [math.log(exact) if log else exact])


def test_det_invalid_matrix_skips_adjoint():
    # det.hpp:58-60: beval returns early when the decomposition is not valid (singular / not PD)
    for policy, A in (("fullpivlu", np.array([[1., 2], [2, 4]])), ("llt", np.array([[1., 2], [2, 1]])),
                      ("ldlt", np.zeros((2, 2)))):
        e = O.OracleExpr()
        a = e.var(A)
        e.bind(e.det(a, policy))
        e.autodiff(1.0)
        assert np.all(e.adj(a) == 0), policy


def test_wishart_goldens():
    # test/reverse/stat/wishart_unittest.cpp:62-106
    e = O.OracleExpr()
    x, v = e.var(kats.WISHART_X), e.var(kats.WISHART_V)
    assert e.bind(e.wishart(x, v, kats.WISHART_N)) == (1, 0)
    assert_double_eq(e.autodiff(1.0), [kats.WISHART_VALUE])
    n, p = kats.WISHART_N, 3
    vi = np.linalg.inv(kats.WISHART_V)
    dX = 0.5 * ((n - p - 1) * np.linalg.inv(kats.WISHART_X) - vi)
    dV = 0.5 * (vi @ kats.WISHART_X @ vi - n * vi)
    assert np.abs(e.adj(x).reshape(3, 3, order="F") - dX).max() <= 1e-15
    assert np.abs(e.adj(v).reshape(3, 3, order="F") - dV).max() <= 1e-15
    # feval_invalid (:69-75): V = 0 is not positive definite
    e = O.OracleExpr()
    x, v = e.var(kats.WISHART_X), e.var(np.zeros((3, 3)))
    e.bind(e.wishart(x, v, kats.WISHART_N))
    assert e.autodiff(1.0)[0] == -np.inf and np.all(e.adj(x) == 0) and np.all(e.adj(v) == 0)
    # n + 1 > p is required (wishart.hpp:203-207)
    e = O.OracleExpr()
    x, v = e.var(kats.WISHART_X), e.var(kats.WISHART_V)
    e.bind(e.wishart(x, v, 2))
    assert e.autodiff(1.0)[0] == -np.inf
