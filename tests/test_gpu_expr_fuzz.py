"""Seeded random expressions through the whole generic path (builder -> planner -> fast-path routing -> specialised /
interpreted stage kernels, dot stages, reductions, placeholders, if_else) against the CPU oracle.

The hand-picked cases of tests/exprs.py pin the shapes the reference's own tests use; this file adds breadth: every
seed builds a different DAG out of the same node vocabulary -- shared sub-expressions included (the reference copies a
sub-expression into each use, the planner shares producer stages inside a consumer) -- on the oracle and on the GPU,
and compares value and every adjoint.  Operand ranges keep every function well inside its domain, so differences are
rounding only: 1e-11 relative to the largest magnitude of the compared array (reordered reductions, <= 4 ulp
transcendentals), the tolerance of tests/test_gpu_expr.py."""
import numpy as np
import pytest

import oracle_py as O

pytestmark = pytest.mark.gpu

UNARY_SAFE = ("sin", "cos", "tanh", "sigmoid", "exp", "atan", "neg", "erf", "sinh")   # any real argument of modest size
UNARY_POS = ("log", "sqrt")                                                        # positive argument


@pytest.fixture(scope="module")
def fb():
    import torch
    assert torch.cuda.is_available()
    import fastad_b200 as F
    F.load()
    F.check(F.lib().fadb_init(0))
    F.use_torch_stream()
    return F


class Gen:
    """Builds the same random expression on any backend: all random draws come from one seeded generator, in an
    order that does not depend on the backend."""

    def __init__(self, seed, n):
        self.rng = np.random.default_rng(seed)
        self.n = n

    def build(self, e):
        rng, n = self.rng, self.n
        nv, ns = int(rng.integers(2, 4)), int(rng.integers(1, 3))
        vec_vars = [e.var(rng.uniform(0.6, 1.6, n)) for _ in range(nv)]
        scl_vars = [e.var(float(rng.uniform(0.7, 1.4))) for _ in range(ns)]
        consts = [e.const(rng.uniform(0.5, 1.5, n))]
        pool = list(vec_vars) + consts               # vector-valued, values in a modest positive range
        spool = list(scl_vars)                       # scalar-valued

        def bounded(x):      # map anything into (0.5, 1.5): keeps later log / sqrt / div arguments positive and tame
            return e.binary("add", e.const(1.0), e.binary("mul", e.const(0.5), e.unary("tanh", x)))

        def operand():
            if rng.random() < 0.25:
                return spool[int(rng.integers(len(spool)))]
            return pool[int(rng.integers(len(pool)))]        # may pick the SAME node twice: shared sub-expression

        for _ in range(int(rng.integers(3, 9))):
            kind = rng.random()
            if kind < 0.35:
                op = UNARY_SAFE[int(rng.integers(len(UNARY_SAFE)))]
                node = bounded(e.unary(op, pool[int(rng.integers(len(pool)))]))
            elif kind < 0.45:
                op = UNARY_POS[int(rng.integers(len(UNARY_POS)))]
                node = bounded(e.unary(op, pool[int(rng.integers(len(pool)))]))
            elif kind < 0.85:
                op = ("add", "sub", "mul", "div")[int(rng.integers(4))]
                a, b = operand(), pool[int(rng.integers(len(pool)))]
                node = bounded(e.binary(op, a, b))
            else:
                node = bounded(e.pow(int(rng.integers(2, 4)), pool[int(rng.integers(len(pool)))]))
            pool.append(node)
        body = pool[-1]
        variables = vec_vars + scl_vars
        shape = rng.random()
        if shape < 0.3:
            # a dot stage feeding the map: X constant (the one-pass route when the row count allows) or built from a map
            cols = int(rng.integers(2, 9))
            X = np.asfortranarray(rng.normal(size=(n, cols)) / np.sqrt(cols))
            w = e.var(rng.normal(size=(cols, 1)) * 0.3)
            variables.append(w)
            eta = e.dot(e.const(X), w)
            if shape < 0.18:     # only constant row operands beside the product: the one-pass route (fast=6) when n is even
                body = bounded(e.unary(UNARY_SAFE[int(rng.integers(len(UNARY_SAFE)))], consts[0]))
            lin = e.binary("add", eta, spool[0])
            body = e.binary("mul", body, e.binary("add", e.unary("tanh", lin), e.binary("mul", e.const(0.1), lin)))   # eta used twice
        elif shape < 0.45:
            # nested reduction: a scalar produced by one stage, consumed (twice) by the next
            s = e.sum(pool[int(rng.integers(len(pool)))])
            sc = e.binary("div", s, e.const(float(n)))
            body = e.binary("add", e.binary("mul", body, sc), e.unary("sin", sc))
        elif shape < 0.6:
            cond = e.binary("lt", spool[0], e.const(1.05))
            body = e.if_else(cond, body, e.binary("mul", body, pool[0]))
        root_kind = rng.random()
        if root_kind < 0.6:
            root, seed = e.sum(body), float(rng.uniform(0.5, 1.5))
        elif root_kind < 0.8:
            root, seed = e.normal(e.const(rng.uniform(0.5, 1.5, n)), body, spool[-1]), float(rng.uniform(0.5, 1.5))
        elif root_kind < 0.9:
            ph = e.var(np.zeros(n))
            variables.append(ph)
            root, seed = e.glue(e.eq(ph, body), e.sum(e.binary("mul", ph, ph))), float(rng.uniform(0.5, 1.5))
        else:
            # a posterior: likelihood-like term + prior terms, combined by +, - and constants (the additive-root execution)
            lik = e.sum(body)
            prior = e.normal(vec_vars[0], e.const(1.0), e.const(2.0))
            pen = e.binary("mul", e.const(-0.5), e.binary("mul", spool[0], spool[0]))
            root = e.binary("add", e.binary("sub", e.binary("mul", e.const(0.25), lik), e.unary("neg", prior)), pen)
            seed = float(rng.uniform(0.5, 1.5))
        return root, variables, seed


def _run(backend, seed, n):
    e = backend()
    root, variables, s = Gen(seed, n).build(e)
    e.bind(root)
    v = e.autodiff(s)
    return np.asarray(v).reshape(-1).copy(), [np.asarray(e.adj(x)).reshape(-1).copy() for x in variables], e


def _close(a, b, what):
    scale = max(float(np.max(np.abs(b))) if b.size else 0.0, 1e-300)
    np.testing.assert_allclose(a, b, rtol=1e-11, atol=1e-11 * scale, err_msg=what)


@pytest.mark.parametrize("engine", ["specialised", "interpreter"])
@pytest.mark.parametrize("seed", range(60))
def test_random_expression_matches_oracle(fb, engine, seed):
    n = (4100, 66, 4099, 128)[seed % 4]          # vector form / tiny / odd (no one-pass route, scalar tail) / even small
    prev = fb.jit_enable(engine == "specialised")
    try:
        try:
            gv, gadj, ge = _run(fb.Expr, seed, n)
        except fb.FadbError as ex:
            if engine == "interpreter" and "exceeds the interpreter" in str(ex):
                pytest.skip("beyond the interpreter's fixed tables (fails loudly, by design)")
            raise
        assert fb.jit_stats()["failed"] == 0 or engine == "interpreter", fb.lib().fadb_jit_diagnostics().decode()
    finally:
        fb.jit_enable(prev)
    ov, oadj, _ = _run(O.OracleExpr, seed, n)
    _close(gv, ov, f"seed {seed} value\n{ge.describe()}")
    for k, (g, o) in enumerate(zip(gadj, oadj)):
        _close(g, o, f"seed {seed} adj[{k}]\n{ge.describe()}")


class PlaceholderGen:
    """Random programs of the reference's placeholder idiom over vectors (`c_k = f_k(leaves, c_1 .. c_{k-1})`, glued, then a
    root that reads the placeholders): the stages the operand ring of the specialised kernel serves (stage_kernel.cuh, JIT_RING).
    Each placeholder is assigned once."""

    def __init__(self, seed, n):
        self.rng = np.random.default_rng(1000 + seed)
        self.n = n

    def build(self, e):
        rng, n = self.rng, self.n
        vec_vars = [e.var(rng.uniform(0.6, 1.6, n)) for _ in range(int(rng.integers(1, 4)))]
        scl_vars = [e.var(float(rng.uniform(0.7, 1.4)))]
        consts = [e.const(rng.uniform(0.5, 1.5, n)) for _ in range(int(rng.integers(1, 3)))]
        pool = list(vec_vars) + consts

        def bounded(x):
            return e.binary("add", e.const(1.0), e.binary("mul", e.const(0.5), e.unary("tanh", x)))

        def pick():
            return pool[int(rng.integers(len(pool)))]

        def expr():
            kind = rng.random()
            if kind < 0.4:
                return bounded(e.unary(UNARY_SAFE[int(rng.integers(len(UNARY_SAFE)))], pick()))
            if kind < 0.5:
                return bounded(e.unary(UNARY_POS[int(rng.integers(len(UNARY_POS)))], pick()))
            op = ("add", "sub", "mul", "div")[int(rng.integers(4))]
            a = scl_vars[0] if rng.random() < 0.2 else pick()
            return bounded(e.binary(op, a, pick()))

        holders, assigns = [], []
        for _ in range(int(rng.integers(1, 4))):
            ph = e.var(np.zeros(n))
            assigns.append(e.eq(ph, expr()))
            holders.append(ph)
            pool.append(ph)                       # later expressions (and the root) read the placeholder
        body = e.binary("sub", e.binary("mul", holders[-1], pick()), e.binary("mul", e.const(0.3), holders[0]))
        variables = vec_vars + scl_vars + holders
        if rng.random() < 0.5:
            return e.glue(*assigns, body), variables, rng.uniform(-1, 1, n)          # vector-valued root, per-element seed
        return e.glue(*assigns, e.sum(body)), variables, float(rng.uniform(0.5, 1.5))


@pytest.mark.parametrize("seed", range(12))
def test_random_placeholder_program_on_the_operand_ring(fb, seed):
    # several trips per thread (the specialised kernel's grid holds 151 552 threads), an odd element count; against the oracle,
    # and bit for bit against the plain kernel of the same program
    n = 320001 if seed % 3 else 4099
    res = {}
    for ring in (1, 0):
        prev = fb.lib().fadb_jit_ring_enable(ring)
        try:
            e = fb.Expr()
            root, variables, s = PlaceholderGen(seed, n).build(e)
            e.bind(root)
            j0 = fb.jit_stats()
            v = e.autodiff(s)
            assert fb.jit_stats()["failed"] == j0["failed"], fb.lib().fadb_jit_diagnostics().decode()
            res[ring] = (np.asarray(v).reshape(-1).copy(), [np.asarray(e.adj(x)).reshape(-1).copy() for x in variables],
                         [np.asarray(e.value(x)).reshape(-1).copy() for x in variables], e.describe())
        finally:
            fb.lib().fadb_jit_ring_enable(prev)
    oe = O.OracleExpr()
    root, variables, s = PlaceholderGen(seed, n).build(oe)
    oe.bind(root)
    ov = np.asarray(oe.autodiff(s)).reshape(-1)
    _close(res[1][0], ov, f"seed {seed} value\n{res[1][3]}")
    for k, x in enumerate(variables):
        _close(res[1][1][k], np.asarray(oe.adj(x)).reshape(-1), f"seed {seed} adj[{k}]\n{res[1][3]}")
        _close(res[1][2][k], np.asarray(oe.value(x)).reshape(-1), f"seed {seed} val[{k}]\n{res[1][3]}")
    np.testing.assert_array_equal(res[1][0], res[0][0])
    for a, b in zip(res[1][1] + res[1][2], res[0][1] + res[0][2]):
        np.testing.assert_array_equal(a, b)
