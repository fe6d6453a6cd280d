"""GPU parity of the generic expression path (builder -> planner -> stage kernels, and the fast-path
dispatch) against the CPU oracle, on the same seeded inputs.  Every case runs twice: on the
bind-time specialised kernels (NVRTC, csrc/jit.cu -- the default) and on the tape interpreter."""
import numpy as np
import pytest

import exprs
import oracle_py as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fb():
    import torch
    assert torch.cuda.is_available()
    import fastad_b200 as F
    F.load()
    F.check(F.lib().fadb_init(0))
    F.use_torch_stream()
    return F


def _run(build, backend, twice=False):
    exprs.RNG = np.random.default_rng(20261019)    # identical inputs for both backends
    e = backend()
    root, variables, seed = build(e)
    e.bind(root)
    val = e.autodiff(seed)
    if twice:                                    # adjoints accumulate (eval_unittest.cpp:37-42)
        val = e.autodiff(seed)
    return val, [e.adj(v).copy() for v in variables], [e.value(v).copy() for v in variables], e


def _close(a, b, rtol, what, floor=1e-300):
    a, b = np.asarray(a), np.asarray(b)
    scale = np.max(np.abs(b)) if b.size else 0.0
    np.testing.assert_allclose(a, b, rtol=rtol, atol=rtol * max(scale, floor), err_msg=what)


@pytest.fixture(params=["specialised", "interpreter"])
def engine(fb, request):
    prev = fb.jit_enable(request.param == "specialised")
    yield request.param
    fb.jit_enable(prev)


@pytest.mark.parametrize("name", sorted(set(exprs.CASES) - exprs.NO_ORACLE))
def test_expr_matches_oracle(fb, engine, name):
    build = exprs.CASES[name]
    if engine == "interpreter" and name in exprs.JIT_ONLY:
        # beyond the interpreter's fixed tables: must fail loudly, not silently truncate
        with pytest.raises(fb.FadbError, match="exceeds the interpreter"):
            _run(build, fb.Expr)
        return
    ov, oadj, oval, _ = _run(build, O.OracleExpr)
    before = fb.jit_stats()
    gv, gadj, gval, ge = _run(build, fb.Expr)
    after = fb.jit_stats()
    assert after["failed"] == before["failed"], fb.lib().fadb_jit_diagnostics().decode()
    if engine == "interpreter":
        assert after["launched"] == before["launched"]
    dense = name.startswith(("dense_normal", "det_", "log_det", "wishart"))
    rtol = 1e-11 if name.startswith("black_scholes") or dense else 1e-12
    _close(gv, ov, rtol, name + " value")
    # Black-Scholes: the adjoints that flow through the placeholders are differences of nearly
    # equal terms of size ~S*phi(d1) (the residual S*phi(d1) - PV*phi(d2) is ~1e-14): the floor is
    # relative to those terms, not to the residual.
    floor = 100.0 if name.startswith("black_scholes") else 1e-300
    if dense:                                # factorisation + inverse: conditioning-limited, not ulp-limited
        for k, (g, o) in enumerate(zip(gadj, oadj)):
            _close(g, o, 1e-8, f"{name} adj[{k}]")
        return
    for k, (g, o) in enumerate(zip(gadj, oadj)):
        _close(g, o, 1e-10 if name.startswith(("black_scholes", "mlp3")) else 1e-11, f"{name} adj[{k}]", floor)
    for k, (g, o) in enumerate(zip(gval, oval)):          # placeholders hold their values
        _close(g, o, 1e-12, f"{name} val[{k}]")


@pytest.mark.parametrize("name", ["placeholders", "vec_sum_chain_100k", "normal_expr_mean", "dot_norm_sum"])
def test_expr_adjoints_accumulate(fb, name):
    build = exprs.CASES[name]
    _, oadj, _, _ = _run(build, O.OracleExpr, twice=True)
    _, gadj, _, _ = _run(build, fb.Expr, twice=True)
    for g, o in zip(gadj, oadj):
        _close(g, o, 1e-11, name)


@pytest.mark.parametrize("name", ["scalar_chain", "vec_nested_sum", "vec_prod_one_zero", "normal_vec_sigma_expr",
                                  "opeq_scl_alias", "opeq_scl_noalias", "opeq_vec_scl", "opeq_vec_noalias", "opeq_chain",
                                  "opeq_mul_then_sum", "opeq_div_self_then_sum", "opeq_then_dot"])
def test_expr_feval_then_beval_equals_autodiff(fb, engine, name):
    # ad::evaluate followed by ad::evaluate_adj (eval.hpp:18-73): forward-only and backward-only launches.  op= stages
    # overwrite w in the forward launch; the backward launch must still see the previous value (eq.hpp:240-271)
    build = exprs.CASES[name]
    exprs.RNG = np.random.default_rng(20261019)
    oe = O.OracleExpr()
    oroot, ovars, oseed = build(oe)
    oe.bind(oroot)
    ov = oe.autodiff(oseed)
    exprs.RNG = np.random.default_rng(20261019)
    e = fb.Expr()
    root, variables, seed = build(e)
    e.bind(root)
    v1 = e.feval()
    e.beval(seed)
    a1 = [e.adj(v).copy() for v in variables]
    x1 = [e.value(v).copy() for v in variables]
    np.testing.assert_allclose(v1, ov, rtol=1e-12)
    for k, (g, o) in enumerate(zip(a1, [oe.adj(v) for v in ovars])):
        _close(g, o, 1e-11, f"{name} adj[{k}] after feval + beval")
    for k, (g, o) in enumerate(zip(x1, [oe.value(v) for v in ovars])):      # op= restores the previous value
        _close(g, o, 1e-12, f"{name} val[{k}] after feval + beval")
    e.reset_adj()
    v2 = e.autodiff(seed)
    a2 = [e.adj(v).copy() for v in variables]
    np.testing.assert_array_equal(v1, v2)
    for x, y in zip(a1, a2):
        np.testing.assert_allclose(x, y, rtol=1e-14, atol=0)


def test_logpdf_root_seed_zero_leaves_producers_untouched(fb, engine):
    # autodiff(1.0) then autodiff(0.0) on normal(y, dot(X, w), sigma_vec): the second call must not
    # accumulate the first call's boundary adjoints again (normal.hpp:160: beval(0) returns before any sweep)
    rng = np.random.default_rng(3)
    e = fb.Expr()
    X, y = e.const(rng.normal(size=(40, 3))), e.const(rng.normal(size=40))
    w, ls = e.var(rng.normal(size=(3, 1))), e.var(rng.uniform(-0.3, 0.3, 40))
    e.bind(e.normal(y, e.dot(X, w), e.unary("exp", ls)))
    v1 = e.autodiff(1.0)
    a1 = [e.adj(w).copy(), e.adj(ls).copy()]
    v2 = e.autodiff(0.0)
    np.testing.assert_array_equal(v1, v2)
    np.testing.assert_array_equal(e.adj(w), a1[0])
    np.testing.assert_array_equal(e.adj(ls), a1[1])
    e.beval(0.0)
    np.testing.assert_array_equal(e.adj(w), a1[0])


@pytest.mark.parametrize("name", ["glm_logistic", "glm_logistic_weighted_ragged", "glm_logistic_tiny", "glm_least_squares_inner",
                                  "glm_linreg_intercept", "glm_linreg_intercept_link_const_sigma"])
def test_one_pass_dot_map_sum_route(fb, name):
    # sum(f(dot(X, w), row constants, scalars)): ONE specialised launch (glm_kernel.cuh) + the scalar epilogue instead of
    # gemv -> stage -> gemv^T; value and adjoints against the oracle AND against the three-launch route of the same plan
    build = exprs.CASES[name]
    ov, oadj, _, _ = _run(build, O.OracleExpr)
    res = {}
    for on in (True, False):
        prev = fb.lib().fadb_glm_enable(1 if on else 0)
        try:
            exprs.RNG = np.random.default_rng(20261019)
            e = fb.Expr()
            root, variables, seed = build(e)
            e.bind(root)
            assert "fast=6" in e.describe()
            l0, j0 = fb.lib().fadb_launch_count(), fb.jit_stats()
            v = e.autodiff(seed)
            l1, j1 = fb.lib().fadb_launch_count(), fb.jit_stats()
            v2 = e.autodiff(seed)                    # adjoints accumulate; same kernel again
            res[on] = (v, [e.adj(x).copy() for x in variables], l1 - l0, j1["launched"] - j0["launched"])
            np.testing.assert_array_equal(v, v2)
        finally:
            fb.lib().fadb_glm_enable(prev)
    assert res[True][2] == 2 and res[True][3] == 1, res[True][2:]          # the tile kernel + glm_finish_kernel
    assert res[False][2] >= 3
    _close(res[True][0], ov, 1e-12, name + " value")
    _close(res[True][0], res[False][0], 1e-12, name + " value vs three-launch route")
    for k, (g, o, h) in enumerate(zip(res[True][1], oadj, res[False][1])):
        _close(g, 2.0 * np.asarray(o), 1e-11, f"{name} adj[{k}] (two sweeps)")
        _close(g, h, 1e-11, f"{name} adj[{k}] vs three-launch route")


def test_additive_root_terms_run_as_roots(fb):
    # log-likelihood + log-priors: each term is swept once (fused, seed known up front), the likelihood on the one-pass tile
    # kernel; value and adjoints against the oracle come from test_expr_matches_oracle[posterior_linreg], here the launches
    exprs.RNG = np.random.default_rng(20261019)
    e = fb.Expr()
    root, variables, seed = exprs.CASES["posterior_linreg"](e)
    e.bind(root)
    d = e.describe()
    assert "additive root: 3 term(s)" in d and "fast=6" in d
    e.autodiff(seed)                                     # first call compiles
    l0 = fb.lib().fadb_launch_count()
    v = e.autodiff(seed)
    # tile kernel + its epilogue, two prior stages (one fused launch each), the root stage
    assert fb.lib().fadb_launch_count() - l0 == 5, fb.lib().fadb_launch_count() - l0
    exprs.RNG = np.random.default_rng(20261019)
    oe = O.OracleExpr()
    oroot, ovars, oseed = exprs.CASES["posterior_linreg"](oe)
    oe.bind(oroot)
    ov = oe.autodiff(oseed)
    _close(v, ov, 1e-12, "posterior value")
    for g, o in zip(variables, ovars):
        _close(e.adj(g), 2.0 * np.asarray(oe.adj(o)), 1e-11, "posterior adjoints after two sweeps")


def test_one_pass_route_invalid_sigma(fb):
    # normal(y, dot(X, w) + b, sigma) with sigma <= 0 through the one-pass kernel: -inf and NO adjoint (normal.hpp:344)
    exprs.RNG = np.random.default_rng(20261019)
    e = fb.Expr()
    root, variables, seed = exprs.CASES["glm_linreg_intercept_invalid_sigma"](e)
    e.bind(root)
    assert "fast=6" in e.describe()
    j0 = fb.jit_stats()
    v = e.autodiff(seed)
    assert fb.jit_stats()["launched"] - j0["launched"] == 1
    assert v[0] == -np.inf
    for x in variables:
        assert np.all(e.adj(x) == 0.0)


def test_one_pass_route_falls_back_on_unaligned_operands(fb):
    # X offset by one double (a view into a larger buffer): the tile path needs 16-byte alignment, the call must take the
    # three-launch route and still be right
    import torch
    import ctypes as C
    rng = np.random.default_rng(11)
    rows, cols = 1024, 6
    X = rng.normal(size=(rows, cols)); y = rng.normal(size=rows); w0 = rng.normal(size=cols)
    buf = torch.zeros(rows * cols + 1, dtype=torch.float64, device="cuda")
    buf[1:] = torch.from_numpy(np.asfortranarray(X).reshape(-1, order="F")).cuda()
    e = fb.Expr()
    e.keep += [buf]
    Xc = e._id(fb.lib().fadb_expr_const_view(e.h, fb.MAT, rows, cols, C.c_void_p(buf.data_ptr() + 8)))
    w = e.var(w0.reshape(cols, 1))
    r = e.unary("tanh", e.binary("sub", e.const(y), e.dot(Xc, w)))
    e.bind(e.sum(e.pow(2, r)))
    assert "fast=6" in e.describe()
    j0 = fb.jit_stats()
    v = e.autodiff(1.0)
    t = np.tanh(y - X @ w0)
    np.testing.assert_allclose(v[0], np.sum(t * t), rtol=1e-12)
    np.testing.assert_allclose(e.adj(w).reshape(-1), -X.T @ (2 * t * (1 - t * t)), rtol=1e-11, atol=1e-12)


def test_expr_specialised_kernels_are_the_default_and_bit_identical_to_interpreter(fb):
    # same source, same flags (-fmad=false): the specialised kernel and the interpreter must agree to the bit
    res = {}
    for on in (True, False):
        prev = fb.jit_enable(on)
        try:
            s0 = fb.jit_stats()
            _, adj, val, e = _run(exprs.CASES["vec_sum_chain_100k"], fb.Expr)
            v = e.autodiff(0.0)                      # second sweep with seed 0: cache hit, adjoints unchanged
            s1 = fb.jit_stats()
            res[on] = (v, adj, val, s1["launched"] - s0["launched"], s1["compiled"] + s1["hits"] - s0["compiled"] - s0["hits"])
        finally:
            fb.jit_enable(prev)
    assert res[True][3] == 2 and res[True][4] >= 1, (res[True][3:], fb.lib().fadb_jit_diagnostics().decode())
    assert res[False][3] == 0
    # per-element adjoints: identical bits; reduced quantities (the value, the two scalar adjoints): the
    # vector path hands elements to threads in groups of 4, so the summation order differs
    np.testing.assert_allclose(res[True][0], res[False][0], rtol=1e-13)
    for a, b in zip(res[True][1], res[False][1]):
        if a.size > 1:
            np.testing.assert_array_equal(a, b)
        else:
            np.testing.assert_allclose(a, b, rtol=1e-12)


def test_expr_specialised_vector_path_unaligned_and_ragged(fb):
    # 4 elements per thread with 256-bit accesses needs 32-byte aligned operands: an operand offset by one
    # double (a VarView into the middle of a buffer) and sizes that are not multiples of 4 must give
    # the same adjoint bits as the interpreter
    import torch
    import ctypes as C
    rng = np.random.default_rng(7)
    for n, off in ((4096, 0), (4099, 0), (5001, 1), (8192, 3)):
        xs = rng.uniform(0.5, 2.0, n + 4); cs = rng.uniform(0.5, 2.0, n + 4)
        out = {}
        for on in (True, False):
            prev = fb.jit_enable(on)
            try:
                xd = torch.from_numpy(xs).cuda(); xa = torch.zeros(n + 4, dtype=torch.float64, device="cuda")
                cd = torch.from_numpy(cs).cuda()
                e = fb.Expr()
                e.keep += [xd, xa, cd]
                x = e._id(fb.lib().fadb_expr_var(e.h, fb.VEC, n, 1, C.c_void_p(xd.data_ptr() + 8 * off), C.c_void_p(xa.data_ptr() + 8 * off)))
                c = e._id(fb.lib().fadb_expr_const_view(e.h, fb.VEC, n, 1, C.c_void_p(cd.data_ptr() + 8 * off)))
                w = e.var(1.7)
                e.bind(e.sum(e.binary("mul", e.unary("log", e.binary("div", x, c)), e.binary("mul", w, x))))
                s0 = fb.jit_stats()
                v = e.autodiff(0.5)
                used = fb.jit_stats()["launched"] - s0["launched"]
                out[on] = (v, xa.cpu().numpy().copy(), e.adj(w).copy(), used)
            finally:
                fb.jit_enable(prev)
        assert out[True][3] == 1 and out[False][3] == 0
        np.testing.assert_allclose(out[True][0], out[False][0], rtol=1e-13)
        np.testing.assert_array_equal(out[True][1], out[False][1])          # includes the untouched head/tail
        np.testing.assert_allclose(out[True][2], out[False][2], rtol=1e-12)
        ref = 0.5 * (1.7 * np.log(xs[off:off + n] / cs[off:off + n]) + 1.7)                # d/dx [log(x/c) w x]
        np.testing.assert_allclose(out[True][1][off:off + n], ref, rtol=1e-13, atol=1e-14)


def test_expr_specialised_kernel_disk_cache(tmp_path):
    # FADB_JIT_CACHE_DIR: the first process compiles and stores the cubin, the second loads it (no NVRTC run)
    import os, subprocess, sys, json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, json, numpy as np; sys.path.insert(0, %r); import fastad_b200 as F\n"
        "F.load(); F.check(F.lib().fadb_init(0))\n"
        "e = F.Expr(); x = e.var(np.linspace(0.5, 2, 777))\n"
        "e.bind(e.sum(e.binary('mul', e.unary('tanh', x), e.unary('atan', x))))\n"
        "v = e.autodiff(1.0); print(json.dumps([float(v[0]), F.jit_stats()]))\n" % root)
    env = dict(os.environ, FADB_JIT_CACHE_DIR=str(tmp_path))
    outs = [json.loads(subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, check=True).stdout.strip().splitlines()[-1])
            for _ in range(2)]
    assert outs[0][1]["compiled"] == 1 and outs[0][1]["launched"] == 1
    assert any(f.endswith(".cubin") for f in os.listdir(tmp_path))
    assert outs[1][1]["compiled"] == 0 and outs[1][1]["hits"] == 1 and outs[1][1]["launched"] == 1
    assert outs[0][0] == outs[1][0]


@pytest.mark.parametrize("name", ["if_else_taken", "if_else_not_taken", "if_else_logical", "if_else_nested_tt", "if_else_nested_tf",
                                  "if_else_nested_f"])
def test_expr_if_else_runs_specialised(fb, name):
    # control flow no longer pins a stage to the interpreter: the specialised kernel carries a jump watermark over its
    # unrolled program; both engines must match the oracle (test_expr_matches_oracle) and each other to the bit
    res = {}
    for on in (True, False):
        prev = fb.jit_enable(on)
        try:
            s0 = fb.jit_stats()
            gv, gadj, _, _ = _run(exprs.CASES[name], fb.Expr)
            s1 = fb.jit_stats()
            res[on] = (gv, gadj, s1["launched"] - s0["launched"], s1["failed"] - s0["failed"])
        finally:
            fb.jit_enable(prev)
    assert res[True][2] >= 1 and res[True][3] == 0, fb.lib().fadb_jit_diagnostics().decode()
    assert res[False][2] == 0
    ov, oadj, _, _ = _run(exprs.CASES[name], O.OracleExpr)
    _close(res[True][0], ov, 1e-12, name)
    np.testing.assert_allclose(res[True][0], res[False][0], rtol=1e-13)     # the 4-elements-per-thread form sums in another order
    for a, b, o in zip(res[True][1], res[False][1], oadj):
        if a.size > 1:
            np.testing.assert_array_equal(a, b)
        else:
            np.testing.assert_allclose(a, b, rtol=1e-12)
        _close(a, o, 1e-11, name)


def test_expr_normal_invalid_sigma_generic_path(fb):
    # vector sigma from an expression with one non-positive element: value -inf, no adjoints
    n = 500
    rng = np.random.default_rng(3)
    for backend in (O.OracleExpr, fb.Expr):
        e = backend()
        x, mu = e.var(rng.normal(size=n)), e.var(rng.normal(size=n))
        sg = rng.uniform(0.5, 2, n); sg[123] = -0.3
        s = e.var(sg)
        root = e.normal(x, mu, e.binary("mul", s, e.const(1.0 + 0 * sg)))
        e.bind(root)
        v = e.autodiff(1.0)
        assert v[0] == -np.inf
        for leaf in (x, mu, s):
            assert np.all(e.adj(leaf) == 0)
        rng = np.random.default_rng(3)


def test_expr_fast_paths_are_selected(fb):
    e = fb.Expr()
    x = e.var(np.linspace(0.5, 2, 1000))
    e.bind(e.sum(e.unary("exp", x)))
    assert "fast=1" in e.describe()
    e = fb.Expr()
    x, mu, s = e.var(np.zeros(100)), e.var(0.1), e.var(1.3)
    e.bind(e.normal(x, mu, s))
    assert "fast=2" in e.describe()
    build = exprs.CASES["linreg_expr"]
    e = fb.Expr()
    root, _, _ = build(e)
    e.bind(root)
    assert "fast=3" in e.describe()


def test_expr_rebind_leaf(fb):
    import torch
    e = fb.Expr()
    x = e.var(np.array([1.0, 2.0, 3.0]))
    e.bind(e.sum(e.binary("mul", x, x)))
    assert e.autodiff()[0] == 14.0
    nv = torch.tensor([2.0, 2.0, 2.0], dtype=torch.float64, device="cuda")
    na = torch.zeros(3, dtype=torch.float64, device="cuda")
    import ctypes as C
    fb.check(fb.lib().fadb_expr_rebind(e.h, x, C.c_void_p(nv.data_ptr()), C.c_void_p(na.data_ptr())))
    assert e.autodiff()[0] == 12.0
    assert na.cpu().tolist() == [4.0, 4.0, 4.0]


RING_CASES = sorted(k for k in exprs.CASES if "_ring" in k)


def _ring_run(fb, build, on, how):
    prev = fb.lib().fadb_jit_ring_enable(1 if on else 0)
    try:
        exprs.RNG = np.random.default_rng(20261019)
        e = fb.Expr()
        root, variables, seed = build(e)
        e.bind(root)
        if how == "autodiff":
            v = e.autodiff(seed)
            v = e.autodiff(seed)                 # adjoints accumulate: the second sweep reads what the first one wrote
        else:
            v = e.feval()
            e.beval(seed)
        return v, [e.adj(x).copy() for x in variables], [e.value(x).copy() for x in variables]
    finally:
        fb.lib().fadb_jit_ring_enable(prev)


@pytest.mark.parametrize("how", ["autodiff", "feval_beval"])
@pytest.mark.parametrize("name", RING_CASES)
def test_operand_ring_is_bit_identical_to_plain_kernel(fb, name, how):
    # Placeholder stages of >= 4096 elements: the specialised kernel with the cp.async operand ring (stage_kernel.cuh, JIT_RING)
    # performs the same arithmetic as the plain one-element kernel -- values, placeholders and adjoints must agree bit for bit,
    # fused (autodiff, twice) and as forward-only + backward-only launches.  The oracle comparison of the same cases runs in
    # test_expr_matches_oracle.
    build = exprs.CASES[name]
    j0 = fb.jit_stats()
    v_on, a_on, x_on = _ring_run(fb, build, True, how)
    j1 = fb.jit_stats()
    v_off, a_off, x_off = _ring_run(fb, build, False, how)
    j2 = fb.jit_stats()
    assert j2["failed"] == j0["failed"], fb.lib().fadb_jit_diagnostics().decode()
    assert j1["compiled"] + j1["hits"] > j0["compiled"] + j0["hits"]
    np.testing.assert_array_equal(v_on, v_off)
    for g, h in zip(a_on, a_off):
        np.testing.assert_array_equal(g, h)
    for g, h in zip(x_on, x_off):
        np.testing.assert_array_equal(g, h)


def test_operand_ring_falls_back_when_operands_overlap(fb):
    # A copy issued one trip early must not be overtaken by a store: two leaves sharing one adjoint buffer (x_adj doubles as
    # the placeholder's adjoint) take the plain kernel, so the result equals the ring-off run bit for bit.
    import ctypes as C
    import torch
    n = 200001
    rng = np.random.default_rng(5)
    xv = rng.uniform(0.5, 2.0, n)
    res = {}
    for on in (True, False):
        prev = fb.lib().fadb_jit_ring_enable(1 if on else 0)
        try:
            e = fb.Expr()
            tx = torch.from_numpy(xv).cuda()
            ta = torch.zeros(n, dtype=torch.float64, device="cuda")
            shared_adj = torch.zeros(n, dtype=torch.float64, device="cuda")
            e.keep += [tx, ta, shared_adj]
            L = fb.lib()
            x = e._id(L.fadb_expr_var(e.h, fb.VEC, n, 1, C.c_void_p(tx.data_ptr()), C.c_void_p(shared_adj.data_ptr())))
            a = e._id(L.fadb_expr_var(e.h, fb.VEC, n, 1, C.c_void_p(ta.data_ptr()), C.c_void_p(shared_adj.data_ptr())))
            root = e.glue(e.eq(a, e.unary("log", x)), e.binary("mul", a, x))
            e.bind(root)
            ones = np.ones(n)
            e.autodiff(ones)
            e.autodiff(ones)
            torch.cuda.synchronize()
            res[on] = (shared_adj.cpu().numpy().copy(), ta.cpu().numpy().copy())
        finally:
            fb.lib().fadb_jit_ring_enable(prev)
    np.testing.assert_array_equal(res[True][0], res[False][0])
    np.testing.assert_array_equal(res[True][1], res[False][1])
    # d/dx [log(x) x] = 1 + log x, accumulated over two sweeps into a buffer that also receives the placeholder's adjoint
    assert np.all(np.isfinite(res[True][0]))
