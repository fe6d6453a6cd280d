"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the
reference's golden vectors.  Run with `-m gpu` on a B200.

Tolerances (BASELINE.json north_star):
  * arithmetic-only per-element adjoints (+,-,*,/ are IEEE on both sides, kernels built
    -fmad=false): BIT-EXACT;
  * per-instance results through libdevice transcendentals (exp/log/erf, <=1-2 ulp vs glibc):
    <= 4 ulp, plus a conditioning floor where the expression subtracts nearly equal terms;
  * reordered reductions: <= 1e-12 relative.
"""
import math
import os

import numpy as np
import pytest

import kats
import oracle_py as O
import synth
from kats import assert_double_eq

pytestmark = pytest.mark.gpu

REL_REDUCE = 1e-12


@pytest.fixture(scope="module")
def fb():
    import torch
    assert torch.cuda.is_available(), "these tests need a CUDA device"
    import fastad_b200 as F
    F.load()
    F.check(F.lib().fadb_init(0))
    F.use_torch_stream()
    return F


def dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()


def zeros(n):
    import torch
    return torch.zeros(n, dtype=torch.float64, device="cuda")


def host(t):
    return t.cpu().numpy()


def rel(a, b):
    return abs(a - b) / max(abs(b), 1e-300)


# ------------------------------------------------------------------ K4 normal log-pdf
def _normal_case(fb, kind, x, mu, sg, const_x=False, const_mu=False, const_sigma=False, seed=1.0,
                 flags=0, adj0=None):
    """Runs GPU and oracle on the same inputs; returns ((v, xa, ma, sa) gpu, same oracle)."""
    x, mu, sg = np.atleast_1d(x), np.atleast_1d(mu), np.atleast_1d(sg)
    init = (lambda n: np.zeros(n)) if adj0 is None else (lambda n: adj0(n))
    xa0, ma0, sa0 = init(len(x)), init(len(mu)), init(len(sg))
    dx, dm, ds = dev(x), dev(mu), dev(sg)
    gxa = None if const_x else dev(xa0)
    gma = None if const_mu else dev(ma0)
    gsa = None if const_sigma else dev(sa0)
    v = fb.normal_lpdf(dx, dm, ds, gxa, gma, gsa, seed=seed, flags=flags)
    oxa = None if const_x else xa0.copy()
    oma = None if const_mu else ma0.copy()
    osa = None if const_sigma else sa0.copy()
    ov = O.normal_lpdf(x, mu if len(mu) > 1 else mu, sg, oxa, oma, osa, seed=seed)
    g = (v, None if gxa is None else host(gxa), None if gma is None else host(gma),
         None if gsa is None else host(gsa))
    return g, (ov, oxa, oma, osa)


@pytest.mark.parametrize("kind", ["sss", "vss", "vvs", "vsv", "vvv"])
def test_normal_reference_goldens(fb, kind):
    # test/reverse/stat/normal_unittest.cpp:105-297 (EXPECT_DOUBLE_EQ = 4 ulp)
    val, xa, ma, sa = kats.NORMAL_KAT[kind]
    x = kats.N_VX if kind[0] == "v" else kats.N_X
    mu = kats.N_VMU if kind[1] == "v" else kats.N_MU
    sg = kats.N_VSIGMA if kind[2] == "v" else kats.N_SIGMA
    g, o = _normal_case(fb, kind, x, mu, sg)
    assert_double_eq([g[0]], [val], what=kind + " value")
    assert_double_eq(g[1], xa, what=kind + " x_adj")
    assert_double_eq(g[2], ma, what=kind + " mu_adj")
    assert_double_eq(g[3], sa, what=kind + " sigma_adj")


def test_normal_constant_operands(fb):
    # normal_unittest.cpp:151-172, 237-257
    val, _, ma, sa = kats.NORMAL_KAT["vss"]
    g, _ = _normal_case(fb, "vss", kats.N_VX, kats.N_MU, kats.N_SIGMA, const_x=True)
    assert_double_eq([g[0]], [val]); assert_double_eq(g[2], ma); assert_double_eq(g[3], sa)
    val, _, ma, _ = kats.NORMAL_KAT["vsv"]
    g, _ = _normal_case(fb, "vsv", kats.N_VX, kats.N_MU, kats.N_VSIGMA, const_x=True, const_sigma=True)
    assert_double_eq([g[0]], [val]); assert_double_eq(g[2], ma)


@pytest.mark.parametrize("mode", ["optimistic", "optimistic_zero_adj", "optimistic_overwrite", "strict"])
@pytest.mark.parametrize("kind,bad", [("vsv", -1.0), ("vvv", 0.0), ("vss", 0.0), ("vvs", -2.0), ("sss", 0.0)])
def test_normal_invalid_sigma(fb, kind, bad, mode):
    # normal_unittest.cpp:207-213, 266-272: value -inf; normal.hpp:160,457: no adjoint update.
    # Vector sigma: the default is the optimistic pass + undo (normal.cu) -- exact when the adjoints start at zero or are
    # stored (FADB_OVERWRITE_ADJ), within one rounding of (previous + contribution) when they accumulate into non-zero
    # values, and the element with the invalid sigma itself is never touched; FADB_STRICT_SIGMA (pre-pass) is bit-exact.
    n = 1000 if kind != "sss" else 1
    x, mu, sg = synth.normal_inputs(n)
    mu = mu if kind[1] == "v" else mu[:1]
    sg = sg.copy() if kind[2] == "v" else sg[:1].copy()
    sg[len(sg) // 2] = bad
    a0 = 0.0 if mode in ("optimistic_zero_adj", "optimistic_overwrite") else 0.25
    flags = {"strict": fb.STRICT_SIGMA, "optimistic_overwrite": fb.OVERWRITE_ADJ}.get(mode, 0)
    g, o = _normal_case(fb, kind, x, mu, sg, adj0=lambda m: np.full(m, a0), flags=flags)
    assert g[0] == -math.inf and o[0] == -math.inf
    exact = mode != "optimistic" or kind[2] == "s"
    for a in g[1:]:
        if exact:
            assert np.all(a == a0), (kind, mode)
        else:
            # |contribution| < 2e3 on these inputs (|x - mu| < 12, sigma >= 0.5): one rounding at that magnitude
            np.testing.assert_allclose(a, a0, rtol=0, atol=2.5e-13)
            if len(a) == n:
                assert a[n // 2] == a0


def test_normal_invalid_sigma_full_width_undo(fb):
    # the undo pass over the vector path, the scalar tail and unaligned operands: 2^20 + 3 elements, three invalid sigmas
    # (first, middle, last element), adjoints start at zero -> must read exactly zero afterwards; then the same
    # operands with the sigmas repaired must give the oracle's bits (nothing of the undone evaluation is left behind)
    import torch
    n = (1 << 20) + 3
    x, mu, sg = synth.normal_inputs(n)
    bad = sg.copy()
    bad[[0, n // 2, n - 1]] = [-1.0, 0.0, -3.0]
    for off in (0, 1):
        dx, dm, db, dg = (dev(np.concatenate([np.zeros(off), a])) [off:] for a in (x, mu, bad, sg))
        xa, ma, sa = (torch.zeros(n + off, dtype=torch.float64, device="cuda")[off:] for _ in range(3))
        v = fb.normal_lpdf(dx, dm, db, xa, ma, sa, seed=0.75)
        assert v == -math.inf
        assert not bool(xa.any()) and not bool(ma.any()) and not bool(sa.any())
        v = fb.normal_lpdf(dx, dm, dg, xa, ma, sa, seed=0.75)
        oxa, oma, osa = np.zeros(n), np.zeros(n), np.zeros(n)
        ov = O.normal_lpdf(x, mu, sg, oxa, oma, osa, seed=0.75)
        assert rel(v, ov) <= REL_REDUCE
        np.testing.assert_array_equal(host(xa), oxa)
        np.testing.assert_array_equal(host(ma), oma)
        np.testing.assert_array_equal(host(sa), osa)


@pytest.mark.parametrize("case", range(30))
def test_normal_vector_sigma_guard_random(fb, case):
    # the optimistic pass + undo over random sizes, operand offsets (32-byte aligned or not), shapes and flags, with 0-3 invalid
    # sigmas at random places.  Valid input: adjoint vectors bit for bit the oracle's.  Invalid input: -inf, and adjoints that
    # started at zero read exactly zero again (undo of the default form, untouched in the strict form).
    import torch
    rng = np.random.default_rng(1000 + case)
    n = int(rng.choice([1, 2, 3, 7, 64, 257, 1000, 4099, 70001]))
    kind = ("vvv", "vsv")[case % 2]
    flags = (0, fb.OVERWRITE_ADJ, fb.STRICT_SIGMA)[case % 3]
    off = int(rng.integers(0, 4))
    x, mu, sg = rng.normal(size=n), rng.normal(size=n if kind == "vvv" else 1), rng.uniform(0.5, 2.0, n)
    nbad = int(rng.integers(0, 4)) if n > 3 else int(rng.integers(0, 2))
    bad_at = rng.choice(n, size=min(nbad, n), replace=False)
    sg[bad_at] = rng.choice([0.0, -1.0, -0.25], size=len(bad_at))
    seed = float(rng.uniform(0.5, 1.5))

    def devoff(a):
        t = torch.zeros(len(a) + off, dtype=torch.float64, device="cuda")
        t[off:] = torch.from_numpy(np.ascontiguousarray(a)).cuda()
        return t[off:]
    dx, dm, ds = devoff(x), devoff(mu), devoff(sg)
    xa, ma, sa = devoff(np.zeros(n)), devoff(np.zeros(len(mu))), devoff(np.zeros(n))
    v = fb.normal_lpdf(dx, dm, ds, xa, ma, sa, seed=seed, flags=flags)
    oxa, oma, osa = np.zeros(n), np.zeros(len(mu)), np.zeros(n)
    ov = O.normal_lpdf(x, mu, sg, oxa, oma, osa, seed=seed)
    if len(bad_at):
        assert v == -math.inf and ov == -math.inf
        assert not bool(xa.any()) and not bool(ma.any()) and not bool(sa.any()), (case, kind, flags, n, off)
    elif n == 1:       # the sss node (normal.hpp:141-171): scalar formulas, 4 ulp like the reference's own tests
        assert rel(v, ov) <= REL_REDUCE
        for g, o in ((xa, oxa), (ma, oma), (sa, osa)):
            assert rel(host(g)[0], o[0]) <= 1e-15
    else:
        assert rel(v, ov) <= REL_REDUCE
        np.testing.assert_array_equal(host(xa), oxa)
        np.testing.assert_array_equal(host(sa), osa)
        if kind == "vvv":
            np.testing.assert_array_equal(host(ma), oma)
        else:
            assert rel(host(ma)[0], oma[0]) <= 1e-11


@pytest.mark.parametrize("kind", ["vss", "vvs", "vsv", "vvv"])
@pytest.mark.parametrize("n", [5, 1 << 10, (1 << 20) + 3])
def test_normal_vs_oracle(fb, kind, n):
    x, mu, sg = synth.normal_inputs(n)
    mu = mu if kind[1] == "v" else mu[:1]
    sg = sg if kind[2] == "v" else sg[:1]
    g, o = _normal_case(fb, kind, x, mu, sg, seed=0.75, adj0=lambda m: np.linspace(-1, 1, m))
    assert rel(g[0], o[0]) <= REL_REDUCE
    # vector adjoints: arithmetic only -> bit exact; scalar adjoints: reductions -> 1e-12
    for ga, oa in zip(g[1:], o[1:]):
        if len(oa) == n:
            np.testing.assert_array_equal(ga, oa)
        else:
            assert rel(ga[0], oa[0]) <= REL_REDUCE


def test_normal_overwrite_and_trust_flags(fb):
    n = 4099
    x, mu, sg = synth.normal_inputs(n)
    g0, o = _normal_case(fb, "vvv", x, mu, sg)
    g1, _ = _normal_case(fb, "vvv", x, mu, sg, flags=fb.OVERWRITE_ADJ | fb.TRUST_SIGMA,
                         adj0=lambda m: np.full(m, 123.0))
    for a, b in zip(g0[1:], g1[1:]):
        np.testing.assert_array_equal(a, b)
    assert g0[0] == g1[0]


def test_normal_seed_zero_is_noop_on_adjoints(fb):
    x, mu, sg = synth.normal_inputs(100)
    g, o = _normal_case(fb, "vvv", x, mu, sg, seed=0.0, adj0=lambda m: np.full(m, 2.0))
    assert rel(g[0], o[0]) <= REL_REDUCE
    for a in g[1:]:
        assert np.all(a == 2.0)


def test_normal_unaligned_and_empty(fb):
    import torch
    n = 1001
    x, mu, sg = synth.normal_inputs(n + 1)
    bx, bm, bs = dev(x), dev(mu), dev(sg)
    xa, ma, sa = zeros(n + 1), zeros(n + 1), zeros(n + 1)
    v = fb.normal_lpdf(bx[1:], bm[1:], bs[1:], xa[1:], ma[1:], sa[1:])   # 8-byte aligned only
    oxa, oma, osa = np.zeros(n), np.zeros(n), np.zeros(n)
    ov = O.normal_lpdf(x[1:].copy(), mu[1:].copy(), sg[1:].copy(), oxa, oma, osa)
    assert rel(v, ov) <= REL_REDUCE
    np.testing.assert_array_equal(host(xa)[1:], oxa)
    np.testing.assert_array_equal(host(sa)[1:], osa)
    e = torch.zeros(0, dtype=torch.float64, device="cuda")
    one = dev([1.0])
    assert fb.normal_lpdf(e, one, one, None, zeros(1), zeros(1)) == 0.0


@pytest.mark.parametrize("log2n", [26, 28])
def test_normal_full_size_properties(fb, log2n):
    # BASELINE config 3 sizes (2^26 and 2^28 elements, vvv).  The oracle does not finish in seconds here, so the inputs
    # are built to make every per-element quantity exactly representable: z = (x - mu) / sigma cycles through dyadic
    # values with sum z^2 = 15 per 8 elements, sigma cycles through powers of two whose logs cancel per 8 elements.
    # Then value = -0.5 * 15 * n / 8 and all three adjoint vectors are exact dyadic numbers: compared bit for bit.
    import torch
    n = 1 << log2n
    f64 = dict(dtype=torch.float64, device="cuda")
    i = torch.arange(n, device="cuda")
    z = torch.tensor([0.5, -1.0, 1.5, -2.0, 2.0, -1.5, 1.0, -0.5], **f64)[i & 7]
    sigma = torch.tensor([1.0, 2.0, 0.5, 4.0, 0.25, 2.0, 0.5, 1.0], **f64)[i & 7]
    mu = (i & 1023).to(torch.float64) / 1024.0
    del i
    x = mu + sigma * z                                   # exact
    seed = 0.75
    want_x = 0.5 - seed * z / sigma                      # adjoints accumulate onto 0.5
    want_m = 0.5 + seed * z / sigma
    want_s = 0.5 + seed * (z * z - 1.0) / sigma
    del z
    for flags in (0, fb.TRUST_SIGMA):
        xa, ma, sa = (torch.full((n,), 0.5, **f64) for _ in range(3))
        v = fb.normal_lpdf(x, mu, sigma, xa, ma, sa, seed=seed, flags=flags)
        assert rel(v, -0.5 * 15.0 * n / 8.0) <= REL_REDUCE
        assert torch.equal(xa, want_x) and torch.equal(ma, want_m) and torch.equal(sa, want_s)
        del xa, ma, sa
    # one invalid sigma anywhere in 2^28 elements: -inf and no adjoint touched (normal.hpp:440-457)
    sigma[n - 3] = -1.0
    xa = torch.zeros(n, **f64)
    assert fb.normal_lpdf(x, mu, sigma, xa, None, None) == -math.inf
    assert not bool(xa.any())
    del x, mu, sigma, xa, want_x, want_m, want_s
    torch.cuda.empty_cache()


# ------------------------------------------------------------------ K2 sum of unary maps
@pytest.mark.parametrize("op", ["exp", "log", "sqrt", "sin", "cos", "tan", "atan", "erf", "sigmoid",
                                "sinh", "cosh", "tanh", "neg", ("pow", 2), ("pow", 3), ("pow", 5), ("pow", -2)])
def test_sum_unary_vs_oracle(fb, op):
    n = (1 << 16) + 5
    x = synth.sum_inputs(n)
    gx, ga = dev(x), dev(np.full(n, 0.5))
    v = fb.sum_unary(op, gx, ga, seed=1.5)
    oa = np.full(n, 0.5)
    ov = O.sum_unary(op, x, oa, seed=1.5)
    assert rel(v, ov) <= REL_REDUCE
    # per-element adjoints: <= 4 ulp of the increment (libdevice vs glibc; the reference's array
    # pow is std::pow per element, pow.hpp:97, ours is the PowFunc multiplication tree), bit-exact
    # for arithmetic-only maps.  tanh's 1 - f*f cancels: absolute floor of a few ulp(1)*seed.
    inc_g, inc_o = host(ga) - 0.5, oa - 0.5
    if op == "neg":
        np.testing.assert_array_equal(host(ga), oa)
    else:
        tol = 4 * np.spacing(np.abs(inc_o)) + 2 * np.spacing(0.5 + np.abs(inc_o))
        if op == "tanh":
            tol = tol + 8 * 1.5 * np.finfo(float).eps
        assert np.all(np.abs(inc_g - inc_o) <= tol), np.max(np.abs(inc_g - inc_o) / tol)


@pytest.mark.parametrize("op", ["asin", "acos"])
def test_sum_unary_inverse_trig(fb, op):
    n = 4097
    x = np.linspace(-0.95, 0.95, n)
    gx, ga = dev(x), zeros(n)
    v = fb.sum_unary(op, gx, ga)
    oa = np.zeros(n)
    ov = O.sum_unary(op, x, oa)
    assert abs(v - ov) <= 1e-12 * max(1.0, np.abs(np.arcsin(x)).sum())
    np.testing.assert_array_equal(host(ga), oa)     # seed/sqrt(1-x*x): arithmetic only


def test_sum_fixture_vector_and_erf_kat(fb):
    # unary_unittest.cpp:320-324 (erf KAT); fixture vector base_fixture.hpp:109-115
    x0, f, df = kats.ERF_KAT
    gx, ga = dev([x0]), zeros(1)
    v = fb.sum_unary("erf", gx, ga, seed=3.14)
    assert_double_eq([v], [f])
    assert_double_eq(host(ga), [3.14 * df])
    gx, ga = dev(kats.FIX_VEC), zeros(5)
    v = fb.sum_unary(("pow", 2), gx, ga)
    assert_double_eq([v], [float((kats.FIX_VEC ** 2).sum())])
    assert_double_eq(host(ga), 2 * kats.FIX_VEC)


def test_sum_full_size_properties(fb):
    # BASELINE config 2 size (N = 2^24): linearity in the seed and accumulate semantics
    n = 1 << 24
    x = synth.sum_inputs(n)
    gx, ga = dev(x), zeros(n)
    v1 = fb.sum_unary("exp", gx, ga, seed=1.0)
    a1 = ga.clone()
    v2 = fb.sum_unary("exp", gx, ga, seed=1.0)          # second sweep doubles (eval_unittest.cpp:37-42)
    assert v1 == v2
    assert bool((ga == 2 * a1).all())
    ref = float(np.exp(x).sum())
    assert rel(v1, ref) <= REL_REDUCE
    assert_double_eq(host(a1)[:4096], np.exp(x[:4096]), ulps=2)


# ------------------------------------------------------------------ K3 prod
@pytest.mark.parametrize("zeros_at", [[], [3], [3, 7]])
def test_prod_vs_oracle(fb, zeros_at):
    n = 300
    rng = np.random.default_rng(5)
    x = rng.uniform(0.9, 1.1, n) * rng.choice([-1.0, 1.0], n)
    for z in zeros_at:
        x[z] = 0.0
    gx, ga = dev(x), dev(np.ones(n))
    v = fb.prod(gx, ga, seed=2.0)
    e = O.OracleExpr()
    xv = e.var(x)
    e.adj(xv)[:] = 1.0
    e.bind(e.prod(xv))
    ov = e.autodiff(2.0)[0]
    assert rel(v, ov) <= REL_REDUCE if ov != 0 else v == 0
    np.testing.assert_allclose(host(ga), e.adj(xv), rtol=1e-12, atol=0)


def test_prod_fixture_zero(fb):
    # base_fixture.hpp:113 has an exact zero; prod.hpp:197-207
    gx, ga = dev(kats.FIX_VEC), zeros(5)
    v = fb.prod(gx, ga, seed=2.0)
    assert v == 0.0
    others = np.prod(np.delete(kats.FIX_VEC, 3))
    np.testing.assert_allclose(host(ga), [0, 0, 0, 2.0 * others, 0], rtol=1e-15)


@pytest.mark.parametrize("zeros_where", [(), ("first",), ("last",), ("middle",), ("first", "last"), ("middle", "last")])
def test_prod_full_size_exact(fb, zeros_where):
    # prod_benchmark size (N = 2^24, the size bench.py times): every x_i is a signed power of two and every aligned group
    # of 4 multiplies to 1, so each thread's running product, each CTA partial and the grid total are exact whatever the
    # order: value and all 2^24 adjoints are compared BIT FOR BIT.  Zeros are placed in the first / middle / last CTA's
    # range: one zero -> only its adjoint is non-zero (seed x product of the others), two zeros (in different CTAs, so the
    # zero COUNT crosses the ticket reduction) -> every adjoint is zero (prod.hpp:188-212).
    import torch
    n = 1 << 24
    f64 = dict(dtype=torch.float64, device="cuda")
    cyc = torch.tensor([2.0, 0.5, -4.0, -0.25, 0.5, 2.0, 0.25, 4.0], **f64)
    x = cyc[torch.arange(n, device="cuda") & 7].contiguous()
    pos = {"first": 5, "middle": n // 2 + 1, "last": n - 3}
    zs = [pos[w] for w in zeros_where]
    orig = [float(x[z]) for z in zs]
    for z in zs:
        x[z] = 0.0
    seed = 1.5
    adj = torch.full((n,), 0.5, **f64)
    v = fb.prod(x, adj, seed=seed)
    if not zs:
        assert v == 1.0
        assert torch.equal(adj, 0.5 + seed / x)
    elif len(zs) == 1:
        assert v == 0.0
        want = torch.full((n,), 0.5, **f64)
        want[zs[0]] = 0.5 + seed * (1.0 / orig[0])          # product of all the others = 1 / (the value that was there)
        assert torch.equal(adj, want)
    else:
        assert v == 0.0
        assert torch.equal(adj, torch.full((n,), 0.5, **f64))
    del x, adj
    torch.cuda.empty_cache()


@pytest.mark.parametrize("kind", ["vss", "vsv", "vvs"])
def test_normal_scalar_operand_shapes_full_size(fb, kind):
    # BASELINE config 3 size (2^26) for the shapes with a SCALAR operand (normal.hpp:225-478): the scalar's adjoint is a
    # sum over all elements.  Dyadic z and power-of-two sigma make every term and every partial sum exactly
    # representable: value to 1e-12 (it contains n log sigma), all adjoints bit for bit.
    import torch
    n = 1 << 26
    f64 = dict(dtype=torch.float64, device="cuda")
    i = torch.arange(n, device="cuda")
    z = torch.tensor([0.5, -1.0, 1.5, -2.0, 2.0, -1.5, 1.0, 0.5], **f64)[i & 7]      # sum z = 1, sum z^2 = 15 per cycle
    cyc = n / 8.0
    seed = 0.75
    one = lambda v: torch.tensor([v], **f64)
    if kind == "vss":
        mu, sg = one(0.25), one(2.0)
        x = 0.25 + 2.0 * z
        want_v = -0.5 * 15.0 * cyc - n * math.log(2.0)
        want_x, want_m, want_s = 0.5 - seed * z / 2.0, 0.5 + seed * cyc / 2.0, 0.5 + seed * (15.0 - 8.0) * cyc / 2.0
    elif kind == "vvs":
        mu, sg = (i & 1023).to(torch.float64) / 1024.0, one(2.0)
        x = mu + 2.0 * z
        want_v = -0.5 * 15.0 * cyc - n * math.log(2.0)
        want_x, want_m, want_s = 0.5 - seed * z / 2.0, 0.5 + seed * z / 2.0, 0.5 + seed * (15.0 - 8.0) * cyc / 2.0
    else:   # vsv: scalar mu, vector sigma whose logs cancel per cycle
        mu = one(0.25)
        sg = torch.tensor([1.0, 2.0, 0.5, 4.0, 0.25, 2.0, 0.5, 1.0], **f64)[i & 7]
        x = 0.25 + sg * z
        want_v = -0.5 * 15.0 * cyc
        zs = (z / sg)
        per_cycle = float(zs[:8].sum())                      # sum over one cycle of z / sigma (dyadic)
        want_x, want_m, want_s = 0.5 - seed * zs, 0.5 + seed * per_cycle * cyc, 0.5 + seed * (z * z - 1.0) / sg
    del i
    xa = torch.full((n,), 0.5, **f64)
    ma = torch.full((mu.numel(),), 0.5, **f64)
    sa = torch.full((sg.numel(),), 0.5, **f64)
    v = fb.normal_lpdf(x, mu, sg, xa, ma, sa, seed=seed)
    assert rel(v, want_v) <= REL_REDUCE
    assert torch.equal(xa, want_x)
    assert torch.equal(ma, want_m) if ma.numel() > 1 else float(ma[0]) == want_m
    assert torch.equal(sa, want_s) if sa.numel() > 1 else float(sa[0]) == want_s
    del x, xa, z
    torch.cuda.empty_cache()


# ------------------------------------------------------------------ K1 Black-Scholes
BS_C = 8.0      # |gpu - oracle| <= BS_C x (eps x magnitude of the output's terms), tests/bs_tol.py


@pytest.mark.parametrize("n", [1, 2, 1001, 1 << 16])
def test_black_scholes_vs_oracle(fb, n):
    import bs_tol
    inp = synth.black_scholes_inputs(n)
    ref = O.black_scholes(inp["S"], inp["K"], inp["sigma"], inp["tau"], inp["r"], nthreads=4)
    out = fb.black_scholes(*(dev(inp[k]) for k in ("S", "K", "sigma", "tau", "r")))
    for name, g, o, sc in zip(bs_tol.NAMES, out, ref, bs_tol.scales(inp)):
        err = np.abs(host(g) - o)
        assert np.all(err <= BS_C * sc), (name, float(np.max(err / sc)), int(np.argmax(err / sc)))


def test_black_scholes_ulp_report_full_size(fb, capsys):
    # BASELINE config 1 at full size against the oracle: 2^20 options, every output, error in units of eps x the
    # magnitude of its terms (bs_tol.scales), plus the two prices in ulps OF THE PRICE for |S/K - 1| < 5 %
    import bs_tol
    n = 1 << 20
    inp = synth.black_scholes_inputs(n)
    ref = O.black_scholes(inp["S"], inp["K"], inp["sigma"], inp["tau"], inp["r"], nthreads=os.cpu_count() or 4)
    out = [host(t) for t in fb.black_scholes(*(dev(inp[k]) for k in ("S", "K", "sigma", "tau", "r")))]
    ratios, atm_ulps = bs_tol.report(inp, out, ref)
    with capsys.disabled():
        print("\nblack-scholes 2^20 vs oracle, max |err| / (eps x term magnitude): "
              + ", ".join(f"{nm} {r:.2f}" for nm, r in zip(bs_tol.NAMES, ratios))
              + f" | near-the-money prices: call {atm_ulps[0]:.1f} ulp, put {atm_ulps[1]:.1f} ulp")
    assert max(ratios) <= BS_C, dict(zip(bs_tol.NAMES, ratios))


def test_black_scholes_example_point(fb):
    # example/black_scholes.cpp:43-70 inputs; closed form in SURVEY.md §6
    p = kats.BS_POINT
    out = fb.black_scholes(*(dev([p[k]]) for k in ("S", "K", "sigma", "tau", "r")))
    k = kats.BS_KAT
    want = [k["call"], k["put"], k["delta_call"], k["vega"], k["rho_call"], k["delta_put"], k["vega"], k["rho_put"]]
    for g, w in zip(out, want):
        assert abs(host(g)[0] - w) <= 1.6e-13 * max(1.0, abs(w)), (host(g)[0], w)


def test_black_scholes_accumulate_and_host_entry(fb):
    n = 3001
    inp = synth.black_scholes_inputs(n)
    d = [dev(inp[k]) for k in ("S", "K", "sigma", "tau", "r")]
    out1 = fb.black_scholes(*d)
    out2 = [o.clone() for o in out1]
    fb.black_scholes(*d, out=out2, flags=0)                # accumulate: greeks double, prices stored
    for k in range(2):
        assert bool((out2[k] == out1[k]).all())
    for k in range(2, 8):
        assert bool((out2[k] == 2 * out1[k]).all())
    hout = [np.zeros(n) for _ in range(8)]
    fb.black_scholes_host(inp["S"], inp["K"], inp["sigma"], inp["tau"], inp["r"], hout)
    for h, g in zip(hout, out1):
        np.testing.assert_array_equal(h, host(g))


def test_black_scholes_host_entry_page_locked_direct_path(fb):
    # all 13 arrays page-locked: the resident kernel runs on their device aliases (one launch, no copies);
    # mixed page-locked / pageable arrays: staged path.  Both must return the resident call's bits.
    import torch
    n = (1 << 18) + 77                                      # more than one staged chunk is not needed here
    inp = synth.black_scholes_inputs(n)
    keys = ("S", "K", "sigma", "tau", "r")
    out1 = fb.black_scholes(*(dev(inp[k]) for k in keys))
    pin_in = [torch.from_numpy(inp[k]).pin_memory() for k in keys]
    pin_out = [torch.full((n,), -7.0, dtype=torch.float64).pin_memory() for _ in range(8)]
    l0 = fb.lib().fadb_launch_count()
    fb.black_scholes_host(*[t.numpy() for t in pin_in], [t.numpy() for t in pin_out])
    assert fb.lib().fadb_launch_count() - l0 == 1          # one kernel launch, nothing staged
    for h, g in zip(pin_out, out1):
        np.testing.assert_array_equal(h.numpy(), host(g))
    mixed_out = [t.numpy() if k % 2 else np.zeros(n) for k, t in enumerate(pin_out)]
    for t in pin_out:
        t.fill_(-7.0)
    fb.black_scholes_host(*[t.numpy() for t in pin_in], mixed_out)
    for h, g in zip(mixed_out, out1):
        np.testing.assert_array_equal(h, host(g))


def test_black_scholes_host_entry_registered_user_arrays(fb):
    # plain numpy arrays (what a user of the reference holds), page-locked in place with fadb_host_register:
    # the host entry takes the one-launch path; after unregistering it stages again.  Same bits every time.
    import ctypes as C
    n = 50000
    inp = synth.black_scholes_inputs(n)
    keys = ("S", "K", "sigma", "tau", "r")
    out1 = fb.black_scholes(*(dev(inp[k]) for k in keys))
    ins = [np.ascontiguousarray(inp[k]) for k in keys]
    outs = [np.zeros(n) for _ in range(8)]
    for a in ins + outs:
        fb.check(fb.lib().fadb_host_register(C.c_void_p(a.ctypes.data), a.nbytes))
    try:
        l0 = fb.lib().fadb_launch_count()
        fb.black_scholes_host(*ins, outs)
        assert fb.lib().fadb_launch_count() - l0 == 1
        for h, g in zip(outs, out1):
            np.testing.assert_array_equal(h, host(g))
    finally:
        for a in ins + outs:
            fb.check(fb.lib().fadb_host_unregister(C.c_void_p(a.ctypes.data)))
    for o in outs:
        o[:] = 0
    fb.black_scholes_host(*ins, outs)                        # pageable again: staged copies
    for h, g in zip(outs, out1):
        np.testing.assert_array_equal(h, host(g))


@pytest.mark.parametrize("n", [777, (1 << 16) + 5, 1 << 20])
def test_black_scholes_batches_early_start_chain(fb, n):
    # fadb_black_scholes_batches: batches 1.. start without waiting for the batch before them (bs_kernel_fast_pf<PDL_EARLY>)
    # and wait at their END.  (a) every batch returns the single call's bits, at sizes where many grids are co-resident
    # (777) and at the BASELINE size; (b) the chain stays ordered after the stream work that precedes the call (inputs
    # produced by a kernel enqueued right before it) and before the work that follows it (a reduction of the results
    # enqueued right after it), repeated to give a broken ordering the chance to show.
    import torch
    keys = ("S", "K", "sigma", "tau", "r")
    nsets = 4
    host_in = [synth.black_scholes_inputs(n, seed=synth.SEED + 31 * s) for s in range(nsets)]
    want = [[t.clone() for t in fb.black_scholes(*(dev(hi[k]) for k in keys))] for hi in host_in]
    src = [[dev(hi[k]) for k in keys] for hi in host_in]
    sets = [([torch.empty(n, dtype=torch.float64, device="cuda") for _ in keys],
             [torch.empty(n, dtype=torch.float64, device="cuda") for _ in range(8)]) for _ in range(nsets)]
    B = fb.BlackScholesBatches(sets)
    big = torch.empty(1 << 24, dtype=torch.float64, device="cuda")
    for rep in range(6):
        for ins, outs in sets:
            for t in ins:
                t.fill_(1.0)                 # valid but wrong operands
            for t in outs:
                t.fill_(-3.0)
        big.fill_(float(rep))                # a long kernel, then the producers of the real inputs, then the call
        for s in range(nsets):
            for t, a in zip(sets[s][0], src[s]):
                t.copy_(a)
        B.run(2 * nsets + 3, rep)
        sums = [torch.stack([o.sum() for o in outs]) for _, outs in sets]      # consumers right behind the chain
        torch.cuda.synchronize()
        for s in range(nsets):
            for k in range(8):
                assert torch.equal(sets[s][1][k], want[s][k]), (rep, s, k)
            assert torch.equal(sums[s], torch.stack([w.sum() for w in want[s]])), (rep, s)


def test_black_scholes_full_size_put_call_parity(fb):
    # BASELINE config 1 size: 2^20 options; size-independent property C - P = S - PV
    n = 1 << 20
    inp = synth.black_scholes_inputs(n)
    out = fb.black_scholes(*(dev(inp[k]) for k in ("S", "K", "sigma", "tau", "r")))
    call, put = host(out[0]), host(out[1])
    pv = inp["K"] * np.exp(-inp["r"] * inp["tau"])
    assert np.max(np.abs((call - put) - (inp["S"] - pv))) <= 1e-11
    dS_c, dS_p = host(out[2]), host(out[5])
    assert np.max(np.abs(dS_c - dS_p - 1.0)) <= 1e-11     # delta_call - delta_put = 1
    assert np.max(np.abs(host(out[3]) - host(out[6]))) == 0.0   # vega identical by construction


def test_peer_memory_allreduce_single_rank(fb):
    # csrc/p2p.cu with a world of one (the multi-rank exchange is exercised by tools/p2p_probe.py under torchrun:
    # bit-identical to the rank-ordered sum on 2 and 8 GPUs): mailbox, IPC handle, push / flag / wait / reduce
    import ctypes as C
    L = fb.lib()
    h = C.create_string_buffer(64)
    fb.check(L.fadb_p2p_create(1, 0, h))
    fb.check(L.fadb_p2p_connect(C.c_char_p(h.raw)))
    for n in (1, 26, 128):
        x = np.random.default_rng(n).normal(size=n)
        t = dev(x)
        for _ in range(3):                       # sequence numbers advance, both payload buffers get used
            fb.check(L.fadb_p2p_allreduce(fb._ptr(t), n))
        np.testing.assert_array_equal(host(t), x)
    err = C.c_ulonglong(7)
    fb.check(L.fadb_p2p_status(C.byref(err)))
    assert err.value == 0
    assert L.fadb_p2p_allreduce(fb._ptr(t), 129) != 0          # payload limit is enforced
    # the regression pass with the exchange fused in (push from the last CTA, pull in the finish kernel) must
    # give the bits of partial + finish; TMA shape (64 columns) and the register-tile shape (7 columns)
    for rows, cols in ((1 << 14, 64), (5002, 7)):
        X, y, w_true = synth.linreg_inputs(rows, cols, seed=3)
        Xd, yd, wd, sg = dev(X.reshape(-1, order="F")), dev(y), dev(w_true + 0.01), dev([1.3])
        out = []
        for fused in (False, True):
            wa, sa, part = zeros(cols), zeros(1), zeros(cols + 1)
            val = C.c_double()
            if fused:
                fb.check(L.fadb_linreg_partial_push(rows, cols, fb._ptr(Xd), fb._ptr(yd), fb._ptr(wd), fb._ptr(part)))
                fb.check(L.fadb_linreg_finish_pull(rows, cols, fb._ptr(sg), fb._ptr(wa), fb._ptr(sa), 0.75, 0, C.byref(val)))
            else:
                fb.check(L.fadb_linreg_partial(rows, cols, fb._ptr(Xd), fb._ptr(yd), fb._ptr(wd), fb._ptr(part)))
                fb.check(L.fadb_linreg_finish(rows, cols, fb._ptr(part), fb._ptr(sg), fb._ptr(wa), fb._ptr(sa), 0.75, 0, C.byref(val)))
            out.append((val.value, host(wa), host(sa)))
        assert out[0][0] == out[1][0]
        np.testing.assert_array_equal(out[0][1], out[1][1])
        np.testing.assert_array_equal(out[0][2], out[1][2])
    # the same for the normal log-pdf: vvv (vector adjoints in the pass) and vss (scalar adjoints in the finish kernel)
    n = 5003
    x, mu, sg = synth.normal_inputs(n)
    for shape, m, s_ in ((3, mu, sg), (0, np.array([0.1]), np.array([1.3]))):
        out = []
        for fused in (False, True):
            xd, md, sd = dev(x), dev(m), dev(s_)
            xa, ma, sa, part = zeros(n), zeros(len(m)), zeros(len(s_)), zeros(4)
            val = C.c_double()
            args = (shape, n, fb._ptr(xd), fb._ptr(md), fb._ptr(sd), fb._ptr(xa), fb._ptr(ma), fb._ptr(sa), 0.75, fb.TRUST_SIGMA, None, fb._ptr(part))
            if fused:
                fb.check(L.fadb_normal_lpdf_partial_push(*args))
                fb.check(L.fadb_normal_lpdf_finish_pull(shape, n, fb._ptr(xd), fb._ptr(md), fb._ptr(sd), None, fb._ptr(ma), fb._ptr(sa), 0.75,
                                                        fb.TRUST_SIGMA, C.byref(val)))
            else:
                fb.check(L.fadb_normal_lpdf_partial(*args))
                fb.check(L.fadb_normal_lpdf_finish(shape, n, fb._ptr(part), fb._ptr(xd), fb._ptr(md), fb._ptr(sd), None, fb._ptr(ma), fb._ptr(sa),
                                                   0.75, fb.TRUST_SIGMA, C.byref(val)))
            out.append((val.value, host(xa), host(ma), host(sa)))
        assert out[0][0] == out[1][0]
        for a, b in zip(out[0][1:], out[1][1:]):
            np.testing.assert_array_equal(a, b)
    # sum(f(x)) and the network with the WHOLE all-reduce fused into the one launch (last CTA pushes, waits, sums):
    # one rank must give the bits of the plain calls; odd sizes, table-log functor, all three network variants
    import torch
    for op, n in (("exp", 5003), ("log", 1 << 16), (("pow", 2), 70001), (("pow", 3), 129)):
        xs = synth.sum_inputs(n)
        out = []
        for call in (L.fadb_sum_unary_async, L.fadb_sum_unary_allreduce_async):
            xd, xa, dv = dev(xs), dev(np.full(n, 0.25)), zeros(1)
            for _ in range(2):                   # accumulating adjoints, both payload buffers
                fb.check(call(fb._op_code(op), n, fb._ptr(xd), fb._ptr(xa), 0.75, 0, fb._ptr(dv)))
            out.append((host(dv), host(xa)))
        np.testing.assert_array_equal(out[0][0], out[1][0])
        np.testing.assert_array_equal(out[0][1], out[1][1])
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    for batch in (257, sms * 256 + 77, sms * 1024 + 5):      # <1,1,256>, <1,1,512> and <1,2,256> variants
        Xn, yn = synth.mlp3_inputs(batch)
        gX, gy, gp = torch.from_numpy(np.ascontiguousarray(Xn.T)).cuda(), dev(yn), dev(kats.NN_PARM)
        out = []
        for call in (L.fadb_mlp3_async, L.fadb_mlp3_allreduce_async):
            gl = dev(np.full(26, 0.5))
            for flags in (0, fb.OVERWRITE_ADJ, 0):
                fb.check(call(batch, fb._ptr(gX), fb._ptr(gy), fb._ptr(gp), fb._ptr(gl), flags, C.c_void_p(gl.data_ptr() + 25 * 8)))
            out.append(host(gl))
        np.testing.assert_array_equal(out[0], out[1])
    fb.check(L.fadb_p2p_status(C.byref(err)))
    assert err.value == 0
    fb.check(L.fadb_p2p_destroy())
    # without connected mailboxes the fused calls fail loudly instead of skipping the exchange
    assert L.fadb_sum_unary_allreduce_async(fb._op_code("exp"), 4, fb._ptr(t), fb._ptr(t), 1.0, 0, fb._ptr(t)) != 0
    assert L.fadb_mlp3_allreduce_async(1, fb._ptr(t), fb._ptr(t), fb._ptr(t), None, 0, fb._ptr(t)) != 0


def test_least_squares_one_pass_kernel(fb):
    # fadb_lsq: sum((y - X w)^2) and -2 X^T (y - X w) from one pass over X; odd sizes, accumulate / overwrite
    import ctypes as C
    for rows, cols in ((1 << 14, 64), (5003, 7), (129, 1), (70000, 33)):
        X, y, w_true = synth.linreg_inputs(rows, cols, seed=rows)
        w = w_true + 0.1
        Xd, yd, wd = dev(X.reshape(-1, order="F")), dev(y), dev(w)
        wadj = dev(np.full(cols, 0.5))
        val = C.c_double()
        fb.check(fb.lib().fadb_lsq(rows, cols, fb._ptr(Xd), fb._ptr(yd), fb._ptr(wd), fb._ptr(wadj), 0.75, 0, C.byref(val)))
        r = y - X @ w
        np.testing.assert_allclose(val.value, r @ r, rtol=1e-12)
        ref = 0.5 + 0.75 * (-2.0) * (X.T @ r)
        np.testing.assert_allclose(host(wadj), ref, rtol=1e-11, atol=1e-11 * np.max(np.abs(ref)))
        fb.check(fb.lib().fadb_lsq(rows, cols, fb._ptr(Xd), fb._ptr(yd), fb._ptr(wd), fb._ptr(wadj), 1.0, fb.OVERWRITE_ADJ, C.byref(val)))
        np.testing.assert_allclose(host(wadj), -2.0 * (X.T @ r), rtol=1e-11, atol=1e-11 * np.max(np.abs(ref)))


# ------------------------------------------------------------------ K7 dense-covariance normal
def _dense_cov(n, rng):
    # benchmark/normal_benchmark.cpp:15-33 builds Sigma = L L^T, L lower with diagonal 10 and U(-1,1)
    # below.  The reference forms log(det L) as the log of a PRODUCT (normal.hpp:661), which overflows
    # to +inf beyond n ~ 300 with that diagonal (10^n); the kernel sums log L_ii instead.  To compare
    # against the oracle (which restates the product) large cases use a diagonal of 1.2.
    diag, scale = (10.0, 1.0) if n <= 200 else (1.2, 1.0 / np.sqrt(n))
    Lo = np.tril(rng.uniform(-1, 1, (n, n)), -1) * scale + diag * np.eye(n)
    return Lo @ Lo.T


def test_dense_normal_reference_goldens(fb):
    import torch
    # normal_unittest.cpp:298-395: only the lower triangle of Sigma is set (and read)
    Sig = np.array([[1.0, 0.0, 0.0], [0.3, 2.0, 0.0], [0.2, -0.3, 3.0]])
    gS = torch.from_numpy(np.ascontiguousarray(Sig.T)).cuda()         # column-major image
    xa, ma, Sa = zeros(3), zeros(1), zeros(9)
    v = fb.normal_dense(dev(kats.N_VX), dev([kats.N_MU]), gS, xa, ma, Sa)
    assert_double_eq([v], [-8.8105250497069019])
    assert_double_eq(host(xa), [-3.7624909485879803, 1.6010137581462704, -0.0890658942795076])
    assert_double_eq(host(ma), [2.2505430847212176])
    sa = host(Sa).reshape(3, 3, order="F")
    np.testing.assert_allclose([sa[0, 0], sa[1, 0], sa[2, 0], sa[1, 1], sa[2, 1], sa[2, 2]],
                               [6.5432306187049774, -2.9250063314004424, 0.2119067294266190, 1.0137007310866775,
                                -0.1038829443345370, -0.1689156028253514], rtol=1e-14, atol=1e-15)
    xa, ma, Sa = zeros(3), zeros(3), zeros(9)
    v = fb.normal_dense(dev(kats.N_VX), dev(kats.N_VMU), gS, xa, ma, Sa)
    assert_double_eq([v], [-7.3649692930088602])
    assert_double_eq(host(xa), [-3.4158218682114407, 0.4279507603186097, -0.5628167994207096])
    assert_double_eq(host(ma), -host(xa))
    # not positive definite (normal_unittest.cpp:304-310): -inf, adjoints untouched
    v = fb.normal_dense(dev(kats.N_VX), dev(kats.N_VMU), zeros(9).reshape(3, 3), xa, ma, Sa)
    assert v == -math.inf
    assert_double_eq(host(ma), -host(xa))


@pytest.mark.parametrize("n,mu_vec", [(7, 1), (32, 0), (33, 1), (100, 1), (257, 0), (1000, 1)])
def test_dense_normal_vs_oracle(fb, n, mu_vec):
    import torch
    rng = np.random.default_rng(n)
    Sig = _dense_cov(n, rng)
    x, mu = rng.uniform(-1, 1, n), (rng.uniform(-1, 1, n) if mu_vec else np.array([0.3]))
    e = O.OracleExpr()
    ex, em, eS = e.var(x), e.var(mu if mu_vec else float(mu[0])), e.var(Sig)
    e.bind(e.normal(ex, em, eS))
    ov = e.autodiff(0.9)[0]
    xa, ma, Sa = zeros(n), zeros(len(mu)), zeros(n * n)
    v = fb.normal_dense(dev(x), dev(mu), torch.from_numpy(np.ascontiguousarray(Sig.T)).cuda(), xa, ma, Sa, seed=0.9)
    assert rel(v, ov) <= 1e-11
    np.testing.assert_allclose(host(xa), e.adj(ex), rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(host(ma), e.adj(em), rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(host(Sa), e.adj(eS), rtol=1e-8, atol=1e-12 * np.abs(e.adj(eS)).max())


def test_dense_normal_logdet_does_not_overflow(fb):
    # the reference's own benchmark covariance at n = 1000: log(det L) = log(10^1000) = +inf there
    import torch
    n = 1000
    rng = np.random.default_rng(1)
    Lo = np.tril(rng.uniform(-1, 1, (n, n)), -1) + 10 * np.eye(n)
    Sig = Lo @ Lo.T
    x, mu = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
    v = fb.normal_dense(dev(x), dev(mu), torch.from_numpy(np.ascontiguousarray(Sig.T)).cuda(), zeros(n), zeros(n), zeros(n * n))
    z = np.linalg.solve(Sig, x - mu)
    want = -0.5 * (x - mu) @ z - np.log(np.diag(np.linalg.cholesky(Sig))).sum()
    assert rel(v, want) <= 1e-11


# ------------------------------------------------------------------ det / log_det / wishart (8f row 3)
def _colmajor(M):
    import torch
    return torch.from_numpy(np.ascontiguousarray(M.T)).cuda()


@pytest.mark.parametrize("log", [False, True])
@pytest.mark.parametrize("policy", ["fullpivlu", "ldlt", "llt"])
def test_det_reference_goldens(fb, policy, log):
    # det_unittest.cpp:57-155 / log_det_unittest.cpp:57-152 (exact determinants 153 and 65)
    A = kats.DET_FPLU if policy == "fullpivlu" else kats.DET_LLT
    exact = 153.0 if policy == "fullpivlu" else 65.0
    Aa = zeros(16)
    v = fb.det(_colmajor(A), Aa, seed=kats.DET_SEED, policy=policy, log=log)
    assert_double_eq([v], [math.log(exact) if log else exact])
    want = kats.DET_SEED * (1.0 if log else exact) * np.linalg.inv(A).T
    assert np.abs(host(Aa).reshape(4, 4, order="F") - want).max() <= 3e-13
    # adjoints accumulate
    fb.det(_colmajor(A), Aa, seed=kats.DET_SEED, policy=policy, log=log)
    assert np.abs(host(Aa).reshape(4, 4, order="F") - 2 * want).max() <= 6e-13


# n <= 160: the whole elimination in one launch of one CTA, matrix in shared memory (gj_small_kernel); above: pivot + update
# launches per column
@pytest.mark.parametrize("policy,n", [("fullpivlu", 1), ("fullpivlu", 5), ("fullpivlu", 67), ("fullpivlu", 160), ("fullpivlu", 161),
                                      ("fullpivlu", 300), ("ldlt", 33), ("ldlt", 150), ("ldlt", 170), ("llt", 31), ("llt", 130),
                                      ("llt", 500)])
def test_det_vs_oracle(fb, policy, n):
    rng = np.random.default_rng(n)
    if policy == "fullpivlu":
        A = rng.uniform(-1, 1, (n, n)) + 0.5 * np.eye(n)
    else:
        Lo = np.tril(rng.uniform(-1, 1, (n, n)), -1) / np.sqrt(n) + 1.1 * np.eye(n)
        A = Lo @ Lo.T
    for log in (False, True):
        e = O.OracleExpr()
        a = e.var(A)
        e.bind(e.det(a, policy, log))
        ov = e.autodiff(0.7)[0]
        Aa = zeros(n * n)
        v = fb.det(_colmajor(A), Aa, seed=0.7, policy=policy, log=log)
        assert rel(v, ov) <= 1e-10, (policy, n, log)
        oa = e.adj(a)
        np.testing.assert_allclose(host(Aa), oa, rtol=1e-7, atol=1e-10 * np.abs(oa).max())


def test_det_single_launch_path_is_bit_identical_to_the_launch_chain():
    # same pivot rule, same fma per element: tools/det_probe.py prints the value and a digest of the adjoint; one process per
    # setting (the switch is read once)
    import hashlib, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for small in ("1", "0"):
        p = subprocess.run([sys.executable, os.path.join(root, "tools", "det_probe.py"), "97", "160"], capture_output=True, text=True,
                           env=dict(os.environ, FADB_GJ_SMALL=small), timeout=300)
        assert p.returncode == 0, p.stderr[-2000:]
        outs.append([l for l in p.stdout.splitlines() if l.startswith("digest")])
    assert outs[0] == outs[1] and len(outs[0]) == 4, outs


def test_det_invalid_leaves_adjoint_untouched(fb):
    for policy, A in (("fullpivlu", np.array([[1., 2], [2, 4]])), ("llt", np.array([[1., 2], [2, 1]])),
                      ("ldlt", np.zeros((2, 2)))):
        Aa = zeros(4)
        v = fb.det(_colmajor(A), Aa, seed=1.0, policy=policy)
        assert np.all(host(Aa) == 0), policy
        if policy != "llt":
            assert v == 0.0


def test_wishart_reference_goldens(fb):
    # wishart_unittest.cpp:62-106
    Xa, Va = zeros(9), zeros(9)
    v = fb.wishart(_colmajor(kats.WISHART_X), _colmajor(kats.WISHART_V), kats.WISHART_N, Xa, Va)
    assert_double_eq([v], [kats.WISHART_VALUE])
    n, p = kats.WISHART_N, 3
    vi = np.linalg.inv(kats.WISHART_V)
    dX = 0.5 * ((n - p - 1) * np.linalg.inv(kats.WISHART_X) - vi)
    dV = 0.5 * (vi @ kats.WISHART_X @ vi - n * vi)
    assert np.abs(host(Xa).reshape(3, 3, order="F") - dX).max() <= 1e-15
    assert np.abs(host(Va).reshape(3, 3, order="F") - dV).max() <= 1e-15
    v = fb.wishart(_colmajor(kats.WISHART_X), zeros(9), kats.WISHART_N, Xa, Va)     # feval_invalid
    assert v == -math.inf
    assert np.abs(host(Xa).reshape(3, 3, order="F") - dX).max() <= 1e-15
    assert fb.wishart(_colmajor(kats.WISHART_X), _colmajor(kats.WISHART_V), 2, Xa, Va) == -math.inf


@pytest.mark.parametrize("p", [2, 40, 129, 400])
def test_wishart_vs_oracle(fb, p):
    rng = np.random.default_rng(p)
    def spd():
        Lo = np.tril(rng.uniform(-1, 1, (p, p)), -1) / np.sqrt(p) + 1.2 * np.eye(p)
        return Lo @ Lo.T
    X, V, df = spd(), spd(), p + 3
    e = O.OracleExpr()
    x, vv = e.var(X), e.var(V)
    e.bind(e.wishart(x, vv, df))
    ov = e.autodiff(1.3)[0]
    Xa, Va = zeros(p * p), zeros(p * p)
    v = fb.wishart(_colmajor(X), _colmajor(V), df, Xa, Va, seed=1.3)
    assert rel(v, ov) <= 1e-11
    np.testing.assert_allclose(host(Xa), e.adj(x), rtol=1e-8, atol=1e-11 * np.abs(e.adj(x)).max())
    np.testing.assert_allclose(host(Va), e.adj(vv), rtol=1e-8, atol=1e-11 * np.abs(e.adj(vv)).max())


# ------------------------------------------------------------------ other diagonal log-pdfs
def _lp_case(fb, dist, n, bvec, cvec, bad=False, seed=0.8, flags=0):
    rng = np.random.default_rng(n + 7 * bvec + 13 * cvec)
    x = rng.uniform(-1, 1, n)
    if dist == "cauchy":
        b = rng.uniform(-1, 1, n if bvec else 1); c = rng.uniform(0.2, 2, n if cvec else 1)
        if bad: c[len(c) // 2] = -0.5
    elif dist == "uniform":
        b = -1.5 - rng.uniform(0, 1, n if bvec else 1); c = 1.5 + rng.uniform(0, 1, n if cvec else 1)
        if bad: x[n // 3] = 7.0
    else:
        x = rng.integers(0, 2, n).astype(float); b = rng.uniform(0.05, 0.95, n if bvec else 1); c = None
        if bad: x[n // 3] = 2.0
    e = O.OracleExpr()
    ex = e.var(x) if dist == "cauchy" else e.const(x)
    eb = e.var(b if bvec else float(b[0]))
    ec = None if c is None else e.var(c if cvec else float(c[0]))
    root = e.logpdf(dist, ex, eb, -1 if ec is None else ec)
    e.bind(root)
    ov = e.autodiff(seed)[0]
    gx, gb = dev(x), dev(b)
    gc = None if c is None else dev(c)
    xa = zeros(n) if dist == "cauchy" else None
    ba = zeros(len(b)); ca = None if c is None else zeros(len(c))
    v = fb.logpdf(dist, gx, gb, gc, xa, ba, ca, seed=seed, flags=flags)
    outs = [(host(ba), e.adj(eb))]
    if ca is not None: outs.append((host(ca), e.adj(ec)))
    if xa is not None: outs.append((host(xa), e.adj(ex)))
    return v, ov, outs


@pytest.mark.parametrize("dist", ["cauchy", "uniform", "bernoulli"])
@pytest.mark.parametrize("bvec,cvec", [(0, 0), (1, 0), (0, 1), (1, 1)])
@pytest.mark.parametrize("n", [3, 4097, (1 << 18) + 1])
def test_logpdf_kernels_vs_oracle(fb, dist, bvec, cvec, n):
    if dist == "bernoulli" and cvec:
        pytest.skip("bernoulli has two operands")
    v, ov, outs = _lp_case(fb, dist, n, bvec, cvec)
    assert rel(v, ov) <= REL_REDUCE
    for g, o in outs:
        if len(o) == n and n > 1:
            np.testing.assert_allclose(g, o, rtol=4e-16, atol=0)      # arithmetic only: <= 1 ulp
        else:
            np.testing.assert_allclose(g, o, rtol=1e-11, atol=1e-12 * max(1.0, np.abs(o).max()))


@pytest.mark.parametrize("dist,bvec,cvec", [("cauchy", 1, 1), ("cauchy", 0, 0), ("uniform", 1, 1), ("uniform", 0, 0),
                                             ("bernoulli", 1, 0), ("bernoulli", 0, 0)])
@pytest.mark.parametrize("strict", [False, True])
def test_logpdf_kernels_out_of_range(fb, dist, bvec, cvec, strict):
    # cauchy.hpp:151, uniform.hpp:159, bernoulli.hpp:145: -inf and no adjoint update -- through the optimistic pass +
    # undo (default; adjoints start at zero here, so "taken back" is exact) and through the FADB_STRICT_SIGMA pre-pass
    v, ov, outs = _lp_case(fb, dist, 2049, bvec, cvec, bad=True, flags=fb.STRICT_SIGMA if strict else 0)
    assert v == -math.inf and ov == -math.inf
    for g, o in outs:
        assert np.all(g == 0) and np.all(o == 0)


# ------------------------------------------------------------------ K5 regression
def test_linreg_full_size_properties(fb):
    # BASELINE config 4 size: X 2^20 x 64 (512 MiB), the TMA kernel.  Integer-valued X and w, dyadic residuals and
    # sigma = 2 make X w, X^T r and every adjoint exactly representable whatever the summation order: w_adj and
    # sigma_adj are compared bit for bit with the same quantities formed by torch, the value to 1e-12.
    import torch
    rows, cols = 1 << 20, 64
    f64 = dict(dtype=torch.float64, device="cuda")
    i = torch.arange(rows, device="cuda")
    j = torch.arange(cols, device="cuda")
    X = (((i[None, :] * 7 + j[:, None] * 3) % 5) - 2).to(torch.float64)         # (cols, rows) == rows x cols column-major
    w = ((j % 3) - 1).to(torch.float64)
    r = torch.tensor([0.5, -1.0, 1.5, -2.0, 2.0, -1.5, 1.0, -0.5], **f64)[i & 7]
    y = w @ X + r
    sigma, seed = torch.tensor([2.0], **f64), 0.75
    sum_r2 = 15.0 * rows / 8.0
    w_adj, s_adj = torch.full((cols,), 0.5, **f64), torch.full((1,), 0.5, **f64)
    v = fb.linreg_nll(X, y, w, w_adj, sigma, s_adj, seed=seed)
    assert rel(v, -0.5 * sum_r2 / 4.0 - rows * math.log(2.0)) <= REL_REDUCE
    assert torch.equal(w_adj, 0.5 + seed * 0.25 * (X @ r))
    assert float(s_adj[0]) == 0.5 + seed * (sum_r2 / 4.0 - rows) / 2.0


@pytest.mark.parametrize("rows,cols", [(5, 2), (1000, 64), (4097, 64), (777, 7), (300, 130)])
def test_linreg_vs_oracle(fb, rows, cols):
    import torch
    if (rows, cols) == (5, 2):
        X, y, w = np.asfortranarray(kats.LINREG_X), kats.LINREG_Y.copy(), kats.LINREG_THETA.copy()
    else:
        X, y, w_true = synth.linreg_inputs(rows, cols)
        w = w_true + 0.01
    sg = np.array([1.3])
    gX = torch.from_numpy(np.ascontiguousarray(X.T)).cuda()    # (cols, rows) C-order == col-major
    gw_adj, gs_adj = zeros(cols), zeros(1)
    v = fb.linreg_nll(gX, dev(y), dev(w), gw_adj, dev(sg), gs_adj, seed=1.0)
    owa, osa = np.zeros(cols), np.zeros(1)
    ov = O.linreg(X, y, w, owa, sg, osa)
    assert rel(v, ov) <= REL_REDUCE
    np.testing.assert_allclose(host(gw_adj), owa, rtol=1e-11, atol=1e-9 * np.abs(owa).max())
    assert rel(host(gs_adj)[0], osa[0]) <= 1e-11


def test_linreg_readme_golden(fb):
    # README.md:623-662: loss 6655, adj [-1210, -12100]; here as -0.5*loss at sigma=1
    import torch
    gX = torch.from_numpy(np.ascontiguousarray(kats.LINREG_X.T)).cuda()
    wa, sa = zeros(2), zeros(1)
    v = fb.linreg_nll(gX, dev(kats.LINREG_Y), dev(kats.LINREG_THETA), wa, dev([1.0]), sa)
    assert_double_eq([v], [-0.5 * kats.LINREG_LOSS])
    assert_double_eq(host(wa), -0.5 * kats.LINREG_ADJ)


# ------------------------------------------------------------------ K6 network
def test_mlp3_mathematica_goldens(fb):
    import torch
    gX = torch.from_numpy(np.ascontiguousarray(kats.NN_X.T)).cuda()
    grad = zeros(25)
    loss = fb.mlp3(gX, dev(kats.NN_Y), dev(kats.NN_PARM), grad)
    g = host(grad)
    assert rel(loss, 92030.04305391687) <= REL_REDUCE
    assert rel(g[0], kats.NN_GRAD_A1_11) <= REL_REDUCE      # GenTestData.nb:691-700
    assert rel(g[3], kats.NN_GRAD_A1_12) <= REL_REDUCE


@pytest.mark.parametrize("batch", [1, 257, 1 << 16])
def test_mlp3_vs_oracle(fb, batch):
    import torch
    X, y = synth.mlp3_inputs(batch)
    gX = torch.from_numpy(np.ascontiguousarray(X.T)).cuda()
    grad = zeros(25)
    loss = fb.mlp3(gX, dev(y), dev(kats.NN_PARM), grad)
    og = np.zeros(25)
    ol = O.mlp3(X, y, kats.NN_PARM.copy(), og, nthreads=4)
    assert rel(loss, ol) <= REL_REDUCE
    np.testing.assert_allclose(host(grad), og, rtol=1e-10, atol=1e-12 * np.abs(og).max())


def test_shutdown_releases_workspaces_and_reinit_reproduces(fb):
    # fadb_shutdown frees the kernels' lazily allocated per-device workspaces (regression partial rows, network
    # partial rows, dense-normal factors, mailboxes); after fadb_init every path must allocate afresh and give
    # the same bits
    import torch
    L = fb.lib()

    def run():
        rows, cols = 5002, 7
        X, y, w_true = synth.linreg_inputs(rows, cols, seed=3)
        wa, sa = zeros(cols), zeros(1)
        v1 = fb.linreg_nll(torch.from_numpy(np.ascontiguousarray(X.T)).cuda(), dev(y), dev(w_true + 0.01), wa, dev([1.3]), sa)
        Xn, yn = synth.mlp3_inputs(300)
        grad = zeros(25)
        v2 = fb.mlp3(torch.from_numpy(np.ascontiguousarray(Xn.T)).cuda(), dev(yn), dev(kats.NN_PARM), grad)
        return v1, host(wa), host(sa), v2, host(grad)
    a = run()
    torch.cuda.synchronize()
    L.fadb_shutdown()
    fb.check(L.fadb_init(0))
    fb.use_torch_stream()
    b = run()
    assert a[0] == b[0] and a[3] == b[3]
    for i in (1, 2, 4):
        np.testing.assert_array_equal(a[i], b[i])


# ------------------------------------------------------------------ ABI behaviour
def test_error_reporting(fb):
    import ctypes as C
    L = fb.lib()
    v = C.c_double()
    rc = L.fadb_sum_unary(12345, 4, None, None, 1.0, 0, C.byref(v))
    assert rc == -1 and b"fadb_sum_unary" in L.fadb_last_error()
    rc = L.fadb_normal_lpdf(9, 4, None, None, None, None, None, None, 1.0, 0, C.byref(v))
    assert rc == -1
    before = L.fadb_launch_count()
    fb.sum_unary("exp", dev([1.0, 2.0]), zeros(2))
    assert L.fadb_launch_count() == before + 1
