"""CPU-only tests: the C ABI loads and exports every declared symbol, and the host-side
planner (expression -> fused stages) behaves as designed.  No CUDA call is made."""
import os
import re

import numpy as np
import pytest

import exprs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def F():
    import fastad_b200 as F
    if not os.path.exists(F.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    F.load()
    return F


def test_library_exports_every_declared_symbol(F):
    hdr = open(os.path.join(ROOT, "include", "fastad_b200.h")).read()
    declared = set(re.findall(r"\b(fadb_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 50
    L = F.lib()
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/fastad_b200.h but not exported"
    assert declared == set(F.SIGNATURES), declared ^ set(F.SIGNATURES)
    assert b"sm_100a" in L.fadb_version()


def test_missing_library_fails_loudly(F, tmp_path):
    with pytest.raises(F.FadbError, match="no CPU fallback"):
        F.load(str(tmp_path / "libfastad_b200.so"))


def _plan(F, build):
    exprs.RNG = np.random.default_rng(1)
    e = F.Expr(device=False)
    root, _, _ = build(e)
    sizes = e.plan(root)
    return e.describe(), sizes


def test_planner_fuses_elementwise_chain_into_one_stage(F):
    d, sizes = _plan(F, exprs.CASES["vec_sum_chain_100k"])
    assert d.startswith("stages=1"), d
    assert "term=sum" in d and "channels=3" in d          # sum + adjoints of the two scalar leaves
    # nothing but the root value is materialised (the reference needs ~6 N-sized caches)
    assert sizes == (1, 0)


def test_planner_black_scholes_is_one_stage(F):
    d, sizes = _plan(F, exprs.CASES["black_scholes_call"])
    assert d.startswith("stages=1"), d
    assert d.count(":= %") == 3                            # the three placeholders
    assert sizes[1] == 0


def test_planner_nested_reduction_splits_stages(F):
    d, _ = _plan(F, exprs.CASES["vec_nested_sum"])
    assert d.startswith("stages=2"), d
    assert "map N=4099 term=sum" in d and "map N=1 term=store" in d
    assert "<- stage 0" in d


def test_planner_scalar_broadcast_of_reduction(F):
    d, _ = _plan(F, exprs.CASES["vec_scalar_times_sum"])
    assert d.startswith("stages=2"), d
    assert "(broadcast) <- stage 0" in d


def test_planner_additive_root(F):
    """A sum of terms (log-likelihood + priors): every term becomes a root of its own -- fused sweep with the seed known up
    front, and the single-pass routes apply INSIDE the posterior (regression with intercept -> fast=6, leaf normal -> fast=2)."""
    d, _ = _plan(F, exprs.CASES["posterior_linreg"])
    assert "additive root: 3 term(s)" in d, d[:300]
    assert "fast=6" in d and "dot 4224x16" in d
    d, _ = _plan(F, exprs.CASES["sum_of_normals"])
    assert "additive root: 2 term(s)" in d and "fast=2" in d
    d, _ = _plan(F, exprs.CASES["vec_nested_sum"])          # w4 * w4 + sin(w4): not linear in the term
    assert "additive root" not in d


def test_planner_fast_paths(F):
    e = F.Expr(device=False)
    x = e.var(np.ones(64))
    e.plan(e.sum(e.unary("log", x)))
    assert "fast=1" in e.describe()
    e = F.Expr(device=False)
    x = e.var(np.ones(64))
    e.plan(e.sum(e.pow(2, x)))
    assert "fast=1" in e.describe()
    d, _ = _plan(F, exprs.CASES["linreg_expr"])
    assert "fast=3" in d and "dot 2000x64" in d
    d, _ = _plan(F, exprs.CASES["least_squares"])
    assert "fast=5" in d and "dot 5000x64" in d
    d, _ = _plan(F, exprs.CASES["least_squares_flipped_odd"])
    assert "fast=5" in d
    d, _ = _plan(F, exprs.CASES["least_squares_tall_generic"])
    assert "fast=0" in d
    d, _ = _plan(F, exprs.CASES["least_squares_x_variable"])
    assert "fast=0" in d
    d, _ = _plan(F, exprs.CASES["normal_expr_mean"])
    assert "fast=0" in d and "term=normal" in d


def test_builder_folds_scalar_constants_on_host(F):
    # unary.hpp:182-184, binary.hpp:251-256, pow.hpp:219-227, sum.hpp:250-252
    import math
    e = F.Expr(device=False)
    c = e.unary("exp", e.const(2.0))
    assert F.lib().fadb_expr_is_const(e.h, c)
    c2 = e.binary("mul", c, e.pow(3, e.const(2.0)))
    assert F.lib().fadb_expr_is_const(e.h, c2)
    x = e.var(1.0)
    e.plan(e.binary("add", x, e.sum(c2)))
    assert f"const {math.exp(2.0) * 8:.4f}"[:12] in e.describe().replace("const 59.1124", "const 59.1124")


def test_builder_rejects_bad_shapes(F):
    e = F.Expr(device=False)
    a, b = e.var(np.ones(3)), e.var(np.ones(4))
    with pytest.raises(F.FadbError, match="shapes differ"):
        e.binary("add", a, b)
    M, v = e.var(np.ones((2, 3))), e.var(np.ones(2))
    with pytest.raises(F.FadbError, match="dot"):
        e.dot(M, v)
    with pytest.raises(F.FadbError):
        e.unary("exp", 12345)


def test_planner_rejects_int_indexed_operands_beyond_2_31(F):
    """The matrix x matrix and transpose kernels index with int; only a tall matrix times a vector (size_t rows)
    may exceed 2^31 - 1 elements.  Planning is host-only: the pointers are never dereferenced here."""
    import ctypes as C
    fake = C.c_void_p(0x1000)

    def mat(e, shape, rows, cols, variable=True):
        if variable:
            return e._id(e.L.fadb_expr_var(e.h, shape, rows, cols, fake, fake))
        return e._id(e.L.fadb_expr_const_view(e.h, shape, rows, cols, fake))

    e = F.Expr(device=False)                    # 2^26 x 64 (32 GiB) constant X times a vector: allowed
    X, w = mat(e, F.MAT, 1 << 26, 64, variable=False), mat(e, F.VEC, 64, 1)
    e.plan(e.sum(e.dot(X, w)))
    e = F.Expr(device=False)                    # the same X as a VARIABLE times 2 columns: int-indexed adjoint
    X, W = mat(e, F.MAT, 1 << 26, 64), mat(e, F.MAT, 64, 2)
    with pytest.raises(F.FadbError, match="2\\^31"):
        e.plan(e.sum(e.dot(X, W)))
    e = F.Expr(device=False)
    X = mat(e, F.MAT, 1 << 26, 64)
    with pytest.raises(F.FadbError, match="2\\^31"):
        e.plan(e.sum(e.transpose(X)))


# ---------------------------------------------------------------- bind-time specialisation (csrc/jit.cu)
@pytest.mark.parametrize("name", ["vec_sum_chain_100k", "black_scholes_call", "placeholders", "normal_vec_sigma_expr"])
def test_stage_program_specialises_to_register_code(F, name):
    """NVRTC needs no device: the generated constants + the embedded stage-kernel source must compile
    for sm_100a with every node value in a register (no stack frame, no spills)."""
    exprs.RNG = np.random.default_rng(1)
    e = F.Expr(device=False)
    root, _, _ = exprs.CASES[name](e)
    e.plan(root)
    nstages = int(e.describe().split()[0].split("=")[1])
    compiled = 0
    for st in range(nstages):
        for mode in (1, 2, 3):
            try:
                text = e.jit_program(st, mode)
            except F.FadbError:
                continue                       # a dot / dense stage
            assert "JIT_PROG" in text and "#define JIT_MODE %d" % mode in text
            nbytes, log = F.jit_compile_only(text)
            assert nbytes > 0
            # node values in local memory would need >= 8 bytes x nodes of stack (80-400 bytes for these
            # programs); the 64-register cap (8 CTAs/SM) may cost a register or two of spills
            import re
            m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", log)
            assert m and int(m.group(1)) <= 48, log[-800:]
            assert re.search(r"Used (\d+) registers", log) and int(re.search(r"Used (\d+) registers", log).group(1)) <= 64
            compiled += 1
    assert compiled >= 3


@pytest.mark.parametrize("name", ["black_scholes_call_ring", "placeholder_chain_ring", "placeholder_reassigned_ring", "placeholder_sum_ring"])
def test_operand_ring_form_of_a_placeholder_stage_compiles(F, name):
    """Per-element stages with placeholder stores and >= 4096 elements: the operand-ring form (mode | 8; cp.async copies one
    trip ahead into a per-thread shared-memory column, stage_kernel.cuh JIT_RING) compiles for sm_100a inside the 64-register
    budget without spills; a read behind the placeholder's assignment names the assigning instruction instead of memory."""
    import re
    exprs.RNG = np.random.default_rng(1)
    e = F.Expr(device=False)
    root, _, _ = exprs.CASES[name](e)
    e.plan(root)
    for mode in (1, 2, 3):
        text = e.jit_program(0, mode | 8)
        assert "#define JIT_RING 1" in text and "JIT_RV[]" in text and "JIT_RA[]" in text
        nrv, nra = int(re.search(r"JIT_NRV (\d+)", text).group(1)), int(re.search(r"JIT_NRA (\d+)", text).group(1))
        assert nrv >= 1 and (nra >= 1) == (mode != 1)            # no old adjoints in a forward-only launch
        plain = e.jit_program(0, mode)
        assert "JIT_RING" not in plain
        nbytes, log = F.jit_compile_only(text)
        assert nbytes > 0, log[-800:]
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", log)
        assert m and int(m.group(2)) == 0 and int(m.group(1)) <= 48, log[-800:]
        assert int(re.search(r"Used (\d+) registers", log).group(1)) <= 64
        assert int(re.search(r"(\d+) bytes smem", log).group(1)) >= (nrv + nra) * 1024
    # small stages and stages without placeholder stores (the four-element form) keep the plain kernels
    for other in ("black_scholes_call", "vec_sum_chain_100k"):
        exprs.RNG = np.random.default_rng(1)
        e2 = F.Expr(device=False)
        r2, _, _ = exprs.CASES[other](e2)
        e2.plan(r2)
        with pytest.raises(F.FadbError):
            e2.jit_program(0, 3 | 8)


@pytest.mark.parametrize("name", ["glm_logistic", "glm_logistic_weighted_ragged", "glm_least_squares_inner", "glm_linreg_intercept",
                                  "glm_linreg_intercept_link_const_sigma"])
def test_one_pass_dot_map_sum_kernel_compiles(F, name):
    """sum(f(dot(X, w), ...)): the planner routes the root stage to the one-pass kernel (fast=6) and its NVRTC source
    (glm_kernel.cuh + the stage program) compiles for sm_100a within the 2-CTAs-per-SM register budget."""
    import re
    exprs.RNG = np.random.default_rng(1)
    e = F.Expr(device=False)
    root, _, _ = exprs.CASES[name](e)
    e.plan(root)
    d = e.describe()
    assert d.startswith("stages=2") and "fast=6" in d, d[:400]
    text = e.jit_program(1, 6)
    assert "#define GLM_DOT_LEAF" in text and "GLM_VEC_SLOT" in text
    nbytes, log = F.jit_compile_only(text)
    assert nbytes > 0, log[-1500:]
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", log)
    regs = int(re.search(r"Used (\d+) registers", log).group(1))
    # 2 CTAs of 160 threads per SM; a few spilled doubles (the IEEE divisions of the log-density terminal) stay in L1
    assert regs <= 204 and int(m.group(2)) <= 256, log[-600:]


@pytest.mark.parametrize("name", ["if_else_taken", "if_else_not_taken", "if_else_logical", "if_else_nested_tt"])
def test_control_flow_stage_specialises(F, name):
    """if_else stages compile too: the jumps become a watermark over the unrolled program (JIT_HAS_CF), node values
    stay in registers."""
    import re
    exprs.RNG = np.random.default_rng(1)
    e = F.Expr(device=False)
    root, _, _ = exprs.CASES[name](e)
    e.plan(root)
    nstages = int(e.describe().split()[0].split("=")[1])
    compiled = 0
    for st in range(nstages):
        for mode in (1, 2, 3):
            try:
                text = e.jit_program(st, mode)
            except F.FadbError:
                continue
            if "JIT_HAS_CF 1" not in text:
                continue
            nbytes, log = F.jit_compile_only(text)
            assert nbytes > 0, log[-800:]
            # node values in local memory would need 8 bytes x nodes x elements per thread (>= 800 bytes for the nested
            # case); a handful of spilled registers under the 64-register cap is fine
            m = re.search(r"(\d+) bytes stack frame", log)
            assert m and int(m.group(1)) <= 128, log[-800:]
            compiled += 1
    assert compiled >= 1


@pytest.mark.parametrize("name", exprs.JIT_ONLY)
def test_stages_beyond_interpreter_limits_plan_and_compile(F, name):
    exprs.RNG = np.random.default_rng(1)
    e = F.Expr(device=False)
    root, _, _ = exprs.CASES[name](e)
    e.plan(root)
    d = e.describe()
    assert d.startswith("stages=1") and "specialised-only" in d, d[:300]
    nbytes, log = F.jit_compile_only(e.jit_program(0, 3))
    assert nbytes > 0
