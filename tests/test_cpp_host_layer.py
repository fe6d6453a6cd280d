"""Builds and runs the C++17 host-layer tests (tests/cpp): the ad:: expression-template API of
include/fastad over the C ABI.  test_types is CPU-only (type layer + host planner); the others
need a B200."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tests", "_build")


@pytest.fixture(scope="module")
def built():
    if not os.path.exists(os.path.join(ROOT, "fastad_b200", "lib", "libfastad_b200.so")):
        import __graft_entry__ as g
        g.build()
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "cpp"), "-j4"], stdout=subprocess.DEVNULL)
    return BUILD


def _run(built, name):
    p = subprocess.run([os.path.join(built, name)], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-4000:] + p.stderr[-2000:]
    assert "0 failures" in p.stdout
    return p.stdout


def test_cpp_type_layer_and_planner(built):
    out = _run(built, "test_types")
    assert "test_types:" in out


def test_cpp_fastmath_accuracy_against_glibc(built):
    out = _run(built, "test_fastmath")
    assert "test_fastmath:" in out


def test_cpp_forward_mode_reference_tests(built):
    out = _run(built, "test_forward")
    assert "test_forward:" in out


@pytest.mark.gpu
def test_cpp_forward_mode_batched_on_device(built):
    out = _run(built, "test_forward_batch")
    assert "test_forward_batch:" in out


@pytest.mark.gpu
def test_cpp_host_layer_reference_tests(built):
    out = _run(built, "test_host_layer")
    assert "test_host_layer:" in out


@pytest.mark.gpu
def test_cpp_reference_black_scholes_example_unchanged(built):
    out = _run(built, "example_black_scholes")
    assert "example_black_scholes:" in out


@pytest.mark.gpu
def test_cpp_reference_main_example_patterns(built):
    out = _run(built, "example_main")
    assert "example_main:" in out


@pytest.mark.gpu
def test_cpp_type_fused_kernels_match_c_abi_path(built):
    out = _run(built, "test_fused")
    assert "test_fused:" in out


@pytest.mark.gpu
def test_fastmath_accuracy_on_the_device(built, capsys):
    # fastmath.cuh was only validated on the host; this runs exp_t / log_t / phi_pair_pw / div_ /
    # sqrt_ / rcp_ ON the GPU, built -fmad=true like black_scholes.cu, over 2^21 points against glibc (<= 1 ulp)
    out = _run(built, "test_fastmath_device")
    assert "test_fastmath_device:" in out
    with capsys.disabled():
        print("\n" + out.strip().splitlines()[0])
