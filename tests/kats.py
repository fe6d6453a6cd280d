"""Known-answer data transcribed from the reference's own tests (data, not code).

All `file:line` are relative to /root/reference/.
"""
import math

import numpy as np

# test/testutil/base_fixture.hpp:101-127
FIX_SCL = 2.31
FIX_VEC = np.array([3.1, -2.3, 1.3, 0.0, 5.1])
FIX_MAT = np.array([[3.1, -2.3, 1.3], [0.2, 5.1, -0.9]])

# test/reverse/stat/normal_unittest.cpp:71-88
N_X, N_MU, N_SIGMA = 2.31, -0.2, 0.01
N_VX = np.array([3.1, -2.3, 1.3])
N_VMU = np.array([-0.3, -2.3, -1.2])
N_VSIGMA = np.array([0.01, 1.03, 2.41])

# normal_unittest.cpp:105-297 — (value, x_adj, mu_adj, sigma_adj) at seed 1
NORMAL_KAT = {
    "sss": (-31495.8948298140203406, [-25100.0], [25100.0], [6300000.0000000009313226]),
    "vss": (-87736.1844894420355558, [-33000.0, 20999.9999999999963620, -15000.0],
            [27000.0000000000036380], [17549700.0]),
    "vvs": (-89036.1844894420355558, [-34000.0, 0.0, -25000.0], [34000.0, -0.0, 25000.0],
            [17809700.0]),
    "vsv": (-54448.5761343555350322, [-33000.0, 1.9794514091808839, -0.2582600161842943],
            [32998.2788086070067948], [10889900.0, 3.0649009313396647, -0.2541950106736757]),
    "vvv": (-57796.8420570641465019, [-34000.0, 0.0, -0.4304333603071572],
            [34000.0, -0.0, 0.4304333603071572],
            [11559900.0, -0.9708737864077670, 0.0315698758372999]),
}

# test/reverse/core/unary_unittest.cpp:320-324
ERF_KAT = (4.0, 0.9999999845827420, 1.2698234671866558e-7)

# README.md:582-607 quadratic form; README.md:623-662 + test/util/GenTestData.nb:216-221 linreg
LINREG_X = np.array([[1, 10], [2, 20], [3, 30], [4, 40], [5, 50]], dtype=float)
LINREG_Y = np.array([32, 64, 96, 128, 160], dtype=float)
LINREG_THETA = np.array([1.0, 2.0])
LINREG_LOSS = 6655.0
LINREG_ADJ = np.array([-1210.0, -12100.0])

# test/util/GenTestData.nb:330ff (full-precision parameters, column-major order of
# example/ceres_3layer_neural_net/src.cpp:21-27) and :691-700 (gradients)
NN_PARM = np.array([
    0.04398400506863087, 0.9601256114996133, -0.5209410945118056,
    -0.8005256013595985, -0.028791364112533024, 0.6358093220436887,       # A1 3x2
    0.5846033162880331, -0.4433822901819533, 0.2243037009072486,          # b1
    0.9750495768482685, -0.8240837134824699, 0.23629950009691214,
    0.6663923726986263, -0.4988280521070929, -0.7814276530077855,
    -0.9110531473087247, -0.23015593969194725, -0.13636652333316235,      # A2 3x3
    0.26342534615305047, 0.8415352118678241, 0.9203421214171716,          # b2
    0.6562896796607487, 0.848247693352548, -0.7486969470512386,           # A3 1x3
    0.21522047772523667])                                                  # b3
NN_X = LINREG_X
NN_Y = LINREG_Y
NN_GRAD_A1_11 = -2953.051309234279   # GenTestData.nb:691-700
NN_GRAD_A1_12 = -29530.51309234279

# example/black_scholes.cpp:43-70 inputs; closed-form values
BS_POINT = dict(S=105.0, K=100.0, sigma=5.0, tau=30.0 / 365, r=0.0125)
BS_KAT = dict(call=56.513603067773943, delta_call=0.77381844492127361,
              put=51.410916100933306, delta_put=-0.22618155507872639,
              vega=9.0549342770573631, rho_call=2.0332055053939553,
              rho_put=-6.1775325521259914)


def ulp_diff(a, b):
    """|a-b| in units of ulp(b) (EXPECT_DOUBLE_EQ passes at <= 4)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    sp = np.spacing(np.maximum(np.abs(b), np.finfo(np.float64).tiny))
    return np.abs(a - b) / sp


def assert_double_eq(a, b, ulps=4, what=""):
    a = np.atleast_1d(np.asarray(a, dtype=np.float64))
    b = np.atleast_1d(np.asarray(b, dtype=np.float64))
    assert a.shape == b.shape, (what, a.shape, b.shape)
    same = (a == b) | (np.isnan(a) & np.isnan(b))
    d = np.where(same, 0.0, ulp_diff(a, b))
    assert np.all(d <= ulps), f"{what}: max ulp diff {d.max()} at {int(d.argmax())}: {a[d.argmax()]!r} vs {b[d.argmax()]!r}"

# ---- det / log_det fixtures, test/reverse/core/det_unittest.cpp:34-54 (log_det_unittest.cpp same), seed :15
DET_SEED = 9.2313
DET_FPLU = np.array([[2., 3, 1, 5], [3, 5, -1, 3], [-2, 3, 1, 0], [-1, -1, 2, 7]])
DET_LLT = np.array([[8., 3, 1, 5], [3, 5, 1, 3], [1, 1, 1, 0], [5, 3, 0, 7]])
# ---- wishart fixture, test/reverse/stat/wishart_unittest.cpp:37-55 (lower triangle set, then mirrored), n = 4
WISHART_X = np.array([[10., 2, 3], [2, 10, 1], [3, 1, 10]])
WISHART_V = np.array([[5., 1, 0], [1, 5, 1], [0, 1, 5]])
WISHART_N = 4
WISHART_VALUE = -12.55942947411780252764      # wishart_unittest.cpp:66 (ref/wishart_ref.py)
