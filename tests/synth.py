"""Fixed-seed synthetic inputs shared by tests and bench.py (SURVEY.md §8d distributions)."""
import numpy as np

SEED = 0xFA57AD


def black_scholes_inputs(n, seed=SEED):
    rng = np.random.default_rng(seed + 1)
    return dict(S=rng.uniform(50, 150, n), K=rng.uniform(50, 150, n), sigma=rng.uniform(0.05, 0.6, n),
                tau=rng.uniform(0.05, 2.0, n), r=rng.uniform(0.0, 0.08, n))


def sum_inputs(n, seed=SEED):
    rng = np.random.default_rng(seed + 2)
    return rng.uniform(0.5, 2.0, n)


def normal_inputs(n, seed=SEED):
    rng = np.random.default_rng(seed + 3)
    return rng.normal(size=n), rng.normal(size=n), rng.uniform(0.5, 2.0, n)


def linreg_inputs(rows, cols, seed=SEED):
    rng = np.random.default_rng(seed + 4)
    X = np.asfortranarray(rng.normal(size=(rows, cols)))
    w_true = rng.normal(size=cols)
    y = X @ w_true + 0.1 * rng.normal(size=rows)
    return X, y, w_true


def mlp3_inputs(batch, seed=SEED):
    rng = np.random.default_rng(seed + 5)
    X = np.asfortranarray(rng.uniform(0, 50, size=(batch, 2)))
    y = 2 * X[:, 0] + 3 * X[:, 1]           # example/ceres_3layer_neural_net/src.cpp:76-79
    return X, np.ascontiguousarray(y)
