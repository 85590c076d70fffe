"""Deterministic synthetic inputs of SURVEY.md section 8d for the developer tools (same generator as bench.py and
oracle.synthetic_problem; the tools that only time or trace the GPU path do not import the oracle)."""
import numpy as np


def synthetic_problem(N, D, seed_x=0, seed_y=1, sigma_n=0.3):
    """X ~ U[0,1)^(N,D); y = sin 3x0 + cos 2x1 + sum_{d>=2} x_d + sigma_n N(0,1), shape (N, 1)."""
    X = np.random.default_rng(seed_x).random((N, D))
    f = np.sin(3 * X[:, 0])
    if D > 1:
        f = f + np.cos(2 * X[:, 1])
    if D > 2:
        f = f + X[:, 2:].sum(axis=1)
    y = f + sigma_n * np.random.default_rng(seed_y).standard_normal(N)
    return X, y.reshape(-1, 1)
