"""Oracle restatement of the quadrature rules (test infrastructure).

Gauss-Legendre: autopdex/seeder.py:370-1044 tabulates roots/weights on [0,1] with
ceil((order+1)/2) points; seeder.py:1046-1062 maps them to [-1,1] (2x-1, 2w) and
seeder.py:324-367 builds tensor rules with x fastest, then y, then z.  Here the
roots are computed (numpy leggauss) instead of tabulated; agreement with the
reference tables to <= 2 ulp is pinned by tests/golden/reference_tables.json.
"""
import math

import numpy as np


def gauss_legendre_1d(order):
    """Points/weights on [0,1] exact for polynomials up to `order` (seeder.py:370-377)."""
    n = int(math.ceil((order + 1) / 2))
    x, w = np.polynomial.legendre.leggauss(n)
    return 0.5 * (x + 1.0), 0.5 * w


def gauss_legendre_nd(dimension, order):
    x01, w01 = gauss_legendre_1d(order)
    x, w = 2.0 * x01 - 1.0, 2.0 * w01          # seeder.py:1053-1057
    if dimension == 1:
        return x, w
    n = x.shape[0]
    if dimension == 2:
        pts = np.array([[x[a], x[b]] for b in range(n) for a in range(n)])
        wts = np.array([w[a] * w[b] for b in range(n) for a in range(n)])
        return pts, wts
    if dimension == 3:
        pts = np.array([[x[a], x[b], x[c]] for c in range(n) for b in range(n) for a in range(n)])
        wts = np.array([w[a] * w[b] * w[c] for c in range(n) for b in range(n) for a in range(n)])
        return pts, wts
    raise NotImplementedError(dimension)
