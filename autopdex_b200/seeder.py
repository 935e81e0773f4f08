"""Quadrature rules used by the b200 backend (host side, NumPy).

Mirrors the names of autopdex.seeder that the hot path needs:
gauss_legendre_1d (seeder.py:370-1044), gauss_legendre_nd (seeder.py:1046-1062),
tensor ordering of seeder.tensor_product_rule (seeder.py:324-367: x fastest).
Roots are computed with Golub-Welsch (numpy) rather than tabulated.
"""
import math

import numpy as np


def gauss_legendre_1d(order):
    """Gauss-Legendre rule on [0, 1], exact up to polynomial degree `order`
    (ceil((order+1)/2) points), as autopdex.seeder.gauss_legendre_1d."""
    if order < 1:
        raise ValueError("Quadrature order not implemented")
    n = int(math.ceil((order + 1) / 2))
    x, w = np.polynomial.legendre.leggauss(n)
    return (x + 1.0) / 2.0, w / 2.0


def gauss_legendre_nd(dimension, order):
    """Tensor Gauss-Legendre rule on [-1, 1]^dimension -> (points, weights)."""
    p, w = gauss_legendre_1d(order)
    p, w = 2.0 * p - 1.0, 2.0 * w
    if dimension == 1:
        return p, w
    if dimension not in (2, 3):
        raise NotImplementedError("Not implemented for this dimensionality!")
    n = p.shape[0]
    grids = np.meshgrid(*([np.arange(n)] * dimension), indexing="ij")
    # flat index = ix + n*iy (+ n*n*iz): x runs fastest
    idx = [g.transpose(*reversed(range(dimension))).ravel() for g in grids]
    pts = np.stack([p[i] for i in idx], axis=1)
    wts = np.ones(n ** dimension)
    for i in idx:
        wts = wts * w[i]
    return pts, wts
