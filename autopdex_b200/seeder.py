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


# ---- integration points in simplex / line meshes ('sparse' assembling mode) ------------------------
def int_pts_ref_tri(order):
    """Reference-triangle rule (weights sum to 1/2).  Orders 1-2 coincide with autopdex.seeder.int_pts_ref_tri
    (seeder.py:1813-1835); higher orders are not tabulated here -- pass your own points in `settings`."""
    if order == 1:
        return np.array([[1 / 3, 1 / 3]]), np.array([0.5])
    if order == 2:
        return np.array([[1 / 6, 1 / 6], [1 / 6, 2 / 3], [2 / 3, 1 / 6]]), np.full(3, 1 / 6)
    raise NotImplementedError("triangle rule of order %d: supply 'integration coordinates' yourself" % order)


def int_pts_ref_tet(order):
    """Reference-tetrahedron rule (weights sum to 1/6).  Order 1 coincides with the reference; order 2 is the
    classical symmetric 4-point rule (the reference tabulates an asymmetric one, seeder.py:2303-2325)."""
    if order == 1:
        return np.array([[0.25, 0.25, 0.25]]), np.array([1 / 6])
    if order == 2:
        a, b = 0.1381966011250105, 0.5854101966249685
        return np.array([[a, a, a], [b, a, a], [a, b, a], [a, a, b]]), np.full(4, 1 / 24)
    raise NotImplementedError("tetrahedron rule of order %d: supply 'integration coordinates' yourself" % order)


def _int_pts_in_mesh(x_nodes, elem, ref_pts, ref_w, nv):
    """Affine map of a reference rule into every element using its first `nv` nodes; weights scaled by the
    absolute size ratio (seeder.py:3456-3488).  Returns (x_int, weights, n_int, connectivity) with one row per
    integration point, element-major (seeder.py:3573-3583)."""
    x_nodes, elem = np.asarray(x_nodes, dtype=np.float64), np.asarray(elem)
    X = x_nodes[elem[:, :nv]]
    edges = X[:, 1:, :] - X[:, :1, :]                              # (n_e, nv-1, dim)
    if nv == 2:
        ratio = np.linalg.norm(edges[:, 0], axis=1)
    elif nv == 3:
        if x_nodes.shape[1] == 2:
            ratio = np.abs(edges[:, 0, 0] * edges[:, 1, 1] - edges[:, 0, 1] * edges[:, 1, 0])
        else:
            ratio = np.linalg.norm(np.cross(edges[:, 0], edges[:, 1]), axis=1)
    else:
        ratio = np.abs(np.einsum("ni,ni->n", np.cross(edges[:, 0], edges[:, 1]), edges[:, 2]))
    x_int = X[:, None, 0, :] + np.einsum("pk,nkd->npd", np.atleast_2d(ref_pts), edges)
    w = ratio[:, None] * np.asarray(ref_w)[None, :]
    n_pts = len(ref_w)
    conn = np.repeat(elem, n_pts, axis=0)
    return x_int.reshape(-1, x_nodes.shape[1]), w.ravel(), w.size, conn


def int_pts_in_line_mesh(x_nodes, elem, order):
    p, w = gauss_legendre_1d(order)
    return _int_pts_in_mesh(x_nodes, elem, p[:, None], w, 2)


def int_pts_in_tri_mesh(x_nodes, elem, order):
    return _int_pts_in_mesh(x_nodes, elem, *int_pts_ref_tri(order), 3)


def int_pts_in_tet_mesh(x_nodes, elem, order):
    return _int_pts_in_mesh(x_nodes, elem, *int_pts_ref_tet(order), 4)
