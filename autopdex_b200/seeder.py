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
_RULES = None


def _simplex_rule(name, order):
    """Tabulated reference-simplex rules of autopdex.seeder (triangle: seeder.py:1813-2285, rules of
    mathsfromnothing.au; tetrahedron: seeder.py:2288-3421, Jaskowiec and Sukumar 2020), orders 1-10, same points in the
    same order, same weights to the last digit: autopdex_b200/data/simplex_rules.npz holds the outputs of the
    reference functions themselves (tools/make_simplex_rules.py)."""
    global _RULES
    if _RULES is None:
        import os
        _RULES = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "simplex_rules.npz")))
    key = "%s_%d_x" % (name, order)
    if key not in _RULES:
        raise ValueError("Quadrature order not implemented")
    return _RULES[key].copy(), _RULES["%s_%d_w" % (name, order)].copy()


def int_pts_ref_tri(order):
    """autopdex.seeder.int_pts_ref_tri: (points (n, 2), weights (n,)) on the reference triangle, weights sum to 1/2."""
    return _simplex_rule("tri", order)


def int_pts_ref_tet(order):
    """autopdex.seeder.int_pts_ref_tet: (points (n, 3), weights (n,)) on the reference tetrahedron, weights sum to
    1/6 -- the reference's (asymmetric) Jaskowiec-Sukumar tables, order 2 included."""
    return _simplex_rule("tet", order)


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
