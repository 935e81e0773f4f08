"""Lagrange shape-function tables for the element families of the b200 backend.

`fem_iso_line_quad_brick` and `fem_iso_line_tri_tet` are the *tokens* a user passes as
`ansatz_fun` (same names as autopdex.spaces, spaces.py:302 and :12193); the backend never
calls them -- it recognises them and tabulates N and dN/dxi at the Gauss points with the
explicit formulas below (node orders of SURVEY.md 8a row a12).
"""
import numpy as np


class _SpaceToken:
    def __init__(self, family):
        self.family = family
        self.__name__ = "fem_iso_line_" + family

    def __call__(self, *a, **k):
        raise RuntimeError("%s is evaluated on the device by the b200 backend; it cannot be called on the host"
                           % self.__name__)

    def __repr__(self):
        return "<autopdex_b200.spaces.%s>" % self.__name__


fem_iso_line_quad_brick = _SpaceToken("quad_brick")
fem_iso_line_tri_tet = _SpaceToken("tri_tet")

# reference positions in {-1, 0, +1} of the tensor-product elements, in the reference's node order
_QUAD = [(-1, -1), (1, -1), (1, 1), (-1, 1)]
_HEX = [(-1, -1, -1), (1, -1, -1), (1, 1, -1), (-1, 1, -1), (-1, -1, 1), (1, -1, 1), (1, 1, 1), (-1, 1, 1)]
_TENSOR_NODES = {
    (1, 2): [(-1,), (1,)],
    (1, 3): [(-1,), (1,), (0,)],
    (2, 4): _QUAD,
    (2, 9): _QUAD + [(0, -1), (1, 0), (0, 1), (-1, 0), (0, 0)],
    (3, 8): _HEX,
    (3, 27): _HEX + [(0, -1, -1), (1, 0, -1), (0, 1, -1), (-1, 0, -1), (0, -1, 1), (1, 0, 1), (0, 1, 1), (-1, 0, 1),
                     (-1, -1, 0), (1, -1, 0), (1, 1, 0), (-1, 1, 0),
                     (-1, 0, 0), (1, 0, 0), (0, -1, 0), (0, 1, 0), (0, 0, -1), (0, 0, 1), (0, 0, 0)],
}


def _lagrange_1d(order, pos, x):
    """value and derivative at x of the 1-D Lagrange polynomial attached to node `pos`."""
    if order == 1:
        return 0.5 * (1.0 + pos * x), 0.5 * pos * np.ones_like(x)
    if pos == 0:
        return 1.0 - x * x, -2.0 * x
    return 0.5 * x * (x + pos), x + 0.5 * pos


def _tensor_tables(dim, nen, xi):
    nodes = _TENSOR_NODES[(dim, nen)]
    order = 1 if nen == 2 ** dim else 2
    n_pts = xi.shape[0]
    N = np.ones((n_pts, nen))
    dN = np.ones((n_pts, nen, dim))
    for a, node in enumerate(nodes):
        for d in range(dim):
            v, dv = _lagrange_1d(order, node[d], xi[:, d])
            N[:, a] *= v
            for k in range(dim):
                dN[:, a, k] *= dv if k == d else v
    return N, dN


def _simplex_tables(dim, nen, xi):
    n_pts = xi.shape[0]
    lam = np.concatenate([1.0 - xi.sum(axis=1, keepdims=True), xi], axis=1)      # barycentric, node 0 first
    dlam = np.concatenate([-np.ones((1, dim)), np.eye(dim)], axis=0)             # (dim+1, dim)
    nv = dim + 1
    if nen == nv:
        return lam.copy(), np.broadcast_to(dlam, (n_pts, nv, dim)).copy()
    # quadratic: vertices lam(2 lam - 1), then mid-edge nodes 4 lam_i lam_j in the reference's edge order
    edges = [(0, 1), (1, 2), (0, 2)] if dim == 2 else [(0, 1), (1, 2), (0, 2), (0, 3), (1, 3), (2, 3)]
    N = np.empty((n_pts, nen))
    dN = np.empty((n_pts, nen, dim))
    for v in range(nv):
        N[:, v] = lam[:, v] * (2.0 * lam[:, v] - 1.0)
        dN[:, v, :] = (4.0 * lam[:, v] - 1.0)[:, None] * dlam[v]
    for e, (i, j) in enumerate(edges):
        N[:, nv + e] = 4.0 * lam[:, i] * lam[:, j]
        dN[:, nv + e, :] = 4.0 * (lam[:, i][:, None] * dlam[j] + lam[:, j][:, None] * dlam[i])
    return N, dN


def shape_tables(family, nen, dim_ref, xi):
    """N (n_pts, nen) and dN/dxi (n_pts, nen, dim_ref) of the element (family, nen, dim_ref)."""
    xi = np.asarray(xi, dtype=np.float64).reshape(-1, dim_ref)
    if dim_ref == 1 or family == "quad_brick":
        if (dim_ref, nen) not in _TENSOR_NODES:
            raise ValueError("element with %d nodes in %d-D is not a supported Q1/Q2 line/quad/brick" % (nen, dim_ref))
        return _tensor_tables(dim_ref, nen, xi)
    if family == "tri_tet":
        if (dim_ref, nen) not in ((2, 3), (2, 6), (3, 4), (3, 10)):
            raise ValueError("element with %d nodes in %d-D is not a supported P1/P2 triangle/tetrahedron" % (nen, dim_ref))
        return _simplex_tables(dim_ref, nen, xi)
    raise ValueError("unknown element family %r" % family)


def simplex_physical_tables(x_eval, x_nodes):
    """P1/P2 shape values and PHYSICAL gradients of 'fem simplex' in integration-point mode.

    spaces.fem_ini_simplex (spaces.py:15194-15296) fits a complete polynomial of order 1 or 2 in
    coordinates shifted to the evaluation point through the nodal values; with as many nodes as
    monomials the fit interpolates, so N = row of the inverse Vandermonde matrix belonging to the
    constant monomial and dN/dx_d = row of the monomial x_d.  Batched over integration points.
    x_eval (n, dim), x_nodes (n, nen, dim) -> N (n, nen), dNdx (n, nen, dim).
    """
    x_eval = np.asarray(x_eval, dtype=np.float64)
    x_nodes = np.asarray(x_nodes, dtype=np.float64)
    n, nen, dim = x_nodes.shape
    order = {(2, 3): 1, (2, 6): 2, (3, 4): 1, (3, 10): 2}.get((dim, nen))
    if order is None:
        raise ValueError("'fem simplex' with %d nodes in %d-D is not supported (P1/P2 only)" % (nen, dim))
    y = x_nodes - x_eval[:, None, :]
    cols = [np.ones((n, nen))] + [y[:, :, d] for d in range(dim)]
    if order == 2:
        for d in range(dim):
            for e in range(d, dim):
                cols.append(y[:, :, d] * y[:, :, e])
    V = np.stack(cols, axis=2)                      # (n, nen, n_monomials == nen)
    Cinv = np.linalg.inv(V)                         # coefficients = Cinv @ nodal values
    return Cinv[:, 0, :].copy(), np.stack([Cinv[:, 1 + d, :] for d in range(dim)], axis=2)
