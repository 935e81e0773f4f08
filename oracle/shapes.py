"""Oracle restatement of the Lagrange shape functions (test infrastructure).

Reference: autopdex/spaces.py:302-12191 (fem_iso_line_quad_brick) and
:12193-15012 (fem_iso_line_tri_tet) hold sympy-generated closed forms.  The oracle
does not re-type them: it builds the unique Lagrange basis on the reference's node
positions by inverting a Vandermonde matrix on the matching polynomial space.
Node orders: line2/3 spaces.py:354,360; quad4 :1913; quad9 :1924;
hex8 :12057; hex27 :12083-12151; tri3/6 :13750,13755; tet4/10 :14627,14632.
Agreement with the reference's generated code is pinned by
tests/golden/reference_tables.json (values evaluated from the reference source).
"""
import itertools

import numpy as np

_Q = [(-1, -1), (1, -1), (1, 1), (-1, 1)]
_H = [(-1, -1, -1), (1, -1, -1), (1, 1, -1), (-1, 1, -1), (-1, -1, 1), (1, -1, 1), (1, 1, 1), (-1, 1, 1)]

REF_NODES = {
    "line2": [(-1,), (1,)],
    "line3": [(-1,), (1,), (0,)],
    "quad4": _Q,
    "quad9": _Q + [(0, -1), (1, 0), (0, 1), (-1, 0), (0, 0)],
    "hex8": _H,
    "hex27": _H + [(0, -1, -1), (1, 0, -1), (0, 1, -1), (-1, 0, -1),
                   (0, -1, 1), (1, 0, 1), (0, 1, 1), (-1, 0, 1),
                   (-1, -1, 0), (1, -1, 0), (1, 1, 0), (-1, 1, 0),
                   (-1, 0, 0), (1, 0, 0), (0, -1, 0), (0, 1, 0), (0, 0, -1), (0, 0, 1),
                   (0, 0, 0)],
    "tri3": [(0, 0), (1, 0), (0, 1)],
    "tri6": [(0, 0), (1, 0), (0, 1), (.5, 0), (.5, .5), (0, .5)],
    "tet4": [(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1)],
    "tet10": [(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1),
              (.5, 0, 0), (.5, .5, 0), (0, .5, 0), (0, 0, .5), (.5, 0, .5), (0, .5, .5)],
}
_ORDER = {"line2": 1, "line3": 2, "quad4": 1, "quad9": 2, "hex8": 1, "hex27": 2,
          "tri3": 1, "tri6": 2, "tet4": 1, "tet10": 2}
_SIMPLEX = {"tri3", "tri6", "tet4", "tet10"}


def element_name(nen, dim, family):
    """family: 'quad_brick' (spaces.fem_iso_line_quad_brick) or 'tri_tet'."""
    table = {("quad_brick", 1, 2): "line2", ("quad_brick", 1, 3): "line3",
             ("tri_tet", 1, 2): "line2", ("tri_tet", 1, 3): "line3",
             ("quad_brick", 2, 4): "quad4", ("quad_brick", 2, 9): "quad9",
             ("quad_brick", 3, 8): "hex8", ("quad_brick", 3, 27): "hex27",
             ("tri_tet", 2, 3): "tri3", ("tri_tet", 2, 6): "tri6",
             ("tri_tet", 3, 4): "tet4", ("tri_tet", 3, 10): "tet10"}
    return table[(family, dim, nen)]


def _exponents(name):
    dim = len(REF_NODES[name][0])
    p = _ORDER[name]
    if name in _SIMPLEX:
        return [e for e in itertools.product(range(p + 1), repeat=dim) if sum(e) <= p]
    return list(itertools.product(range(p + 1), repeat=dim))


def _monomials(expo, xi):
    xi = np.atleast_2d(xi)
    return np.stack([np.prod(xi ** np.asarray(e), axis=1) for e in expo], axis=1)


def _dmonomials(expo, xi):
    xi = np.atleast_2d(xi)
    dim = xi.shape[1]
    out = np.zeros((xi.shape[0], len(expo), dim))
    for m, e in enumerate(expo):
        for d in range(dim):
            if e[d] == 0:
                continue
            ee = list(e)
            ee[d] -= 1
            out[:, m, d] = e[d] * np.prod(xi ** np.asarray(ee), axis=1)
    return out


def shape_tables(name, xi):
    """N (n_pts, nen) and dN/dxi (n_pts, nen, dim_ref) at reference points xi."""
    nodes = np.asarray(REF_NODES[name], dtype=np.float64)
    xi = np.asarray(xi, dtype=np.float64).reshape(-1, nodes.shape[1])
    expo = _exponents(name)
    coef = np.linalg.inv(_monomials(expo, nodes))      # column a = coefficients of N_a
    N = _monomials(expo, xi) @ coef
    dN = np.einsum("pmd,ma->pad", _dmonomials(expo, xi), coef)
    return N, dN


def simplex_physical_tables(x_eval, x_nodes):
    """P1/P2 Lagrange values and PHYSICAL gradients for 'sparse' (integration point) mode.

    Restates spaces.fem_ini_simplex (spaces.py:15194-15296): a complete polynomial of
    order 1 (3/4 nodes) or 2 (6/10 nodes) in coordinates shifted to the evaluation point
    is fitted through the nodal values; with as many nodes as monomials the least-squares
    fit interpolates, so value = constant coefficient and gradient = linear coefficients.
    x_eval (n_pts, dim), x_nodes (n_pts, nen, dim) -> N (n_pts, nen), dNdx (n_pts, nen, dim).
    """
    x_eval = np.asarray(x_eval, dtype=np.float64)
    x_nodes = np.asarray(x_nodes, dtype=np.float64)
    n_pts, nen, dim = x_nodes.shape
    p = {3: 1, 6: 2}[nen] if dim == 2 else {4: 1, 10: 2}[nen]
    expo = [e for e in itertools.product(range(p + 1), repeat=dim) if sum(e) <= p]
    N = np.empty((n_pts, nen))
    dN = np.empty((n_pts, nen, dim))
    i0 = expo.index((0,) * dim)
    lin = [expo.index(tuple(1 if q == d else 0 for q in range(dim))) for d in range(dim)]
    for q in range(n_pts):
        V = _monomials(expo, x_nodes[q] - x_eval[q])   # (nen, n_monomials)
        C = np.linalg.inv(V)                           # coefficients = C @ f
        N[q] = C[i0]
        for d in range(dim):
            dN[q, :, d] = C[lin[d]]
    return N, dN
