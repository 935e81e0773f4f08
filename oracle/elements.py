"""Oracle element residuals/tangents in closed form (test infrastructure).

The reference obtains these by JAX AD of the weak form / potential per Gauss point
(models.py:1654-1720 domain element, :1768-1850 surface element, :1230-1264 mixed
potential; variational_schemes.py:185-252 for integration-point mode).  Here the same
quantities are written out in closed form and vectorised
over elements with einsum.  Local dof order: node-major, component-minor
(assembler.py:129-131).

A *set* is a dict:
  kind   'domain' | 'surface' | 'intpoint'
  etype  'quad4', 'hex8', ... (oracle.shapes.REF_NODES)      [domain/surface]
  conn   (n_rows, nen) int   (rows = elements, or integration points for 'intpoint')
  nf     dofs per node
  gp     (xi (n_gp, dim_ref), w (n_gp,))                     [domain/surface]
  N, dNdx, w   per-row tables (n_rows, nen), (n_rows, nen, dim), (n_rows,)   [intpoint]
  model  dict(name=..., **parameters); parameters broadcast against (n_rows, n_gp[, ncomp])
"""
import numpy as np

from . import shapes


def _inv_det(J):
    """Closed-form inverse / determinant, as utility.matrix_inv/matrix_det (utility.py:656-817)."""
    d = J.shape[-1]
    if d == 1:
        det = J[..., 0, 0]
        return 1.0 / J, det
    if d == 2:
        a, b, c, e = J[..., 0, 0], J[..., 0, 1], J[..., 1, 0], J[..., 1, 1]
        det = a * e - b * c
        inv = np.empty_like(J)
        inv[..., 0, 0], inv[..., 0, 1], inv[..., 1, 0], inv[..., 1, 1] = e, -b, -c, a
        return inv / det[..., None, None], det
    a = J
    adj = np.empty_like(J)
    adj[..., 0, 0] = a[..., 1, 1] * a[..., 2, 2] - a[..., 1, 2] * a[..., 2, 1]
    adj[..., 0, 1] = a[..., 0, 2] * a[..., 2, 1] - a[..., 0, 1] * a[..., 2, 2]
    adj[..., 0, 2] = a[..., 0, 1] * a[..., 1, 2] - a[..., 0, 2] * a[..., 1, 1]
    adj[..., 1, 0] = a[..., 1, 2] * a[..., 2, 0] - a[..., 1, 0] * a[..., 2, 2]
    adj[..., 1, 1] = a[..., 0, 0] * a[..., 2, 2] - a[..., 0, 2] * a[..., 2, 0]
    adj[..., 1, 2] = a[..., 0, 2] * a[..., 1, 0] - a[..., 0, 0] * a[..., 1, 2]
    adj[..., 2, 0] = a[..., 1, 0] * a[..., 2, 1] - a[..., 1, 1] * a[..., 2, 0]
    adj[..., 2, 1] = a[..., 0, 1] * a[..., 2, 0] - a[..., 0, 0] * a[..., 2, 1]
    adj[..., 2, 2] = a[..., 0, 0] * a[..., 1, 1] - a[..., 0, 1] * a[..., 1, 0]
    det = a[..., 0, 0] * adj[..., 0, 0] + a[..., 0, 1] * adj[..., 1, 0] + a[..., 0, 2] * adj[..., 2, 0]
    return adj / det[..., None, None], det


def _par(model, key, n_rows, g, default=None, ncomp=None):
    """Fetch parameter `key` for Gauss point g, shape (n_rows,) or (n_rows, ncomp)."""
    v = model.get(key, default)
    if v is None:
        return None
    v = np.asarray(v, dtype=np.float64)
    if ncomp is None:
        if v.ndim == 0:
            return np.full(n_rows, float(v))
        if v.ndim == 1:                     # per Gauss point (same for all elements)
            return np.full(n_rows, v[g])
        return v[:, g]
    if v.ndim == 1:
        return np.broadcast_to(v, (n_rows, ncomp))
    if v.ndim == 2:                         # (n_gp, ncomp)
        return np.broadcast_to(v[g], (n_rows, ncomp))
    return v[:, g, :]


def elasticity_matrix(mode, Em, nu):
    """Voigt material matrix EXACTLY as models.py:570-601 (note plain strain uses
    C33 = 2*mu, i.e. coeff*(1-2nu) without the customary 1/2 -- reproduced)."""
    Em, nu = np.asarray(Em, float), np.asarray(nu, float)
    z = np.zeros_like(Em)
    if mode == "plain strain":
        mu = Em / (2 * (1 + nu))
        c1, c2 = 1 - 2 * nu, 1 - nu
        co = 2 * mu / c1
        C = np.array([[co * c2, co * nu, z], [co * nu, co * c2, z], [z, z, co * c1]])
    elif mode == "plain stress":
        co = Em / (1 - nu ** 2)
        C = np.array([[co, co * nu, z], [co * nu, co, z], [z, z, co * (1 - nu) / 2]])
    elif mode == "3d":
        co = Em / (1 + nu)
        c1 = 1 - 2 * nu
        c2, c3, c4 = co * (1 - nu) / c1, co * nu / c1, co * 0.5
        C = np.array([[c2, c3, c3, z, z, z], [c3, c2, c3, z, z, z], [c3, c3, c2, z, z, z],
                      [z, z, z, c4, z, z], [z, z, z, z, c4, z], [z, z, z, z, z, c4]])
    elif mode == "lame 2d":
        # P = d psi/dF of models.linear_elastic_strain_energy (models.py:1167-1185) inside hyperelastic_steady_state_weak
        # 'plain strain' (models.py:917-1000): sigma = lam tr(eps) 1 + 2 mu eps with eps_33 = 0 -- the customary plane-strain
        # matrix (shear mu on the engineering shear), NOT the matrix of linear_elasticity_weak above
        lam, mu = Em * nu / ((1 + nu) * (1 - 2 * nu)), Em / (2 * (1 + nu))
        C = np.array([[lam + 2 * mu, lam, z], [lam, lam + 2 * mu, z], [z, z, mu]])
    else:
        raise ValueError(mode)
    return np.moveaxis(C, (0, 1), (-2, -1))            # (..., nv, nv)


def _bmatrix(G):
    """Strain-displacement matrix, Voigt [11,22,12] / [11,22,33,12,13,23] with engineering
    shear (models.py:546-559).  G (n, nen, dim) -> B (n, nv, nen*dim)."""
    n, nen, dim = G.shape
    nv = 3 if dim == 2 else 6
    B = np.zeros((n, nv, nen, dim))
    if dim == 2:
        B[:, 0, :, 0] = G[:, :, 0]
        B[:, 1, :, 1] = G[:, :, 1]
        B[:, 2, :, 0] = G[:, :, 1]
        B[:, 2, :, 1] = G[:, :, 0]
    else:
        B[:, 0, :, 0] = G[:, :, 0]
        B[:, 1, :, 1] = G[:, :, 1]
        B[:, 2, :, 2] = G[:, :, 2]
        B[:, 3, :, 0] = G[:, :, 1]
        B[:, 3, :, 1] = G[:, :, 0]
        B[:, 4, :, 0] = G[:, :, 2]
        B[:, 4, :, 2] = G[:, :, 0]
        B[:, 5, :, 1] = G[:, :, 2]
        B[:, 5, :, 2] = G[:, :, 1]
    return B.reshape(n, nv, nen * dim)


def point_contribution(model, N, G, s, u, g, settings, want_tangent=True):
    """Residual/tangent contribution of one integration point for all rows.
    want_tangent=False: the Poisson models skip the element matrix (K is returned as None), as the reference's
    residual-only assembly does (assembler.py:587-637); the other models still form it.

    N (n, nen) shape values, G (n, nen, dim) physical gradients, s (n,) weight (signed
    w*detJ for domain elements, models.py:1691-1694), u (n, nen, nf) local dofs.
    Returns R (n, nen*nf), K (n, nen*nf, nen*nf).
    """
    name = model["name"]
    n, nen = N.shape
    nf = u.shape[2]
    if name in ("poisson_potential", "poisson_weak"):
        # potential: (1/2) c grad(phi).grad(phi) - f phi   (tests/test_dicts_as_dofs_user_potential.py:24-35)
        # weak:      -c grad(theta).grad(dtheta) + f dtheta (models.py:124-130)
        c = _par(model, "coefficient", n, g, 1.0)
        f = _par(model, "source", n, g, 0.0)
        if not want_tangent:
            gu = np.einsum("nad,na->nd", G, u[:, :, 0])
            R = (s * c)[:, None] * np.einsum("nad,nd->na", G, gu) - (s * f)[:, None] * N
            return (-R if name == "poisson_weak" else R), None
        GG = np.matmul(G, np.swapaxes(G, 1, 2))
        K = (s * c)[:, None, None] * GG
        R = np.einsum("nab,nb->na", K, u[:, :, 0]) - (s * f)[:, None] * N
        if name == "poisson_weak":
            K, R = -K, -R
        return R, K
    if name == "capacity":
        # -c dtheta (theta - theta_n)/dt   (models.py:1981-2008)
        c = _par(model, "coefficient", n, g, 1.0)
        dt = float(settings["time increment"])
        un = model["_dofs_n_local"]                    # (n, nen, 1), gathered by the caller
        NN = np.einsum("na,nb->nab", N, N)
        K = -(s * c / dt)[:, None, None] * NN
        R = np.einsum("nab,nb->na", K, u[:, :, 0] - un[:, :, 0])
        return R, K
    if name == "neumann":
        # -du . t   (models.py:766-777); zero tangent (models.py:1828-1832)
        t = _par(model, "traction", n, g, ncomp=nf)
        R = -(s[:, None, None] * N[:, :, None] * t[:, None, :]).reshape(n, nen * nf)
        return R, np.zeros((n, nen * nf, nen * nf))
    dim = G.shape[2]
    Em = _par(model, "youngs_modulus", n, g)
    nu = _par(model, "poisson_ratio", n, g)
    b = _par(model, "body_load", n, g, ncomp=nf) if model.get("body_load") is not None else None
    if name == "linear_elasticity":
        # sigma_voigt . deps_voigt - b . du   (models.py:605-633)
        # mode 'lame': the isotropic tensor lam 1x1 + 2 mu I_sym in the mesh's dimension (3-D: identical to '3d')
        mode = model["mode"] if model["mode"] != "lame" else ("3d" if dim == 3 else "lame 2d")
        C = elasticity_matrix(mode, Em, nu)
        B = _bmatrix(G)
        K = s[:, None, None] * np.einsum("npi,npq,nqj->nij", B, C, B)
        R = np.einsum("nij,nj->ni", K, u.reshape(n, nen * nf))
    elif name == "neo_hooke":
        # P : dF - b . du, P = d psi/dF (models.py:940-998), psi of models.py:1137-1146.
        lam = Em * nu / ((1 + nu) * (1 - 2 * nu))      # models.py:955-956
        mu = Em / (2 * (1 + nu))
        F = np.einsum("nai,naJ->niJ", u, G) + np.eye(dim)   # plain strain: in-plane block, F33 = 1
        Finv, J = _inv_det(F)
        FinvT = np.swapaxes(Finv, 1, 2)
        c1 = mu - 0.5 * lam * (J * J - 1.0)
        c2 = lam * J * J
        P = mu[:, None, None] * F - c1[:, None, None] * FinvT
        R = (s[:, None, None] * np.einsum("niJ,naJ->nai", P, G)).reshape(n, nen * nf)
        gp = np.einsum("niJ,naJ->nai", FinvT, G)       # pushed-forward gradients
        GG = np.einsum("naJ,nbJ->nab", G, G)
        K = (mu[:, None, None, None, None] * GG[:, :, None, :, None] * np.eye(dim)[None, None, :, None, :]
             + c1[:, None, None, None, None] * np.einsum("nak,nbi->naibk", gp, gp)
             + c2[:, None, None, None, None] * np.einsum("nai,nbk->naibk", gp, gp))
        K = (s[:, None, None, None, None] * K).reshape(n, nen * nf, nen * nf)
    else:
        raise ValueError("oracle: unknown model %r" % name)
    if b is not None:
        R = R - (s[:, None, None] * N[:, :, None] * b[:, None, :]).reshape(n, nen * nf)
    return R, K


def set_contributions(st, coords, dofs, settings, want_tangent=True, rows=None):
    """Element (or integration-point) residuals and tangents of one set.
    coords (n_nodes, dim), dofs (n_nodes, nf) -> Re (n_rows, ndof_e), Ke (n_rows, ndof_e, ndof_e) (None if the model
    skipped it, see point_contribution).  rows: slice of the set's rows (chunked / threaded assembly; only for sets whose
    model parameters are scalars or per-Gauss-point tables)."""
    conn = np.asarray(st["conn"])
    if rows is not None:
        conn = conn[rows]
    n, nen = conn.shape
    nf = st["nf"]
    if st["model"]["name"] == "pattern_only":
        # structural entries only: the explicit zero blocks the reference's BCOO holds for the field pairs of a
        # multi-field problem that an integrand does not couple (assembler.py:79-117)
        return np.zeros((n, nen * nf)), (np.zeros((n, nen * nf, nen * nf)) if want_tangent else None)
    u = dofs[conn].reshape(n, nen, nf)
    model = dict(st["model"])
    if model["name"] == "capacity":
        model["_dofs_n_local"] = np.asarray(settings["dofs n"], float)[conn].reshape(n, nen, nf)
    ndof = nen * nf
    if st["kind"] == "intpoint":
        # one row per integration point (assembler.py:874-924,978-1035); weights are
        # physical and positive (seeder.py:3473-3488), gradients are physical.
        assert rows is None
        return point_contribution(model, st["N"], st["dNdx"], np.asarray(st["w"], float), u, 0, settings)
    X = coords[conn]                                    # (n, nen, dim)
    xi, w = st["gp"]
    Nt, dNt = shapes.shape_tables(st["etype"], xi)
    Re, Ke = np.zeros((n, ndof)), (np.zeros((n, ndof, ndof)) if want_tangent else None)
    for g in range(len(w)):
        Jm = np.einsum("nad,ak->ndk", X, dNt[g])        # dX_d / dxi_k
        Nn = np.broadcast_to(Nt[g], (n, nen))
        if st["kind"] == "domain":
            Jinv, det = _inv_det(Jm)
            G = np.einsum("ak,nkd->nad", dNt[g], Jinv)
            s = w[g] * det                              # SIGNED (models.py:1691-1694)
        else:                                           # surface: models.py:1806-1821
            if X.shape[2] == 2:
                scal = np.sqrt(Jm[:, 0, 0] ** 2 + Jm[:, 1, 0] ** 2)
            else:
                scal = np.linalg.norm(np.cross(Jm[:, :, 0], Jm[:, :, 1]), axis=1)
            G = np.zeros((n, nen, X.shape[2]))
            s = w[g] * scal
        r, k = point_contribution(model, Nn, G, s, u, g, settings, want_tangent)
        Re += r
        if k is not None and Ke is not None:
            Ke += k
        elif Ke is not None:
            Ke = None
    return Re, Ke


def gauss_point_coordinates(st, coords):
    """Physical coordinates of every Gauss point, (n_rows, n_gp, dim): the value of
    ansatz_fun['physical coor'](x_int) in models.py:1244-1247."""
    X = np.asarray(coords, float)[np.asarray(st["conn"])]
    Nt, _ = shapes.shape_tables(st["etype"], st["gp"][0])
    return np.einsum("ga,nad->ngd", Nt, X)
