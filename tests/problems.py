"""Problem builders shared by the oracle tests and the GPU parity tests.

Each builder returns plain NumPy data (coords, connectivity, masks, parameters) in the
reference's conventions, so that the same inputs feed the oracle and the b200 backend.
"""
import numpy as np

from oracle import mesher as omesh
from oracle import quadrature as oquad
from oracle import elements as oel


def readme_source(x):
    """Source term of examples/miscellaneous/short_example.py:23-33 and
    tests/test_dicts_as_dofs_user_potential.py:31-34 (x: (..., 2))."""
    x2 = x - np.array([1.0, 0.5])
    return 20.0 * (np.sin(10.0 * np.sum(x * x, axis=-1)) - np.cos(10.0 * np.sum(x2 * x2, axis=-1)))


def readme_poisson(n):
    """README / G1 problem on an n x n quad mesh of the unit square.  Dirichlet mask = boundary
    nodes minus the 4 corners (geometry.psdf_polygon is NaN at corners)."""
    pts = [[0., 0.], [1., 0.], [1., 1.], [0., 1.]]
    coords, elems = omesh.structured_mesh((n, n), pts, "quad")
    gp = oquad.gauss_legendre_nd(2, 2)
    st = dict(kind="domain", etype="quad4", conn=elems, nf=1, gp=gp,
              model=dict(name="poisson_potential"))
    st["model"]["source"] = readme_source(oel.gauss_point_coordinates(st, coords))
    tol = 1e-9
    on_b = ((np.abs(coords[:, 0]) < tol) | (np.abs(coords[:, 0] - 1) < tol)
            | (np.abs(coords[:, 1]) < tol) | (np.abs(coords[:, 1] - 1) < tol))
    corner = ((np.abs(coords[:, 0]) < tol) | (np.abs(coords[:, 0] - 1) < tol)) & \
             ((np.abs(coords[:, 1]) < tol) | (np.abs(coords[:, 1] - 1) < tol))
    mask = (on_b & ~corner)[:, None]
    return dict(sets=[st], coords=coords, mask=mask, values=np.zeros(mask.shape), nf=1)


# --- G2: Cook's membrane, hard-coded mesh of
# tests/test_user_elem_impl_diff_and_adaptive_load_step.py:31-55 (mesh DATA, not code) ---
COOK_NODES = np.array([
    [0., 0.], [48., 44.], [48., 60.], [0., 44.], [12., 11.], [24., 22.], [36., 33.], [6., 5.5], [18., 16.5],
    [30., 27.5], [42., 38.5], [48., 52.], [48., 48.], [48., 56.], [36., 56.], [24., 52.], [12., 48.], [42., 58.],
    [30., 54.], [18., 50.], [6., 46.], [0., 33.], [0., 22.], [0., 11.], [0., 38.5], [0., 27.5], [0., 16.5],
    [0., 5.5], [22.78099159, 39.33936105], [10.25815892, 28.78971905], [40.00000145, 47.99999935],
    [11.44621914, 36.55774821], [30.81792875, 42.11576444], [7.62152819, 16.21744291], [23.3904958, 30.66968052],
    [17.11360537, 37.94855463], [10.85218903, 32.67373363], [17.12907946, 25.39485952], [17.12134241, 31.67170708],
    [8.93984355, 22.50358098], [9.81076409, 13.60872146], [13.46992178, 19.50179049], [38.00000072, 51.99999968],
    [44.00000072, 49.99999968], [43.00000036, 53.99999984], [33.40896438, 37.55788222], [26.79946017, 40.72756274],
    [28.39973009, 34.11378137], [38.00000072, 40.49999968], [43.00000036, 44.24999984], [11.72310957, 42.2788741],
    [23.3904958, 45.66968052], [17.55680268, 43.97427731], [5.72310957, 34.7788741], [5.86155479, 40.38943705],
    [33.40896438, 49.05788222], [35.70448255, 44.77894095], [5.12907946, 25.39485952], [5.42609452, 30.08686681],
    [28.39973009, 47.36378137], [3.81076409, 13.60872146], [4.46992178, 19.50179049], [4.90538205, 9.55436073]])
COOK_ELEMS = np.array([
    [5, 28, 31, 29, 34, 35, 36, 37, 38], [4, 5, 29, 33, 8, 37, 39, 40, 41], [2, 14, 30, 11, 17, 42, 43, 13, 44],
    [5, 6, 32, 28, 9, 45, 46, 34, 47], [6, 1, 11, 30, 10, 12, 43, 48, 49], [15, 16, 31, 28, 19, 50, 35, 51, 52],
    [16, 3, 21, 31, 20, 24, 53, 50, 54], [6, 30, 14, 32, 48, 42, 55, 45, 56], [22, 29, 31, 21, 57, 36, 53, 25, 58],
    [14, 15, 28, 32, 18, 51, 46, 55, 59], [23, 33, 29, 22, 60, 39, 57, 26, 61], [33, 23, 0, 4, 60, 27, 7, 40, 62]])
COOK_SURF = np.array([
    [0, 4, 7], [4, 5, 8], [5, 6, 9], [6, 1, 10], [1, 11, 12], [11, 2, 13], [2, 14, 17],
    [14, 15, 18], [15, 16, 19], [16, 3, 20], [3, 21, 24], [21, 22, 25], [22, 23, 26], [23, 0, 27]])


def cook_g2(q0=20.0, Em=100.0, nu=0.3):
    """G2 set-up: Q2 plain-strain neo-Hooke + line3 Neumann on x=48."""
    coords = COOK_NODES
    mask = np.repeat((np.abs(coords[:, 0]) < 1e-6)[:, None], 2, axis=1)
    neumann = COOK_SURF[np.all(np.abs(coords[COOK_SURF][:, :, 0] - 48.0) < 1e-6, axis=1)]
    dom = dict(kind="domain", etype="quad9", conn=COOK_ELEMS, nf=2, gp=oquad.gauss_legendre_nd(2, 4),
               model=dict(name="neo_hooke", mode="plain strain", youngs_modulus=Em, poisson_ratio=nu))
    sur = dict(kind="surface", etype="line3", conn=neumann, nf=2, gp=oquad.gauss_legendre_nd(1, 4),
               model=dict(name="neumann", traction=np.array([0.0, q0])))
    return dict(sets=[dom, sur], coords=coords, mask=mask, values=np.zeros(mask.shape), nf=2, q0=q0)


def _boundary_mask_box(coords, lo, hi, tol=1e-9):
    on = np.zeros(coords.shape[0], dtype=bool)
    for d in range(coords.shape[1]):
        on |= (np.abs(coords[:, d] - lo[d]) < tol) | (np.abs(coords[:, d] - hi[d]) < tol)
    return on


UNIT_CUBE = [[0., 0., 0.], [1., 0., 0.], [1., 1., 0.], [0., 1., 0.], [0., 0., 1.], [1., 0., 1.], [1., 1., 1.], [0., 1., 1.]]


def poisson_hex(n, etype="hex8", distort=0.0, source=1.0):
    """BASELINE config 4 family: 3-D Poisson (models.poisson_weak as 'user element'), n^3 bricks on the
    unit cube, unit coefficient, constant source, homogeneous Dirichlet on all six faces."""
    coords, elems = omesh.structured_mesh((n, n, n), UNIT_CUBE, "brick")
    order = 2
    if etype == "hex27":
        coords, elems = omesh.elevate_bricks(coords, elems)
        order = 4
    if distort:
        rng = np.random.default_rng(3)
        inner = ~_boundary_mask_box(coords, [0, 0, 0], [1, 1, 1])
        coords = coords + inner[:, None] * rng.uniform(-distort, distort, coords.shape) / n
    st = dict(kind="domain", etype=etype, conn=elems, nf=1, gp=oquad.gauss_legendre_nd(3, order),
              model=dict(name="poisson_weak", coefficient=1.0, source=source))
    mask = _boundary_mask_box(coords, [0, 0, 0], [1, 1, 1])[:, None]
    return dict(sets=[st], coords=coords, mask=mask, values=np.zeros(mask.shape), nf=1)


def boundary_faces_x1(n, etype="quad4"):
    """quad4 faces of the structured n^3 brick mesh on the plane x = 1 (i = n), nodes (j,k),(j+1,k),(j+1,k+1),(j,k+1)."""
    sy, sx = n + 1, (n + 1) * (n + 1)
    J, K = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    J, K = J.ravel(), K.ravel()
    nid = lambda dj, dk: n * sx + (J + dj) * sy + (K + dk)
    return np.stack([nid(0, 0), nid(1, 0), nid(1, 1), nid(0, 1)], axis=1).astype(np.int64)


def neo_hooke_brick(n, traction=(0.0, 0.0, -2.0), Em=100.0, nu=0.3, model="neo_hooke"):
    """BASELINE config 5 family: n^3 hex8 brick, clamped at x=0, traction on the face x=1."""
    coords, elems = omesh.structured_mesh((n, n, n), UNIT_CUBE, "brick")
    dom = dict(kind="domain", etype="hex8", conn=elems, nf=3, gp=oquad.gauss_legendre_nd(3, 2),
               model=dict(name=model, mode="3d", youngs_modulus=Em, poisson_ratio=nu))
    sur = dict(kind="surface", etype="quad4", conn=boundary_faces_x1(n), nf=3, gp=oquad.gauss_legendre_nd(2, 2),
               model=dict(name="neumann", traction=np.asarray(traction, float)))
    mask = np.repeat((np.abs(coords[:, 0]) < 1e-9)[:, None], 3, axis=1)
    return dict(sets=[dom, sur], coords=coords, mask=mask, values=np.zeros(mask.shape), nf=3)


def elasticity_quad(n, mode="plain strain", etype="quad4", Em=100.0, nu=0.3, body=(0.0, -1.0)):
    """2-D linear elasticity (models.linear_elasticity_weak) on Cook's membrane geometry."""
    pts = [[0., 0.], [48., 44.], [48., 60.], [0., 44.]]
    coords, elems = omesh.structured_mesh((n, n), pts, "quad")
    order = 2
    if etype == "quad9":
        coords, elems = omesh.elevate_quads(coords, elems)
        order = 4
    dom = dict(kind="domain", etype=etype, conn=elems, nf=2, gp=oquad.gauss_legendre_nd(2, order),
               model=dict(name="linear_elasticity", mode=mode, youngs_modulus=Em, poisson_ratio=nu,
                          body_load=np.asarray(body, float)))
    mask = np.repeat((np.abs(coords[:, 0]) < 1e-9)[:, None], 2, axis=1)
    values = np.zeros(mask.shape)
    return dict(sets=[dom], coords=coords, mask=mask, values=values, nf=2)


# ---- 'sparse' (integration point) mode: transient heat conduction on simplices -----------------------
def _simplex_intpoints(coords, elems, ref_pts, ref_w, nv):
    """Oracle-side restatement of seeder.int_pts_in_{line,tri,tet}_mesh (seeder.py:3456-3619): affine map through
    the first nv nodes, weight = |size ratio| * w_ref, one connectivity row per integration point (element-major)."""
    X = coords[elems[:, :nv]]
    x_int, w, conn = [], [], []
    for e in range(elems.shape[0]):
        d = X[e, 1:] - X[e, 0]
        if nv == 2:
            ratio = np.linalg.norm(d[0])
        elif nv == 3:
            ratio = abs(np.linalg.det(d)) if d.shape[1] == 2 else np.linalg.norm(np.cross(d[0], d[1]))
        else:
            ratio = abs(np.linalg.det(d))
        for p, wr in zip(np.atleast_2d(ref_pts), ref_w):
            x_int.append(X[e, 0] + p @ d)
            w.append(ratio * wr)
            conn.append(elems[e])
    return np.asarray(x_int), np.asarray(w), np.asarray(conn)


TRI_RULE = (np.array([[1 / 6, 1 / 6], [1 / 6, 2 / 3], [2 / 3, 1 / 6]]), np.full(3, 1 / 6))
_A, _B = 0.1381966011250105, 0.5854101966249685
TET_RULE = (np.array([[_A, _A, _A], [_B, _A, _A], [_A, _B, _A], [_A, _A, _B]]), np.full(4, 1 / 24))


def heat_sparse(dim=2, n=4, order=1, dt=0.2, inflow=-3.0, capacity=0.1, conduct=1.5, seed=0):
    """Transient heat conduction as in examples/heat_conduction/maze_backward_euler.py:308-351: three 'sparse'
    sets (poisson_weak conduction, forward_backward_euler_weak capacity, neumann_weak inflow on the face x=1)
    on P1/P2 triangles or tetrahedra from the structured mesher.  Dirichlet theta=0 on the face x=0."""
    from oracle import shapes as oshapes
    if dim == 2:
        coords, elems = omesh.structured_mesh((n, n), [[0., 0.], [1., 0.], [1.2, 1.], [0., 1.]], "tri")
        if order == 2:
            coords, elems = omesh.elevate_triangles(coords, elems)
        rule, nv = TRI_RULE, 3
    else:
        coords, elems = omesh.structured_mesh((n, n, n), UNIT_CUBE, "tet")
        rule, nv = TET_RULE, 4
    x_int, w_int, conn = _simplex_intpoints(coords, elems, rule[0], rule[1], nv)
    N, dN = oshapes.simplex_physical_tables(x_int, coords[conn])
    # boundary facets on x = max: sub-simplices of the elements whose nodes all lie on that face
    xmax = coords[:, 0].max()
    if dim == 2:
        on = np.abs(coords[:, 0] - (1.0 + 0.2 * coords[:, 1])) < 1e-9      # right edge of the trapezoid
        face_of = lambda el: [el[list(c)] for c in ([0, 1], [1, 2], [2, 0]) if on[el[list(c)]].all()]
        frule, fnv = (np.array([[0.21132486540518713], [0.7886751345948129]]), np.array([0.5, 0.5])), 2
    else:
        on = np.abs(coords[:, 0] - xmax) < 1e-9
        face_of = lambda el: [el[list(c)] for c in ([0, 1, 2], [0, 1, 3], [0, 2, 3], [1, 2, 3]) if on[el[list(c)]].all()]
        frule, fnv = TRI_RULE, 3
    faces = np.asarray([f for el in elems[:, :nv] for f in face_of(el)])
    xs, ws, sconn = _simplex_intpoints(coords, faces, frule[0], frule[1], fnv)
    # surface shape functions: P1 on the facet in its own affine coordinates (exact for straight facets)
    Ns = np.zeros((xs.shape[0], fnv))
    for q in range(xs.shape[0]):
        Xf = coords[sconn[q]]
        A = np.vstack([np.ones(fnv), (Xf - Xf[0]).T])
        Ns[q] = np.linalg.lstsq(A, np.concatenate([[1.0], xs[q] - Xf[0]]), rcond=None)[0]
    rng = np.random.default_rng(seed)
    theta_n = rng.uniform(0.0, 1.0, (coords.shape[0], 1))
    cond = dict(kind="intpoint", conn=conn, nf=1, N=N, dNdx=dN, w=w_int,
                model=dict(name="poisson_weak", coefficient=conduct, source=0.0))
    cap = dict(kind="intpoint", conn=conn, nf=1, N=N, dNdx=dN, w=w_int, model=dict(name="capacity", coefficient=capacity))
    sur = dict(kind="intpoint", conn=sconn, nf=1, N=Ns, dNdx=np.zeros(Ns.shape + (dim,)), w=ws,
               model=dict(name="neumann", traction=np.array([inflow])))
    mask = (np.abs(coords[:, 0]) < 1e-9)[:, None]
    settings = {"time increment": dt, "dofs n": theta_n}
    return dict(sets=[cond, cap, sur], coords=coords, mask=mask, values=np.zeros(mask.shape), nf=1,
                settings=settings, x_int=(x_int, x_int, xs), w_int=(w_int, w_int, ws))
