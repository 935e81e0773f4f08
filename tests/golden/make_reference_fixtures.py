#!/usr/bin/env python
"""Generate tests/golden/reference_fixtures.npz by EXECUTING the unmodified reference (AutoPDEx under
/root/reference) on top of the NumPy stand-in for JAX in fakejax.py.  Runs only in the build container
(the reference tree does not travel to the GPU box); the .npz it writes is committed.

    python tests/golden/make_reference_fixtures.py [case ...]

Every array saved here is an OUTPUT OF THE REFERENCE'S OWN CODE (assembler.assemble_residual /
assemble_tangent / _get_indices, solver.solver, mesher, seeder, spaces, geometry selectors); see fakejax.py
for what is substituted (the AD engine is numerical: tangents ~1e-9, everything else ~1e-14).
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import fakejax  # noqa: E402

jax = fakejax.install()
import flax  # noqa: E402
import jax.numpy as jnp  # noqa: E402
from autopdex import assembler, geometry, mesher, models, seeder, solver, spaces, utility  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests import problems  # noqa: E402  (mesh DATA of the Cook test only)

OUT = os.path.join(HERE, "reference_fixtures.npz")
A = lambda x: np.asarray(x)


def bcoo_arrays(mat):
    return A(mat.data).astype(float), A(mat.indices)[:, 0].astype(np.int64), A(mat.indices)[:, 1].astype(np.int64)


def case_tables(out):
    for d in (1, 2, 3):
        for o in (1, 2, 3, 4, 5, 6):
            x, w = seeder.gauss_legendre_nd(dimension=d, order=o)
            out["gauss_%d_%d_x" % (d, o)], out["gauss_%d_%d_w" % (d, o)] = A(x), A(w)
    for o in (1, 2):
        x, w = seeder.int_pts_ref_tri(o)
        out["tri_rule_%d_x" % o], out["tri_rule_%d_w" % o] = A(x), A(w)
        x, w = seeder.int_pts_ref_tet(o)
        out["tet_rule_%d_x" % o], out["tet_rule_%d_w" % o] = A(x), A(w)
    rng = np.random.default_rng(11)
    for fam, fn, cases in (("quad_brick", spaces.fem_iso_line_quad_brick, [(1, 2), (1, 3), (2, 4), (2, 9), (3, 8), (3, 27)]),
                           ("tri_tet", spaces.fem_iso_line_tri_tet, [(2, 3), (2, 6), (3, 4), (3, 10)])):
        for dim, nen in cases:
            xi = rng.uniform(-0.8, 0.8, (5, dim)) if fam == "quad_brick" else rng.uniform(0.05, 0.3, (5, dim))
            N = np.stack([A(fn(jnp.asarray(p[0] if dim == 1 else p), jnp.zeros((nen, dim)), jnp.eye(nen), None, False, dim))
                          for p in xi])   # 1-D elements take a scalar xi (vmap over gauss_legendre_nd(1, .))
            out["shape_%s_%d_%d_xi" % (fam, dim, nen)], out["shape_%s_%d_%d_N" % (fam, dim, nen)] = xi, N
    cube = [[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1.2], [1, 1, 1], [0, 1, 1]]
    quad = [[0, 0], [2, 0], [2.5, 1.5], [0, 1]]
    for name, args in (("quad", ((3, 4), quad, "quad")), ("tri", ((3, 4), quad, "tri")),
                       ("brick", ((2, 3, 4), cube, "brick")), ("tet", ((2, 3, 2), cube, "tet"))):
        c, e = mesher.structured_mesh(*args)
        out["mesh_%s_coords" % name], out["mesh_%s_elems" % name] = A(c), A(e).astype(np.int64)
    c, e = mesher.structured_mesh((2, 2, 2), cube, "brick")
    c2, e2 = mesher.elevate_mesh_order(c, e)
    out["mesh_hex27_coords"], out["mesh_hex27_elems"] = A(c2), A(e2).astype(np.int64)
    c, e = mesher.structured_mesh((3, 2), quad, "tri")
    c2, e2 = mesher.elevate_mesh_order(c, e)
    out["mesh_tri6_coords"], out["mesh_tri6_elems"] = A(c2), A(e2).astype(np.int64)
    # integration points in a triangle mesh ('sparse' mode input, seeder.py:3555-3585)
    x_int, w_int, n_int, conn = seeder.int_pts_in_tri_mesh(c, e, 2)
    out["intpts_tri_x"], out["intpts_tri_w"], out["intpts_tri_conn"] = A(x_int), A(w_int), A(conn).astype(np.int64)
    # array-dof COO index emission (assembler.py:123-141)
    conn = np.array([[0, 3, 4, 1], [1, 4, 5, 2]])
    out["indices_nf2"] = A(assembler._get_indices(jnp.asarray(conn), jnp.zeros((6, 2)))).astype(np.int64)


def case_indices_dict(out):
    """dict-dof COO index emission with TWO fields of different dofs per node and different connectivities
    (assembler.py:61-121: field offsets in key order, blocks `for field_i: for field_j:`)."""
    from autopdex import assembler
    cu = np.array([[0, 3, 4, 1], [1, 4, 5, 2], [3, 6, 7, 4]])
    cp = np.array([[0, 1, 2], [1, 2, 3], [2, 3, 0]])
    dofs = {"u": jnp.zeros((8, 2)), "p": jnp.zeros(4)}
    out["indices_dict_up"] = A(assembler._get_indices({"u": jnp.asarray(cu), "p": jnp.asarray(cp)}, dofs)).astype(np.int64)


def readme_case(n, out, tag):
    pts = [[0., 0.], [1., 0.], [1., 1.], [0., 1.]]
    coords, elems = mesher.structured_mesh((n, n), pts, "quad")
    node_coordinates, connectivity = {"phi": coords}, {"phi": elems}
    dirichlet_nodes = geometry.in_sdfs(node_coordinates["phi"], lambda x: geometry.psdf_polygon(x, pts))
    dirichlet_dofs = {"phi": dirichlet_nodes}
    dirichlet_conditions = utility.dict_zeros_like(dirichlet_dofs, dtype=jnp.float64)

    def integrand_fun(x_int, ansatz_fun, settings, static_settings, elem_number, set):   # short_example.py:23-33
        x = ansatz_fun["physical coor"](x_int)
        phi_fun = ansatz_fun["phi"]
        phi = phi_fun(x_int)
        dphi_dx = jax.jacrev(phi_fun)(x_int)
        x_1 = x
        x_2 = x - jnp.array([1., 0.5])
        source_term = 20 * (jnp.sin(10 * x_1 @ x_1) - jnp.cos(10 * x_2 @ x_2))
        return (1 / 2) * dphi_dx @ dphi_dx - source_term * phi

    user_potential = models.mixed_reference_domain_potential(
        integrand_fun, {"phi": spaces.fem_iso_line_quad_brick}, *seeder.gauss_legendre_nd(dimension=2, order=2), "phi")
    static_settings = flax.core.FrozenDict({"assembling mode": ("user potential",), "solution structure": ("nodal imposition",),
                                            "model": (user_potential,), "solver type": "newton", "solver backend": "scipy",
                                            "solver": "lapack", "verbose": -1})
    settings = {"connectivity": (connectivity,), "dirichlet dofs": dirichlet_dofs, "node coordinates": node_coordinates,
                "dirichlet conditions": dirichlet_conditions}
    rng = np.random.default_rng(5)
    dofs = {"phi": jnp.asarray(rng.uniform(-1, 1, coords.shape[0]))}
    R = assembler.assemble_residual(dofs, settings, static_settings)
    K = assembler.assemble_tangent(dofs, settings, static_settings)
    data, rows, cols = bcoo_arrays(K)
    sol, infos = solver.solver(utility.dict_zeros_like(dirichlet_dofs, dtype=jnp.float64), settings, static_settings)
    out.update({tag + "_mask": A(dirichlet_nodes), tag + "_dofs": A(dofs["phi"]), tag + "_R": A(R["phi"]),
                tag + "_K_data": data, tag + "_K_rows": rows, tag + "_K_cols": cols, tag + "_sol": A(sol["phi"]),
                tag + "_infos": np.array([float(infos[0]), float(infos[1]), float(infos[2])])})


def element_case(out, tag, coords, elems, nf, weak, ansatz, gp, surf=None, settings_extra=None, dofs_scale=0.02,
                 solve=False, dirichlet_dofs=None):
    """'user element' sets: domain element (+ optional surface set) -> R, K (BCOO) at random dofs, optional solve."""
    elem = models.isoparametric_domain_element_galerkin(weak, ansatz, *gp)
    model_list, conn_list, modes = [elem], [jnp.asarray(elems)], ["user element"]
    if surf is not None:
        s_elems, s_weak, s_gp = surf
        model_list.append(models.isoparametric_surface_element_galerkin(s_weak, ansatz, *s_gp, tangent_contributions=False))
        conn_list.append(jnp.asarray(s_elems))
        modes.append("user element")
    n = len(model_list)
    static_settings = flax.core.FrozenDict({"number of fields": (nf,) * n, "assembling mode": tuple(modes),
                                            "solution structure": ("nodal imposition",) * n, "model": tuple(model_list),
                                            "solver type": "newton", "solver backend": "scipy", "solver": "lapack",
                                            "verbose": -1})
    if dirichlet_dofs is None:
        dirichlet_dofs = np.zeros((coords.shape[0], nf), dtype=bool)
    settings = {"dirichlet dofs": jnp.asarray(dirichlet_dofs), "connectivity": tuple(conn_list),
                "node coordinates": jnp.asarray(coords), "dirichlet conditions": jnp.zeros((coords.shape[0], nf))}
    settings.update(settings_extra or {})
    rng = np.random.default_rng(7)
    dofs = jnp.asarray(rng.uniform(-dofs_scale, dofs_scale, (coords.shape[0], nf)))
    t = time.time()
    R = assembler.assemble_residual(dofs, settings, static_settings)
    K = assembler.assemble_tangent(dofs, settings, static_settings)
    data, rows, cols = bcoo_arrays(K)
    out.update({tag + "_dofs": A(dofs), tag + "_R": A(R), tag + "_K_data": data, tag + "_K_rows": rows, tag + "_K_cols": cols})
    if solve:
        sol, infos = solver.solver(jnp.zeros((coords.shape[0], nf)), settings, static_settings)
        out[tag + "_sol"] = A(sol)
        out[tag + "_infos"] = np.array([float(infos[0]), float(infos[1]), float(infos[2])])
    print("  %s: %d dofs, %.1f s" % (tag, dofs.size, time.time() - t), flush=True)


def case_elements(out):
    E = lambda x, settings: settings["youngs modulus"]
    nu = lambda x, settings: settings["poisson ratio"]
    mat = {"youngs modulus": 100.0, "poisson ratio": 0.3, "load multiplier": 4.0}
    # Cook's membrane, first 2 of the 12 Q2 elements + the two line3 Neumann elements (G2 set-up)
    coords, elems = problems.COOK_NODES, problems.COOK_ELEMS
    neumann = problems.COOK_SURF[[4, 5]]
    weak = models.hyperelastic_steady_state_weak(models.neo_hooke, E, nu, "plain strain")
    trac = models.neumann_weak(lambda x, settings: jnp.asarray([0., settings["load multiplier"]]))
    element_case(out, "cook_q9", coords, elems[[2, 4]], 2, weak, spaces.fem_iso_line_quad_brick,
                 seeder.gauss_legendre_nd(dimension=2, order=4),
                 surf=(neumann, trac, seeder.gauss_legendre_nd(dimension=1, order=4)), settings_extra=mat, dofs_scale=0.3)
    # hex8 neo-Hooke 3-D + quad4 traction face
    c, e = mesher.structured_mesh((1, 1, 2), [[0, 0, 0], [1, 0, 0], [1.1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1.2], [1, 1, 1], [0, 1, 1]], "brick")
    weak3 = models.hyperelastic_steady_state_weak(models.neo_hooke, E, nu, "3d")
    trac3 = models.neumann_weak(lambda x, settings: jnp.asarray([0., 0.5, -settings["load multiplier"]]))
    top = np.array([[2, 8, 11, 5]])                     # nodes with k = 2 (z top), any order is a valid quad4 input
    element_case(out, "hex8_neo", A(c), A(e), 3, weak3, spaces.fem_iso_line_quad_brick,
                 seeder.gauss_legendre_nd(dimension=3, order=2),
                 surf=(top, trac3, seeder.gauss_legendre_nd(dimension=2, order=2)), settings_extra=mat, dofs_scale=0.03)
    # linear elasticity: the three modes (plain strain carries the reference's C33 quirk)
    cq, eq = mesher.structured_mesh((2, 1), [[0, 0], [2, 0], [2.5, 1.5], [0, 1]], "quad")
    for mode in ("plain strain", "plain stress"):
        w = models.linear_elasticity_weak(E, nu, mode, lambda x: jnp.asarray([0.3, -1.0]))
        element_case(out, "linel_" + mode.replace(" ", "_"), A(cq), A(eq), 2, w, spaces.fem_iso_line_quad_brick,
                     seeder.gauss_legendre_nd(dimension=2, order=2), settings_extra=mat, dofs_scale=0.05)
    w = models.linear_elasticity_weak(E, nu, "3d", lambda x: jnp.asarray([0.3, -1.0, 0.2]))
    element_case(out, "linel_3d", A(c), A(e)[:1], 3, w, spaces.fem_iso_line_quad_brick,
                 seeder.gauss_legendre_nd(dimension=3, order=2), settings_extra=mat, dofs_scale=0.05)
    # P2 triangles as isoparametric user elements (spaces.fem_iso_line_tri_tet), neo-Hooke plain strain
    ct, et = mesher.structured_mesh((1, 1), [[0, 0], [2, 0], [2.3, 1.0], [0, 1]], "tri")
    ct, et = mesher.elevate_mesh_order(ct, et)
    element_case(out, "tri6_neo", A(ct), A(et), 2, weak, spaces.fem_iso_line_tri_tet, seeder.int_pts_ref_tri(2),
                 settings_extra=mat, dofs_scale=0.05)


def potential_case(out, tag, coords, elems, ansatz, gp, dim):
    """README-style 'user potential' with dict dofs on a 3-D (or any) isoparametric element: the route on which the
    reference itself can run the scalar Poisson problem of BASELINE config 4 on hex elements (poisson_weak as a 'user
    element' with array dofs fails inside the reference: jnp.dot of (1,3) gradients for (n,1) dofs, dofs.shape[-1] in
    assembler._get_residual for (n,) dofs)."""
    shift = jnp.asarray([0.3, -0.2, 0.5][:dim])

    def integrand_fun(x_int, ansatz_fun, settings, static_settings, elem_number, set):   # short_example.py:23-33, in 3-D
        x = ansatz_fun["physical coor"](x_int)
        phi_fun = ansatz_fun["phi"]
        phi = phi_fun(x_int)
        dphi_dx = jax.jacrev(phi_fun)(x_int)
        source_term = 3.0 * jnp.sin(2.0 * x @ x) - jnp.cos((x - shift) @ (x - shift))
        return (1 / 2) * dphi_dx @ dphi_dx - source_term * phi

    user_potential = models.mixed_reference_domain_potential(integrand_fun, {"phi": ansatz}, *gp, "phi")
    static_settings = flax.core.FrozenDict({"assembling mode": ("user potential",), "solution structure": ("nodal imposition",),
                                            "model": (user_potential,), "solver type": "newton", "solver backend": "scipy",
                                            "solver": "lapack", "verbose": -1})
    n = coords.shape[0]
    settings = {"connectivity": ({"phi": jnp.asarray(elems)},), "dirichlet dofs": {"phi": jnp.zeros(n, dtype=bool)},
                "node coordinates": {"phi": jnp.asarray(coords)}, "dirichlet conditions": {"phi": jnp.zeros(n)}}
    rng = np.random.default_rng(11)
    dofs = {"phi": jnp.asarray(rng.uniform(-1, 1, n))}
    t = time.time()
    R = assembler.assemble_residual(dofs, settings, static_settings)
    K = assembler.assemble_tangent(dofs, settings, static_settings)
    data, rows, cols = bcoo_arrays(K)
    out.update({tag + "_coords": A(coords), tag + "_elems": A(elems), tag + "_gp_x": A(gp[0]), tag + "_gp_w": A(gp[1]),
                tag + "_dofs": A(dofs["phi"]), tag + "_R": A(R["phi"]), tag + "_K_data": data, tag + "_K_rows": rows,
                tag + "_K_cols": cols})
    print("  %s: %d dofs, %.1f s" % (tag, n, time.time() - t), flush=True)


def case_two_fields(out):
    """MULTI-FIELD dict dofs executed by the reference: fields 'phi' and 'psi' (scalar, 2-D Q1 quads, the same element
    topology but different node coordinates), two 'user potential' domains -- the first integrand touches phi only, the
    second psi only -- so that the BCOO holds explicit zero blocks for the uncoupled field pairs (assembler.py:79-117)."""
    c_phi, e = mesher.structured_mesh((3, 2), [[0, 0], [2, 0], [2.5, 1.5], [0, 1]], "quad")
    c_phi, e = A(c_phi), A(e)
    c_psi = c_phi * np.array([1.0, 0.7]) + np.array([0.1, 0.2]) + 0.05 * np.sin(3.0 * c_phi[:, ::-1])
    gp = seeder.gauss_legendre_nd(dimension=2, order=2)

    def make(field, coef, a, b):
        def integrand_fun(x_int, ansatz_fun, settings, static_settings, elem_number, set):
            x = ansatz_fun["physical coor"](x_int)
            f_fun = ansatz_fun[field]
            df = jax.jacrev(f_fun)(x_int)
            return 0.5 * coef * df @ df - (a * jnp.sin(2.0 * x @ x) + b) * f_fun(x_int)
        return integrand_fun

    ans = {"phi": spaces.fem_iso_line_quad_brick, "psi": spaces.fem_iso_line_quad_brick}
    pot1 = models.mixed_reference_domain_potential(make("phi", 1.0, 3.0, -1.0), ans, *gp, "phi")
    pot2 = models.mixed_reference_domain_potential(make("psi", 2.5, -2.0, 0.5), ans, *gp, "psi")
    static_settings = flax.core.FrozenDict({"assembling mode": ("user potential", "user potential"),
                                            "solution structure": ("nodal imposition", "nodal imposition"),
                                            "model": (pot1, pot2), "solver type": "newton", "solver backend": "scipy",
                                            "solver": "lapack", "verbose": -1})
    n = c_phi.shape[0]
    conn = {"phi": jnp.asarray(e), "psi": jnp.asarray(e)}
    settings = {"connectivity": (conn, conn), "dirichlet dofs": {"phi": jnp.zeros(n, dtype=bool), "psi": jnp.zeros(n, dtype=bool)},
                "node coordinates": {"phi": jnp.asarray(c_phi), "psi": jnp.asarray(c_psi)},
                "dirichlet conditions": {"phi": jnp.zeros(n), "psi": jnp.zeros(n)}}
    rng = np.random.default_rng(5)
    dofs = {"phi": jnp.asarray(rng.uniform(-1, 1, n)), "psi": jnp.asarray(rng.uniform(-1, 1, n))}
    R = assembler.assemble_residual(dofs, settings, static_settings)
    K = assembler.assemble_tangent(dofs, settings, static_settings)
    data, rows, cols = bcoo_arrays(K)
    out.update({"two_fields_c_phi": c_phi, "two_fields_c_psi": c_psi, "two_fields_elems": e, "two_fields_gp_x": A(gp[0]),
                "two_fields_gp_w": A(gp[1]), "two_fields_dofs_phi": A(dofs["phi"]), "two_fields_dofs_psi": A(dofs["psi"]),
                "two_fields_R_phi": A(R["phi"]), "two_fields_R_psi": A(R["psi"]), "two_fields_K_data": data,
                "two_fields_K_rows": rows, "two_fields_K_cols": cols})


def case_potential3d(out):
    cube2 = [[0, 0, 0], [1, 0, 0], [1.1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1.2], [1, 1, 1], [0, 1, 1]]
    c, e = mesher.structured_mesh((1, 2, 2), cube2, "brick")
    potential_case(out, "pot_hex8", A(c), A(e), spaces.fem_iso_line_quad_brick, seeder.gauss_legendre_nd(dimension=3, order=2), 3)
    c1, e1 = mesher.structured_mesh((1, 1, 1), cube2, "brick")
    c27, e27 = mesher.elevate_mesh_order(c1, e1)
    potential_case(out, "pot_hex27", A(c27), A(e27), spaces.fem_iso_line_quad_brick, seeder.gauss_legendre_nd(dimension=3, order=4), 3)
    # tetrahedra: the 5 hard-coded nodes of two tets sharing a face (the mesher has no tet meshes)
    ct = np.array([[0., 0., 0.], [1., 0., 0.1], [0.1, 1., 0.], [0., 0.2, 1.], [1.1, 1.2, 0.9]])
    et = np.array([[0, 1, 2, 3], [1, 2, 3, 4]])
    potential_case(out, "pot_tet4", ct, et, spaces.fem_iso_line_tri_tet, seeder.int_pts_ref_tet(2), 3)
    cq, eq = mesher.structured_mesh((2, 1), [[0, 0], [2, 0], [2.5, 1.5], [0, 1]], "tri")
    potential_case(out, "pot_tri3", A(cq), A(eq), spaces.fem_iso_line_tri_tet, seeder.int_pts_ref_tri(2), 2)
    cq6, eq6 = mesher.elevate_mesh_order(cq, eq)
    potential_case(out, "pot_tri6", A(cq6), A(eq6), spaces.fem_iso_line_tri_tet, seeder.int_pts_ref_tri(2), 2)


def case_potential_more(out):
    """quad9 and tet10 on the same route (tet10 mid-edge nodes in the order of spaces.fem_iso_line_tri_tet: edges
    (0,1), (1,2), (0,2), (0,3), (1,3), (2,3); slightly curved edges)."""
    cq, eq = mesher.structured_mesh((2, 1), [[0, 0], [2, 0], [2.5, 1.5], [0, 1]], "quad")
    c9, e9 = problems_elevate_quads(A(cq), A(eq))
    potential_case(out, "pot_quad9", c9, e9, spaces.fem_iso_line_quad_brick, seeder.gauss_legendre_nd(dimension=2, order=4), 2)
    ct = np.array([[0., 0., 0.], [1., 0., 0.1], [0.1, 1., 0.], [0., 0.2, 1.], [1.1, 1.2, 0.9]])
    et = np.array([[0, 1, 2, 3], [1, 2, 3, 4]])
    pts, edges, rows = [p for p in ct], {}, []
    rng = np.random.default_rng(4)
    for el in et:
        row = list(el)
        for a, b in ((0, 1), (1, 2), (0, 2), (0, 3), (1, 3), (2, 3)):
            key = (min(el[a], el[b]), max(el[a], el[b]))
            if key not in edges:
                edges[key] = len(pts)
                pts.append(0.5 * (ct[el[a]] + ct[el[b]]) + rng.uniform(-0.03, 0.03, 3))
            row.append(edges[key])
        rows.append(row)
    potential_case(out, "pot_tet10", np.array(pts), np.array(rows), spaces.fem_iso_line_tri_tet, seeder.int_pts_ref_tet(2), 3)


def problems_elevate_quads(coords, elems):
    from oracle import mesher as omesh          # quad4 -> quad9 in the node order of spaces.py:1924 (DATA for the reference run)
    return omesh.elevate_quads(coords, elems)


def case_elements_more(out):
    """Q1 quads + line2 traction (Cook Q1 set-up), tet4 neo-Hooke + tri3 traction face, tri3 plain-stress elasticity."""
    E = lambda x, settings: settings["youngs modulus"]
    nu = lambda x, settings: settings["poisson ratio"]
    mat = {"youngs modulus": 100.0, "poisson ratio": 0.3, "load multiplier": 4.0}
    cq, eq = mesher.structured_mesh((2, 2), [[0., 0.], [48., 44.], [48., 60.], [0., 44.]], "quad")
    weak = models.hyperelastic_steady_state_weak(models.neo_hooke, E, nu, "plain strain")
    trac = models.neumann_weak(lambda x, settings: jnp.asarray([0., settings["load multiplier"]]))
    right = np.array([[6, 7], [7, 8]])                      # nodes with i = 2 (x = 48)
    element_case(out, "quad4_neo_line2", A(cq), A(eq), 2, weak, spaces.fem_iso_line_quad_brick,
                 seeder.gauss_legendre_nd(dimension=2, order=2),
                 surf=(right, trac, seeder.gauss_legendre_nd(dimension=1, order=2)), settings_extra=mat, dofs_scale=0.5)
    ct = np.array([[0., 0., 0.], [1., 0., 0.1], [0.1, 1., 0.], [0., 0.2, 1.], [1.1, 1.2, 0.9]])
    et = np.array([[0, 1, 2, 3], [1, 2, 3, 4]])
    weak3 = models.hyperelastic_steady_state_weak(models.neo_hooke, E, nu, "3d")
    trac3 = models.neumann_weak(lambda x, settings: jnp.asarray([0.2, 0.5, -settings["load multiplier"]]))
    element_case(out, "tet4_neo_tri3", ct, et, 3, weak3, spaces.fem_iso_line_tri_tet, seeder.int_pts_ref_tet(2),
                 surf=(np.array([[1, 2, 4]]), trac3, seeder.int_pts_ref_tri(2)), settings_extra=mat, dofs_scale=0.03)
    c3, e3 = mesher.structured_mesh((2, 1), [[0, 0], [2, 0], [2.5, 1.5], [0, 1]], "tri")
    w = models.linear_elasticity_weak(E, nu, "plain stress", lambda x: jnp.asarray([0.3, -1.0]))
    element_case(out, "tri3_linel", A(c3), A(e3), 2, w, spaces.fem_iso_line_tri_tet, seeder.int_pts_ref_tri(2),
                 settings_extra=mat, dofs_scale=0.05)
    out["tet_rule_2_x"], out["tet_rule_2_w"] = A(seeder.int_pts_ref_tet(2)[0]), A(seeder.int_pts_ref_tet(2)[1])


def case_sparse_compiled(out):
    """'sparse' assembling mode (assembler.py:874-1035 -> variational_schemes.weak_form_galerkin ->
    solution_structures 'compiled' shape functions): conduction + Euler capacity + surface inflow on P1 triangles,
    one linear (backward-Euler) step through solver.solver with 'solver type': 'linear'."""
    p = problems.heat_sparse(2, 2, 1)
    cond, cap, sur = p["sets"]
    static_settings = flax.core.FrozenDict({
        "assembling mode": ("sparse",) * 3, "solution structure": ("nodal imposition",) * 3,
        "variational scheme": ("weak form galerkin",) * 3, "solution space": ("fem simplex",) * 3,
        "shape function mode": "compiled", "number of fields": (1, 1, 1), "maximal number of neighbors": (3, 3, 2),
        "model": (models.poisson_weak(lambda x, settings: 1.5, lambda x: 0.0),
                  models.forward_backward_euler_weak(lambda x, settings: 0.1),
                  models.neumann_weak(lambda x: -3.0)),
        "solver type": "linear", "solver backend": "scipy", "solver": "lapack", "verbose": -1})
    dofs_n = jnp.asarray(p["settings"]["dofs n"])
    settings = {"connectivity": tuple(jnp.asarray(s["conn"]) for s in p["sets"]), "node coordinates": jnp.asarray(p["coords"]),
                "dirichlet dofs": jnp.asarray(p["mask"]), "dirichlet conditions": jnp.asarray(p["values"]),
                "integration coordinates": tuple(jnp.asarray(x) for x in p["x_int"]),
                "integration weights": tuple(jnp.asarray(w) for w in p["w_int"]),
                "compiled shape functions": tuple((jnp.asarray(s["N"]), jnp.asarray(s["dNdx"])) for s in p["sets"]),
                "time increment": 0.2, "dofs n": dofs_n}
    R = assembler.assemble_residual(dofs_n, settings, static_settings)
    K = assembler.assemble_tangent(dofs_n, settings, static_settings)
    data, rows, cols = bcoo_arrays(K)
    delta, _ = solver.solver(dofs_n, settings, static_settings)
    out.update({"sparse_R": A(R), "sparse_K_data": data, "sparse_K_rows": rows, "sparse_K_cols": cols,
                "sparse_delta": A(delta), "sparse_dofs_n": A(dofs_n)})


def case_simplex_direct(out):
    """spaces.fem_ini_simplex (spaces.py:15194-15296), the 'direct' shape functions of 'fem simplex': values exactly,
    physical gradients by Richardson differences of the reference function."""
    from oracle import shapes as oshapes
    rng = np.random.default_rng(21)
    st = flax.core.FrozenDict({"shape function mode": "compiled"})
    for name in ("tri3", "tri6", "tet4"):
        ref = np.asarray(oshapes.REF_NODES[name], dtype=float)
        dim = ref.shape[1]
        Amat = np.eye(dim) + 0.2 * rng.standard_normal((dim, dim))
        xI = ref @ Amat.T + rng.uniform(-1, 1, dim)
        x = xI.mean(axis=0) + 0.05 * rng.standard_normal(dim)
        fun = lambda xx: spaces.fem_ini_simplex(xx, jnp.asarray(xI), jnp.zeros((xI.shape[0], 1)), st, 0)
        fakejax.POINT_MODE = "fd"       # nested derivatives w.r.t. small vectors: differences, not complex steps
        N = A(fun(jnp.asarray(x)))
        dN = A(jax.jacfwd(fun)(jnp.asarray(x)))
        fakejax.POINT_MODE = "complex"
        out["simplex_%s_xI" % name], out["simplex_%s_x" % name] = xI, x
        out["simplex_%s_N" % name], out["simplex_%s_dN" % name] = N, dN


def case_newton_semantics(out):
    """solver.damped_newton (solver.py:837-948) driven by synthetic residual sequences."""
    def run(norms, newton_tol=1e-8, maxiter=30):
        state = {"k": 0}

        def lin_solve(d):
            return jnp.zeros(2)

        def residual(d):
            r = norms[min(state["k"], len(norms) - 1)]
            state["k"] += 1
            return jnp.asarray([r, 0.0])
        sol, (it, rn, div) = solver.damped_newton(jnp.zeros(2), residual, lin_solve, jnp.asarray([True, True]), newton_tol,
                                                  maxiter, 1.0, verbose=-1)
        return [float(it), float(rn), float(div)]
    out["newton_converge"] = np.array(run([1.0, 1e-3, 1e-9]))
    out["newton_diverge"] = np.array(run([1.0, 0.5, 0.4, 8.0, 1e-9]))
    out["newton_early_jump_ok"] = np.array(run([1.0, 50.0, 1e-9]))         # ratio > 10 at itt <= 1 is tolerated
    out["newton_nan"] = np.array(run([1.0, float("nan"), 1e-9]))
    out["newton_maxiter"] = np.array(run([1.0] * 10, maxiter=3))


def case_dae_rules(out):
    """autopdex.dae time integrators (dae.py:288-318 BackwardEuler, :420-481 AdamsMoulton, :537-587 BackwardDiffFormula,
    :707-766 DiagonallyImplicitRungeKutta): the discrete value / first derivative of every stage (`_rule`) and the
    end-of-step update (`_update`) evaluated by the reference's own classes on random stage values and histories."""
    from autopdex import dae
    rng = np.random.default_rng(5)
    n, dt = 7, 0.37
    cases = ([("be", dae.BackwardEuler())] + [("bdf%d" % k, dae.BackwardDiffFormula(k)) for k in range(1, 7)]
             + [("am%d" % k, dae.AdamsMoulton(k)) for k in range(1, 7)]
             + [("dirk%d" % k, dae.DiagonallyImplicitRungeKutta(k)) for k in (1, 2, 3)])
    for tag, integ in cases:
        q_n = rng.normal(size=(integ.num_steps, n))
        q_t_n = rng.normal(size=(integ.num_steps, 1, n))
        stages = rng.normal(size=(integ.num_stages, n))
        # one-stage integrators are called with the stage value itself by `update`, with q_stages[...] by the residual
        arg = jnp.asarray(stages if integ.num_stages > 1 else stages[0])
        val, q_t = integ.value_and_derivatives(arg, jnp.asarray(q_n), jnp.asarray(q_t_n), dt)[:2]
        q_n1, q_t_n1 = integ.update(jnp.asarray(stages), jnp.asarray(q_n), jnp.asarray(q_t_n), dt)
        q_t_n1 = q_t_n1[0] if isinstance(q_t_n1, tuple) else q_t_n1
        pre = "dae_%s_" % tag
        out[pre + "q_n"], out[pre + "q_t_n"], out[pre + "stages"], out[pre + "dt"] = q_n, q_t_n, stages, np.array(dt)
        out[pre + "value"], out[pre + "q_t"] = A(val).reshape(integ.num_stages, n), A(q_t).reshape(integ.num_stages, n)
        out[pre + "q_n1"], out[pre + "q_t_n1"] = A(q_n1), A(q_t_n1)
        out[pre + "positions"] = A(integ.stage_positions).astype(float)
        out[pre + "order"] = np.array(int(integ.order))


DAE_SAVE_TIMES = [0.0, 0.2, 0.35, 0.5, 0.7, 0.9, 1.0]
DAE_CTRL_SEQ = [(True, 2, 0.1), (True, 9, 0.2), (False, 4, 0.25), (True, 1, 0.29), (False, 3, 0.02), (False, 3, 0.011)]


def case_dae_control(out):
    """dae.SaveEquidistantPolicy / SaveAllPolicy (dae.py:1186-1311) fed with a sequence of accepted steps, and
    dae.RootIterationController / ConstantStepSizeController (dae.py:1474-1573) fed with a sequence of
    (converged, iterations, dt): the histories and controller states the reference's own classes produce."""
    from autopdex import dae
    q0 = np.arange(4.0)
    for tag, pol, max_steps in (("equi3", dae.SaveEquidistantPolicy(num_points=3, tol=1e-6), 10),
                                ("equi_default", dae.SaveEquidistantPolicy(), 5), ("all", dae.SaveAllPolicy(), 10),
                                ("all_clipped", dae.SaveAllPolicy(), 4)):
        st = pol.initialize({"a": jnp.asarray(q0)}, 1.0, max_steps, {"u": jnp.asarray(0.0)})
        for t in DAE_SAVE_TIMES:
            st = pol.save_step(st, t, {"a": jnp.asarray(q0 + t)}, {"u": jnp.asarray(2 * t)})
        h = pol.finalize(st)
        out["dae_save_%s_t" % tag], out["dae_save_%s_q" % tag], out["dae_save_%s_u" % tag] = A(h.t), A(h.q["a"]), A(h.user["u"])
    for tag, ctrl in (("root", dae.RootIterationController(target_niters=6, gamma=0.5, max_step_size=0.3, min_step_size=0.01)),
                      ("const", dae.ConstantStepSizeController())):
        st, rows = ctrl.initialize(0.0), []
        for conv, its, dt in DAE_CTRL_SEQ:
            st = ctrl.compute_scaler(None, None, None, st, 1, conv, its, dt, -1)
            st = ctrl.check_accept(st, conv, -1)
            rows.append([float(st.step_scaler), float(getattr(st, "dt", 1.0)), float(st.accept), float(st.interrupt)])
        out["dae_ctrl_%s" % tag] = np.array(rows)


LOAD_STEP_SCRIPTS = {
    # (newton steps, diverged) per call of the Newton solver, repeated from the last entry on
    "smooth": [(6, False)],
    "halving": [(5, False), (30, True), (30, True), (4, False), (9, False), (30, True), (3, False)],
    "stall": [(30, True)],                     # never converges: increments shrink below min_increment
    "slow": [(12, False)],                     # more iterations than the target: increments shrink
}


def case_load_stepping(out):
    """solver.adaptive_load_stepping (solver.py:155-457) driven by a scripted Newton solver: the sequence of load
    multipliers it tries, the multipliers it accepts, and the final carry -- the rollback / halving / capping control flow
    of :296-379 as the reference's own code executes it."""
    from autopdex import solver as rsolver
    real = rsolver.solver
    for tag, script in LOAD_STEP_SCRIPTS.items():
        calls = []

        def fake_solver(dofs, settings, static_settings, newton_tol=1e-8, **kwargs):
            steps, div = script[min(len(calls), len(script) - 1)]
            calls.append(float(settings["load multiplier"]))
            return dofs + 1.0, (jnp.asarray(steps), jnp.asarray(1e-12), jnp.asarray(div))
        rsolver.solver = fake_solver
        try:
            st = flax.core.FrozenDict({"solver type": "newton", "verbose": -1})
            carry = rsolver.adaptive_load_stepping(jnp.zeros(3), {"load multiplier": 0.0}, st, path_dependent=True,
                                                   implicit_diff_mode=None, max_load_steps=1000)
        finally:
            rsolver.solver = real
        out["loadstep_%s_tried" % tag] = np.array(calls)
        out["loadstep_%s_final" % tag] = np.array([float(A(carry[0])[0]), float(carry[1]), float(carry[2])])


DAE_MANAGER_SCHEMES = ("backward_euler", "bdf2", "am2", "dirk2", "dirk3")


def case_dae_manager(out):
    """The reference's TimeSteppingManager.run (dae.py:2087-2268: stage loop, history roll, integrator updates,
    constrained dofs in dae.newton_solver) EXECUTED on the semi-discrete heat conduction system M q_t + K q = F of
    tests/test_zz_gpu_r02_dae._settings(2) (27 nodes, Dirichlet face x = 0 with non-zero values) given as a dense
    'dae' callable: the trajectories the b200 backend must reproduce through its device solves."""
    from autopdex import dae
    from tests import test_zz_gpu_r02_dae as T
    coords, K, M, F, mask, values, res, settings = T._settings(2)
    Kd, Md = jnp.asarray(K.toarray()), jnp.asarray(M.toarray())
    q0 = 0.3 * np.cos(coords[:, 1])
    dt, n_steps = 0.05, 3

    def dae_fun(q_fun, t, settings):
        return Md @ jax.jacfwd(q_fun)(t)["theta"] + Kd @ q_fun(t)["theta"] - jnp.asarray(F)
    tight = lambda *a, **k: dae.newton_solver(*a, **dict(k, atol=1e-12, max_iter=10))
    for scheme in DAE_MANAGER_SCHEMES:
        integ = (dae.BackwardEuler() if scheme == "backward_euler" else dae.BackwardDiffFormula(int(scheme[3:])) if scheme.startswith("bdf")
                 else dae.AdamsMoulton(int(scheme[2:])) if scheme.startswith("am") else dae.DiagonallyImplicitRungeKutta(int(scheme[4:])))
        st = flax.core.FrozenDict({"dae": dae_fun, "time integrators": {"theta": integ}, "verbose": -1, "implicit diff mode": None})
        mgr = dae.TimeSteppingManager(st, root_solver=tight, save_policy=dae.SaveAllPolicy())
        state = mgr.run({"theta": jnp.asarray(q0)}, dt, dt * n_steps, n_steps,
                        {"current time": 0.0, "dirichlet dofs": {"theta": jnp.asarray(mask)},
                         "dirichlet conditions": {"theta": jnp.asarray(values)}})
        assert int(state.num_accepted) == n_steps
        out["dae_manager_%s_q" % scheme] = A(state.history.q["theta"])[:n_steps + 1]
        out["dae_manager_%s_t" % scheme] = A(state.history.t)[:n_steps + 1]
    # the Newton-iteration step-size controller inside the reference's loop (accepted steps grow by 1 + gamma (target - 1) / target
    # for this linear system, capped at max_step_size, the last step cut at t_max)
    st = flax.core.FrozenDict({"dae": dae_fun, "time integrators": {"theta": dae.BackwardEuler()}, "verbose": -1, "implicit diff mode": None})
    # (default root solver, atol 1e-8: the stand-in's difference-quotient Jacobian is exact to ~1e-10 only, and a tighter
    # tolerance would cost the reference a second Newton iteration that the exact tangent does not need)
    mgr = dae.TimeSteppingManager(st, save_policy=dae.SaveAllPolicy(),
                                  step_size_controller=dae.RootIterationController(target_niters=6, gamma=0.5, max_step_size=0.05))
    state = mgr.run({"theta": jnp.asarray(q0)}, 0.02, 0.2, 12, {"current time": 0.0, "dirichlet dofs": {"theta": jnp.asarray(mask)},
                                                                 "dirichlet conditions": {"theta": jnp.asarray(values)}})
    na = int(state.num_accepted)
    out["dae_manager_root_controller_t"] = A(state.history.t)[:na + 1]
    out["dae_manager_root_controller_q"] = A(state.history.q["theta"])[:na + 1]
    out["dae_manager_root_controller_counts"] = np.array([na, int(state.num_rejected)])


def utility_inputs():
    rng = np.random.default_rng(21)
    d = {"u": rng.normal(size=(5, 2)), "p": rng.normal(size=(5,)), "T": rng.normal(size=(3, 1))}
    flat = rng.normal(size=5 * 2 + 5 + 3)
    nodes = rng.random(6) < 0.5
    return d, flat, nodes


def case_utility(out):
    """utility.dict_flatten / reshape_as / dict_zeros_like / dof_select (utility.py:82-199, 448-462) on a three-field dict
    (insertion order u, p, T -- the order the reference concatenates in)."""
    d, flat, nodes = utility_inputs()
    dj = {k: jnp.asarray(v) for k, v in d.items()}
    out["util_flat"] = A(utility.dict_flatten(dj))
    back = utility.reshape_as(jnp.asarray(flat), dj)
    for k in d:
        out["util_reshape_" + k] = A(back[k])
        out["util_zeros_" + k] = A(utility.dict_zeros_like(dj)[k])
    out["util_dofsel_list"] = A(utility.dof_select(jnp.asarray(nodes), jnp.asarray([True, False, True])))
    out["util_dofsel_bool"] = A(utility.dof_select(jnp.asarray(nodes), True))


def case_hyper_linear(out):
    """hyperelastic_steady_state_weak with models.linear_elastic_strain_energy (models.py:917-1000, 1167-1185) as user
    elements: 'plain strain' on two Q1 quads, '3d' on one distorted hex8."""
    E = lambda x, settings: settings["youngs modulus"]
    nu = lambda x, settings: settings["poisson ratio"]
    mat = {"youngs modulus": 100.0, "poisson ratio": 0.3}
    cq, eq = mesher.structured_mesh((2, 1), [[0, 0], [2, 0], [2.5, 1.5], [0, 1]], "quad")
    # (no volume load: models.py:996-997 takes jnp.dot(b, virt_u) with virt_u the test FUNCTION -- the reference cannot
    # evaluate a hyperelastic weak form with a volume load at all)
    w2 = models.hyperelastic_steady_state_weak(models.linear_elastic_strain_energy, E, nu, "plain strain")
    element_case(out, "hyperlin_plain_strain", A(cq), A(eq), 2, w2, spaces.fem_iso_line_quad_brick,
                 seeder.gauss_legendre_nd(dimension=2, order=2), settings_extra=mat, dofs_scale=0.05)
    cube2 = [[0, 0, 0], [1, 0, 0], [1.1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1.2], [1, 1, 1], [0, 1, 1]]
    cb, eb = mesher.structured_mesh((1, 1, 2), cube2, "brick")
    w3 = models.hyperelastic_steady_state_weak(models.linear_elastic_strain_energy, E, nu, "3d")
    element_case(out, "hyperlin_3d", A(cb), A(eb)[:1], 3, w3, spaces.fem_iso_line_quad_brick,
                 seeder.gauss_legendre_nd(dimension=3, order=2), settings_extra=mat, dofs_scale=0.05)


CASES = {"dae_rules": case_dae_rules, "hyper_linear": case_hyper_linear, "utility": case_utility, "dae_manager": case_dae_manager, "load_stepping": case_load_stepping, "dae_control": case_dae_control, "tables": case_tables, "indices_dict": case_indices_dict, "two_fields": case_two_fields, "readme3": lambda o: readme_case(3, o, "readme3"),
         "readme5": lambda o: readme_case(5, o, "readme5"), "elements": case_elements, "potential3d": case_potential3d, "potential_more": case_potential_more, "elements_more": case_elements_more, "sparse": case_sparse_compiled, "simplex": case_simplex_direct,
         "newton": case_newton_semantics}

if __name__ == "__main__":
    out = dict(np.load(OUT)) if os.path.exists(OUT) else {}
    for name in (sys.argv[1:] or list(CASES)):
        t = time.time()
        print("case", name, flush=True)
        CASES[name](out)
        print("  done in %.1f s" % (time.time() - t), flush=True)
        np.savez_compressed(OUT, **out)
    print("wrote", OUT, "with", len(out), "arrays")
