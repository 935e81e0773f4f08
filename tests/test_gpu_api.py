"""GPU parity through the reference-facing API: autopdex_b200.solver.solver / adaptive_load_stepping /
assembler.* with `'solver backend': 'b200'`.  These read like the reference's own tests
(tests/test_dicts_as_dofs_user_potential.py, tests/test_user_elem_impl_diff_and_adaptive_load_step.py,
tests/test_backward_euler.py) with the backend swapped."""
import numpy as np
import pytest

from oracle import assemble as oasm
from oracle import solve as osolve
from tests import problems

pytestmark = pytest.mark.gpu


def test_readme_dict_dofs_user_potential_golden():
    from autopdex_b200 import mesher, models, seeder, solver, spaces, utility
    pts = [[0., 0.], [1., 0.], [1., 1.], [0., 1.]]
    coords, elems = mesher.structured_mesh((5, 5), pts, "quad")
    node_coordinates = {"phi": coords}
    connectivity = {"phi": elems}
    p = problems.readme_poisson(5)            # Dirichlet selection: boundary minus corners (geometry.psdf_polygon quirk)
    dirichlet_dofs = {"phi": p["mask"][:, 0]}
    dirichlet_conditions = utility.dict_zeros_like(dirichlet_dofs, dtype=np.float64)
    integrand = models.poisson_potential("phi", source_fun=problems.readme_source)
    user_potential = models.mixed_reference_domain_potential(integrand, {"phi": spaces.fem_iso_line_quad_brick},
                                                             *seeder.gauss_legendre_nd(dimension=2, order=2), "phi")
    static_settings = {"assembling mode": ("user potential",), "solution structure": ("nodal imposition",),
                       "model": (user_potential,), "solver type": "newton", "solver backend": "b200", "solver": "cg",
                       "type of preconditioner": "jacobi", "verbose": -1}
    settings = {"connectivity": (connectivity,), "dirichlet dofs": dirichlet_dofs, "node coordinates": node_coordinates,
                "dirichlet conditions": dirichlet_conditions}
    initial_guess = utility.dict_zeros_like(dirichlet_dofs, dtype=np.float64)
    sol, infos = solver.solver(initial_guess, settings, static_settings)
    assert infos[0] == 1 and not infos[2]
    assert np.isclose(sol["phi"].sum(), 1.9066412530282952, rtol=1e-9, atol=0)     # reference golden value G1


def _cook_settings(krylov="bicgstab"):
    from autopdex_b200 import models, seeder, spaces
    p = problems.cook_g2()
    weak1 = models.hyperelastic_steady_state_weak(models.neo_hooke, lambda x, settings: settings["youngs modulus"],
                                                  lambda x, settings: settings["poisson ratio"], "plain strain")
    elem1 = models.isoparametric_domain_element_galerkin(weak1, spaces.fem_iso_line_quad_brick,
                                                         *seeder.gauss_legendre_nd(dimension=2, order=4))
    weak2 = models.neumann_weak(lambda x, settings: np.asarray([0.0, settings["load multiplier"]]))
    elem2 = models.isoparametric_surface_element_galerkin(weak2, spaces.fem_iso_line_quad_brick,
                                                          *seeder.gauss_legendre_nd(dimension=1, order=4),
                                                          tangent_contributions=False)
    static_settings = {"number of fields": (2, 2), "assembling mode": ("user element", "user element"),
                       "solution structure": ("nodal imposition", "nodal imposition"), "model": (elem1, elem2),
                       "solver type": "newton", "solver backend": "b200", "solver": krylov,
                       "type of preconditioner": "jacobi", "verbose": -1}
    settings = {"dirichlet dofs": p["mask"], "connectivity": (p["sets"][0]["conn"], p["sets"][1]["conn"]),
                "load multiplier": p["q0"], "node coordinates": p["coords"], "dirichlet conditions": p["values"],
                "youngs modulus": 100.0, "poisson ratio": 0.3}
    return p, settings, static_settings


def test_cook_adaptive_load_stepping_golden():
    from autopdex_b200 import solver
    p, settings, static_settings = _cook_settings()
    q0 = p["q0"]

    def multiplier_settings(settings, multiplier):
        settings["load multiplier"] = multiplier * q0
        return settings

    dofs0 = np.zeros(p["mask"].shape)
    out = solver.adaptive_load_stepping(dofs0, settings, static_settings, multiplier_settings, False, None,
                                        newton_tol=1e-8, tol=1e-13)
    dofs, multiplier = out[0], out[1]
    assert np.isclose(multiplier, 1.0)
    assert np.isclose(dofs.ravel() @ dofs.ravel(), 19390.35027108, rtol=1e-8, atol=0)   # reference golden value G2


def test_cook_sensitivities_reproduce_the_reference_golden_values():
    """SURVEY 8f N3 end to end on the reference's own numbers: tests/test_user_elem_impl_diff_and_adaptive_load_step.py
    pins d(u.u)/dE = -216.0310416, d(u.u)/dnu = -1859.43760286 and the second derivatives (:152-164, jnp.allclose: rtol 1e-5) for the converged
    Cook's membrane state, computed there by JAX AD through implicit_diff (forward: _root_jvp, implicit_diff.py:274-304;
    reverse: _root_vjp, :139-183).  Here the linear solves those rules need run on the device (solver.tangent_solve) and the
    parameter derivative of the residual, which JAX AD supplies in the reference, is a central difference of
    assembler.assemble_residual:  forward  du/dp = -K^-1 dR/dp, dJ/dp = 2 u . du/dp;  reverse  dJ/dp = -(K^-T 2u) . dR/dp."""
    from autopdex_b200 import assembler, solver
    p, settings, static_settings = _cook_settings()
    q0 = p["q0"]

    def multiplier_settings(settings, multiplier):
        settings["load multiplier"] = multiplier * q0
        return settings

    out = solver.adaptive_load_stepping(np.zeros(p["mask"].shape), settings, static_settings, multiplier_settings, False, None,
                                        newton_tol=1e-10, tol=1e-13)
    u, settings = out[0], dict(out[4])
    assert np.isclose(out[1], 1.0) and np.isclose(u.ravel() @ u.ravel(), 19390.35027108, rtol=1e-8, atol=0)
    free = ~p["mask"]

    def dR_dp(key, h):
        rp = assembler.assemble_residual(u, dict(settings, **{key: settings[key] + h}), static_settings)
        rm = assembler.assemble_residual(u, dict(settings, **{key: settings[key] - h}), static_settings)
        return np.where(free, (rp - rm) / (2 * h), 0.0)
    lam = solver.tangent_solve(u, 2.0 * u, settings, static_settings, transpose=True, tol=1e-13)      # adjoint: one solve
    for key, h, golden in (("youngs modulus", 1e-3, -216.0310416), ("poisson ratio", 1e-6, -1859.43760286)):
        g = dR_dp(key, h)
        du = -solver.tangent_solve(u, g, settings, static_settings, transpose=False, tol=1e-13)         # forward: one per parameter
        forward = 2.0 * (u.ravel() @ du.ravel())
        reverse = -(lam.ravel() @ g.ravel())
        print("sensitivity %s: forward %.10f reverse %.10f reference golden %.10f" % (key, forward, reverse, golden))
        assert np.isclose(forward, golden, rtol=2e-6, atol=0), (key, forward, golden)
        assert np.isclose(reverse, golden, rtol=2e-6, atol=0), (key, reverse, golden)
        assert np.isclose(forward, reverse, rtol=1e-9, atol=0)

    # second order (the test's remaining golden values, :157-160: 4.17157824, -51.42695148 twice, -45854.7002203): central
    # differences of the adjoint first-order sensitivities between re-converged neighbouring parameter values
    steps = {"youngs modulus": 1e-3, "poisson ratio": 1e-6}

    def gradient(st):
        uu, info = solver.solver(u, st, static_settings, newton_tol=1e-11, tol=1e-14)
        assert not info[2]
        adj = solver.tangent_solve(uu, 2.0 * uu, st, static_settings, transpose=True, tol=1e-14)
        g = {}
        for key, h in steps.items():
            rp = assembler.assemble_residual(uu, dict(st, **{key: st[key] + h}), static_settings)
            rm = assembler.assemble_residual(uu, dict(st, **{key: st[key] - h}), static_settings)
            g[key] = -(adj.ravel() @ np.where(free, (rp - rm) / (2 * h), 0.0).ravel())
        return g
    golden2 = {("youngs modulus", "youngs modulus"): 4.17157824, ("youngs modulus", "poisson ratio"): -51.42695148,
               ("poisson ratio", "youngs modulus"): -51.42695148, ("poisson ratio", "poisson ratio"): -45854.7002203}
    for key, h in (("youngs modulus", 0.05), ("poisson ratio", 1e-4)):
        gp, gm = gradient(dict(settings, **{key: settings[key] + h})), gradient(dict(settings, **{key: settings[key] - h}))
        for other in steps:
            second = (gp[other] - gm[other]) / (2 * h)
            print("second-order sensitivity d2/d(%s)d(%s): %.8f reference golden %.8f" % (key, other, second, golden2[key, other]))
            assert np.isclose(second, golden2[key, other], rtol=5e-6, atol=0), (key, other, second)   # the reference: rtol 1e-5
    solver.clear_plan_cache()


def test_linear_solver_type_returns_mixed_vector():
    from autopdex_b200 import solver
    p, settings, static_settings = _cook_settings("bicgstab")
    static_settings = dict(static_settings, **{"solver type": "linear"})
    settings["dirichlet conditions"] = np.where(p["mask"], 0.25, 0.0)          # non-homogeneous Dirichlet values
    dofs0 = np.random.default_rng(2).uniform(-0.01, 0.01, p["mask"].shape)
    sol, infos = solver.solver(dofs0, settings, static_settings, tol=1e-13)
    assert infos is None
    prob = osolve.Problem(p["sets"], p["coords"], p["mask"], settings["dirichlet conditions"])
    ref = osolve.solve_linear(prob, dofs0)
    assert np.allclose(sol[p["mask"]], 0.25)                                    # Dirichlet entries = imposed values
    assert np.linalg.norm(sol - ref) / np.linalg.norm(ref) < 1e-8


def test_damped_newton_iteration_count_matches_oracle():
    from autopdex_b200 import solver
    p, settings, static_settings = _cook_settings()
    static_settings = dict(static_settings, **{"solver type": "damped newton"})
    settings["load multiplier"] = 3.0
    p["sets"][1]["model"]["traction"] = np.array([0.0, 3.0])
    dofs0 = np.zeros(p["mask"].shape)
    sol, (it, rn, div) = solver.solver(dofs0, settings, static_settings, newton_tol=1e-6, damping_coefficient=0.8,
                                       maxiter=30, tol=1e-13)
    prob = osolve.Problem(p["sets"], p["coords"], p["mask"], p["values"])
    ref, (rit, rrn, rdiv) = osolve.damped_newton(prob, dofs0, newton_tol=1e-6, damping=0.8)
    assert (it, div) == (rit, rdiv)
    assert np.linalg.norm(sol - ref) / np.linalg.norm(ref) < 1e-8


def test_newton_maxiter_flags_divergence_like_reference():
    from autopdex_b200 import solver
    p, settings, static_settings = _cook_settings()
    dofs0 = np.zeros(p["mask"].shape)
    sol, (it, rn, div) = solver.solver(dofs0, settings, static_settings, newton_tol=1e-30, maxiter=2, tol=1e-13)
    prob = osolve.Problem(p["sets"], p["coords"], p["mask"], p["values"])
    ref, (rit, rrn, rdiv) = osolve.damped_newton(prob, dofs0, newton_tol=1e-30, maxiter=2)
    assert (it, div) == (rit, rdiv) == (3, True)


@pytest.mark.parametrize("dim,order,mode", [(2, 1, "direct"), (2, 2, "direct"), (3, 1, "direct"), (2, 1, "compiled")])
def test_sparse_mode_backward_euler_steps(dim, order, mode):
    """Three 'sparse' sets (conduction, capacity, surface inflow), 'solver type': 'linear', three backward-Euler
    steps with the plan (pattern) reused -- the structure of tests/test_backward_euler.py:281-373."""
    from autopdex_b200 import models, solver
    p = problems.heat_sparse(dim, 3, order)
    cond, cap, sur = p["sets"]
    static_settings = {"assembling mode": ("sparse",) * 3, "solution structure": ("nodal imposition",) * 3,
                       "variational scheme": ("weak form galerkin",) * 3, "solution space": ("fem simplex",) * 3,
                       "shape function mode": mode,
                       "model": (models.poisson_weak(lambda x, settings: 1.5, lambda x: 0.0),
                                 models.forward_backward_euler_weak(lambda x, settings: 0.1),
                                 models.neumann_weak(lambda x: -3.0)),
                       "solver type": "linear", "solver backend": "b200", "solver": "cg",
                       "type of preconditioner": "jacobi", "verbose": -1}
    settings = {"connectivity": tuple(s["conn"] for s in p["sets"]), "node coordinates": p["coords"],
                "dirichlet dofs": p["mask"], "dirichlet conditions": p["values"],
                "integration coordinates": p["x_int"], "integration weights": p["w_int"],
                "time increment": 0.2, "dofs n": p["settings"]["dofs n"].copy()}
    if mode == "compiled":
        settings["compiled shape functions"] = tuple((s["N"], s["dNdx"]) for s in p["sets"])
    else:
        # the surface set's physical-space P1 fit needs dim+1 nodes; feed its facet tables as 'compiled' is the
        # reference's route for lower-dimensional facets, so test the direct route on the two volume sets only
        static_settings["assembling mode"] = ("sparse",) * 2
        for k in ("solution structure", "variational scheme", "solution space", "model"):
            static_settings[k] = static_settings[k][:2]
        for k in ("connectivity", "integration coordinates", "integration weights"):
            settings[k] = settings[k][:2]
        p["sets"] = p["sets"][:2]
    prob = osolve.Problem(p["sets"], p["coords"], p["mask"], p["values"], dict(p["settings"]))
    dofs = p["settings"]["dofs n"].copy()
    ref = dofs.copy()
    for step in range(3):
        settings["dofs n"] = dofs
        dofs = dofs + solver.solver(dofs, settings, static_settings, tol=1e-13)[0]          # maze_backward_euler.py:362
        prob.settings["dofs n"] = ref
        ref = ref + osolve.solve_linear(prob, ref)
        assert np.linalg.norm(dofs - ref) / np.linalg.norm(ref) < 1e-8
    assert len(solver._PLAN_CACHE) >= 1


def test_assembler_module_matches_oracle():
    from autopdex_b200 import assembler
    p, settings, static_settings = _cook_settings()
    dofs = np.random.default_rng(0).uniform(-0.05, 0.05, p["mask"].shape)
    R = assembler.assemble_residual(dofs, settings, static_settings)
    B = assembler.assemble_tangent(dofs, settings, static_settings)          # the reference's wire format (BCOO)
    Ro, data = oasm.assemble(p["sets"], p["coords"], dofs, {})
    rows, cols = oasm.coo_indices(p["sets"])
    # element-local pairs in reference order, duplicates not summed: indices bit-exact, values <= 1e-12
    assert B.indices.shape == (rows.size, 2) and B.indices.dtype == np.int64 and B.nse == rows.size
    assert np.array_equal(B.indices[:, 0], rows) and np.array_equal(B.indices[:, 1], cols)
    assert B.shape == (dofs.size, dofs.size)
    assert np.abs(B.data - data).max() / np.abs(data).max() < 1e-12
    full = oasm.scipy_assembling(data, rows, cols, dofs.size)
    # what solver.scipy_assembling makes of it (solver.py:1207-1211): the reference's own SciPy calls on OUR BCOO
    mine = oasm.scipy_assembling(B.data, B.indices[:, 0], B.indices[:, 1], dofs.size)
    assert np.array_equal(mine.indptr, full.indptr) and np.array_equal(mine.indices, full.indices)
    assert np.abs(mine.data - full.data).max() / np.abs(full.data).max() < 1e-12
    for K in (B.sum_duplicates(), assembler.assemble_tangent(dofs, settings, static_settings, format="csr")):
        assert np.array_equal(K.indptr, full.indptr) and np.array_equal(K.indices, full.indices)
        assert np.abs(K.data - full.data).max() / np.abs(full.data).max() < 1e-12
    assert np.abs(R.ravel() - Ro).max() / np.abs(Ro).max() < 1e-12


def test_bcoo_of_register_kernel_set_matches_oracle():
    """hex8 Poisson goes through the register kernel, whose element stream is entry-major and upper-triangle only:
    apdx_get_coo_values must hand back the reference's element-major full matrices."""
    from autopdex_b200 import assembler
    import bench
    settings, static_settings, _ = bench.build_problem(5, 0, 1)
    p = problems.poisson_hex(5)
    dofs = np.random.default_rng(3).uniform(-1, 1, p["mask"].shape)
    B = assembler.assemble_tangent(dofs, settings, static_settings)
    _, data = oasm.assemble(p["sets"], p["coords"], dofs, {})
    rows, cols = oasm.coo_indices(p["sets"])
    assert np.array_equal(B.indices[:, 0], rows) and np.array_equal(B.indices[:, 1], cols)
    assert np.abs(B.data - data).max() / np.abs(data).max() < 1e-12


@pytest.mark.parametrize("transpose", [False, True])
def test_tangent_solve_matches_reference_sensitivity_solve(transpose):
    """`solve_fun(mat or mat.T, rhs, free_dofs_flat)` of implicit_diff._root_vjp/_root_jvp
    (implicit_diff.py:139-183, 225-234) with mat = tangent at the converged Cook's membrane state, against SciPy."""
    import scipy.sparse.linalg as spla
    from autopdex_b200 import solver
    p, settings, static_settings = _cook_settings("bicgstab")
    settings = dict(settings, **{"load multiplier": 4.0})
    sol, infos = solver.solver(np.zeros(p["mask"].shape), settings, static_settings, tol=1e-12)
    assert not infos[2]
    rhs = np.random.default_rng(11).standard_normal(p["mask"].shape)
    u = solver.tangent_solve(sol, rhs, settings, static_settings, transpose=transpose, tol=1e-13)
    # oracle tangent at the same state (sets carry the load multiplier through the traction parameter)
    p4 = problems.cook_g2(q0=4.0)
    n = p["mask"].size
    free = ~p["mask"].ravel()
    _, data = oasm.assemble(p4["sets"], p4["coords"], sol, {})
    rows, cols = oasm.coo_indices(p4["sets"])
    K = oasm.scipy_assembling(data, rows, cols, n, free)
    ref = np.zeros(n)
    ref[free] = spla.spsolve((K.T if transpose else K).tocsc(), rhs.ravel()[free])
    assert np.all(u.ravel()[~free] == 0.0)
    assert np.linalg.norm(u.ravel() - ref) < 1e-8 * np.linalg.norm(ref)
