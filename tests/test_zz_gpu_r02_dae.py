"""(Written after the round's GPU budget was spent: the host half of every scenario here runs on the CPU against the
oracle-backed stand-in, tests/test_host_logic_oracle_backend.py; this file sorts last so that a first-run surprise cannot
mask the validated suite under `pytest -x`.)

GPU parity of the dae.TimeSteppingManager hook (reference: autopdex/dae.py:1734-2240, PDE
path :1809-1876 / :1878-2085) and of the steady 'user residual' assembling mode (assembler.py:1217-1308) with tagged
integrands: transient heat conduction c theta_t + div(-k grad theta) = f advanced by BackwardEuler / BDF2, against a
SciPy time loop on the oracle's mass and stiffness matrices."""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

from oracle import assemble as oasm
from oracle import quadrature as oquad
from tests import problems

pytestmark = pytest.mark.gpu

K_COND, C_CAP = 2.0, 0.5


def _source(x):
    return 3.0 * x[..., 1] + np.sin(2.0 * x[..., 0])


def _oracle_matrices(n):
    from autopdex_b200 import mesher
    coords, elems = mesher.structured_mesh((n, n, n), problems.UNIT_CUBE, "brick")
    gp = oquad.gauss_legendre_nd(3, 2)
    dom = dict(kind="domain", etype="hex8", conn=elems, nf=1, gp=gp, model=dict(name="poisson_potential", coefficient=K_COND))
    from oracle import elements as oel
    f = _source(oel.gauss_point_coordinates(dom, coords))
    dom["model"]["source"] = f
    nn = coords.shape[0]
    R0, dK = oasm.assemble([dom], coords, np.zeros((nn, 1)), {})
    rows, cols = oasm.coo_indices([dom])
    K = oasm.scipy_assembling(dK, rows, cols, nn)
    cap = dict(kind="domain", etype="hex8", conn=elems, nf=1, gp=gp, model=dict(name="capacity", coefficient=C_CAP))
    _, dM = oasm.assemble([cap], coords, np.zeros((nn, 1)), {"time increment": -1.0, "dofs n": np.zeros((nn, 1))})
    M = oasm.scipy_assembling(dM, rows, cols, nn)
    return coords, elems, K, M, -R0


def _scipy_steps(K, M, F, mask, values, q0, integ_coeffs, dt, n_steps):
    """(a M + K) q = F - M b with q_t = a q + b, Dirichlet rows eliminated."""
    free = ~mask
    hist = [q0.copy() for _ in range(len(integ_coeffs) - 1)]
    out = []
    for _ in range(n_steps):
        a = integ_coeffs[0] / dt
        b = sum(c / dt * h for c, h in zip(integ_coeffs[1:], hist))
        A = (a * M + K).tocsr()
        rhs = F - M @ b
        q = np.where(mask, values, 0.0)
        q[free] = spla.spsolve(A[free][:, free].tocsc(), rhs[free] - A[free][:, mask] @ values[mask])
        hist = [q] + hist[:-1]
        out.append(q)
    return out


def _settings(n):
    from autopdex_b200 import models, seeder, spaces
    coords, elems, K, M, F = _oracle_matrices(n)
    mask = np.abs(coords[:, 0]) < 1e-12
    values = np.where(mask, 1.0 + coords[:, 2], 0.0)
    integrand = models.heat_conduction_time("theta", conductivity_fun=lambda x: K_COND, capacity_fun=lambda x: C_CAP,
                                            source_fun=_source)
    res = models.mixed_reference_domain_residual_time(integrand, {"theta": spaces.fem_iso_line_quad_brick},
                                                      *seeder.gauss_legendre_nd(3, 2), "theta")
    settings = {"connectivity": ({"theta": elems},), "node coordinates": {"theta": coords}, "dirichlet dofs": {"theta": mask},
                "dirichlet conditions": {"theta": values}, "current time": 0.0}
    return coords, K, M, F, mask, values, res, settings


@pytest.mark.parametrize("scheme", ["backward_euler", "bdf2"])
def test_time_stepping_manager_heat_conduction_matches_scipy_loop(scheme):
    from autopdex_b200 import dae, solver
    n, dt, n_steps = 6, 0.05, 3
    coords, K, M, F, mask, values, res, settings = _settings(n)
    integ, coeffs = (dae.BackwardEuler(), [1.0, -1.0]) if scheme == "backward_euler" else (dae.BackwardDiffFormula(2), [1.5, -2.0, 0.5])
    static_settings = {"assembling mode": ("user residual",), "model": (res,), "time integrators": {"theta": integ},
                       "solver backend": "b200", "solver": "cg", "type of preconditioner": "jacobi", "verbose": -1}
    q0 = 0.3 * np.cos(coords[:, 1])
    save = dae.SaveAllPolicy()
    mgr = dae.TimeSteppingManager(static_settings, save_policy=save, tol=1e-13)
    out = mgr.run({"theta": q0}, dt, dt * n_steps, 100, settings)
    assert out.num_accepted == n_steps and out.num_rejected == 0 and out.num_steps == n_steps
    assert all(it == 1 for it in out.newton_iterations)            # linear problem: one solve per step (dae.py:1640-1695)
    ref = _scipy_steps(K, M, F, mask, values, q0, coeffs, dt, n_steps)
    assert len(save.q) == n_steps + 1 and np.isclose(save.t[-1], dt * n_steps)
    for k in range(n_steps):
        assert np.linalg.norm(save.q[k + 1]["theta"] - ref[k]) / np.linalg.norm(ref[k]) < 1e-9, (scheme, k)
    assert np.array_equal(out.q["theta"], save.q[-1]["theta"])
    assert len(solver._PLAN_CACHE) == 1                            # one plan for all steps
    solver.clear_plan_cache()


def test_time_stepping_manager_with_multigrid():
    from autopdex_b200 import dae, solver
    n, dt = 8, 0.1
    coords, K, M, F, mask, values, res, settings = _settings(n)
    settings["b200 multigrid"] = {"n_elements": (n, n, n)}
    static_settings = {"assembling mode": ("user residual",), "model": (res,), "time integrators": {"theta": dae.BackwardEuler()},
                       "solver backend": "b200", "solver": "cg", "type of preconditioner": "multigrid", "verbose": -1}
    q0 = np.zeros(coords.shape[0])
    out = dae.TimeSteppingManager(static_settings, tol=1e-13).run({"theta": q0}, dt, 2 * dt, 10, settings)
    ref = _scipy_steps(K, M, F, mask, values, q0, [1.0, -1.0], dt, 2)
    assert out.num_accepted == 2
    assert np.linalg.norm(out.q["theta"] - ref[-1]) / np.linalg.norm(ref[-1]) < 1e-9
    solver.clear_plan_cache()


def test_time_stepping_manager_rejections():
    from autopdex_b200 import dae, models, seeder, spaces
    coords, K, M, F, mask, values, res, settings = _settings(2)
    base = {"assembling mode": ("user residual",), "model": (res,), "time integrators": {"theta": dae.BackwardEuler()},
            "solver backend": "b200", "solver": "cg", "type of preconditioner": "jacobi"}
    with pytest.raises(ValueError):
        dae.TimeSteppingManager(dict(base, **{"time integrators": {"theta": object()}}))
    with pytest.raises(ValueError):
        dae.TimeSteppingManager(dict(base, **{"solver backend": "scipy"}))
    with pytest.raises(ValueError):
        dae.TimeSteppingManager(base, root_solver=lambda *a: None)
    with pytest.raises(ValueError):                                 # a user-written time-dependent integrand
        models.mixed_reference_domain_residual_time(lambda *a: 0.0, {"theta": spaces.fem_iso_line_quad_brick},
                                                    *seeder.gauss_legendre_nd(3, 2), "theta")
    steady = models.mixed_reference_domain_residual(models.poisson_residual("theta"), {"theta": spaces.fem_iso_line_quad_brick},
                                                    *seeder.gauss_legendre_nd(3, 2), "theta")
    with pytest.raises(ValueError):                                 # nothing transient to advance
        dae.TimeSteppingManager(dict(base, model=(steady,)))


def test_user_residual_mode_matches_user_potential_and_oracle():
    """'user residual' (assembler.py:1217-1308) with the tagged weak form c grad(phi).grad(dphi) - f dphi: the same
    residual, BCOO tangent and solution as the 'user potential' route of the README problem."""
    from autopdex_b200 import assembler, mesher, models, seeder, solver, spaces
    n = 6
    pts = [[0., 0.], [1., 0.], [1., 1.], [0., 1.]]
    coords, elems = mesher.structured_mesh((n, n), pts, "quad")
    p = problems.readme_poisson(n)
    mask = {"phi": p["mask"][:, 0]}
    gp = seeder.gauss_legendre_nd(dimension=2, order=2)
    pot = models.mixed_reference_domain_potential(models.poisson_potential("phi", source_fun=problems.readme_source),
                                                  {"phi": spaces.fem_iso_line_quad_brick}, *gp, "phi")
    res = models.mixed_reference_domain_residual(models.poisson_residual("phi", source_fun=problems.readme_source),
                                                 {"phi": spaces.fem_iso_line_quad_brick}, *gp, "phi")
    settings = {"connectivity": ({"phi": elems},), "dirichlet dofs": mask, "node coordinates": {"phi": coords},
                "dirichlet conditions": {"phi": np.zeros(coords.shape[0])}}
    out = {}
    for mode, model in (("user potential", pot), ("user residual", res)):
        st = {"assembling mode": (mode,), "solution structure": ("nodal imposition",), "model": (model,), "solver type": "newton",
              "solver backend": "b200", "solver": "cg", "type of preconditioner": "jacobi", "verbose": -1}
        sol, infos = solver.solver({"phi": np.zeros(coords.shape[0])}, settings, st, tol=1e-13)
        assert infos[0] == 1 and not infos[2]
        d = {"phi": np.random.default_rng(1).uniform(-1, 1, coords.shape[0])}
        out[mode] = (sol["phi"], assembler.assemble_residual(d, settings, st)["phi"], assembler.assemble_tangent(d, settings, st))
    a, b = out["user potential"], out["user residual"]
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert np.array_equal(a[2].indices, b[2].indices) and np.array_equal(a[2].data, b[2].data)
    with pytest.raises(ValueError):                                 # mode / model mismatch is rejected when settings are read
        solver.solver({"phi": np.zeros(coords.shape[0])}, settings,
                      {"assembling mode": ("user potential",), "solution structure": ("nodal imposition",), "model": (res,),
                       "solver type": "newton", "solver backend": "b200", "solver": "cg"})
    solver.clear_plan_cache()


def test_verbose_prints_one_line_per_newton_iteration(capsys):
    """solver.py:906-909: "Residual after Newton iteration {i}: {res}" for every iteration when verbose > 0."""
    from autopdex_b200 import solver
    from tests.test_gpu_api import _cook_settings
    p, settings, static_settings = _cook_settings()
    settings = dict(settings, **{"load multiplier": 1.0})
    sol, (n_it, res, div) = solver.solver(np.zeros(p["mask"].shape), settings, dict(static_settings, verbose=1), tol=1e-12)
    lines = [l for l in capsys.readouterr().out.splitlines() if l.startswith("Residual after Newton iteration")]
    assert len(lines) == n_it and n_it > 1 and not div
    assert lines[0].startswith("Residual after Newton iteration 1: ")
    assert np.isclose(float(lines[-1].split(": ")[1]), res)
    hist = [float(l.split(": ")[1]) for l in lines]
    assert hist[-1] < 1e-8 < hist[0]
