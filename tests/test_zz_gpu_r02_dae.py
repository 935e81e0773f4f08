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


def _dirichlet_solve(A, rhs, mask, values):
    free = ~mask
    A = A.tocsr()
    q = np.where(mask, values, 0.0)
    q[free] = spla.spsolve(A[free][:, free].tocsc(), rhs[free] - A[free][:, mask] @ values[mask])
    return q


def _scipy_steps_adams_moulton(K, M, F, mask, values, q0, num_steps, dt, n_steps):
    """dae.AdamsMoulton (dae.py:420-481) written out for M q_t + K q = F: q_t = ((q - q_n) / dt - sum_j c_j q_t_hist[j-1]) / c_0,
    derivative history starting at zero (dae.py:1805)."""
    from autopdex_b200.dae import AdamsMoulton
    c = AdamsMoulton._COEFFS[num_steps]
    q, hist, out = q0.copy(), [np.zeros_like(q0) for _ in range(num_steps)], []
    for _ in range(n_steps):
        past = sum(cj * h for cj, h in zip(c[1:], hist))
        q_new = _dirichlet_solve(M / (dt * c[0]) + K, F + M @ (q / (dt * c[0]) + past / c[0]), mask, values)
        q_t = ((q_new - q) / dt - past) / c[0]
        hist = [q_t] + hist[:-1]
        q = q_new
        out.append(q)
    return out


def _scipy_steps_dirk(K, M, F, mask, values, q0, num_stages, dt, n_steps):
    """dae.DiagonallyImplicitRungeKutta (dae.py:707-766) written out in the reference's own variables: stage unknown
    x_i (constrained to the Dirichlet values), slope K_i = (x_i - q_n) / (dt a_ii), residual at q_n + dt sum_j a_ij K_j."""
    r, al = np.sqrt(3.0), 2 * np.cos(np.pi / 18) / np.sqrt(3.0)
    A, b = {1: ([[0.5]], [1.0]),
            2: ([[0.5 + r / 6, 0.0], [-r / 3, 0.5 + r / 6]], [0.5, 0.5]),
            3: ([[(1 + al) / 2, 0, 0], [-al / 2, (1 + al) / 2, 0], [1 + al, -(1 + 2 * al), (1 + al) / 2]],
                [1 / (6 * al ** 2), 1 - 1 / (3 * al ** 2), 1 / (6 * al ** 2)])}[num_stages]
    q, out = q0.copy(), []
    for _ in range(n_steps):
        slopes = []
        for i in range(num_stages):
            d = dt * sum(A[i][j] * slopes[j] for j in range(i)) if i else np.zeros_like(q)
            g = 1.0 / (dt * A[i][i])
            x = _dirichlet_solve(g * M + K, F + M @ (g * q) - K @ d, mask, values)
            slopes.append((x - q) * g)
        q = q + dt * sum(bi * ki for bi, ki in zip(b, slopes))
        out.append(q)
    return out


def _integrator(scheme):
    from autopdex_b200 import dae
    if scheme == "backward_euler":
        return dae.BackwardEuler()
    if scheme.startswith("bdf"):
        return dae.BackwardDiffFormula(int(scheme[3:]))
    if scheme.startswith("am"):
        return dae.AdamsMoulton(int(scheme[2:]))
    return dae.DiagonallyImplicitRungeKutta(int(scheme[4:]))


def _reference_steps(scheme, K, M, F, mask, values, q0, dt, n_steps):
    """(integrator of autopdex_b200.dae, SciPy time loop) for a scheme name."""
    from autopdex_b200 import dae
    integ = _integrator(scheme)
    if scheme == "backward_euler":
        return integ, _scipy_steps(K, M, F, mask, values, q0, [1.0, -1.0], dt, n_steps)
    if scheme.startswith("bdf"):
        return integ, _scipy_steps(K, M, F, mask, values, q0, dae.BackwardDiffFormula._COEFFS[integ.num_steps], dt, n_steps)
    if scheme.startswith("am"):
        return integ, _scipy_steps_adams_moulton(K, M, F, mask, values, q0, integ.num_steps, dt, n_steps)
    return integ, _scipy_steps_dirk(K, M, F, mask, values, q0, integ.num_stages, dt, n_steps)


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


@pytest.mark.parametrize("scheme", ["backward_euler", "bdf2", "am1", "am3", "dirk1", "dirk2", "dirk3"])
def test_time_stepping_manager_heat_conduction_matches_scipy_loop(scheme):
    from autopdex_b200 import dae, solver
    n, dt, n_steps = 6, 0.05, 3
    solver.clear_plan_cache()                                      # plans of earlier tests must not count below
    coords, K, M, F, mask, values, res, settings = _settings(n)
    q0 = 0.3 * np.cos(coords[:, 1])
    integ, ref = _reference_steps(scheme, K, M, F, mask, values, q0, dt, n_steps)
    static_settings = {"assembling mode": ("user residual",), "model": (res,), "time integrators": {"theta": integ},
                       "solver backend": "b200", "solver": "cg", "type of preconditioner": "jacobi", "verbose": -1}
    save = dae.SaveAllPolicy()
    mgr = dae.TimeSteppingManager(static_settings, save_policy=save, tol=1e-13)
    out = mgr.run({"theta": q0}, dt, dt * n_steps, 100, settings)
    assert out.num_accepted == n_steps and out.num_rejected == 0 and out.num_steps == n_steps
    assert all(it == 1 for it in out.newton_iterations)            # linear problem: one solve per stage (dae.py:1640-1695)
    assert len(save.q) == n_steps + 1 and np.isclose(save.t[-1], dt * n_steps)
    for k in range(n_steps):
        assert np.linalg.norm(save.q[k + 1]["theta"] - ref[k]) / np.linalg.norm(ref[k]) < 1e-9, (scheme, k)
    assert np.array_equal(out.q["theta"], save.q[-1]["theta"])
    assert len(solver._PLAN_CACHE) == 1                            # one plan for all steps
    solver.clear_plan_cache()


def test_root_iteration_controller_grows_and_rejects_like_the_reference():
    """dae.RootIterationController (dae.py:1509-1573) around the device solve: the linear problem converges in one Newton
    iteration, so every accepted step scales dt by 1 + gamma (target - 1) / target up to max_step_size; with an
    unreachable tolerance every attempt is rejected, dt is halved down to min_step_size and the run is interrupted
    (q -> NaN, dae.py:2245-2249)."""
    from autopdex_b200 import dae, solver
    n, dt0 = 4, 0.02
    coords, K, M, F, mask, values, res, settings = _settings(n)
    static_settings = {"assembling mode": ("user residual",), "model": (res,), "time integrators": {"theta": dae.BackwardEuler()},
                       "solver backend": "b200", "solver": "cg", "type of preconditioner": "jacobi", "verbose": -1}
    q0 = 0.3 * np.cos(coords[:, 1])
    ctrl = dae.RootIterationController(target_niters=6, gamma=0.5, max_step_size=0.05)
    save = dae.SaveAllPolicy()
    out = dae.TimeSteppingManager(static_settings, save_policy=save, step_size_controller=ctrl, tol=1e-13).run(
        {"theta": q0}, dt0, 0.2, 100, settings)
    # the reference's step sequence, written out
    t, dt, q, times = 0.0, dt0, q0, []
    while t < 0.2 * (1 - 1e-14):
        t_new = min(t + dt, 0.2)
        q = _scipy_steps(K, M, F, mask, values, q, [1.0, -1.0], t_new - t, 1)[0]
        dt = float(np.clip((t_new - t) * (1 + 0.5 * (6 - 1) / 6), 1e-6, 0.05))
        t = t_new
        times.append(t)
    assert out.num_accepted == len(times) and out.num_rejected == 0
    assert np.allclose(save.t[1:], times, rtol=1e-13)
    assert np.linalg.norm(out.q["theta"] - q) / np.linalg.norm(q) < 1e-9
    # never converging: rejected, halved, interrupted at the minimum step size
    ctrl = dae.RootIterationController(min_step_size=dt0 / 4)
    out = dae.TimeSteppingManager(static_settings, step_size_controller=ctrl, atol=1e-300, max_iter=1, tol=1e-13).run(
        {"theta": q0}, dt0, 0.2, 100, settings)
    assert (out.num_accepted, out.num_rejected) == (0, 2) and np.isnan(out.q["theta"]).all()
    # the constant controller cannot repeat a failed step: interrupted at once
    out = dae.TimeSteppingManager(static_settings, atol=1e-300, max_iter=1, tol=1e-13).run({"theta": q0}, dt0, 0.2, 100, settings)
    assert (out.num_accepted, out.num_rejected) == (0, 1) and np.isnan(out.q["theta"]).all()
    solver.clear_plan_cache()


@pytest.mark.parametrize("scheme", ["backward_euler", "bdf2", "am2", "dirk2", "dirk3"])
def test_time_stepping_manager_reproduces_trajectories_of_the_reference_manager(scheme):
    """Against the REFERENCE'S OWN TimeSteppingManager.run (dae.py:2087-2268), executed by
    tests/golden/make_reference_fixtures.py (case `dae_manager`) on the same semi-discrete system given as a dense 'dae'
    callable: stage loop, history roll, integrator updates and the constrained dofs of dae.newton_solver are the
    reference's code there, assembly + Newton + Krylov are the device's here."""
    import os
    from autopdex_b200 import dae, solver
    fix = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_fixtures.npz"))
    coords, K, M, F, mask, values, res, settings = _settings(2)
    static_settings = {"assembling mode": ("user residual",), "model": (res,), "time integrators": {"theta": _integrator(scheme)},
                       "solver backend": "b200", "solver": "cg", "type of preconditioner": "jacobi", "verbose": -1}
    q0 = 0.3 * np.cos(coords[:, 1])
    save = dae.SaveAllPolicy()
    out = dae.TimeSteppingManager(static_settings, save_policy=save, atol=1e-12, tol=1e-14).run({"theta": q0}, 0.05, 0.15, 3, settings)
    ref_q, ref_t = fix["dae_manager_%s_q" % scheme], fix["dae_manager_%s_t" % scheme]
    assert out.num_accepted == 3 and np.allclose(out.history.t[:4], ref_t, rtol=1e-14)
    for k in range(4):
        assert np.linalg.norm(out.history.q["theta"][k] - ref_q[k]) / np.linalg.norm(ref_q[k]) < 1e-9, (scheme, k)
    solver.clear_plan_cache()


def test_root_iteration_controller_inside_the_loop_reproduces_the_reference_manager():
    """The step-size controller inside the time loop, against the reference's own manager run (fixture case `dae_manager`,
    RootIterationController(target 6, gamma 0.5, max step 0.05), dt0 = 0.02, t_max = 0.2): the same accepted times (growth
    by 17/12 per step, capped, the last step cut at t_max) and the same states (the reference run carries the ~1e-10
    error of the stand-in's difference-quotient Jacobian at its default tolerance)."""
    import os
    from autopdex_b200 import dae, solver
    fix = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_fixtures.npz"))
    coords, K, M, F, mask, values, res, settings = _settings(2)
    static_settings = {"assembling mode": ("user residual",), "model": (res,), "time integrators": {"theta": dae.BackwardEuler()},
                       "solver backend": "b200", "solver": "cg", "type of preconditioner": "jacobi", "verbose": -1}
    ctrl = dae.RootIterationController(target_niters=6, gamma=0.5, max_step_size=0.05)
    out = dae.TimeSteppingManager(static_settings, save_policy=dae.SaveAllPolicy(), step_size_controller=ctrl, tol=1e-14).run(
        {"theta": 0.3 * np.cos(coords[:, 1])}, 0.02, 0.2, 12, settings)
    ref_t, ref_q = fix["dae_manager_root_controller_t"], fix["dae_manager_root_controller_q"]
    assert [out.num_accepted, out.num_rejected] == list(fix["dae_manager_root_controller_counts"])
    n = out.num_accepted + 1
    assert np.allclose(out.history.t[:n], ref_t, rtol=1e-13) and np.isnan(out.history.t[n:]).all()
    for k in range(n):
        assert np.linalg.norm(out.history.q["theta"][k] - ref_q[k]) / np.linalg.norm(ref_q[k]) < 1e-7, k
    solver.clear_plan_cache()


def test_save_policies_and_postprocessing_mirror_the_reference():
    """dae.SaveAllPolicy / SaveEquidistantPolicy / SaveNothingPolicy (dae.py:1160-1311) and the user data of
    `postprocessing_fun` (dae.py:2140, 2188): pre-allocated arrays of max_steps + 1 rows padded with NaN, equidistant
    targets hit by the first accepted step at or after them."""
    from autopdex_b200 import dae, solver
    n, dt, n_steps = 4, 0.05, 4
    coords, K, M, F, mask, values, res, settings = _settings(n)
    static_settings = {"assembling mode": ("user residual",), "model": (res,), "time integrators": {"theta": dae.BackwardEuler()},
                       "solver backend": "b200", "solver": "cg", "type of preconditioner": "jacobi", "verbose": -1}
    q0 = 0.3 * np.cos(coords[:, 1])
    ref = _scipy_steps(K, M, F, mask, values, q0, [1.0, -1.0], dt, n_steps)
    post = lambda q_fun, t, settings: {"mean": np.asarray(q_fun(t)["theta"]).mean(), "time": np.asarray(t)}
    out = dae.TimeSteppingManager(static_settings, save_policy=dae.SaveAllPolicy(), postprocessing_fun=post, tol=1e-13).run(
        {"theta": q0}, dt, dt * n_steps, 6, settings)
    h = out.history
    assert isinstance(h, dae.HistoryState) and h.t.shape == (7,) and h.q["theta"].shape == (7, coords.shape[0])
    assert np.allclose(h.t[:5], dt * np.arange(5)) and np.isnan(h.t[5:]).all() and np.isnan(h.q["theta"][5:]).all()
    assert np.array_equal(h.q["theta"][0], q0)
    for k in range(n_steps):
        assert np.linalg.norm(h.q["theta"][k + 1] - ref[k]) / np.linalg.norm(ref[k]) < 1e-9
    assert np.allclose(h.user["mean"][:5], h.q["theta"][:5].mean(axis=1)) and np.allclose(h.user["time"][:5], h.t[:5])
    # two equidistant intervals over four steps: rows at t = 0, 2 dt, 4 dt
    out = dae.TimeSteppingManager(static_settings, save_policy=dae.SaveEquidistantPolicy(num_points=2), tol=1e-13).run(
        {"theta": q0}, dt, dt * n_steps, 6, settings)
    assert np.allclose(out.history.t, [0.0, 2 * dt, 4 * dt])
    assert np.linalg.norm(out.history.q["theta"][1] - ref[1]) / np.linalg.norm(ref[1]) < 1e-9
    assert np.linalg.norm(out.history.q["theta"][2] - ref[3]) / np.linalg.norm(ref[3]) < 1e-9
    out = dae.TimeSteppingManager(static_settings, save_policy=dae.SaveNothingPolicy(), tol=1e-13).run({"theta": q0}, dt, dt, 6, settings)
    assert out.history is None and out.num_accepted == 1
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
    dae.TimeSteppingManager(dict(base, dae="call pde"))                 # the reference's key for the PDE route is accepted
    with pytest.raises(ValueError, match="call pde"):                   # its ODE mode (a user-written residual) is not
        dae.TimeSteppingManager(dict(base, dae=lambda q_fun, t, settings: 0.0))
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


# ---- multi-field dict dofs -------------------------------------------------------------
def _two_field_case():
    """The problem of tests/golden/make_reference_fixtures.py:case_two_fields with the tagged integrands."""
    from autopdex_b200 import models, seeder, spaces
    from tests.test_reference_fixtures import FIX
    e = FIX["two_fields_elems"]
    gp = seeder.gauss_legendre_nd(dimension=2, order=2)
    ans = {"phi": spaces.fem_iso_line_quad_brick, "psi": spaces.fem_iso_line_quad_brick}
    src = lambda a, b: (lambda x: a * np.sin(2.0 * np.sum(x * x, axis=-1)) + b)
    pot1 = models.mixed_reference_domain_potential(models.poisson_potential("phi", source_fun=src(3.0, -1.0)), ans, *gp, "phi")
    pot2 = models.mixed_reference_domain_potential(
        models.poisson_potential("psi", source_fun=src(-2.0, 0.5), coefficient_fun=lambda x: 2.5 + 0.0 * x[..., 0]), ans, *gp, "psi")
    static_settings = {"assembling mode": ("user potential", "user potential"),
                       "solution structure": ("nodal imposition", "nodal imposition"), "model": (pot1, pot2),
                       "solver type": "newton", "solver backend": "b200", "solver": "cg", "type of preconditioner": "jacobi",
                       "verbose": -1}
    n = FIX["two_fields_c_phi"].shape[0]
    conn = {"phi": e, "psi": e}
    settings = {"connectivity": (conn, conn), "dirichlet dofs": {"phi": np.zeros(n, dtype=bool), "psi": np.zeros(n, dtype=bool)},
                "node coordinates": {"phi": FIX["two_fields_c_phi"], "psi": FIX["two_fields_c_psi"]},
                "dirichlet conditions": {"phi": np.zeros(n), "psi": np.zeros(n)}}
    dofs = {"phi": FIX["two_fields_dofs_phi"], "psi": FIX["two_fields_dofs_psi"]}
    return FIX, settings, static_settings, dofs, n


def test_two_field_dict_dofs_match_reference_run():
    """Residual, BCOO (order, explicit zero blocks, values) and CSR pattern of a two-field dict-dof problem against the
    outputs of the unmodified reference (reference-run fixture `two_fields_*`)."""
    import scipy.sparse as sp
    from autopdex_b200 import assembler
    FIX, settings, static_settings, dofs, n = _two_field_case()
    R = assembler.assemble_residual(dofs, settings, static_settings)
    assert set(R.keys()) == {"phi", "psi"}
    for k in ("phi", "psi"):
        assert np.abs(R[k] - FIX["two_fields_R_" + k]).max() / np.abs(FIX["two_fields_R_" + k]).max() < 1e-11
    B = assembler.assemble_tangent(dofs, settings, static_settings)
    assert np.array_equal(B.indices[:, 0], FIX["two_fields_K_rows"]) and np.array_equal(B.indices[:, 1], FIX["two_fields_K_cols"])
    ref = FIX["two_fields_K_data"]
    assert np.array_equal(B.data == 0.0, ref == 0.0)                       # the same explicit zero blocks
    assert np.abs(B.data - ref).max() / np.abs(ref).max() < 2e-7           # numerical AD in the fixture generator
    want = sp.csr_matrix(sp.coo_matrix((ref, (FIX["two_fields_K_rows"], FIX["two_fields_K_cols"])), shape=(2 * n, 2 * n)))
    want.sort_indices()
    K = B.sum_duplicates()
    assert np.array_equal(K.indptr, want.indptr) and np.array_equal(K.indices, want.indices)   # pattern bit-exact
    assert np.abs(K.data - want.data).max() / np.abs(want.data).max() < 2e-7


def test_two_field_dict_dofs_newton_solve():
    """The coupled (block-diagonal) system solved as ONE Newton problem: each field equals its stand-alone solve."""
    from autopdex_b200 import models, seeder, solver, spaces
    FIX, settings, static_settings, dofs, n = _two_field_case()
    c = FIX["two_fields_c_phi"]
    m_phi = np.abs(c[:, 0]) < 1e-12
    m_psi = np.abs(c[:, 1]) < 1e-12
    settings = dict(settings, **{"dirichlet dofs": {"phi": m_phi, "psi": m_psi},
                                 "dirichlet conditions": {"phi": np.where(m_phi, 0.5, 0.0), "psi": np.where(m_psi, -1.0, 0.0)}})
    zero = {"phi": np.zeros(n), "psi": np.zeros(n)}
    sol, (steps, res, div) = solver.solver(zero, settings, static_settings, tol=1e-13)
    assert steps == 1 and not div and res < 1e-9
    for k, j in (("phi", 0), ("psi", 1)):
        st1 = dict(static_settings, **{"assembling mode": ("user potential",), "solution structure": ("nodal imposition",)})
        gp = seeder.gauss_legendre_nd(dimension=2, order=2)
        integ = static_settings["model"][j]
        st1["model"] = (integ,)
        s1 = {"connectivity": ({k: settings["connectivity"][j][k]},), "dirichlet dofs": {k: settings["dirichlet dofs"][k]},
              "node coordinates": {k: settings["node coordinates"][k]}, "dirichlet conditions": {k: settings["dirichlet conditions"][k]}}
        # the single-field factory checks are the same object: build a one-field model of the same integrand
        w = integ.weak
        one = models.mixed_reference_domain_potential(
            models.poisson_potential(k, source_fun=w.funs["source"], coefficient_fun=w.funs["coefficient"]),
            {k: spaces.fem_iso_line_quad_brick}, *gp, k)
        st1["model"] = (one,)
        alone, _ = solver.solver({k: np.zeros(n)}, s1, st1, tol=1e-13)
        assert np.linalg.norm(sol[k] - alone[k]) / np.linalg.norm(alone[k]) < 1e-9
    with pytest.raises(ValueError):                                   # different dofs per node across the fields
        solver.solver({"phi": np.zeros(n), "psi": np.zeros((n, 2))}, settings, static_settings)
    solver.clear_plan_cache()
