"""Host logic of the product without a GPU: backend.Plan / DeviceArray are replaced by the oracle-backed stand-ins of
tests/oracle_backend.py, so everything the Python layer does between the reference-facing API and the C ABI --
settings parsing, device-set expansion, coefficient evaluation, the multigrid hierarchy, the dae time loop, verbose
output -- is executed and checked numerically.  The same scenarios run against the CUDA path in the `-m gpu` tests."""
import numpy as np
import pytest

from tests import oracle_backend, problems
from tests import test_zz_gpu_r02_dae as dae_cases


@pytest.fixture
def fake(monkeypatch):
    plan_cls = oracle_backend.install(monkeypatch)
    yield plan_cls
    from autopdex_b200 import solver
    solver._PLAN_CACHE.clear()


def test_stand_in_reproduces_readme_golden_value(fake):
    """Self-check of the stand-in: README problem (G1, tests/test_dicts_as_dofs_user_potential.py:62-63) through
    solver.solver."""
    from autopdex_b200 import mesher, models, seeder, solver, spaces, utility
    coords, elems = mesher.structured_mesh((5, 5), [[0., 0.], [1., 0.], [1., 1.], [0., 1.]], "quad")
    p = problems.readme_poisson(5)
    dd = {"phi": p["mask"][:, 0]}
    pot = models.mixed_reference_domain_potential(models.poisson_potential("phi", source_fun=problems.readme_source),
                                                  {"phi": spaces.fem_iso_line_quad_brick}, *seeder.gauss_legendre_nd(2, 2), "phi")
    st = {"assembling mode": ("user potential",), "solution structure": ("nodal imposition",), "model": (pot,),
          "solver type": "newton", "solver backend": "b200", "solver": "cg", "type of preconditioner": "jacobi", "verbose": -1}
    settings = {"connectivity": ({"phi": elems},), "dirichlet dofs": dd, "node coordinates": {"phi": coords},
                "dirichlet conditions": utility.dict_zeros_like(dd, dtype=np.float64)}
    sol, infos = solver.solver(utility.dict_zeros_like(dd, dtype=np.float64), settings, st)
    assert infos[0] == 1 and not infos[2]
    assert np.isclose(sol["phi"].sum(), 1.9066412530282952, rtol=1e-10, atol=0)


@pytest.mark.parametrize("scheme", ["backward_euler", "bdf2", "bdf3", "am1", "am2", "am4", "dirk1", "dirk2", "dirk3"])
def test_time_stepping_manager_host_loop(fake, scheme):
    """dae.TimeSteppingManager: integrator coefficients -> ('time increment', 'dofs n') of the capacity set, history
    roll, save policy, plan reuse -- against a SciPy loop on the oracle's mass / stiffness matrices."""
    from autopdex_b200 import dae, solver
    n, dt, n_steps = 4, 0.05, 4
    coords, K, M, F, mask, values, res, settings = dae_cases._settings(n)
    q0 = 0.3 * np.cos(coords[:, 1])
    integ, ref = dae_cases._reference_steps(scheme, K, M, F, mask, values, q0, dt, n_steps)
    static_settings = {"assembling mode": ("user residual",), "model": (res,), "time integrators": {"theta": integ},
                       "solver backend": "b200", "solver": "cg", "type of preconditioner": "jacobi", "verbose": -1}
    save = dae.SaveAllPolicy()
    out = dae.TimeSteppingManager(static_settings, save_policy=save).run({"theta": q0}, dt, dt * n_steps, 100, settings)
    assert (out.num_accepted, out.num_rejected, out.num_steps) == (n_steps, 0, n_steps)
    assert out.newton_iterations == [1] * n_steps
    for k in range(n_steps):
        assert np.linalg.norm(save.q[k + 1]["theta"] - ref[k]) / np.linalg.norm(ref[k]) < 1e-10, (scheme, k)
    assert len(fake.instances) == 1                     # one plan, two device sets on the same connectivity
    plan = fake.instances[0]
    assert [s.model for s in plan.sets] == ["poisson_potential", "capacity"]
    a_last = integ.stage_rule(integ.num_stages - 1, dt, np.zeros((integ.num_stages, 1)), np.zeros((integ.num_steps, 1)),
                              np.zeros((integ.num_steps, 1, 1)))[0]
    assert np.isclose(plan.dt, -1.0 / a_last)           # 'time increment' of the last stage solved: -1/a
    # the last pre_step_updates call of a step is the one of its last stage (dae.py:1900-1903): t_n + c_s dt
    assert out.settings["current time"] == pytest.approx(dt * (n_steps - 1) + dt * integ.stage_positions[-1])
    # t_max cuts the last step short and num_time_steps bounds the loop (dae.py:2133-2136)
    out2 = dae.TimeSteppingManager(static_settings).run({"theta": q0}, dt, 2.5 * dt, 100, settings)
    assert out2.num_accepted == 3 and out2.settings["current time"] == pytest.approx((2 + 0.5 * integ.stage_positions[-1]) * dt)
    out3 = dae.TimeSteppingManager(static_settings).run({"theta": q0}, dt, 10 * dt, 2, settings)
    assert out3.num_steps == 2


def test_root_iteration_controller_host(fake):
    dae_cases.test_root_iteration_controller_grows_and_rejects_like_the_reference()


@pytest.mark.parametrize("scheme", ["backward_euler", "bdf2", "am2", "dirk2", "dirk3"])
def test_time_stepping_manager_against_reference_manager_host(fake, scheme):
    dae_cases.test_time_stepping_manager_reproduces_trajectories_of_the_reference_manager(scheme)


@pytest.mark.parametrize("tag", ["hyperlin_plain_strain", "hyperlin_3d"])
def test_linear_elastic_strain_energy_route_host(fake, tag):
    from tests import test_zz_gpu_r02_hyper_linear as hl
    hl.test_linear_elastic_strain_energy_route_against_reference_run(tag)
    hl.test_other_strain_energies_are_rejected()
    hl.test_newton_solve_with_the_lame_mode_matches_oracle("cg")


def test_root_iteration_controller_against_reference_manager_host(fake):
    dae_cases.test_root_iteration_controller_inside_the_loop_reproduces_the_reference_manager()


def test_save_policies_and_postprocessing_host(fake):
    dae_cases.test_save_policies_and_postprocessing_mirror_the_reference()


def test_time_stepping_manager_rejections_host(fake):
    dae_cases.test_time_stepping_manager_rejections()


def test_user_residual_mode_host(fake):
    dae_cases.test_user_residual_mode_matches_user_potential_and_oracle()


def test_multigrid_hierarchy_host_construction(fake):
    """'type of preconditioner': 'multigrid': chain of coarse plans on the injected meshes, transfer operators sized to
    the free dofs of both levels, surface sets on the finest level only, per-call coarse field upload, destroy order."""
    from autopdex_b200 import solver
    from tests.multi_gpu_worker import neohooke_api_problem
    m = 8
    p = problems.neo_hooke_brick(m)
    static_settings = dict(neohooke_api_problem(p), solver="cg", **{"type of preconditioner": "multigrid"})
    settings = {"connectivity": tuple(s["conn"] for s in p["sets"]), "node coordinates": p["coords"], "dirichlet dofs": p["mask"],
                "dirichlet conditions": p["values"], "b200 multigrid": {"n_elements": (m, m, m), "pre": 3, "ratio": 5.0}}
    sol, (steps, res, div) = solver.solver(np.zeros(p["mask"].shape), settings, static_settings)
    assert not div and res < 1e-8
    plans = fake.instances
    assert [pl.n_nodes for pl in plans] == [9 ** 3, 5 ** 3, 3 ** 3]           # 8 -> 4 -> 2 elements per direction
    assert [len(pl.sets) for pl in plans] == [2, 1, 1]                       # the traction face stays on the finest level
    assert plans[0].coarse is plans[1] and plans[1].coarse is plans[2] and plans[2].coarse is None
    assert plans[0].mg_options == (3, 0, 0, 5.0, 0.0)
    from autopdex_b200 import mesher
    c4, e4 = mesher.structured_mesh((4, 4, 4), problems.UNIT_CUBE, "brick")
    assert np.allclose(plans[1].coords, c4) and np.array_equal(plans[1].sets[0].conn, e4)
    assert np.array_equal(plans[1].mask.reshape(-1, 3), np.repeat((np.abs(c4[:, 0]) < 1e-9)[:, None], 3, axis=1))
    # the prolongation reproduces linear fields: P x_c = x_f on the free dofs
    P = plans[0].transfer[0]
    import scipy.sparse as sp
    Pm = sp.csr_matrix((P[2], P[1], P[0]), shape=(plans[0].n_free, plans[1].n_free))
    lin_f = (p["coords"] @ np.array([1.0, 2.0, -1.0]))[:, None].repeat(3, axis=1).ravel()[~plans[0].mask]
    lin_c = (c4 @ np.array([1.0, 2.0, -1.0]))[:, None].repeat(3, axis=1).ravel()[~plans[1].mask]
    interior = np.asarray(np.abs(Pm).sum(axis=1)).ravel() > 1 - 1e-12         # rows whose coarse neighbours are all free
    assert interior.sum() > 0 and np.allclose((Pm @ lin_c)[interior], lin_f[interior])
    # second call: plans reused, coarse arrays not rebuilt
    cache_obj = next(iter(solver._PLAN_CACHE.values()))._mg_cache["coords"]
    solver.solver(np.zeros(p["mask"].shape), settings, static_settings)
    assert len(fake.instances) == 3 and next(iter(solver._PLAN_CACHE.values()))._mg_cache["coords"] is cache_obj
    solver.clear_plan_cache()
    assert all(pl.destroyed for pl in plans)


def test_multigrid_rejections_host(fake):
    from tests import test_zz_gpu_multigrid as mgt
    mgt.test_multigrid_rejections()


def test_verbose_lines_host(fake, capsys):
    dae_cases.test_verbose_prints_one_line_per_newton_iteration(capsys)


def test_two_field_dict_dofs_host(fake):
    """Multi-field dict dofs: concatenated node space, shifted connectivities, one pattern-only set per domain -- residual,
    BCOO order / zero blocks / values and the CSR pattern against the outputs of the unmodified reference."""
    dae_cases.test_two_field_dict_dofs_match_reference_run()
    plan = fake.instances[0]
    assert [s.model for s in plan.sets] == ["poisson_potential", "poisson_potential", "pattern_only", "pattern_only"]
    assert plan.sets[2].conn.shape[1] == 8 and plan.n_nodes == 24 and plan.nf == 1


def test_two_field_dict_dofs_newton_host(fake):
    dae_cases.test_two_field_dict_dofs_newton_solve()
