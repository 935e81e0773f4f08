"""hyperelastic_steady_state_weak with models.linear_elastic_strain_energy (reference: models.py:917-1000, 1167-1185)
through the reference-facing API: the b200 backend maps it onto the linear-elasticity kernel with the isotropic tensor in
the mesh's dimension ('plain strain': eps_33 = 0, the customary matrix -- not the one of linear_elasticity_weak, whose
shear entry the reference doubles).  Checked against outputs of the reference's own assembler (fixture case
`hyper_linear`, tests/golden/make_reference_fixtures.py); the host half runs in the CPU suite against the oracle-backed
stand-in (tests/test_host_logic_oracle_backend.py)."""
import numpy as np
import pytest

from tests import test_reference_fixtures as trf

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag", trf.ELEMENT_TAGS_HYPERLIN)
def test_linear_elastic_strain_energy_route_against_reference_run(tag):
    from autopdex_b200 import assembler, models, seeder, solver, spaces
    sets, coords, nf = trf._element_problem(tag)
    dim = coords.shape[1]
    weak = models.hyperelastic_steady_state_weak(models.linear_elastic_strain_energy, lambda x, settings: settings["youngs modulus"],
                                                 lambda x, settings: settings["poisson ratio"], "3d" if dim == 3 else "plain strain")
    elem = models.isoparametric_domain_element_galerkin(weak, spaces.fem_iso_line_quad_brick, *seeder.gauss_legendre_nd(dimension=dim, order=2))
    static_settings = {"number of fields": (nf,), "assembling mode": ("user element",), "solution structure": ("nodal imposition",),
                       "model": (elem,), "solver type": "newton", "solver backend": "b200", "solver": "cg",
                       "type of preconditioner": "jacobi", "verbose": -1}
    settings = {"dirichlet dofs": np.zeros((coords.shape[0], nf), dtype=bool), "connectivity": (sets[0]["conn"],),
                "node coordinates": coords, "dirichlet conditions": np.zeros((coords.shape[0], nf)),
                "youngs modulus": 100.0, "poisson ratio": 0.3}
    dofs = trf.FIX[tag + "_dofs"]
    R = assembler.assemble_residual(dofs, settings, static_settings)
    K = assembler.assemble_tangent(dofs, settings, static_settings)
    assert np.array_equal(np.asarray(K.indices)[:, 0], trf.FIX[tag + "_K_rows"])
    assert np.array_equal(np.asarray(K.indices)[:, 1], trf.FIX[tag + "_K_cols"])
    assert trf.rel(np.asarray(R).ravel(), trf.FIX[tag + "_R"].ravel()) < 1e-11
    assert trf.rel(np.asarray(K.data), trf.FIX[tag + "_K_data"]) < trf.TANGENT_RTOL
    solver.clear_plan_cache()


def test_other_strain_energies_are_rejected():
    from autopdex_b200 import models

    def isochoric_neo_hooke(F, mu):
        return 0.0
    with pytest.raises(ValueError, match="strain energy"):
        models.hyperelastic_steady_state_weak(isochoric_neo_hooke, lambda x: 1.0, lambda x: 0.3, "3d")


@pytest.mark.parametrize("krylov", ["cg", "bicgstab"])
def test_newton_solve_with_the_lame_mode_matches_oracle(krylov):
    """The whole path (assembly, symmetric sliced-ELL matrix, Jacobi-Krylov, Newton) for the new material mode: a clamped
    plane-strain membrane under a body load (an extension: the reference cannot evaluate a hyperelastic weak form with a
    volume load, models.py:996-997), against the oracle's 'scipy' path -- one Newton step, solution to 1e-8."""
    from autopdex_b200 import models, seeder, solver, spaces
    from oracle import solve as osolve
    from tests import problems
    p = problems.elasticity_quad(6, mode="lame")
    prob = osolve.Problem(p["sets"], p["coords"], p["mask"], p["values"])
    ref, (rsteps, _, rdiv) = osolve.damped_newton(prob, np.zeros(p["mask"].shape))
    weak = models.hyperelastic_steady_state_weak(models.linear_elastic_strain_energy, lambda x: 100.0, lambda x: 0.3, "plain strain",
                                                 lambda x: np.asarray([0.0, -1.0]))
    elem = models.isoparametric_domain_element_galerkin(weak, spaces.fem_iso_line_quad_brick, *seeder.gauss_legendre_nd(dimension=2, order=2))
    static_settings = {"number of fields": (2,), "assembling mode": ("user element",), "solution structure": ("nodal imposition",),
                       "model": (elem,), "solver type": "newton", "solver backend": "b200", "solver": krylov,
                       "type of preconditioner": "jacobi", "verbose": -1}
    settings = {"dirichlet dofs": p["mask"], "connectivity": (p["sets"][0]["conn"],), "node coordinates": p["coords"],
                "dirichlet conditions": p["values"]}
    sol, (steps, res, div) = solver.solver(np.zeros(p["mask"].shape), settings, static_settings, tol=1e-13)
    assert steps == rsteps == 1 and not div and not rdiv
    assert np.linalg.norm(np.asarray(sol).ravel() - ref.ravel()) / np.linalg.norm(ref) < 1e-8
    solver.clear_plan_cache()
