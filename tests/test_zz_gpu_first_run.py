"""GPU tests written in the CPU-only session 3 of round 1: their oracle halves were run on the CPU while they were
written; the device halves had their first run on a B200 with the last 95 GPU-seconds of the round: 12 passed
(profiles/r01i_first_run_tests_12_passed.txt).  The file sorts last so that under `pytest -x` anything new added here
cannot hide the long-verified suite; the tests can move into test_gpu_api.py / test_reference_fixtures.py."""
import numpy as np
import pytest

from oracle import assemble as oasm
from oracle import solve as osolve
from tests import problems
from tests.test_gpu_api import _cook_settings

pytestmark = pytest.mark.gpu


def test_user_element_hex8_backward_euler_steps():
    """BASELINE config 3 at parity size: transient heat conduction on a hex8 mesh as three 'user element' sets
    (poisson_weak conduction, forward_backward_euler_weak capacity, neumann_weak inflow on the face x = 1),
    'solver type': 'linear', backward-Euler steps with the pattern reused (maze_backward_euler.py:308-374)."""
    from autopdex_b200 import mesher, models, seeder, solver, spaces
    from oracle import quadrature as oquad
    n, dt, inflow = 4, 50.0 / 250.0, -1.0e3
    coords, elems = mesher.structured_mesh((n, n, n), problems.UNIT_CUBE, "brick")
    face = problems.boundary_faces_x1(n)
    gp3, gp2 = seeder.gauss_legendre_nd(3, 2), seeder.gauss_legendre_nd(2, 2)
    cond = models.isoparametric_domain_element_galerkin(models.poisson_weak(lambda x, s: 1.0), spaces.fem_iso_line_quad_brick, *gp3)
    cap = models.isoparametric_domain_element_galerkin(models.forward_backward_euler_weak(lambda x, s: 0.1),
                                                       spaces.fem_iso_line_quad_brick, *gp3)
    flux = models.isoparametric_surface_element_galerkin(models.neumann_weak(lambda x: inflow), spaces.fem_iso_line_quad_brick,
                                                         *gp2, tangent_contributions=False)
    static_settings = {"assembling mode": ("user element",) * 3, "solution structure": ("nodal imposition",) * 3,
                       "model": (cond, cap, flux), "solver type": "linear", "solver backend": "b200", "solver": "cg",
                       "type of preconditioner": "jacobi", "verbose": -1}
    mask = (np.abs(coords[:, 0]) < 1e-9)[:, None]
    values = np.zeros(mask.shape)
    dofs = np.zeros(mask.shape)
    settings = {"connectivity": (elems, elems, face), "node coordinates": coords, "dirichlet dofs": mask,
                "dirichlet conditions": values, "time increment": dt, "dofs n": dofs}
    osets = [dict(kind="domain", etype="hex8", conn=elems, nf=1, gp=oquad.gauss_legendre_nd(3, 2),
                  model=dict(name="poisson_weak", coefficient=1.0, source=0.0)),
             dict(kind="domain", etype="hex8", conn=elems, nf=1, gp=oquad.gauss_legendre_nd(3, 2),
                  model=dict(name="capacity", coefficient=0.1)),
             dict(kind="surface", etype="quad4", conn=face, nf=1, gp=oquad.gauss_legendre_nd(2, 2),
                  model=dict(name="neumann", traction=np.asarray([inflow])))]
    prob = osolve.Problem(osets, coords, mask, values, {"time increment": dt, "dofs n": dofs.copy()})
    ref = dofs.copy()
    n_plans = None
    for step in range(3):
        settings["dofs n"] = dofs
        dofs = dofs + solver.solver(dofs, settings, static_settings, tol=1e-13)[0]          # maze_backward_euler.py:362
        prob.settings["dofs n"] = ref
        ref = ref + osolve.solve_linear(prob, ref)
        assert np.linalg.norm(dofs - ref) / np.linalg.norm(ref) < 1e-8
        n_plans = len(solver._PLAN_CACHE) if n_plans is None else n_plans
        assert len(solver._PLAN_CACHE) == n_plans          # the plan (pattern) of step 1 serves the later steps
    assert ref.max() > 1.0                                  # the inflow heats the bar: a non-trivial state was compared


def test_assemble_tangent_diagonal_matches_oracle():
    """assembler.assemble_tangent_diagonal (assembler.py:639-680): flat diagonal of the unreduced tangent."""
    from autopdex_b200 import assembler
    p, settings, static_settings = _cook_settings()
    dofs = np.random.default_rng(0).uniform(-0.05, 0.05, p["mask"].shape)
    _, data = oasm.assemble(p["sets"], p["coords"], dofs, {})
    rows, cols = oasm.coo_indices(p["sets"])
    full = oasm.scipy_assembling(data, rows, cols, dofs.size)
    D = assembler.assemble_tangent_diagonal(dofs, settings, static_settings)
    assert D.shape == (dofs.size,) and np.abs(D - full.diagonal()).max() / np.abs(full.diagonal()).max() < 1e-12


def _gpu_vs_reference_run(tag, p):
    """Full CSR pattern (exact), values and residual of the CUDA path against the reference's own outputs."""
    import scipy.sparse as sp
    from autopdex_b200 import backend
    from tests import gpu_util
    from tests import test_reference_fixtures as trf
    FIX = trf.FIX
    plan = gpu_util.make_plan(p)
    n = p["mask"].size
    d, r = backend.DeviceArray.from_host(FIX[tag + "_dofs"].ravel()), backend.DeviceArray(n)
    plan.assemble(d, True, r)
    ref = sp.csr_matrix(sp.coo_matrix((FIX[tag + "_K_data"], (FIX[tag + "_K_rows"], FIX[tag + "_K_cols"])), shape=(n, n)))
    ref.sort_indices()
    indptr, indices = plan.csr(False)
    assert np.array_equal(indptr, ref.indptr) and np.array_equal(indices, ref.indices)
    assert trf.rel(plan.values(False), ref.data) < trf.TANGENT_RTOL
    assert trf.rel(r.download(), FIX[tag + "_R"].ravel()) < 1e-11
    plan.destroy()


@pytest.mark.parametrize("tag", ["pot_hex8", "pot_hex27", "pot_tet4", "pot_tri3", "pot_tri6", "pot_quad9", "pot_tet10"])
def test_gpu_against_reference_run_user_potential(tag):
    """The CUDA path against the reference's own outputs for the README-style 'user potential' on hex8 (the register
    kernel of BASELINE config 4), hex27, tet4, tet10, tri3, tri6 and quad9 elements (fixtures of case_potential3d)."""
    from tests import test_reference_fixtures as trf
    if tag + "_R" not in trf.FIX:
        pytest.skip("fixture %s not generated yet" % tag)
    p = trf.potential_problem(tag)
    p["mask"][0] = True                       # any mask: the full-CSR values do not depend on it
    _gpu_vs_reference_run(tag, p)


@pytest.mark.parametrize("tag", ["quad4_neo_line2", "tet4_neo_tri3", "tri3_linel"])
def test_gpu_against_reference_run_more_user_elements(tag):
    """Q1 quads + line2 traction, tet4 neo-Hooke + tri3 traction face, tri3 plain-stress elasticity (case_elements_more)."""
    from tests import test_reference_fixtures as trf
    if tag + "_R" not in trf.FIX:
        pytest.skip("fixture %s not generated yet" % tag)
    sets, coords, nf = trf._element_problem(tag)
    mask = np.zeros((coords.shape[0], nf), dtype=bool)
    mask[0] = True
    _gpu_vs_reference_run(tag, dict(sets=sets, coords=coords, mask=mask, values=np.zeros(mask.shape), nf=nf))
