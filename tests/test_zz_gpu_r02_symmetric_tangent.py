"""(Written after the round's GPU budget was spent; executed on the emulated build, tests/emu.)

The generic element kernel integrates the upper node pairs only and stores every block twice, transposed (csrc/elements.cu,
phase C): the assembled tangents must be symmetric to the last bit and still equal the oracle's to 1e-12."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import assemble as oasm
from tests import problems

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", ["neo_hooke_hex8", "elasticity_quad9", "poisson_hex27"])
def test_generic_kernel_tangent_is_bitwise_symmetric_and_matches_oracle(case):
    from autopdex_b200 import backend
    from tests import gpu_util
    if case == "neo_hooke_hex8":
        p = problems.neo_hooke_brick(4)
    elif case == "elasticity_quad9":
        p = problems.elasticity_quad(4, etype="quad9")
    else:
        p = problems.poisson_hex(3, etype="hex27", distort=0.15)
    plan = gpu_util.make_plan(p)
    n = p["mask"].size
    rng = np.random.default_rng(7)
    dofs = 1e-2 * rng.standard_normal(n)
    d, r = backend.DeviceArray.from_host(dofs), backend.DeviceArray(n)
    plan.assemble(d, True, r)
    indptr, indices = plan.csr(False)
    A = sp.csr_matrix((plan.values(False), indices, indptr), shape=(n, n))
    assert (A != A.T).nnz == 0                                   # bitwise: (b, a) stores the transposed (a, b) block
    R, data = oasm.assemble(p["sets"], p["coords"], dofs.reshape(p["mask"].shape), p.get("settings") or {})
    rows, cols = oasm.coo_indices(p["sets"])
    ref = oasm.scipy_assembling(data, rows, cols, n)
    assert np.abs(A.data - ref.data).max() <= 1e-12 * np.abs(ref.data).max()
    # the reference-order element streams (BCOO data of assembler.assemble_tangent) carry the same values
    coo = plan.coo_values()
    assert np.abs(coo - data).max() <= 1e-12 * np.abs(data).max()
    plan.destroy()


def test_cook_membrane_q1_multigrid_pcg_matches_jacobi_bicgstab():
    """BASELINE config 2 family (2-D plane-strain neo-Hooke, a domain and a Neumann line set, adaptive load stepping) with
    the multigrid-preconditioned CG: same load-step history and tip displacement as the Jacobi-BiCGSTAB of the config, in
    a fraction of the Krylov iterations (nf = 2 hierarchy in 2-D; the surface set lives on the finest level only)."""
    from autopdex_b200 import mesher, models, seeder, solver, spaces
    n = 16
    pts = [[0., 0.], [48., 44.], [48., 60.], [0., 44.]]
    coords, elems = mesher.structured_mesh((n, n), pts, "quad")
    line = mesher.boundary_faces((n, n), 0, 1)
    lam, mu = 100.0, 40.0
    Em, nu = mu * (3 * lam + 2 * mu) / (lam + mu), lam / (2 * (lam + mu))
    weak = models.hyperelastic_steady_state_weak(models.neo_hooke, lambda x, s: s["youngs modulus"], lambda x, s: s["poisson ratio"],
                                                 "plain strain")
    el = models.isoparametric_domain_element_galerkin(weak, spaces.fem_iso_line_quad_brick, *seeder.gauss_legendre_nd(2, 2))
    tr = models.neumann_weak(lambda x, s: np.asarray([0.0, s["load multiplier"]]))
    sf = models.isoparametric_surface_element_galerkin(tr, spaces.fem_iso_line_quad_brick, *seeder.gauss_legendre_nd(1, 2),
                                                       tangent_contributions=False)
    mask = np.repeat((np.abs(coords[:, 0]) < 1e-9)[:, None], 2, axis=1)
    q0 = 4.0

    def mult(s, m):
        s["load multiplier"] = m * q0
        return s
    out = {}
    for pc in ("jacobi", "multigrid"):
        st = {"assembling mode": ("user element", "user element"), "solution structure": ("nodal imposition",) * 2, "model": (el, sf),
              "solver type": "newton", "solver backend": "b200", "solver": "bicgstab" if pc == "jacobi" else "cg",
              "type of preconditioner": pc, "verbose": -1}
        settings = {"connectivity": (elems, line), "node coordinates": coords, "dirichlet dofs": mask,
                    "dirichlet conditions": np.zeros(mask.shape), "youngs modulus": Em, "poisson ratio": nu, "load multiplier": q0}
        if pc == "multigrid":
            settings["b200 multigrid"] = {"n_elements": (n, n)}
        res = solver.adaptive_load_stepping(np.zeros(mask.shape), settings, st, mult, False, None, newton_tol=1e-8, tol=1e-10)
        out[pc] = (np.asarray(res[0]), float(res[1]), int(solver.last_stats["krylov_iters"]))
        solver.clear_plan_cache()
    (uj, mj, kj), (um, mm, km) = out["jacobi"], out["multigrid"]
    assert mj == mm == 1.0
    assert np.linalg.norm(um - uj) <= 1e-8 * np.linalg.norm(uj)
    assert km * 3 <= kj, (km, kj)


def test_caller_owned_stream_and_async_entry_points():
    """compute calls on the caller's stream, asynchronous on it (apdx_plan_set_stream, apdx_assemble_async,
    apdx_spmv_async): same residual / product as the synchronising entry points, also for a Newton solve on that stream
    and after switching back to the plan's own stream."""
    from autopdex_b200 import backend
    from tests import gpu_util
    p = problems.neo_hooke_brick(4)
    plan = gpu_util.make_plan(p)
    n = p["mask"].size
    rng = np.random.default_rng(3)
    dofs = 1e-2 * rng.standard_normal(n)
    d, r0, r1 = backend.DeviceArray.from_host(dofs), backend.DeviceArray(n), backend.DeviceArray(n)
    plan.assemble(d, True, r0)                                   # reference: the synchronising entry points
    x = rng.standard_normal(plan.n_free)
    xd, y0, y1 = backend.DeviceArray.from_host(x), backend.DeviceArray(plan.n_free), backend.DeviceArray(plan.n_free)
    plan.spmv(xd, y0)
    vals0 = plan.values(False).copy()
    st = backend.Stream()
    plan.set_stream(st)
    plan.assemble_async(d, True, r1)
    plan.spmv_async(xd, y1)
    st.synchronize()
    assert np.array_equal(r1.download(), r0.download()) and np.array_equal(y1.download(), y0.download())
    assert np.array_equal(plan.values(False), vals0)
    # a whole Newton solve on the caller's stream, then back on the plan's own
    sols = []
    for stream in (st, None):
        plan.set_stream(stream)
        dd = backend.DeviceArray.from_host(np.zeros(n))
        vv = backend.DeviceArray.from_host(p["values"])
        it, rn, div = plan.newton(backend.KrylovOptions("bicgstab", rtol=1e-12), dd, vv)
        assert not div and rn < 1e-8
        sols.append((it, dd.download().copy()))
    assert sols[0][0] == sols[1][0] and np.array_equal(sols[0][1], sols[1][1])
    plan.destroy()
    st.destroy()
