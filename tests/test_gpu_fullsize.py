"""Size-independent properties at BASELINE.json's full size (256^3 hex8 Poisson, 16 974 593 dofs) where the oracle
cannot run: the finite-element patch test (a linear field imposed on the boundary is reproduced exactly in the
interior), symmetry and linearity of the assembled reduced operator, and an independent residual check of the Krylov
solution through apdx_spmv.  Also the multi-set 3-D neo-Hooke configuration at a size beyond the oracle's reach."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CUBE = [[0., 0., 0.], [1., 0., 0.], [1., 1., 0.], [0., 1., 0.], [0., 0., 1.], [1., 0., 1.], [1., 1., 1.], [0., 1., 1.]]


def _poisson_plan(m, distort=0.0):
    from autopdex_b200 import backend, mesher, seeder
    coords, elems = mesher.structured_mesh((m, m, m), CUBE, "brick")
    tol = 1e-12
    onb = (np.abs(coords) < tol).any(axis=1) | (np.abs(coords - 1.0) < tol).any(axis=1)
    if distort:
        rng = np.random.default_rng(3)
        coords = coords + (~onb)[:, None] * rng.uniform(-distort, distort, coords.shape) / m
    st = backend.SetSpec("domain", "poisson_weak", elems.astype(np.int32), family="quad_brick",
                         gp=seeder.gauss_legendre_nd(3, 2), params={"coefficient": 1.0, "source": 0.0})
    plan = backend.Plan(3, coords.shape[0], 1, [st], onb)
    plan.set_coords(coords)
    return plan, coords, onb


def test_patch_test_256_cubed():
    """u = a + b.x on the boundary, f = 0  =>  u_h = a + b.x everywhere (exactly representable by Q1 elements, also on
    the distorted mesh), one Newton step.  Exercises pattern build, assembly, Dirichlet imposition and PCG at full size."""
    from autopdex_b200 import backend
    m = 256
    plan, coords, onb = _poisson_plan(m, distort=0.2)
    assert plan.n_dofs == 16974593 and plan.nnz == 454756609           # SURVEY.md section 8 sizes
    exact = 0.3 + coords @ np.array([1.0, -2.0, 0.5])
    vals = backend.DeviceArray.from_host(np.where(onb, exact, 0.0))
    d = backend.DeviceArray.from_host(np.zeros(plan.n_dofs))
    it, rn, div = plan.newton(backend.KrylovOptions("cg", rtol=1e-11), d, vals, newton_tol=1e-8)
    sol = d.download()
    assert it == 1 and not div and rn < 1e-8
    assert np.abs(sol - exact).max() < 1e-8 * np.abs(exact).max()
    # symmetry and linearity of the reduced operator, independent residual of a fresh solve
    rng = np.random.default_rng(0)
    n = plan.n_free
    x, y = rng.standard_normal(n), rng.standard_normal(n)
    xd, yd, ax, ay, az = (backend.DeviceArray.from_host(v) for v in (x, y, np.zeros(n), np.zeros(n), np.zeros(n)))
    plan.spmv(xd, ax)
    plan.spmv(yd, ay)
    Ax, Ay = ax.download(), ay.download()
    assert abs(x @ Ay - y @ Ax) <= 1e-12 * (np.linalg.norm(x) * np.linalg.norm(Ay))
    zd = backend.DeviceArray.from_host(2.0 * x - 3.0 * y)
    plan.spmv(zd, az)
    assert np.abs(az.download() - (2.0 * Ax - 3.0 * Ay)).max() <= 1e-12 * np.abs(Ax).max()
    b = backend.DeviceArray.from_host(Ax)                          # solve A w = A x  ->  w = x
    w = backend.DeviceArray.from_host(np.zeros(n))
    iters, relres = plan.krylov(backend.KrylovOptions("cg", rtol=1e-10), b, w)
    plan.spmv(w, az)
    assert np.linalg.norm(az.download() - Ax) <= 2e-10 * np.linalg.norm(Ax)
    assert np.linalg.norm(w.download() - x) <= 1e-6 * np.linalg.norm(x)
    plan.destroy()


def test_elasticity_patch_test_hex8_neo_hooke_small_strain_and_linear():
    """Constant-strain patch test for the vector kernels on a distorted 24^3 hex8 mesh: linear elasticity reproduces a
    linear displacement field exactly; neo-Hooke with all boundary dofs prescribed by a homogeneous deformation gives
    that homogeneous deformation in the interior (equilibrium of a constant first Piola-Kirchhoff stress)."""
    from autopdex_b200 import backend, mesher, seeder
    m = 24
    coords, elems = mesher.structured_mesh((m, m, m), CUBE, "brick")
    tol = 1e-12
    onb = (np.abs(coords) < tol).any(axis=1) | (np.abs(coords - 1.0) < tol).any(axis=1)
    rng = np.random.default_rng(5)
    coords = coords + (~onb)[:, None] * rng.uniform(-0.2, 0.2, coords.shape) / m
    H = np.array([[0.02, 0.01, 0.0], [-0.015, 0.03, 0.005], [0.0, 0.01, -0.01]])
    exact = coords @ H.T
    mask = np.repeat(onb[:, None], 3, axis=1)
    for model in ("linear_elasticity", "neo_hooke"):
        st = backend.SetSpec("domain", model, elems.astype(np.int32), family="quad_brick", gp=seeder.gauss_legendre_nd(3, 2),
                             mode="3d", params={"youngs_modulus": 100.0, "poisson_ratio": 0.3})
        plan = backend.Plan(3, coords.shape[0], 3, [st], mask)
        plan.set_coords(coords)
        vals = backend.DeviceArray.from_host(np.where(mask, exact, 0.0))
        # start from a nearby homogeneous state: from u = 0 the boundary layer of elements would be inverted by the
        # prescribed boundary displacements in the first iterate (the reference would diverge in the same way)
        d = backend.DeviceArray.from_host(0.9 * exact)
        it, rn, div = plan.newton(backend.KrylovOptions("bicgstab", rtol=1e-12, maxiter=20000), d, vals, newton_tol=1e-9)
        sol = d.download().reshape(mask.shape)
        assert not div and rn < 1e-9 and it <= (1 if model == "linear_elasticity" else 6)
        assert np.abs(sol - exact).max() < 1e-8 * np.abs(exact).max(), model
        plan.destroy()
