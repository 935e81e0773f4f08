"""Edge cases of the hot path: systems of a single free dof / a single element, element sets with zero elements
(the reference accepts empty connectivity arrays: every per-set loop of assembler.py simply contributes nothing), and a
Newton solve whose right-hand side is zero."""
import numpy as np
import pytest

from oracle import quadrature as oquad
from oracle import solve as osolve
from tests import problems

pytestmark = pytest.mark.gpu


def _newton_vs_oracle(p, method):
    from autopdex_b200 import backend
    from tests import gpu_util
    plan = gpu_util.make_plan(p)
    d = backend.DeviceArray.from_host(np.zeros(p["mask"].size))
    v = backend.DeviceArray.from_host(p["values"])
    it, rn, div = plan.newton(backend.KrylovOptions(method, rtol=1e-13), d, v)
    prob = osolve.Problem(p["sets"], p["coords"], p["mask"], p["values"])
    ref, (steps, _, rdiv) = osolve.damped_newton(prob, np.zeros(p["mask"].shape))
    assert (it, div) == (steps, rdiv)
    scale = max(np.linalg.norm(ref), 1e-300)
    assert np.linalg.norm(d.download() - ref.ravel()) <= 1e-8 * scale
    plan.destroy()


def test_one_free_dof_poisson():
    _newton_vs_oracle(problems.poisson_hex(2), "cg")           # 27 nodes, 26 on the boundary


def test_one_element_neo_hooke():
    _newton_vs_oracle(problems.neo_hooke_brick(1), "bicgstab")  # one hex8, one face clamped: 12 free dofs


def _with_empty_surface_set(p):
    faces = problems.boundary_faces_x1(4)[:0]
    sur = dict(kind="surface", etype="quad4", conn=faces, nf=1, gp=oquad.gauss_legendre_nd(2, 2),
               model=dict(name="neumann", traction=np.array([2.0])))
    return dict(p, sets=[p["sets"][0], sur])


def test_empty_element_set_is_a_no_op():
    from tests.test_gpu_parity import _check_assembly
    p = _with_empty_surface_set(problems.poisson_hex(4, distort=0.1))
    dofs = np.random.default_rng(2).uniform(-1, 1, p["mask"].shape)
    _check_assembly(p, dofs)            # pattern, maps, values, residual, SpMV against the oracle
    _newton_vs_oracle(p, "cg")


def test_zero_right_hand_side_converges_immediately():
    """Source 0, homogeneous Dirichlet values: R = 0, the Krylov solve must return x = 0 without dividing by b.b = 0."""
    from autopdex_b200 import backend
    from tests import gpu_util
    p = problems.poisson_hex(5, source=0.0)
    plan = gpu_util.make_plan(p)
    d = backend.DeviceArray.from_host(np.zeros(p["mask"].size))
    v = backend.DeviceArray.from_host(p["values"])
    it, rn, div = plan.newton(backend.KrylovOptions("cg", rtol=1e-10), d, v)
    assert not div and rn == 0.0 and np.all(d.download() == 0.0)
    prob = osolve.Problem(p["sets"], p["coords"], p["mask"], p["values"])
    _, (steps, _, _) = osolve.damped_newton(prob, np.zeros(p["mask"].shape))
    assert it == steps
    plan.destroy()


def _pcg_iterates(A, b, minv, k):
    """k iterations of Jacobi-preconditioned CG from x0 = 0 (jax.scipy.sparse.linalg.cg's recurrence, solver.py:1093-1126)."""
    x = np.zeros_like(b)
    r = b.copy()
    z = minv * r
    p = z.copy()
    rz = r @ z
    for _ in range(k):
        q = A @ p
        alpha = rz / (p @ q)
        x += alpha * p
        r -= alpha * q
        z = minv * r
        rz_new = r @ z
        p = z + (rz_new / rz) * p
        rz = rz_new
    return x


@pytest.mark.parametrize("k", [1, 5, 31, 32, 33, 40])
def test_cg_stopped_by_maxiter_returns_the_kth_iterate(k):
    """The x-update of an iteration is applied by the kernel that follows the convergence test (k_cg_p); a solve that is
    cut off by maxiter -- inside a replayed 32-iteration chunk or exactly at its end -- must still return iterate k."""
    from autopdex_b200 import backend
    from oracle import assemble as oasm
    from tests import gpu_util
    p = problems.poisson_hex(9, distort=0.15)
    plan = gpu_util.make_plan(p)
    n = p["mask"].size
    free = ~p["mask"].ravel()
    dofs = np.random.default_rng(4).uniform(-1, 1, p["mask"].shape)
    d, r = backend.DeviceArray.from_host(dofs.ravel()), backend.DeviceArray(n)
    plan.assemble(d, True, r)
    _, data = oasm.assemble(p["sets"], p["coords"], dofs, {})
    rows, cols = oasm.coo_indices(p["sets"])
    A = oasm.scipy_assembling(data, rows, cols, n, free)
    b = np.random.default_rng(8).standard_normal(plan.n_free)
    bd, xd = backend.DeviceArray.from_host(b), backend.DeviceArray(plan.n_free)
    xd.zero()
    it, _ = plan.krylov(backend.KrylovOptions("cg", rtol=1e-30, maxiter=k), bd, xd)
    assert it == k
    ref = _pcg_iterates(A, b, 1.0 / A.diagonal(), k)
    assert np.linalg.norm(xd.download() - ref) < 1e-9 * np.linalg.norm(ref)
    plan.destroy()
