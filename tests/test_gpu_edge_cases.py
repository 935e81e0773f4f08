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
