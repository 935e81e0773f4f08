"""GPU parity: the CUDA path (through the C ABI) against the oracle on the same inputs.

Bars (BASELINE.json north_star): sparsity pattern and index maps BIT-EXACT, assembled values
within 1e-12 relative, Newton iteration counts equal, solutions within 1e-8 relative L2.
"""
import numpy as np
import pytest

from oracle import assemble as oasm
from oracle import solve as osolve
from tests import problems

pytestmark = pytest.mark.gpu

VAL_RTOL = 1e-12


def _rel(a, b):
    scale = max(np.abs(b).max(), 1e-300)
    return np.abs(a - b).max() / scale


def _check_assembly(p, dofs, settings=None):
    from autopdex_b200 import backend
    from tests import gpu_util
    plan = gpu_util.make_plan(p, settings)
    n = p["mask"].size
    free = ~p["mask"].ravel()
    # --- pattern and maps: bit-exact ---
    pat = oasm.pattern(p["sets"], n, free)
    indptr, indices = plan.csr(False)
    assert np.array_equal(indptr, pat["indptr"]) and np.array_equal(indices, pat["indices"])
    rptr, rind = plan.csr(True)
    assert np.array_equal(rptr, pat["red_indptr"]) and np.array_equal(rind, pat["red_indices"])
    assert np.array_equal(plan.elem_map(), pat["pos"])
    # --- values and residual ---
    d = backend.DeviceArray.from_host(dofs)
    r = backend.DeviceArray(n)
    plan.assemble(d, True, r)
    R, data = oasm.assemble(p["sets"], p["coords"], dofs, settings or {})
    rows, cols = oasm.coo_indices(p["sets"])
    full = oasm.scipy_assembling(data, rows, cols, n)
    red = oasm.scipy_assembling(data, rows, cols, n, free)
    assert _rel(plan.values(False), full.data) < VAL_RTOL
    assert _rel(plan.values(True), red.data) < VAL_RTOL
    assert _rel(r.download(), R) < VAL_RTOL
    # residual-only pass gives the same residual
    r2 = backend.DeviceArray(n)
    plan.assemble(d, False, r2)
    assert np.array_equal(r.download(), r2.download())
    # --- SpMV on the reduced system ---
    x = np.random.default_rng(1).standard_normal(plan.n_free)
    xd, yd = backend.DeviceArray.from_host(x), backend.DeviceArray(plan.n_free)
    plan.spmv(xd, yd)
    assert _rel(yd.download(), red @ x) < 1e-12
    plan.destroy()


def test_readme_quad4_pattern_values():
    for n in (5, 23):
        p = problems.readme_poisson(n)
        dofs = np.random.default_rng(0).uniform(-1, 1, p["mask"].shape)
        _check_assembly(p, dofs)


def test_cook_quad9_line3_pattern_values():
    p = problems.cook_g2()
    dofs = np.random.default_rng(0).uniform(-0.05, 0.05, p["mask"].shape)
    _check_assembly(p, dofs)


def test_g1_newton_golden():
    from autopdex_b200 import backend
    from tests import gpu_util
    p = problems.readme_poisson(5)
    plan = gpu_util.make_plan(p)
    d = backend.DeviceArray.from_host(np.zeros(p["mask"].size))
    v = backend.DeviceArray.from_host(p["values"])
    it, rn, div = plan.newton(backend.KrylovOptions("cg", rtol=1e-12), d, v)
    assert (it, div) == (1, False) and rn < 1e-8
    assert np.isclose(d.download().sum(), 1.9066412530282952, rtol=1e-10, atol=0)


def test_g2_cook_load_stepping_golden():
    """Adaptive load stepping (solver.py:296-379) driven from the host around apdx_newton."""
    from autopdex_b200 import backend
    from tests import gpu_util
    p = problems.cook_g2()
    plan = gpu_util.make_plan(p)
    v = backend.DeviceArray.from_host(p["values"])
    opts = backend.KrylovOptions("bicgstab", rtol=1e-13)
    dofs = np.zeros(p["mask"].size)
    m, inc, trace = 0.0, 0.2, []
    while m < 1.0 and inc > 0.01:
        m += inc
        plan.set_param(1, "traction", np.array([0.0, m * p["q0"]]))
        d = backend.DeviceArray.from_host(dofs)
        it, rn, div = plan.newton(opts, d, v, newton_tol=1e-8)
        trace.append(it)
        if div:
            m -= inc
            inc *= 0.5
        else:
            inc *= 1 + 0.5 * (7 - it) / 7
            dofs = d.download()
        inc = min(inc, 1.0)
        if m + inc > 1.0:
            inc = 1.0 - m
    assert trace == [6, 6, 5, 5, 4]
    assert np.isclose(dofs @ dofs, 19390.35027108, rtol=1e-8, atol=0)


def test_newton_matches_oracle_readme_40():
    from autopdex_b200 import backend
    from tests import gpu_util
    p = problems.readme_poisson(40)
    prob = osolve.Problem(p["sets"], p["coords"], p["mask"], p["values"])
    ref, (steps, _, div) = osolve.damped_newton(prob, np.zeros(p["mask"].shape))
    plan = gpu_util.make_plan(p)
    d = backend.DeviceArray.from_host(np.zeros(p["mask"].size))
    v = backend.DeviceArray.from_host(p["values"])
    it, rn, dv = plan.newton(backend.KrylovOptions("cg", rtol=1e-12), d, v)
    assert (it, dv) == (steps, div)
    sol = d.download()
    assert np.linalg.norm(sol - ref.ravel()) / np.linalg.norm(ref) < 1e-8


def test_hex8_poisson_pattern_values():
    p = problems.poisson_hex(6, distort=0.2)
    dofs = np.random.default_rng(0).uniform(-1, 1, p["mask"].shape)
    _check_assembly(p, dofs)


def test_hex27_poisson_pattern_values():
    p = problems.poisson_hex(3, etype="hex27")
    dofs = np.random.default_rng(0).uniform(-1, 1, p["mask"].shape)
    _check_assembly(p, dofs)


def test_hex8_neo_hooke_and_neumann_pattern_values():
    p = problems.neo_hooke_brick(4)
    dofs = np.random.default_rng(0).uniform(-0.02, 0.02, p["mask"].shape)
    _check_assembly(p, dofs)


@pytest.mark.parametrize("mode,etype", [("plain strain", "quad4"), ("plain stress", "quad9")])
def test_linear_elasticity_pattern_values(mode, etype):
    p = problems.elasticity_quad(5, mode, etype)
    dofs = np.random.default_rng(0).uniform(-0.1, 0.1, p["mask"].shape)
    _check_assembly(p, dofs)


def test_newton_neo_hooke_brick_matches_oracle():
    from autopdex_b200 import backend
    from tests import gpu_util
    p = problems.neo_hooke_brick(5)
    prob = osolve.Problem(p["sets"], p["coords"], p["mask"], p["values"])
    ref, (steps, _, div) = osolve.damped_newton(prob, np.zeros(p["mask"].shape))
    plan = gpu_util.make_plan(p)
    d = backend.DeviceArray.from_host(np.zeros(p["mask"].size))
    v = backend.DeviceArray.from_host(p["values"])
    it, rn, dv = plan.newton(backend.KrylovOptions("bicgstab", rtol=1e-12), d, v)
    assert (it, dv) == (steps, div)
    assert np.linalg.norm(d.download() - ref.ravel()) / np.linalg.norm(ref) < 1e-8


@pytest.mark.parametrize("dim,order", [(2, 1), (2, 2), (3, 1)])
def test_sparse_intpoint_sets_pattern_values(dim, order):
    p = problems.heat_sparse(dim, 3, order)
    dofs = np.random.default_rng(0).uniform(-1, 1, p["mask"].shape)
    _check_assembly(p, dofs, p["settings"])


def test_tet10_tri6_domain_elements_pattern_values():
    """P2 simplices as isoparametric 'user element's (spaces.fem_iso_line_tri_tet)."""
    from oracle import mesher as omesh, quadrature as oquad
    coords, elems = omesh.structured_mesh((3, 3), [[0., 0.], [2., 0.], [2.3, 1.], [0., 1.]], "tri")
    coords, elems = omesh.elevate_triangles(coords, elems)
    gp = (np.array([[1 / 6, 1 / 6], [1 / 6, 2 / 3], [2 / 3, 1 / 6]]), np.full(3, 1 / 6))
    st = dict(kind="domain", etype="tri6", conn=elems, nf=2, gp=gp,
              model=dict(name="neo_hooke", mode="plain strain", youngs_modulus=50.0, poisson_ratio=0.25))
    mask = np.repeat((np.abs(coords[:, 0]) < 1e-9)[:, None], 2, axis=1)
    p = dict(sets=[st], coords=coords, mask=mask, values=np.zeros(mask.shape), nf=2)
    _check_assembly(p, np.random.default_rng(0).uniform(-0.02, 0.02, mask.shape))
    coords, elems = omesh.structured_mesh((2, 2, 2), problems.UNIT_CUBE, "tet")
    a, b = 0.1381966011250105, 0.5854101966249685
    gp = (np.array([[a, a, a], [b, a, a], [a, b, a], [a, a, b]]), np.full(4, 1 / 24))
    st = dict(kind="domain", etype="tet4", conn=elems, nf=3, gp=gp,
              model=dict(name="linear_elasticity", mode="3d", youngs_modulus=50.0, poisson_ratio=0.25,
                         body_load=np.array([0.0, 0.0, -1.0])))
    mask = np.repeat((np.abs(coords[:, 2]) < 1e-9)[:, None], 3, axis=1)
    p = dict(sets=[st], coords=coords, mask=mask, values=np.zeros(mask.shape), nf=3)
    _check_assembly(p, np.random.default_rng(0).uniform(-0.02, 0.02, mask.shape))
