"""Symmetric (mirrored) sliced-ELL storage of the Krylov matrix (csrc/sell.cu): lower columns of offset-mode slices are
read from the transposed position instead of being stored.  Checks against SciPy's `csr[:, free][free] @ x`
(solver.py:1207-1217) and against the full-storage layout (APDX_SELL_SYM=0) on scalar and vector problems."""
import os

import numpy as np
import pytest

from oracle import assemble as oasm
from tests import problems

pytestmark = pytest.mark.gpu


def _plan(p, sym):
    from tests import gpu_util
    old = os.environ.get("APDX_SELL_SYM")
    os.environ["APDX_SELL_SYM"] = "1" if sym else "0"   # read when the sliced-ELL copy is built (first assembly)
    try:
        from autopdex_b200 import backend
        plan = gpu_util.make_plan(p)
        n = p["mask"].size
        dofs = np.random.default_rng(3).uniform(-1e-2, 1e-2, p["mask"].shape)
        d, r = backend.DeviceArray.from_host(dofs.ravel()), backend.DeviceArray(n)
        plan.assemble(d, True, r)
    finally:
        if old is None:
            os.environ.pop("APDX_SELL_SYM")
        else:
            os.environ["APDX_SELL_SYM"] = old
    return plan, dofs


# builder, and whether the slices are expected to be in offset mode (mirroring only happens there)
CASES = {
    "poisson_hex8_13": (lambda: problems.poisson_hex(13, distort=0.15), True),
    "poisson_hex27_4": (lambda: problems.poisson_hex(4, etype="hex27"), False),
    "neo_hooke_hex8_7": (lambda: problems.neo_hooke_brick(7), True),
    "elasticity_quad4_40": (lambda: problems.elasticity_quad(40, mode="plain strain", etype="quad4"), True),
    "readme_quad4_70": (lambda: problems.readme_poisson(70), True),
}


@pytest.mark.parametrize("case", sorted(CASES))
def test_mirrored_spmv_matches_scipy_and_full_storage(case):
    from autopdex_b200 import backend
    build, expect_mirrored = CASES[case]
    p = build()
    n = p["mask"].size
    free = ~p["mask"].ravel()
    plan_s, dofs = _plan(p, True)
    plan_f, _ = _plan(p, False)
    info_s, info_f = plan_s.sell_info(), plan_f.sell_info()
    assert info_s["symmetric"] == 1 and info_f["symmetric"] == 0 and info_f["mirrored_entries"] == 0
    # structured meshes: most lower columns are mirrored, and exactly those are no longer stored
    if expect_mirrored:
        assert info_s["mirrored_entries"] > 0.25 * info_f["stored_values"], (info_s, info_f)
    assert info_s["stored_values"] + info_s["mirrored_entries"] == info_f["stored_values"]
    _, data = oasm.assemble(p["sets"], p["coords"], dofs, {})
    rows, cols = oasm.coo_indices(p["sets"])
    red = oasm.scipy_assembling(data, rows, cols, n, free)
    assert abs(red - red.T).max() < 1e-12 * abs(red).max()   # the premise: in-scope tangents are symmetric
    rng = np.random.default_rng(5)
    for _ in range(2):
        x = rng.standard_normal(plan_s.n_free)
        xd = backend.DeviceArray.from_host(x)
        ys, yf = backend.DeviceArray(plan_s.n_free), backend.DeviceArray(plan_s.n_free)
        plan_s.spmv(xd, ys)
        plan_f.spmv(xd, yf)
        ref = red @ x
        scale = np.abs(ref).max()
        assert np.abs(ys.download() - ref).max() < 1e-12 * scale
        assert np.abs(yf.download() - ref).max() < 1e-12 * scale
    plan_s.destroy()
    plan_f.destroy()


@pytest.mark.parametrize("method", ["cg", "bicgstab"])
def test_krylov_solution_same_with_mirrored_storage(method):
    from autopdex_b200 import backend
    p = problems.poisson_hex(17, distort=0.1)
    sols, iters = [], []
    for sym in (True, False):
        plan, _ = _plan(p, sym)
        b = np.random.default_rng(9).standard_normal(plan.n_free)
        bd, xd = backend.DeviceArray.from_host(b), backend.DeviceArray(plan.n_free)
        xd.zero()
        it, relres = plan.krylov(backend.KrylovOptions(method, rtol=1e-11), bd, xd)
        sols.append(xd.download())
        iters.append(it)
        plan.destroy()
    assert abs(iters[0] - iters[1]) <= 2
    assert np.linalg.norm(sols[0] - sols[1]) < 1e-8 * np.linalg.norm(sols[1])


def _permute_nodes(p, perm):
    """Renumber the nodes of a problem dict: new id of old node i is perm[i]."""
    inv = np.empty_like(perm)
    inv[perm] = np.arange(perm.size)
    q = dict(p)
    q["coords"] = p["coords"][inv]
    q["mask"] = p["mask"][inv]
    q["values"] = p["values"][inv]
    q["sets"] = [dict(st, conn=perm[st["conn"]]) for st in p["sets"]]
    return q


@pytest.mark.parametrize("fraction", [0.002, 1.0])
def test_mirrored_storage_on_irregular_numbering(fraction):
    """Node numbers of a few nodes (0.2 %) or of all nodes shuffled: slices fall back to explicit columns, lower columns whose partner rows
    live in such slices must stay stored, and the product must still match SciPy."""
    from autopdex_b200 import backend
    p = problems.poisson_hex(14, distort=0.1)
    n_nodes = p["coords"].shape[0]
    rng = np.random.default_rng(17)
    perm = np.arange(n_nodes)
    pick = rng.choice(n_nodes, size=max(6, int(fraction * n_nodes)), replace=False)
    perm[pick] = perm[rng.permutation(pick)]
    q = _permute_nodes(p, perm)
    plan_s, dofs = _plan(q, True)
    plan_f, _ = _plan(q, False)
    info_s, info_f = plan_s.sell_info(), plan_f.sell_info()
    assert info_s["stored_values"] + info_s["mirrored_entries"] == info_f["stored_values"]
    if fraction < 0.5:
        assert 0 < info_s["mirrored_entries"] < 0.5 * info_f["stored_values"]
    n = q["mask"].size
    free = ~q["mask"].ravel()
    _, data = oasm.assemble(q["sets"], q["coords"], dofs, {})
    rows, cols = oasm.coo_indices(q["sets"])
    red = oasm.scipy_assembling(data, rows, cols, n, free)
    x = rng.standard_normal(plan_s.n_free)
    xd = backend.DeviceArray.from_host(x)
    for plan in (plan_s, plan_f):
        yd = backend.DeviceArray(plan.n_free)
        plan.spmv(xd, yd)
        ref = red @ x
        assert np.abs(yd.download() - ref).max() < 1e-12 * np.abs(ref).max()
        plan.destroy()


@pytest.mark.parametrize("nf_case", ["poisson", "neo_hooke"])
def test_mirrored_storage_on_an_owned_row_range(nf_case):
    """A partitioned plan multiplies only its owned rows [f0, f1) and reads ghost columns on both sides.
    Lower columns whose partner row is not owned must stay stored; checked here on ONE GPU by restricting a plan to a
    slab of its rows (no communicator: the caller's x already holds the ghost entries)."""
    from autopdex_b200 import backend
    m = 10
    p = problems.poisson_hex(m, distort=0.1) if nf_case == "poisson" else problems.neo_hooke_brick(6)
    nf = p["nf"]
    n_side = (m if nf_case == "poisson" else 6) + 1
    plane = n_side * n_side * nf                       # dofs per node plane (i slowest, mesher.py:160)
    lo, hi = 3 * plane, (n_side - 2) * plane
    results = {}
    for sym in (True, False):
        from tests import gpu_util
        old = os.environ.get("APDX_SELL_SYM")
        os.environ["APDX_SELL_SYM"] = "1" if sym else "0"
        try:
            plan = gpu_util.make_plan(p)
            plan.set_partition(lo, hi, 0, 1)            # neighbour ranks are only labels without a communicator
            n = p["mask"].size
            dofs = np.random.default_rng(3).uniform(-1e-2, 1e-2, p["mask"].shape)
            d, r = backend.DeviceArray.from_host(dofs.ravel()), backend.DeviceArray(n)
            plan.assemble(d, True, r)
        finally:
            if old is None:
                os.environ.pop("APDX_SELL_SYM")
            else:
                os.environ["APDX_SELL_SYM"] = old
        f0, f1 = plan.f0, plan.f1
        assert 0 < f0 < f1 < plan.n_free
        x = np.random.default_rng(6).standard_normal(plan.n_free)
        xd, yd = backend.DeviceArray.from_host(x), backend.DeviceArray(plan.n_free)
        yd.zero()
        plan.spmv(xd, yd)
        results[sym] = (yd.download()[f0:f1], plan.sell_info(), (f0, f1), dofs)
        plan.destroy()
    (ys, info_s, rng_s, dofs), (yf, info_f, rng_f, _) = results[True], results[False]
    assert rng_s == rng_f and info_s["mirrored_entries"] > 0
    assert info_s["stored_values"] + info_s["mirrored_entries"] == info_f["stored_values"]
    n = p["mask"].size
    free = ~p["mask"].ravel()
    _, data = oasm.assemble(p["sets"], p["coords"], dofs, {})
    rows, cols = oasm.coo_indices(p["sets"])
    red = oasm.scipy_assembling(data, rows, cols, n, free)
    x = np.random.default_rng(6).standard_normal(red.shape[0])
    ref = (red @ x)[rng_s[0]:rng_s[1]]
    assert np.abs(ys - ref).max() < 1e-12 * np.abs(ref).max()
    assert np.abs(yf - ref).max() < 1e-12 * np.abs(ref).max()
