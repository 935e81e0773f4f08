"""Transfer operators of the structured multigrid hierarchy built on the device (apdx_plan_set_coarse_structured,
csrc/multigrid.cu: k_transfer_rows / k_inject_map) against the host statement of the same construction
(autopdex_b200/multigrid.prolongation, itself checked against kron interpolation in test_capi_and_host.py): indptr,
indices, data and the injection map must be IDENTICAL -- 2-D and 3-D, one and three dofs per node, random Dirichlet
masks, and slabs with ghost planes on either side (the local operators of a partitioned hierarchy)."""
import numpy as np
import pytest

from autopdex_b200 import backend, multigrid
from oracle import quadrature as oquad

pytestmark = pytest.mark.gpu


def _plan(dims, nf, mask, rng):
    """A plan on the structured mesh with `dims` nodes per direction: Poisson (nf = 1) or linear elasticity (nf = dim)."""
    dim = len(dims)
    shape = tuple(n - 1 for n in dims)
    conn = multigrid.structured_connectivity(shape)
    grids = np.meshgrid(*[np.linspace(0.0, 1.0, n) for n in dims], indexing="ij")
    coords = np.stack([g.ravel() for g in grids], axis=1)
    gp = oquad.gauss_legendre_nd(dim, 2)
    if nf == 1:
        spec = backend.SetSpec("domain", "poisson_weak", conn, family="quad_brick", gp=gp, params=dict(coefficient=1.0, source=1.0))
    else:
        spec = backend.SetSpec("domain", "linear_elasticity", conn, family="quad_brick", gp=gp,
                               mode="3d" if dim == 3 else "plain strain",
                               params=dict(youngs_modulus=100.0, poisson_ratio=0.3, body_load=np.zeros(dim)))
    plan = backend.Plan(dim, coords.shape[0], nf, [spec], mask)
    plan.set_coords(coords)
    return plan


CASES = [
    # (fine node counts, dofs per node, global plane offsets (fine, coarse) and coarse plane count or None, Dirichlet fraction)
    ((9, 9, 9), 1, None, 0.0),
    ((9, 9, 9), 1, None, 0.3),
    ((9, 5, 13), 3, None, 0.25),
    ((17, 9), 1, None, 0.2),
    ((9, 13), 2, None, 0.3),
    # slabs: local fine planes [g0, g1), local coarse planes [G0, G1) of a 33-plane global mesh
    ((7, 5, 9), 1, (10, 17, 5, 9), 0.2),      # g0 even: the lower coarse ghost plane 5 = fine plane 10 is local
    ((7, 5, 9), 3, (11, 18, 5, 10), 0.2),     # lower coarse ghost plane 5 = fine plane 10 is NOT local; upper one (9 -> 18) neither
    ((6, 9), 1, (0, 6, 0, 4), 0.2),           # first slab of a 2-D mesh: upper coarse ghost plane 3 = fine plane 6 not local
    ((5, 9, 5), 1, (28, 33, 14, 17), 0.1),    # last slab
]


@pytest.mark.parametrize("dims_f, nf, slab, frac", CASES)
def test_device_transfer_operators_equal_the_host_construction(dims_f, nf, slab, frac):
    rng = np.random.default_rng(11)
    dims_f = list(dims_f)
    dim = len(dims_f)
    if slab is None:
        off_f, off_c = 0, 0
        dims_c = [(n - 1) // 2 + 1 for n in dims_f]
    else:
        off_f, g1, off_c, G1 = slab
        assert g1 - off_f == dims_f[0]
        dims_c = [G1 - off_c] + [(n - 1) // 2 + 1 for n in dims_f[1:]]
    nn_f, nn_c = int(np.prod(dims_f)), int(np.prod(dims_c))
    mask_f = rng.random((nn_f, nf)) < frac
    mask_c = rng.random((nn_c, nf)) < frac
    mask_f[0, 0] = mask_c[0, 0] = frac > 0          # keep at least one Dirichlet dof when any are asked for
    fine, coarse = _plan(dims_f, nf, mask_f, rng), _plan(dims_c, nf, mask_c, rng)
    try:
        fine.set_coarse_structured(coarse, dims_f, dims_c, off_f, off_c)
        (p_ptr, p_idx, p_val), inj = fine.get_transfer(0)
        (r_ptr, r_idx, r_val), _ = fine.get_transfer(1)
    finally:
        coarse.destroy()             # the order of solver._State.destroy: coarse levels first
        fine.destroy()
    # host statement; shape_fine = GLOBAL element counts (the slowest direction is only read through the slab planes)
    shape_fine = tuple(2 * (n - 1) for n in dims_c) if slab is not None else tuple(n - 1 for n in dims_f)
    hslab = None if slab is None else ((off_f, off_f + dims_f[0]), (off_c, off_c + dims_c[0]))
    P, R = multigrid.prolongation(shape_fine, nf, ~mask_f, ~mask_c, hslab)
    for got, ref, name in ((p_ptr, P[0], "P.indptr"), (p_idx, P[1], "P.indices"), (p_val, P[2], "P.data"),
                           (r_ptr, R[0], "R.indptr"), (r_idx, R[1], "R.indices"), (r_val, R[2], "R.data")):
        assert got.shape == np.asarray(ref).shape, name
        assert np.array_equal(got, ref), name
    if slab is None:
        nodes = multigrid.fine_node_ids(shape_fine)
    else:
        nodes = multigrid.fine_node_ids_slab(shape_fine, (off_f, off_f + dims_f[0]), (off_c, off_c + dims_c[0]))
    assert np.array_equal(inj, (nodes[:, None] * nf + np.arange(nf)).ravel())


def test_structured_link_rejects_mismatched_meshes():
    rng = np.random.default_rng(0)
    fine, coarse = _plan([9, 9, 9], 1, None, rng), _plan([5, 5, 5], 1, None, rng)
    other = _plan([5, 5, 4], 1, None, rng)
    try:
        with pytest.raises(Exception, match="node counts|fine and .* coarse nodes"):
            fine.set_coarse_structured(other, [9, 9, 9], [5, 5, 4])
        with pytest.raises(Exception, match="node counts"):
            fine.set_coarse_structured(coarse, [9, 9, 5], [5, 5, 3])
        with pytest.raises(Exception, match="do not lie over"):
            fine.set_coarse_structured(coarse, [9, 9, 9], [5, 5, 5], 0, 4)
    finally:
        for pl in (coarse, other, fine):
            pl.destroy()
