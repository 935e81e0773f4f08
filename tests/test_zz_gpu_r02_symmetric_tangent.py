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
