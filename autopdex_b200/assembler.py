"""assembler.assemble_residual / assemble_tangent / assemble_tangent_diagonal of the reference interface on the b200
backend.

assemble_residual (assembler.py:587-637) returns a dofs-shaped array; assemble_tangent
(assembler.py:682-777) returns the tangent with duplicates ALREADY SUMMED, as a CSR triple in
SciPy's canonical form (what solver.scipy_assembling, solver.py:1207-1211, makes of the
reference's BCOO), because the duplicate summation happens on the device.
"""
from collections import namedtuple

import numpy as np

from . import backend
from . import solver as _solver

CSR = namedtuple("CSR", ["data", "indices", "indptr", "shape"])


def _prepare(dofs, settings, static_settings):
    cfg = _solver._Config(static_settings)
    st = _solver._state_for(cfg, dofs, settings)
    st.update_fields(settings)
    d0 = np.ascontiguousarray(st._unwrap(dofs), dtype=np.float64)
    st.dofs_d.upload(d0.ravel())
    return st, d0


def assemble_residual(dofs, settings, static_settings):
    st, d0 = _prepare(dofs, settings, static_settings)
    st.plan.assemble(st.dofs_d, False, st.out_d)
    r = st.out_d.download().reshape(d0.shape)
    return {st.dict_key: r} if st.dict_key is not None else r


def assemble_tangent(dofs, settings, static_settings, reduced=False):
    st, _ = _prepare(dofs, settings, static_settings)
    st.plan.assemble(st.dofs_d, True, None)
    indptr, indices = st.plan.csr(reduced)
    n = st.plan.n_free if reduced else st.plan.n_dofs
    return CSR(st.plan.values(reduced), indices, indptr, (n, n))


def csr_diagonal(csr):
    """Diagonal of a CSR triple in canonical form (absent diagonal entries are zeros)."""
    n = csr.shape[0]
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(csr.indptr))
    on_diag = csr.indices == rows
    diag = np.zeros(n, dtype=np.float64)
    diag[rows[on_diag]] = csr.data[on_diag]
    return diag


def assemble_tangent_diagonal(dofs, settings, static_settings):
    """assembler.assemble_tangent_diagonal (assembler.py:639-680): the flat (dict_flatten order) diagonal of the
    tangent WITHOUT Dirichlet reduction, i.e. the sum over the sets of the element-matrix diagonals scattered to
    their dofs (_get_tangent_diagonal, assembler.py:219-329).  Read off the device-assembled, duplicate-summed CSR
    matrix: index bookkeeping on the host, no arithmetic (the Jacobi preconditioner of the Krylov solve takes its
    diagonal from the sliced-ELL matrix on the device and never comes through here)."""
    return csr_diagonal(assemble_tangent(dofs, settings, static_settings))


__all__ = ["assemble_residual", "assemble_tangent", "assemble_tangent_diagonal", "csr_diagonal", "CSR", "backend"]
