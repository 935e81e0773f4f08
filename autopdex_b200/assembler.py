"""assembler.assemble_residual / assemble_tangent / assemble_tangent_diagonal of the reference interface on the b200
backend.

assemble_residual (assembler.py:587-637) returns a dofs-shaped array.  assemble_tangent (assembler.py:682-777)
returns the reference's wire format: a BCOO-shaped object (`data` (nse,), `indices` (nse, 2) int64, `shape`) holding
every element-local pair in the reference's order with duplicates NOT summed -- the element streams copied off the
device (apdx_get_coo_values) and the index array of _get_indices.  The device keeps the duplicate-summed matrix
anyway (that is what the solver uses); `BCOO.sum_duplicates()` / `format="csr"` / `reduced=True` return it as a CSR
triple in SciPy's canonical form, i.e. what solver.scipy_assembling (solver.py:1207-1217) makes of the BCOO.
"""
from collections import namedtuple

import numpy as np

from . import backend
from . import solver as _solver

CSR = namedtuple("CSR", ["data", "indices", "indptr", "shape"])


class BCOO:
    """Duck type of jax.experimental.sparse.BCOO as far as the reference's callers use it (solver.py:1207-1211:
    `.data`, `.indices[:, 0]`, `.indices[:, 1]`, `.shape`; `nse`, `todense()` for small systems)."""

    def __init__(self, data, indices, shape, plan=None):
        self.data, self.indices, self.shape, self._plan = data, indices, tuple(shape), plan

    @property
    def nse(self):
        return self.data.shape[0]

    def sum_duplicates(self, reduced=False):
        """The duplicate-summed matrix as a CSR triple: read from the device (which summed it in the same assembly
        pass) when this object came from assemble_tangent, else summed here in COO order."""
        if self._plan is not None:
            indptr, indices = self._plan.csr(reduced)
            n = self._plan.n_free if reduced else self._plan.n_dofs
            return CSR(self._plan.values(reduced), indices, indptr, (n, n))
        import scipy.sparse as sp
        m = sp.csr_matrix(sp.coo_matrix((self.data, (self.indices[:, 0], self.indices[:, 1])), shape=self.shape))
        m.sort_indices()
        return CSR(m.data, m.indices.astype(np.int64), m.indptr.astype(np.int64), self.shape)

    def todense(self):
        out = np.zeros(self.shape)
        np.add.at(out, (self.indices[:, 0], self.indices[:, 1]), self.data)
        return out


def _dofs_per_node(a):
    return 1 if np.ndim(a) == 1 else np.shape(a)[-1]


def _get_indices(connectivity, dofs):
    """(nse, 2) int64 row/column indices of every element-local pair of ONE set, in the order of
    assembler._get_indices (assembler.py:41-141).
    array dofs: element-major, local row, local column; local dofs node-major / component-minor (:123-141).
    dict dofs (one connectivity per field): blocks `for field_i: for field_j:` (:79-80), each element-major with the
    rows of field_i and the columns of field_j, global ids offset by the cumulative field sizes in key order (:64-71,
    the dict_flatten numbering of utility.py:104-128)."""
    if callable(dofs):
        dofs = dofs(0.)
    if isinstance(dofs, dict):
        keys = list(dofs.keys())
        off, cur = {}, 0
        for k in keys:
            off[k] = cur
            cur += int(np.size(dofs[k]))
        gd = {}
        for k in keys:
            c = np.asarray(connectivity[k], dtype=np.int64)
            nf = _dofs_per_node(dofs[k])
            gd[k] = (off[k] + c[:, :, None] * nf + np.arange(nf, dtype=np.int64)).reshape(c.shape[0], -1)
        blocks = []
        for ki in keys:
            for kj in keys:
                gi, gj = gd[ki], gd[kj]
                rows = np.repeat(gi, gj.shape[1], axis=1)
                cols = np.tile(gj, (1, gi.shape[1]))
                blocks.append(np.stack([rows.ravel(), cols.ravel()], axis=-1))
        return np.concatenate(blocks, axis=0)
    c = np.asarray(connectivity, dtype=np.int64)
    nf = _dofs_per_node(dofs)
    gd = (c[:, :, None] * nf + np.arange(nf, dtype=np.int64)).reshape(c.shape[0], -1)
    nd = gd.shape[1]
    return np.stack([np.repeat(gd, nd, axis=1).ravel(), np.tile(gd, (1, nd)).ravel()], axis=-1)


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
    return st._wrap(r)


def assemble_tangent(dofs, settings, static_settings, reduced=False, format="bcoo"):
    """format 'bcoo' (default, the reference's return type) or 'csr' (duplicates summed on the device);
    reduced=True: the Dirichlet-reduced CSR matrix csr[:, free][free] the solver works on."""
    st, d0 = _prepare(dofs, settings, static_settings)
    st.plan.assemble(st.dofs_d, True, None)
    if reduced or format == "csr":
        indptr, indices = st.plan.csr(reduced)
        n = st.plan.n_free if reduced else st.plan.n_dofs
        return CSR(st.plan.values(reduced), indices, indptr, (n, n))
    n = st.plan.n_dofs
    if st.fields is not None:
        return _multi_field_bcoo(st, dofs, settings, n)
    # sets in order (assembler.py:715-752); the unwrapped single field has the numbering of dict_flatten
    idx = [_get_indices(np.asarray(c), d0) for c in st.conn_refs]
    indices = np.concatenate(idx, axis=0) if idx else np.zeros((0, 2), dtype=np.int64)
    return BCOO(st.plan.coo_values(), indices, (n, n), st.plan)


def _multi_field_bcoo(st, dofs, settings, n):
    """BCOO of a multi-field dict-dof problem in the reference's order (assembler.py:715-752 per domain, :79-117 inside a
    domain: blocks `for field_i: for field_j:`, each element-major).  The device holds the model's (field, field) block of
    every domain element-major, which is exactly the order inside that block; every other block is an explicit zero."""
    nf = st.nf
    offs, o = [], 0
    for c in st.conn_refs:                                     # COO offsets of the device sets
        offs.append(o)
        o += c.shape[0] * (c.shape[1] * nf) ** 2
    data, indices = [], []
    doms = sorted({d for _, _, d in st.cfg.sets})
    if len(doms) != len(st.cfg.sets):
        raise ValueError("b200 backend: BCOO export of a multi-field problem needs one model per domain")
    for i, (route, m, dom) in enumerate(st.cfg.sets):
        conn = settings["connectivity"][dom]
        indices.append(_get_indices({k: np.asarray(conn[k]) for k in st.fields}, dofs))
        n_e = np.asarray(conn[m.field]).shape[0]
        for fi in st.fields:
            for fj in st.fields:
                cnt = n_e * (np.asarray(conn[fi]).shape[1] * nf) * (np.asarray(conn[fj]).shape[1] * nf)
                if fi == m.field and fj == m.field:
                    data.append(st.plan.coo_values(offs[i], cnt))
                else:
                    data.append(np.zeros(cnt))
    return BCOO(np.concatenate(data), np.concatenate(indices, axis=0), (n, n), st.plan)


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
    return csr_diagonal(assemble_tangent(dofs, settings, static_settings, format="csr"))


__all__ = ["assemble_residual", "assemble_tangent", "assemble_tangent_diagonal", "csr_diagonal", "CSR", "BCOO",
           "_get_indices", "backend"]
