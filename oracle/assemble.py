"""Oracle global assembly, index maps and the SciPy half of the reference path
(test infrastructure).

COO emission order follows assembler._get_indices (assembler.py:123-141) and the
per-set concatenation of assemble_tangent (assembler.py:715-752); duplicate summing and
Dirichlet row/column deletion are the reference's own SciPy calls
(solver.scipy_assembling, solver.py:1207-1217).
"""
import numpy as np
import scipy.sparse as sp

from . import elements


def global_dofs(conn, nf):
    """(n_rows, nen*nf) global dof ids, node-major / component-minor (assembler.py:130-131)."""
    conn = np.asarray(conn, dtype=np.int64)
    return (conn[:, :, None] * nf + np.arange(nf)).reshape(conn.shape[0], conn.shape[1] * nf)   # also for 0 elements


def coo_indices(sets):
    """rows, cols (int64) of every element-local pair in reference order:
    set-major, element-major, local row, local col (assembler.py:134-140,749-752)."""
    rows, cols = [], []
    for st in sets:
        gd = global_dofs(st["conn"], st["nf"])
        nd = gd.shape[1]
        rows.append(np.repeat(gd, nd, axis=1).ravel())
        cols.append(np.tile(gd, (1, nd)).ravel())
    return np.concatenate(rows), np.concatenate(cols)


def _set_contributions_chunked(st, coords, dofs, settings, want_tangent, threads, chunk=16384):
    """set_contributions over row chunks on a thread pool (NumPy's einsum / matmul loops release the GIL): the same
    arithmetic per element, so the results are bit-identical to the unchunked call."""
    n = np.asarray(st["conn"]).shape[0]
    per_row = any(isinstance(v, np.ndarray) and v.ndim >= 2 and v.shape[0] == n for v in st["model"].values())
    if threads <= 1 or st["kind"] == "intpoint" or n <= chunk or per_row:
        return elements.set_contributions(st, coords, dofs, settings, want_tangent)
    from concurrent.futures import ThreadPoolExecutor
    parts = [slice(a, min(a + chunk, n)) for a in range(0, n, chunk)]
    with ThreadPoolExecutor(threads) as ex:
        out = list(ex.map(lambda sl: elements.set_contributions(st, coords, dofs, settings, want_tangent, sl), parts))
    Re = np.concatenate([o[0] for o in out])
    Ke = None if out[0][1] is None else np.concatenate([o[1] for o in out])
    return Re, Ke


def assemble(sets, coords, dofs, settings, want_tangent=True, threads=1):
    """Global residual (n_dofs,) and COO tangent data in reference order (or None).
    threads > 1: element chunks on a thread pool (bench.py's CPU baseline; parameters must not be per-element arrays)."""
    dofs = np.asarray(dofs, dtype=np.float64)
    nf = dofs.shape[1]
    R = np.zeros(dofs.size)
    data = []
    for st in sets:
        Re, Ke = _set_contributions_chunked(st, np.asarray(coords, float), dofs, settings, want_tangent, threads)
        # assembler.py:431 (segment sum of the element vectors); bincount adds in the same ascending order as add.at
        R += np.bincount(global_dofs(st["conn"], nf).ravel(), weights=Re.ravel(), minlength=R.size)
        if want_tangent:
            data.append(Ke.ravel())                                      # assembler.py:1383
    return R, (np.concatenate(data) if want_tangent else None)


def scipy_assembling(data, rows, cols, n, free=None):
    """coo -> csr (sums duplicates, sorted columns, explicit zeros kept), then
    [:, free][free]  (solver.py:1207-1217)."""
    csr = sp.csr_matrix(sp.coo_matrix((data, (rows, cols)), shape=(n, n)))
    if free is not None:
        csr = csr[:, free]
        csr = csr[free]
    csr.sort_indices()
    return csr


def pattern(sets, n, free=None):
    """Pattern-only CSR (all-ones data, duplicates NOT meaningful) plus the element-local
    -> CSR position map pos[k] of SURVEY.md Appendix A.2."""
    rows, cols = coo_indices(sets)
    csr = scipy_assembling(np.ones(rows.shape[0]), rows, cols, n)
    indptr, indices = csr.indptr.astype(np.int64), csr.indices.astype(np.int64)
    # position of (row, col) inside its (sorted) row
    key = rows * n + cols
    csr_key = np.repeat(np.arange(n, dtype=np.int64), np.diff(indptr)) * n + indices
    pos = np.searchsorted(csr_key, key)
    assert np.array_equal(csr_key[pos], key)
    out = {"indptr": indptr, "indices": indices, "pos": pos}
    if free is not None:
        red = scipy_assembling(np.ones(rows.shape[0]), rows, cols, n, free)
        out["red_indptr"] = red.indptr.astype(np.int64)
        out["red_indices"] = red.indices.astype(np.int64)
    return out
