"""Host side of the geometric multigrid preconditioner of the b200 backend (`'type of preconditioner': 'multigrid'`).

The reference's preconditioners beyond Jacobi are algebraic multigrid through pyamg (autopdex/solver.py:1399-1491) and
PETSc's pc types (solver.py:1224-1333).  The b200 backend keeps the solve on the device and needs the hierarchy in a
form its element kernels can assemble: coarser MESHES of the same model.  For meshes of mesher.structured_mesh
('quad' / 'brick') this module derives them from the fine `settings`:

    settings['b200 multigrid'] = {'n_elements': (nx, ny[, nz])}            # the argument of mesher.structured_mesh
                                  [, 'levels': L, 'pre': 2, 'post': 2, 'coarsest': 12, 'ratio': 3.0, 'coarsest ratio': 40.0]

Level l+1 keeps every other node of level l in every direction (node (i, j, k) of the coarse mesh is node (2i, 2j, 2k) of
the fine one: its coordinates, Dirichlet flags and state are INJECTED), its connectivity is the structured connectivity
of the halved element counts, the prolongation is (bi/tri)linear interpolation.  Domain sets whose connectivity has one
row per mesh element are re-discretised on every level; surface sets (Neumann loads: no tangent) exist on the finest
level only.  The device side is csrc/multigrid.cu.
"""
import numpy as np


def coarsenable(n_elements):
    return all(n % 2 == 0 and n >= 4 for n in n_elements)


def level_shapes(n_elements, levels=None):
    """Element counts per level, finest first: halve while every direction stays even and >= 2 elements."""
    shapes = [tuple(int(n) for n in n_elements)]
    while coarsenable(shapes[-1]) and (levels is None or len(shapes) < levels):
        shapes.append(tuple(n // 2 for n in shapes[-1]))
    return shapes


def node_count(shape):
    return int(np.prod([n + 1 for n in shape]))


def fine_node_ids(shape_fine):
    """Fine node id of every coarse node (coarse node (i, j, k) = fine node (2i, 2j, 2k)); node numbering of
    mesher.structured_mesh: the first direction is the slowest."""
    grids = np.meshgrid(*[np.arange(0, n + 1, 2) for n in shape_fine], indexing="ij")
    strides = np.cumprod([1] + [n + 1 for n in shape_fine[::-1]])[::-1][1:]
    return sum(g.ravel().astype(np.int64) * int(s) for g, s in zip(grids, strides))


def structured_connectivity(shape):
    """Connectivity of mesher.structured_mesh(shape, ., 'quad' | 'brick') (node order of mesher.py: counter-clockwise
    bottom face, then top face)."""
    if len(shape) == 2:
        nx, ny = shape
        I, J = [a.ravel() for a in np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")]
        n00 = I * (ny + 1) + J
        return np.stack([n00, n00 + 1, n00 + (ny + 1) + 1, n00 + (ny + 1)], axis=1).astype(np.int32)
    nx, ny, nz = shape
    I, J, K = [a.ravel() for a in np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")]
    sy, sx = nz + 1, (ny + 1) * (nz + 1)
    n0 = I * sx + J * sy + K
    return np.stack([n0, n0 + sx, n0 + sx + sy, n0 + sy, n0 + 1, n0 + sx + 1, n0 + sx + sy + 1, n0 + sy + 1],
                    axis=1).astype(np.int32)


def prolongation(shape_fine, nf, free_fine, free_coarse):
    """(Multi-)linear interpolation from the coarse to the fine level, reduced to the free dofs: CSR arrays
    (indptr int32, indices int32, data) of P [n_free_fine x n_free_coarse] and of R = P^T.
    free_fine / free_coarse: bool (n_nodes, nf), True = free dof.  Built row by row without a sort: the (up to) 2^dim
    coarse neighbours of a fine node are visited in ascending column order."""
    dim = len(shape_fine)
    shape_coarse = tuple(n // 2 for n in shape_fine)
    nn_f = node_count(shape_fine)
    free_fine = np.asarray(free_fine, dtype=bool).reshape(nn_f, nf)
    free_coarse = np.asarray(free_coarse, dtype=bool).reshape(node_count(shape_coarse), nf)
    red_c = np.cumsum(free_coarse.ravel()) - 1                      # coarse full dof -> reduced id
    idx = np.unravel_index(np.arange(nn_f, dtype=np.int64), [n + 1 for n in shape_fine])
    cstr = np.cumprod([1] + [n + 1 for n in shape_coarse[::-1]])[::-1][1:]
    n_free_f = int(free_fine.sum())
    rows_free = free_fine.ravel()
    combos = [tuple((c >> (dim - 1 - d)) & 1 for d in range(dim)) for c in range(1 << dim)]   # lexicographic
    cols, wts, valid = [], [], []
    for combo in combos:
        node = np.zeros(nn_f, dtype=np.int64)
        w = np.ones(nn_f)
        for d in range(dim):
            a = idx[d]
            odd = (a & 1).astype(bool)
            if combo[d] == 0:
                c = a // 2
                wd = np.where(odd, 0.5, 1.0)
            else:
                c = (a + 1) // 2
                wd = np.where(odd, 0.5, 0.0)
            node += c * int(cstr[d])
            w *= wd
        cols.append(node)
        wts.append(w)
    # expand to dofs (node-major, component-minor) and reduce
    ptr_count = np.zeros(nn_f * nf, dtype=np.int32)
    ent = []
    for node, w in zip(cols, wts):
        cd = (node[:, None] * nf + np.arange(nf)).ravel()            # coarse full dof per fine dof
        wd = np.repeat(w, nf)
        ok = (wd != 0.0) & rows_free & free_coarse.ravel()[cd]
        ent.append((ok, red_c[cd], wd))
        ptr_count += ok
    counts = ptr_count[rows_free]
    indptr = np.zeros(n_free_f + 1, dtype=np.int64)
    np.cumsum(counts, out=indptr[1:])
    nnz = int(indptr[-1])
    if nnz >= 2 ** 31:
        raise ValueError("b200 multigrid: prolongation with %d entries exceeds the 32-bit index range" % nnz)
    indices = np.empty(nnz, dtype=np.int32)
    data = np.empty(nnz)
    row_start = np.zeros(nn_f * nf, dtype=np.int64)
    row_start[rows_free] = indptr[:-1]
    filled = np.zeros(nn_f * nf, dtype=np.int64)
    for ok, c, w in ent:
        pos = (row_start + filled)[ok]
        indices[pos] = c[ok]
        data[pos] = w[ok]
        filled += ok
    import scipy.sparse as sp
    n_free_c = int(free_coarse.sum())
    P = sp.csr_matrix((data, indices, indptr), shape=(n_free_f, n_free_c))
    R = P.T.tocsr()
    R.sort_indices()
    return ((indptr.astype(np.int32), indices, data),
            (R.indptr.astype(np.int32), R.indices.astype(np.int32), np.ascontiguousarray(R.data)))


def coarse_level_settings(settings, shape_fine, set_kinds, unwrap, wrap, cache=None):
    """`settings` of the next-coarser level: injected node coordinates / Dirichlet flags, structured connectivity for the
    domain sets, every other entry passed through (coefficient callables read `settings`).  set_kinds: ('domain' |
    'surface', index into settings['connectivity']) per device set of the fine level; returns (coarse settings, kept set indices, fine node ids).
    cache: dict kept by the caller; the injected arrays are reused while the fine coordinate / mask objects are the same
    (no copies per solver call, and the backend can page-lock buffers it sees twice)."""
    cache = cache if cache is not None else {}
    shape_c = tuple(n // 2 for n in shape_fine)
    if "fine_nodes" not in cache:
        cache["fine_nodes"] = fine_node_ids(shape_fine)
        cache["conn"] = structured_connectivity(shape_c)
    fine_nodes, conn_c = cache["fine_nodes"], cache["conn"]
    n_el = int(np.prod(shape_fine))
    kept, conns = [], []
    for i, (kind, dom) in enumerate(set_kinds):
        if kind != "domain":
            continue
        c = unwrap(settings["connectivity"][dom])
        if np.shape(c) != (n_el, 1 << len(shape_fine)):
            raise ValueError("b200 multigrid: domain %d has connectivity %s, expected one %d-node element per cell of the "
                             "structured %s mesh" % (i, np.shape(c), 1 << len(shape_fine), "x".join(map(str, shape_fine))))
        kept.append(i)
        conns.append(wrap(conn_c))
    out = dict(settings)
    out["connectivity"] = tuple(conns)
    cobj = settings["node coordinates"]
    if cache.get("coords_key") != id(cobj):
        coords = np.asarray(unwrap(cobj), dtype=np.float64)
        cache["coords_key"], cache["coords_ref"] = id(cobj), cobj
        cache["coords"] = np.ascontiguousarray(coords[fine_nodes])
        cache["n_nodes"] = coords.shape[0]
    out["node coordinates"] = wrap(cache["coords"])
    n_nodes = cache["n_nodes"]
    if "dirichlet dofs" in settings:
        dobj = settings["dirichlet dofs"]
        if cache.get("dd_key") != id(dobj):
            dd_f = np.asarray(unwrap(dobj))
            dd = dd_f.reshape(n_nodes, -1)[fine_nodes]
            dd = np.ascontiguousarray(dd if dd_f.ndim > 1 else dd.ravel())
            cache["dd_key"], cache["dd_ref"], cache["dd"], cache["dv"] = id(dobj), dobj, dd, np.zeros(dd.shape)
        out["dirichlet dofs"] = wrap(cache["dd"])
        out["dirichlet conditions"] = wrap(cache["dv"])             # coarse levels solve for corrections
    if "dofs n" in settings:
        dn = np.asarray(settings["dofs n"], dtype=np.float64)
        out["dofs n"] = dn.reshape(n_nodes, -1)[fine_nodes].reshape((-1,) + dn.shape[1:])
    out.pop("b200 multigrid", None)
    out.pop("b200 partition", None)
    return out, kept, fine_nodes
