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


def slab_coarse_planes(planes):
    """Planes of the next-coarser level held by a process that holds the fine planes (g0, g1, p0, p1) = local range and
    owned range, global indices along the slowest direction: coarse plane I coincides with fine plane 2I and belongs to
    the rank that owns that fine plane: returns the owned coarse range [P0, P1)."""
    g0, g1, p0, p1 = planes
    P0, P1 = (p0 + 1) // 2, (p1 + 1) // 2
    return P0, P1


def slab_levels(n_elements, planes, levels=None, min_owned=2):
    """[(global element counts, (g0, g1, p0, p1))] per level of a slab, finest first.  A level is added while the global
    mesh can be halved and this process would own at least `min_owned` planes of it (every rank needs free dofs of its
    own on every level; the first and last plane of the mesh usually carry boundary conditions)."""
    out = [(tuple(int(n) for n in n_elements), tuple(int(v) for v in planes))]
    while coarsenable(out[-1][0]) and (levels is None or len(out) < levels):
        shape, pl = out[-1]
        shape_c = tuple(n // 2 for n in shape)
        P0, P1 = slab_coarse_planes(pl)
        if P1 - P0 < min_owned:
            break
        G0, G1 = max(P0 - 1, 0), min(P1 + 1, shape_c[0] + 1)
        out.append((shape_c, (G0, G1, P0, P1)))
    return out


def fine_node_ids_slab(shape_fine, planes_f, planes_c):
    """Local fine node id behind every local coarse node of a slab (coarse plane I = fine plane 2I).  A coarse GHOST
    plane may coincide with a fine plane two planes beyond the owned range, which this process does not hold: it is
    pointed at the nearest local plane and the caller replaces its values by the neighbour's (exchange_ghost_planes)."""
    g0, g1 = planes_f[0], planes_f[1]
    G0, G1 = planes_c[0], planes_c[1]
    fi = np.clip(2 * np.arange(G0, G1, dtype=np.int64) - g0, 0, g1 - g0 - 1)
    grids = np.meshgrid(fi, *[np.arange(0, n + 1, 2) for n in shape_fine[1:]], indexing="ij")
    dims = [g1 - g0] + [n + 1 for n in shape_fine[1:]]
    strides = np.cumprod([1] + dims[::-1])[::-1][1:]
    return sum(g.ravel().astype(np.int64) * int(st) for g, st in zip(grids, strides))


def exchange_ghost_planes(arr, per_plane, planes, rank_lo, rank_hi):
    """Node array (n_local_nodes, ...) of a slab level: the ghost planes are replaced by the neighbours' adjacent owned
    planes (collective between neighbours; backend.comm_exchange_planes)."""
    from . import backend
    a = np.asarray(arr)
    G0, G1, P0, P1 = planes
    item = int(np.prod(a.shape[1:])) if a.ndim > 1 else 1
    out = backend.comm_exchange_planes(a.astype(np.float64), (P0 - G0) * per_plane * item, (G1 - P1) * per_plane * item,
                                       rank_lo, rank_hi)
    return out.astype(a.dtype) if a.dtype != np.float64 else out


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


def prolongation(shape_fine, nf, free_fine, free_coarse, slab=None):
    """(Multi-)linear interpolation from the coarse to the fine level, reduced to the free dofs: CSR arrays
    (indptr int32, indices int32, data) of P [n_free_fine x n_free_coarse] and of R = P^T.
    free_fine / free_coarse: bool (n_nodes, nf), True = free dof.  Built row by row without a sort: the (up to) 2^dim
    coarse neighbours of a fine node are visited in ascending column order.
    slab = ((g0, g1), (G0, G1)): this process holds the node planes [g0, g1) of the fine and [G0, G1) of the coarse level
    (global plane indices along the slowest direction; shape_fine is the GLOBAL element count): rows and columns are the
    local nodes, a coarse neighbour outside the local planes is dropped (it never is for an owned fine plane)."""
    dim = len(shape_fine)
    shape_coarse = tuple(n // 2 for n in shape_fine)
    dims_f = [n + 1 for n in shape_fine]
    dims_c = [n + 1 for n in shape_coarse]
    off_f = off_c = 0
    if slab is not None:
        (off_f, g1), (off_c, G1) = slab
        dims_f[0], dims_c[0] = g1 - off_f, G1 - off_c
    nn_f, nn_c = int(np.prod(dims_f)), int(np.prod(dims_c))
    free_fine = np.asarray(free_fine, dtype=bool).reshape(nn_f, nf)
    free_coarse = np.asarray(free_coarse, dtype=bool).reshape(nn_c, nf)
    red_c = np.cumsum(free_coarse.ravel()) - 1                      # coarse full dof -> reduced id
    idx = list(np.unravel_index(np.arange(nn_f, dtype=np.int64), dims_f))
    idx[0] = idx[0] + off_f                                         # global plane index
    cstr = np.cumprod([1] + dims_c[::-1])[::-1][1:]
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
            if d == 0 and slab is not None:
                c = c - off_c
                inside = (c >= 0) & (c < dims_c[0])
                wd = np.where(inside, wd, 0.0)
                c = np.where(inside, c, 0)
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
    P = sp.csr_matrix((data, indices, indptr.astype(np.int64)), shape=(n_free_f, n_free_c))
    R = P.T.tocsr()
    R.sort_indices()
    return ((indptr.astype(np.int32), indices, data),
            (R.indptr.astype(np.int32), R.indices.astype(np.int32), np.ascontiguousarray(R.data)))


def coarse_level_settings(settings, shape_fine, set_kinds, unwrap, wrap, cache=None, slab=None):
    """`settings` of the next-coarser level: injected node coordinates / Dirichlet flags, structured connectivity for the
    domain sets, every other entry passed through (coefficient callables read `settings`).  set_kinds: ('domain' |
    'surface', index into settings['connectivity']) per device set of the fine level; returns (coarse settings, kept set indices, fine node ids).
    cache: dict kept by the caller; the injected arrays are reused while the fine coordinate / mask objects are the same
    (no copies per solver call, and the backend can page-lock buffers it sees twice).
    slab (multi-GPU): dict(planes_f=(g0, g1, p0, p1), planes_c=(G0, G1, P0, P1), rank_lo, rank_hi) -- shape_fine is the GLOBAL
    element count, the arrays are the local slab's; injected node arrays get their ghost planes from the neighbours
    (collective: every rank of the communicator builds / updates its hierarchy at the same time)."""
    cache = cache if cache is not None else {}
    shape_c = tuple(n // 2 for n in shape_fine)
    if slab is None:
        shape_f_loc, shape_c_loc = tuple(shape_fine), shape_c
        exchange = lambda a: a
    else:
        (g0, g1), (G0, G1) = slab["planes_f"][:2], slab["planes_c"][:2]
        shape_f_loc, shape_c_loc = (g1 - g0 - 1,) + tuple(shape_fine[1:]), (G1 - G0 - 1,) + shape_c[1:]
        per_plane_c = int(np.prod([n + 1 for n in shape_c[1:]]))
        exchange = lambda a: exchange_ghost_planes(a, per_plane_c, slab["planes_c"], slab["rank_lo"], slab["rank_hi"])
    if "fine_nodes" not in cache:
        cache["fine_nodes"] = (fine_node_ids(shape_fine) if slab is None
                               else fine_node_ids_slab(shape_fine, slab["planes_f"], slab["planes_c"]))
        cache["conn"] = structured_connectivity(shape_c_loc)
    fine_nodes, conn_c = cache["fine_nodes"], cache["conn"]
    n_el = int(np.prod(shape_f_loc))
    kept, conns = [], []
    for i, (kind, dom) in enumerate(set_kinds):
        if kind != "domain":
            continue
        c = unwrap(settings["connectivity"][dom])
        if np.shape(c) != (n_el, 1 << len(shape_fine)):
            raise ValueError("b200 multigrid: domain %d has connectivity %s, expected one %d-node element per cell of the "
                             "structured %s mesh" % (i, np.shape(c), 1 << len(shape_fine), "x".join(map(str, shape_f_loc))))
        kept.append(i)
        conns.append(wrap(conn_c))
    out = dict(settings)
    out["connectivity"] = tuple(conns)
    cobj = settings["node coordinates"]
    if cache.get("coords_key") != id(cobj):
        coords = np.asarray(unwrap(cobj), dtype=np.float64)
        cache["coords_key"], cache["coords_ref"] = id(cobj), cobj
        cache["coords"] = np.ascontiguousarray(exchange(coords[fine_nodes]))
        cache["n_nodes"] = coords.shape[0]
    out["node coordinates"] = wrap(cache["coords"])
    n_nodes = cache["n_nodes"]
    if "dirichlet dofs" in settings:
        dobj = settings["dirichlet dofs"]
        if cache.get("dd_key") != id(dobj):
            dd_f = np.asarray(unwrap(dobj))
            dd = dd_f.reshape(n_nodes, -1)[fine_nodes]
            if slab is not None:
                dd = exchange(dd.astype(np.float64)) > 0.5
            dd = np.ascontiguousarray(dd if dd_f.ndim > 1 else dd.ravel())
            cache["dd_key"], cache["dd_ref"], cache["dd"], cache["dv"] = id(dobj), dobj, dd, np.zeros(dd.shape)
        out["dirichlet dofs"] = wrap(cache["dd"])
        out["dirichlet conditions"] = wrap(cache["dv"])             # coarse levels solve for corrections
    if "dofs n" in settings:
        dn = np.asarray(settings["dofs n"], dtype=np.float64)
        out["dofs n"] = exchange(dn.reshape(n_nodes, -1)[fine_nodes]).reshape((-1,) + dn.shape[1:])
    out.pop("b200 multigrid", None)
    out.pop("b200 partition", None)
    if slab is not None:
        G0, G1, P0, P1 = slab["planes_c"]
        out["b200 partition"] = dict(owned_node_begin=(P0 - G0) * per_plane_c, owned_node_end=(P1 - G0) * per_plane_c,
                                     rank_lo=slab["rank_lo"], rank_hi=slab["rank_hi"], planes=(G0, G1, P0, P1))
    return out, kept, fine_nodes
