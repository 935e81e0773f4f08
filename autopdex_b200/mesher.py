"""Structured meshes (host, NumPy) with the node/element numbering of autopdex.mesher.

structured_mesh follows autopdex/mesher.py:29-206 (node id i*(ny+1)+j, resp.
i*(ny+1)*(nz+1)+j*(nz+1)+k; 2-D quads clockwise w.r.t. the quad4 reference nodes, exactly
as the reference emits them); elevate_mesh_order follows mesher.py:208-360.  boundary_faces /
elevate_quads / slab_partition are additions needed by the synthetic BASELINE configs
(the reference has no surface-mesh generator for structured meshes).
"""
import numpy as np


def structured_mesh(n_elements, vertices, element_type, order=1):
    if order != 1:
        raise NotImplementedError("Only order==1 is implemented at this moment.")
    v = np.asarray(vertices, dtype=np.float64)
    dim = v.shape[1]
    if dim == 2:
        if element_type not in ("quad", "tri"):
            raise NotImplementedError("For 2D, element_type must be either 'quad' or 'tri'.")
        nx, ny = n_elements
        S, T = np.meshgrid(np.linspace(-1, 1, nx + 1), np.linspace(-1, 1, ny + 1), indexing="ij")
        s, t = S.reshape(-1, 1), T.reshape(-1, 1)
        coords = ((1 - s) * (1 - t) * v[0] + (1 + s) * (1 - t) * v[1] + (1 + s) * (1 + t) * v[2]
                  + (1 - s) * (1 + t) * v[3]) / 4
        I, J = [a.ravel() for a in np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")]
        n00 = I * (ny + 1) + J
        quads = np.stack([n00, n00 + 1, n00 + (ny + 1) + 1, n00 + (ny + 1)], axis=1).astype(np.int64)
        if element_type == "quad":
            return coords, quads
        return coords, np.concatenate([quads[:, [0, 1, 2]], quads[:, [0, 2, 3]]], axis=0)
    if dim == 3:
        if element_type not in ("brick", "tet"):
            raise NotImplementedError("For 3D, element_type must be either 'brick' or 'tet'.")
        nx, ny, nz = n_elements
        S, T, U = np.meshgrid(np.linspace(-1, 1, nx + 1), np.linspace(-1, 1, ny + 1), np.linspace(-1, 1, nz + 1),
                              indexing="ij")
        s, t, u = S.reshape(-1, 1), T.reshape(-1, 1), U.reshape(-1, 1)
        coords = ((1 - s) * (1 - t) * (1 - u) * v[0] + (1 + s) * (1 - t) * (1 - u) * v[1]
                  + (1 + s) * (1 + t) * (1 - u) * v[2] + (1 - s) * (1 + t) * (1 - u) * v[3]
                  + (1 - s) * (1 - t) * (1 + u) * v[4] + (1 + s) * (1 - t) * (1 + u) * v[5]
                  + (1 + s) * (1 + t) * (1 + u) * v[6] + (1 - s) * (1 + t) * (1 + u) * v[7]) / 8
        I, J, K = [a.ravel() for a in np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")]
        sy, sx = nz + 1, (ny + 1) * (nz + 1)
        n0 = I * sx + J * sy + K
        bricks = np.stack([n0, n0 + sx, n0 + sx + sy, n0 + sy, n0 + 1, n0 + sx + 1, n0 + sx + sy + 1, n0 + sy + 1],
                          axis=1).astype(np.int64)
        if element_type == "brick":
            return coords, bricks
        pick = np.array([[0, 1, 2, 6], [0, 2, 3, 6], [0, 3, 7, 6], [0, 7, 4, 6], [0, 4, 5, 6], [0, 5, 1, 6]])
        return coords, bricks[:, pick].reshape(-1, 4)
    raise ValueError("Unsupported dimension: vertices must have 2 or 3 columns.")


def _elevate(coords, elements, edge_pairs, face_quads, face_creation, face_output, interior):
    base = np.asarray(coords, dtype=np.float64)
    new = list(base)
    edges, faces, out = {}, {}, []
    for loc in np.asarray(elements):
        en = []
        for a, b in edge_pairs:
            key = (min(loc[a], loc[b]), max(loc[a], loc[b]))
            if key not in edges:
                edges[key] = len(new)
                new.append(0.5 * (base[loc[a]] + base[loc[b]]))
            en.append(edges[key])
        fn = {}
        for f in face_creation:
            ids = [loc[q] for q in face_quads[f]]
            key = tuple(sorted(ids))
            if key not in faces:
                faces[key] = len(new)
                new.append(np.mean(base[ids], axis=0))
            fn[f] = faces[key]
        row = list(loc) + en + [fn[f] for f in face_output]
        if interior:
            row.append(len(new))
            new.append(np.mean(base[loc], axis=0))
        out.append(row)
    return np.asarray(new), np.asarray(out, dtype=np.int64)


def elevate_mesh_order(coords, elements):
    """tri3 -> tri6 and hex8 -> hex27 with the reference's node creation order (mesher.py:332-360)."""
    coords, elements = np.asarray(coords), np.asarray(elements)
    dim, nen = coords.shape[1], elements.shape[1]
    if dim == 2 and nen == 3:
        return _elevate(coords, elements, [(0, 1), (1, 2), (2, 0)], {}, [], [], False)
    if dim == 3 and nen == 8:
        e = [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7)]
        fq = {"bottom": (0, 1, 2, 3), "top": (4, 5, 6, 7), "front": (0, 1, 5, 4), "right": (1, 2, 6, 5),
              "back": (2, 3, 7, 6), "left": (3, 0, 4, 7)}
        return _elevate(coords, elements, e, fq, ["bottom", "top", "front", "right", "back", "left"],
                        ["left", "right", "front", "back", "bottom", "top"], True)
    raise NotImplementedError("Mesh elevation to order 2 not implemented for this element type.")


def elevate_quads(coords, elements):
    """quad4 -> quad9 (corners, mid-sides 0-1,1-2,2-3,3-0, centre = quad9 order of spaces.py:1924)."""
    return _elevate(coords, elements, [(0, 1), (1, 2), (2, 3), (3, 0)], {}, [], [], True)


def boundary_faces(n_elements, axis, side):
    """Boundary elements of a structured quad/brick mesh on the face `axis` = const.
    side 0: first node layer, 1: last.  Returns line2 (2-D) or quad4 (3-D) connectivity."""
    n = list(n_elements)
    dim = len(n)
    strides = [int(np.prod([m + 1 for m in n[d + 1:]])) for d in range(dim)]
    fixed = (n[axis] if side else 0) * strides[axis]
    others = [d for d in range(dim) if d != axis]
    if dim == 2:
        d = others[0]
        a = np.arange(n[d])
        return np.stack([fixed + a * strides[d], fixed + (a + 1) * strides[d]], axis=1).astype(np.int64)
    d0, d1 = others
    A, B = [x.ravel() for x in np.meshgrid(np.arange(n[d0]), np.arange(n[d1]), indexing="ij")]
    nid = lambda da, db: fixed + (A + da) * strides[d0] + (B + db) * strides[d1]
    return np.stack([nid(0, 0), nid(1, 0), nid(1, 1), nid(0, 1)], axis=1).astype(np.int64)


def slab_partition(n_elements, rank, nranks):
    """Slab decomposition of a structured brick/quad mesh along the slowest node index i
    (SURVEY.md 8e).  Returns dict with the local node range [node_lo, node_hi) (owned planes plus
    one ghost plane per side), the owned node range, the local element-index slice along i and the
    neighbour ranks."""
    n = list(n_elements)
    planes = n[0] + 1
    per_plane = int(np.prod([m + 1 for m in n[1:]]))
    b = [(planes * r) // nranks for r in range(nranks + 1)]
    p0, p1 = b[rank], b[rank + 1]
    if p1 <= p0:
        raise ValueError("more ranks than node planes")
    g0, g1 = max(p0 - 1, 0), min(p1 + 1, planes)
    # elements between planes e and e+1 touch an owned plane iff e in [p0-1, p1-1]
    e0, e1 = max(p0 - 1, 0), min(p1, n[0])
    return dict(plane_lo=g0, plane_hi=g1, owned_plane_lo=p0, owned_plane_hi=p1, elem_lo=e0, elem_hi=e1,
                node_lo=g0 * per_plane, node_hi=g1 * per_plane, owned_node_lo=p0 * per_plane,
                owned_node_hi=p1 * per_plane, rank_lo=rank - 1 if rank > 0 else -1,
                rank_hi=rank + 1 if rank < nranks - 1 else -1, per_plane=per_plane)


def slab_partition_mesh(coords, connectivity, n_elements, rank, nranks):
    """Slab `rank` of an already generated structured mesh with any number of element sets (domain and surface sets,
    scalar or vector problems): the local mesh keeps the node planes [plane_lo, plane_hi) of slab_partition in global
    order (local id = global id - node_lo) and every element of every set that touches an owned node.  Returns the
    same keys as rcb_partition (nodes, n_owned, elements, element_ids, 'b200 partition')."""
    part = slab_partition(n_elements, rank, nranks)
    lo, hi = part["node_lo"], part["node_hi"]
    olo, ohi = part["owned_node_lo"], part["owned_node_hi"]
    nodes = np.arange(lo, hi, dtype=np.int64)
    elements, element_ids = [], []
    for c in connectivity:
        c = np.asarray(c)
        rows = np.flatnonzero(((c >= olo) & (c < ohi)).any(axis=1))
        loc = c[rows]
        if loc.size and (loc.min() < lo or loc.max() >= hi):
            raise ValueError("slab_partition_mesh: an element spans more than two adjacent node planes")
        elements.append(loc - lo)
        element_ids.append(rows)
    bp = dict(owned_node_begin=olo - lo, owned_node_end=ohi - lo, rank_lo=part["rank_lo"], rank_hi=part["rank_hi"],
              planes=(part["plane_lo"], part["plane_hi"], part["owned_plane_lo"], part["owned_plane_hi"]))
    return {"nodes": nodes, "n_owned": int(ohi - olo), "elements": elements, "element_ids": element_ids,
            "slab": part, "b200 partition": bp}


# ---- general partition: recursive coordinate bisection (north_star: "slab/RCB partitions") -------------
def rcb_owner(coords, nranks):
    """Recursive coordinate bisection of the NODES: owner[n] in [0, nranks).  Every cut is perpendicular to the longest
    extent of the current box and splits the node count in proportion to the ranks on either side; ties are broken by
    the node id, so that every process computes the same owners."""
    coords = np.asarray(coords, dtype=np.float64)
    owner = np.zeros(coords.shape[0], dtype=np.int32)
    stack = [(np.arange(coords.shape[0], dtype=np.int64), 0, int(nranks))]
    while stack:
        ids, r0, nr = stack.pop()
        if nr == 1 or ids.size == 0:
            owner[ids] = r0
            continue
        x = coords[ids]
        ax = int(np.argmax(x.max(axis=0) - x.min(axis=0)))
        nl = nr // 2
        k = (ids.size * nl) // nr
        order = np.lexsort((ids, x[:, ax]))            # by coordinate, then by node id
        ids = ids[order]
        stack.append((ids[:k], r0, nl))
        stack.append((ids[k:], r0 + nl, nr - nl))
    return owner


def rcb_partition(coords, connectivity, rank, nranks, owner=None):
    """Part `rank` of a node-based partition of an arbitrary mesh (default owners: rcb_owner).

    connectivity: tuple of (n_elem, nen) arrays (one per element set, global node ids).  The local mesh holds every
    element that touches an owned node ("ghost elements" are assembled redundantly, so the owned rows need no
    communication) and numbers its nodes [owned, ascending global id | ghosts by (owner, global id)].
    Returns a dict: nodes (global id per local node), n_owned, elements (local connectivity per set), element_ids
    (global element index per local element, per set), neighbours, send_nodes (per neighbour: local ids of the owned
    nodes it ghosts, in ITS ghost order), recv_node_ranges (per neighbour: local id range of the ghosts it owns), and
    the settings['b200 partition'] dict of the b200 backend under 'b200 partition'."""
    coords = np.asarray(coords)
    if owner is None:
        owner = rcb_owner(coords, nranks)
    owner = np.asarray(owner)
    conns = [np.asarray(c) for c in connectivity]
    local_rows = [np.flatnonzero((owner[c] == rank).any(axis=1)) for c in conns]
    owned = np.flatnonzero(owner == rank)
    touched = np.unique(np.concatenate([conns[i][local_rows[i]].ravel() for i in range(len(conns))] + [owned]))
    ghosts = touched[owner[touched] != rank]
    ghosts = ghosts[np.lexsort((ghosts, owner[ghosts]))]
    nodes = np.concatenate([owned, ghosts]).astype(np.int64)
    local_of = np.full(coords.shape[0], -1, dtype=np.int64)
    local_of[nodes] = np.arange(nodes.size)
    elements = [local_of[conns[i][local_rows[i]]] for i in range(len(conns))]
    neighbours = [int(q) for q in np.unique(owner[ghosts])]
    # ghosts of neighbour q are contiguous here; what q ghosts of mine: my nodes in elements that touch a q-owned node
    send_nodes, recv_ranges = [], []
    go = owner[ghosts]
    for q in neighbours:
        b = owned.size + int(np.searchsorted(go, q, side="left"))
        e = owned.size + int(np.searchsorted(go, q, side="right"))
        recv_ranges.append((b, e))
        mine = []
        for i in range(len(conns)):
            c = conns[i][local_rows[i]]
            oc = owner[c]
            rows = (oc == q).any(axis=1)
            mine.append(c[rows][oc[rows] == rank])
        send_nodes.append(local_of[np.unique(np.concatenate(mine))])
    part = dict(owned_node_begin=0, owned_node_end=int(owned.size), neighbours=neighbours, send_nodes=send_nodes,
                recv_node_ranges=recv_ranges)
    return {"nodes": nodes, "n_owned": int(owned.size), "elements": elements, "element_ids": local_rows,
            "neighbours": neighbours, "send_nodes": send_nodes, "recv_node_ranges": recv_ranges, "owner": owner,
            "b200 partition": part}
