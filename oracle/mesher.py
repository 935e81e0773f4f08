"""Oracle restatement of the structured mesh generator (test infrastructure).

Follows autopdex/mesher.py:29-206 (structured_mesh), :208-244 (tri6 elevation),
:246-330 (hex27 elevation).  Node ids: i*(ny+1)+j in 2-D, i*(ny+1)*(nz+1)+j*(nz+1)+k in 3-D
(mesher.py:104-107,160-167).  2-D quads come out clockwise w.r.t. the quad4
reference nodes (det J < 0) -- reproduced on purpose.
"""
import numpy as np


def _lin(n):
    return np.linspace(-1.0, 1.0, n + 1)


def structured_mesh(n_elements, vertices, element_type):
    v = np.asarray(vertices, dtype=np.float64)
    dim = v.shape[1]
    if dim == 2:
        nx, ny = n_elements
        coords = np.empty(((nx + 1) * (ny + 1), 2))
        q = 0
        for s in _lin(nx):          # mesher.py:88-99  meshgrid(indexing='ij') -> i slowest
            for t in _lin(ny):
                coords[q] = ((1 - s) * (1 - t) * v[0] + (1 + s) * (1 - t) * v[1]
                             + (1 + s) * (1 + t) * v[2] + (1 - s) * (1 + t) * v[3]) / 4
                q += 1
        quads = np.empty((nx * ny, 4), dtype=np.int64)
        e = 0
        for i in range(nx):         # mesher.py:102-113
            for j in range(ny):
                quads[e] = (i * (ny + 1) + j, i * (ny + 1) + j + 1,
                            (i + 1) * (ny + 1) + j + 1, (i + 1) * (ny + 1) + j)
                e += 1
        if element_type == "quad":
            return coords, quads
        if element_type == "tri":   # mesher.py:120-122: two blocks, not interleaved
            return coords, np.concatenate([quads[:, [0, 1, 2]], quads[:, [0, 2, 3]]], axis=0)
        raise NotImplementedError(element_type)
    if dim == 3:
        nx, ny, nz = n_elements
        S, T, U = np.meshgrid(_lin(nx), _lin(ny), _lin(nz), indexing="ij")
        s, t, u = S.ravel()[:, None], T.ravel()[:, None], U.ravel()[:, None]
        coords = ((1 - s) * (1 - t) * (1 - u) * v[0] + (1 + s) * (1 - t) * (1 - u) * v[1]
                  + (1 + s) * (1 + t) * (1 - u) * v[2] + (1 - s) * (1 + t) * (1 - u) * v[3]
                  + (1 - s) * (1 - t) * (1 + u) * v[4] + (1 + s) * (1 - t) * (1 + u) * v[5]
                  + (1 + s) * (1 + t) * (1 + u) * v[6] + (1 - s) * (1 + t) * (1 + u) * v[7]) / 8
        I, J, K = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
        I, J, K = I.ravel(), J.ravel(), K.ravel()
        sy, sx = nz + 1, (ny + 1) * (nz + 1)
        nid = lambda di, dj, dk: (I + di) * sx + (J + dj) * sy + (K + dk)
        bricks = np.stack([nid(0, 0, 0), nid(1, 0, 0), nid(1, 1, 0), nid(0, 1, 0),
                           nid(0, 0, 1), nid(1, 0, 1), nid(1, 1, 1), nid(0, 1, 1)],
                          axis=1).astype(np.int64)      # mesher.py:159-168
        if element_type == "brick":
            return coords, bricks
        if element_type == "tet":   # mesher.py:187-201: six tets sharing the n0-n6 diagonal
            pick = [[0, 1, 2, 6], [0, 2, 3, 6], [0, 3, 7, 6], [0, 7, 4, 6], [0, 4, 5, 6], [0, 5, 1, 6]]
            return coords, bricks[:, pick].reshape(-1, 4)
        raise NotImplementedError(element_type)
    raise ValueError("vertices must have 2 or 3 columns")


def elevate_triangles(coords, elements):
    """tri3 -> tri6, mid-side nodes appended in first-seen order (mesher.py:208-244)."""
    coords = [np.asarray(c, dtype=np.float64) for c in np.asarray(coords)]
    seen, out = {}, []
    base = np.asarray(coords)

    def mid(a, b):
        key = (min(a, b), max(a, b))
        if key not in seen:
            seen[key] = len(coords)
            coords.append(0.5 * (base[a] + base[b]))
        return seen[key]

    for n0, n1, n2 in np.asarray(elements):
        out.append([n0, n1, n2, mid(n0, n1), mid(n1, n2), mid(n2, n0)])
    return np.asarray(coords), np.asarray(out, dtype=np.int64)


def elevate_bricks(coords, elements):
    """hex8 -> hex27 (mesher.py:246-330): 8 corners, 12 edges, faces in the order
    left,right,front,back,bottom,top (mesher.py:324-327), interior."""
    base = np.asarray(coords, dtype=np.float64)
    coords = list(base)
    edges, faces, out = {}, {}, []
    e_pairs = [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7)]
    f_quads = [(3, 0, 4, 7), (1, 2, 6, 5), (0, 1, 5, 4), (2, 3, 7, 6), (0, 1, 2, 3), (4, 5, 6, 7)]
    # Creation ORDER of new nodes in the reference: all 12 edges, then faces in the
    # order bottom, top, front, right, back, left (mesher.py:311-316), then the interior node.
    f_creation = [4, 5, 2, 1, 3, 0]
    for loc in np.asarray(elements):
        en = []
        for a, b in e_pairs:
            key = tuple(sorted((loc[a], loc[b])))
            if key not in edges:
                edges[key] = len(coords)
                coords.append(0.5 * (base[loc[a]] + base[loc[b]]))
            en.append(edges[key])
        fn = [None] * 6
        for f in f_creation:
            ids = [loc[q] for q in f_quads[f]]
            key = tuple(sorted(ids))
            if key not in faces:
                faces[key] = len(coords)
                coords.append(np.mean(base[ids], axis=0))
            fn[f] = faces[key]
        centre = len(coords)
        coords.append(np.mean(base[loc], axis=0))
        out.append(list(loc) + en + fn + [centre])
    return np.asarray(coords), np.asarray(out, dtype=np.int64)


def elevate_quads(coords, elements):
    """quad4 -> quad9 in the quad9 node order of spaces.py:1924 (corners, mid-sides
    0-1,1-2,2-3,3-0, centre).  The reference has no quad elevation (mesher.py:355-360);
    this is a harness helper for BASELINE config 2."""
    base = np.asarray(coords, dtype=np.float64)
    coords = list(base)
    edges, out = {}, []
    for loc in np.asarray(elements):
        en = []
        for a, b in [(0, 1), (1, 2), (2, 3), (3, 0)]:
            key = tuple(sorted((loc[a], loc[b])))
            if key not in edges:
                edges[key] = len(coords)
                coords.append(0.5 * (base[loc[a]] + base[loc[b]]))
            en.append(edges[key])
        centre = len(coords)
        coords.append(np.mean(base[loc], axis=0))
        out.append(list(loc) + en + [centre])
    return np.asarray(coords), np.asarray(out, dtype=np.int64)
