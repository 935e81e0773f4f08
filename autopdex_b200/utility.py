"""Array / dict helpers of the reference interface that the hot path touches (host, NumPy).

Names and semantics follow autopdex/utility.py: dict_zeros_like (:82-91), dict_ones_like (:93-102),
dict_flatten (:104-128), reshape_as (:130-199), dof_select (:448-462).
"""
from collections.abc import Mapping

import numpy as np


def _is_dict(x):
    return isinstance(x, Mapping)


def dict_zeros_like(arr, **kw):
    if _is_dict(arr):
        return {k: np.zeros_like(np.asarray(v), **kw) for k, v in arr.items()}
    return np.zeros_like(np.asarray(arr), **kw)


def dict_ones_like(arr, **kw):
    if _is_dict(arr):
        return {k: np.ones_like(np.asarray(v), **kw) for k, v in arr.items()}
    return np.ones_like(np.asarray(arr), **kw)


def dict_flatten(arr):
    """Concatenate the fields in key order -> global dof id = field offset + node*dpn + comp."""
    if _is_dict(arr):
        parts = [dict_flatten(arr[k]) for k in arr.keys()]
        return np.concatenate(parts) if parts else np.array([])
    return np.asarray(arr).ravel()


def reshape_as(flat_array, signature_array):
    flat_array = np.asarray(flat_array)
    if _is_dict(signature_array):
        out, start = {}, 0
        for k, sig in signature_array.items():
            if _is_dict(sig):
                size = dict_flatten(sig).size
                out[k] = reshape_as(flat_array[start:start + size], sig)
            else:
                sig = np.asarray(sig)
                size = sig.size
                out[k] = flat_array[start:start + size].reshape(sig.shape)
            start += size
        if start != flat_array.size:
            raise ValueError("The size of flat_array does not match the total size of signature_array.")
        return out
    sig = np.asarray(signature_array)
    if flat_array.size != sig.size:
        raise ValueError("The size of flat_array does not match the size of signature_array.")
    return flat_array.reshape(sig.shape)


def dof_select(dirichlet_nodes, selected_fields):
    dirichlet_nodes = np.asarray(dirichlet_nodes)
    if isinstance(selected_fields, bool):
        return dirichlet_nodes * selected_fields
    return np.outer(dirichlet_nodes, np.asarray(selected_fields))


# ---- result export -------------------------------------------------------------------------------
# The reference recommends meshio for VTK files (autopdex/plotter.py:16); meshio is not a dependency here, so the
# solutions that come back from the device are written by this small legacy-VTK writer.  The reference's node
# orders (spaces.py:1913-1949, mesher.py:332-360) are VTK's for every element below.
_VTK_CELL = {(1, 2): 3, (1, 3): 21, (2, 3): 5, (2, 6): 22, (2, 4): 9, (2, 9): 28, (3, 4): 10, (3, 10): 24, (3, 8): 12,
             (3, 27): 29}


def write_vtk(path, node_coordinates, connectivity, point_data=None, cell_dim=None):
    """Write an unstructured grid with nodal fields as a binary legacy VTK file.

    node_coordinates (n_nodes, dim); connectivity (n_elem, nen) of ONE element type; point_data
    {name: (n_nodes,) or (n_nodes, k)} with k <= 3 written as a vector, otherwise as k scalars name_0 ...;
    cell_dim: topological dimension of the elements (default: the space dimension; pass dim - 1 for the
    surface sets of a 'user element' problem)."""
    x = np.asarray(node_coordinates, dtype=np.float64)
    conn = np.asarray(connectivity)
    if x.ndim != 2 or conn.ndim != 2:
        raise ValueError("write_vtk: node_coordinates (n_nodes, dim) and connectivity (n_elem, nen) expected")
    n, dim = x.shape
    ne, nen = conn.shape
    key = (dim if cell_dim is None else cell_dim, nen)
    if key not in _VTK_CELL:
        raise ValueError("write_vtk: no VTK cell for a %d-dimensional element with %d nodes" % key)
    if ne and (conn.min() < 0 or conn.max() >= n):
        raise ValueError("write_vtk: connectivity refers to nodes outside node_coordinates")
    pts = np.zeros((n, 3), dtype=">f8")
    pts[:, :dim] = x
    cells = np.empty((ne, nen + 1), dtype=">i4")
    cells[:, 0] = nen
    cells[:, 1:] = conn
    with open(path, "wb") as f:
        f.write(b"# vtk DataFile Version 3.0\nautopdex_b200 result\nBINARY\nDATASET UNSTRUCTURED_GRID\n")
        f.write(b"POINTS %d double\n" % n + pts.tobytes() + b"\n")
        f.write(b"CELLS %d %d\n" % (ne, ne * (nen + 1)) + cells.tobytes() + b"\n")
        f.write(b"CELL_TYPES %d\n" % ne + np.full(ne, _VTK_CELL[key], dtype=">i4").tobytes() + b"\n")
        if point_data:
            f.write(b"POINT_DATA %d\n" % n)
            for name, v in point_data.items():
                v = np.asarray(v, dtype=np.float64).reshape(n, -1)
                tag = str(name).replace(" ", "_").encode()
                if 1 < v.shape[1] <= 3:
                    vec = np.zeros((n, 3), dtype=">f8")
                    vec[:, :v.shape[1]] = v
                    f.write(b"VECTORS " + tag + b" double\n" + vec.tobytes() + b"\n")
                else:
                    for k in range(v.shape[1]):
                        t = tag if v.shape[1] == 1 else tag + b"_%d" % k
                        f.write(b"SCALARS " + t + b" double 1\nLOOKUP_TABLE default\n" + v[:, k].astype(">f8").tobytes() + b"\n")
