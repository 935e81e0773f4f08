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
