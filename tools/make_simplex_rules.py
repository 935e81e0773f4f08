#!/usr/bin/env python
"""Tabulate the reference's simplex quadrature rules into autopdex_b200/data/simplex_rules.npz.

The triangle rules (https://mathsfromnothing.au/triangle-quadrature-rules/) and tetrahedron rules (Jaskowiec and
Sukumar 2020, doi 10.1002/nme.6313) that autopdex.seeder.int_pts_ref_tri / int_pts_ref_tet return (seeder.py:1813-2285,
2288-3421) are published tables of numbers; results identical to the reference's need the same tables, point order
included.  This script CALLS the unmodified reference functions (on the NumPy stand-in for JAX of tests/golden) for every
order they implement and stores their outputs; it runs only in the build container, the .npz is committed.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import fakejax  # noqa: E402

fakejax.install()
from autopdex import seeder  # noqa: E402

out = {}
for name, fun in (("tri", seeder.int_pts_ref_tri), ("tet", seeder.int_pts_ref_tet)):
    orders = []
    for order in range(1, 40):
        try:
            res = fun(order)
        except Exception:
            res = None
        if res is None:
            continue
        x, w = res
        x, w = np.asarray(x, dtype=np.float64), np.asarray(w, dtype=np.float64)
        out["%s_%d_x" % (name, order)], out["%s_%d_w" % (name, order)] = x, w
        orders.append(order)
        print(name, order, x.shape, w.sum())
    out["%s_orders" % name] = np.asarray(orders)
path = os.path.join(ROOT, "autopdex_b200", "data", "simplex_rules.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes")
