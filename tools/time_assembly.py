#!/usr/bin/env python
"""Time the assembly pass (element kernel + deterministic scatter) of a 3-D problem through apdx_assemble.
   python tools/time_assembly.py poisson|neohooke|linel N [reps]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autopdex_b200 import backend, mesher, seeder  # noqa: E402

CUBE = [[0., 0., 0.], [1., 0., 0.], [1., 1., 0.], [0., 1., 0.], [0., 0., 1.], [1., 0., 1.], [1., 1., 1.], [0., 1., 1.]]
model, n = sys.argv[1], int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
coords, elems = mesher.structured_mesh((n, n, n), CUBE, "brick")
nf = 1 if model == "poisson" else 3
mask = np.repeat((np.abs(coords[:, 0]) < 1e-12)[:, None], nf, axis=1)
if model == "poisson":
    st = backend.SetSpec("domain", "poisson_weak", elems.astype(np.int32), family="quad_brick", gp=seeder.gauss_legendre_nd(3, 2),
                         params={"coefficient": 1.0, "source": 1.0})
else:
    st = backend.SetSpec("domain", "neo_hooke" if model == "neohooke" else "linear_elasticity", elems.astype(np.int32),
                         family="quad_brick", gp=seeder.gauss_legendre_nd(3, 2), mode="3d",
                         params={"youngs_modulus": 100.0, "poisson_ratio": 0.3})
plan = backend.Plan(3, coords.shape[0], nf, [st], mask)
plan.set_coords(coords)
rng = np.random.default_rng(0)
d = backend.DeviceArray.from_host(rng.uniform(-1e-3, 1e-3, mask.size))
r = backend.DeviceArray(mask.size)
tan, res = [], []
for i in range(reps + 2):
    plan.assemble(d, True, r)
    tan.append(plan.stats()["assembly_tangent_ms"])
    plan.assemble(d, False, r)
    res.append(plan.stats()["assembly_residual_ms"])
tan, res = tan[2:], res[2:]
n_el = elems.shape[0]
print(json.dumps({"model": model, "n": n, "elements": n_el, "tangent_pass_ms": float(np.mean(tan)), "residual_pass_ms": float(np.mean(res)),
                  "elements_per_s_tangent": n_el / (np.mean(tan) * 1e-3), "note": "apdx_assemble: full CSR + sliced-ELL values + residual"}))
