#!/usr/bin/env python
"""Time the sliced-ELL SpMV of a 3-D problem with symmetric (mirrored) storage on or off (APDX_SELL_SYM) and print a
checksum of y = A x for a fixed x.
   APDX_SELL_SYM=0 python tools/time_spmv.py poisson|neohooke|linel N [reps]
   python tools/time_spmv.py all poisson N [reps]     # both storage modes, each in its own process, one JSON line each"""
import hashlib
import json
import os
import subprocess
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

VARIANTS = ["1", "0"]   # APDX_SELL_SYM: mirrored storage on / off
CUBE = [[0., 0., 0.], [1., 0., 0.], [1., 1., 0.], [0., 1., 0.], [0., 0., 1.], [1., 0., 1.], [1., 1., 1.], [0., 1., 1.]]


def main(model, n, reps):
    from autopdex_b200 import backend, mesher, seeder
    coords, elems = mesher.structured_mesh((n, n, n), CUBE, "brick")
    nf = 1 if model == "poisson" else 3
    if model == "poisson":   # homogeneous Dirichlet on all six faces, as in BASELINE config 4
        onb = (np.abs(coords) < 1e-12).any(axis=1) | (np.abs(coords - 1.0) < 1e-12).any(axis=1)
        mask = onb[:, None]
        st = backend.SetSpec("domain", "poisson_weak", elems.astype(np.int32), family="quad_brick",
                             gp=seeder.gauss_legendre_nd(3, 2), params={"coefficient": 1.0, "source": 1.0})
    else:
        mask = np.repeat((np.abs(coords[:, 0]) < 1e-12)[:, None], nf, axis=1)
        st = backend.SetSpec("domain", "neo_hooke" if model == "neohooke" else "linear_elasticity", elems.astype(np.int32),
                             family="quad_brick", gp=seeder.gauss_legendre_nd(3, 2), mode="3d",
                             params={"youngs_modulus": 100.0, "poisson_ratio": 0.3})
    plan = backend.Plan(3, coords.shape[0], nf, [st], mask)
    plan.set_coords(coords)
    rng = np.random.default_rng(0)
    d = backend.DeviceArray.from_host(rng.uniform(-1e-3, 1e-3, mask.size))
    r = backend.DeviceArray(mask.size)
    plan.assemble(d, True, r)
    n_free, nnz = plan.n_free, plan.nnz_reduced
    xh = rng.uniform(-1.0, 1.0, n_free)
    x = backend.DeviceArray.from_host(xh)
    y = backend.DeviceArray(n_free)
    plan.spmv(x, y)
    yh = y.download()
    ms = plan.time_spmv(reps)
    alg = nnz * 12 + n_free * 16 + (n_free + 1) * 4
    impl = plan.stats().get("sell_bytes", 0.0) + n_free * 16
    print(json.dumps({"sell_sym": os.environ.get("APDX_SELL_SYM", "default"), "sell": plan.sell_info(), "model": model, "n": n, "n_free": n_free,
                      "nnz": nnz, "ms": ms, "algorithmic_gbs": alg / ms * 1e-6, "implementation_gbs": impl / ms * 1e-6,
                      "y_sha1": hashlib.sha1(yh.tobytes()).hexdigest()[:16], "y_sum": float(yh.sum())}))


if __name__ == "__main__":
    if sys.argv[1] == "all":
        for sym in VARIANTS:
            env = dict(os.environ, APDX_SELL_SYM=sym)
            subprocess.run([sys.executable, os.path.abspath(__file__)] + sys.argv[2:], env=env, check=False)
    else:
        main(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 50)
