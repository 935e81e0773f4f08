#!/usr/bin/env python
"""Run the five BASELINE.json configurations through the public API of the b200 backend and print one JSON line
each (timings from the plan's CUDA events, parity guards).  Sizes are arguments so that the same script serves
parity-sized and full-sized runs:

  python tools/run_config.py readme 200        # config 1: README 2-D Poisson, 200x200 Q1, user potential
  python tools/run_config.py cook 64 2         # config 2: Cook's membrane N x N, order 1|2, neo-Hooke, load stepping
  python tools/run_config.py heat 32 10        # config 3: transient heat, N^3 hex8, backward Euler, n steps
  python tools/run_config.py poisson 256       # config 4 (bench.py's workload)
  python tools/run_config.py neohooke 128      # config 5: 3-D neo-Hooke brick, Newton + BiCGSTAB, load steps
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autopdex_b200 import mesher, models, seeder, solver, spaces  # noqa: E402

UNIT_CUBE = [[0., 0., 0.], [1., 0., 0.], [1., 1., 0.], [0., 1., 0.], [0., 0., 1.], [1., 0., 1.], [1., 1., 1.], [0., 1., 1.]]


def base_static(**kw):
    s = {"solution structure": None, "solver type": "newton", "solver backend": "b200", "solver": "cg",
         "type of preconditioner": "jacobi", "verbose": -1}
    s.update(kw)
    n = len(s["assembling mode"])
    s["solution structure"] = ("nodal imposition",) * n
    return s


def readme(n):
    pts = [[0., 0.], [1., 0.], [1., 1.], [0., 1.]]
    coords, elems = mesher.structured_mesh((n, n), pts, "quad")
    tol = 1e-9
    onb = (np.abs(coords) < tol).any(axis=1) | (np.abs(coords - 1) < tol).any(axis=1)
    corner = ((np.abs(coords[:, 0]) < tol) | (np.abs(coords[:, 0] - 1) < tol)) & ((np.abs(coords[:, 1]) < tol) | (np.abs(coords[:, 1] - 1) < tol))
    mask = onb & ~corner                                    # geometry.psdf_polygon is NaN at the corners
    src = lambda x: 20.0 * (np.sin(10.0 * np.sum(x * x, axis=-1)) - np.cos(10.0 * np.sum((x - np.array([1.0, 0.5])) ** 2, axis=-1)))
    pot = models.mixed_reference_domain_potential(models.poisson_potential("phi", source_fun=src),
                                                  {"phi": spaces.fem_iso_line_quad_brick}, *seeder.gauss_legendre_nd(2, 2), "phi")
    st = base_static(**{"assembling mode": ("user potential",), "model": (pot,)})
    settings = {"connectivity": ({"phi": elems},), "node coordinates": {"phi": coords}, "dirichlet dofs": {"phi": mask},
                "dirichlet conditions": {"phi": np.zeros(coords.shape[0])}}
    dofs = {"phi": np.zeros(coords.shape[0])}
    solver.solver(dofs, settings, st, tol=1e-10)            # builds the plan
    t = time.perf_counter()
    sol, info = solver.solver(dofs, settings, st, tol=1e-10)
    return {"config": "README 2D Poisson %dx%d Q1 (user potential)" % (n, n), "e2e_ms": (time.perf_counter() - t) * 1e3,
            "newton": list(map(float, info)), "sum_phi": float(sol["phi"].sum()),
            "expected_sum_phi": {5: 1.9066412530282952, 200: 2233.155221149569}.get(n), **solver.last_stats}


def cook(n, order, precond="jacobi"):
    """BASELINE config 2 (ii).  precond = 'multigrid' (order 1 only): CG preconditioned by the geometric hierarchy."""
    pts = [[0., 0.], [48., 44.], [48., 60.], [0., 44.]]
    coords, elems = mesher.structured_mesh((n, n), pts, "quad")
    line = mesher.boundary_faces((n, n), 0, 1)
    if order == 2:
        c4 = coords
        coords, elems = mesher.elevate_quads(coords, elems)
        # line3 boundary elements: mid-side node of each boundary edge
        mids = {}
        for e in elems:
            for a, b, m in ((0, 1, 4), (1, 2, 5), (2, 3, 6), (3, 0, 7)):
                mids[(min(e[a], e[b]), max(e[a], e[b]))] = e[m]
        line = np.array([[a, b, mids[(min(a, b), max(a, b))]] for a, b in line])
    lam, mu = 100.0, 40.0
    Em, nu = mu * (3 * lam + 2 * mu) / (lam + mu), lam / (2 * (lam + mu))      # quadrilaterals_p_refinement.py:63-66
    weak = models.hyperelastic_steady_state_weak(models.neo_hooke, lambda x, s: s["youngs modulus"], lambda x, s: s["poisson ratio"], "plain strain")
    el = models.isoparametric_domain_element_galerkin(weak, spaces.fem_iso_line_quad_brick, *seeder.gauss_legendre_nd(2, 2 * order))
    tr = models.neumann_weak(lambda x, s: np.asarray([0.0, s["load multiplier"]]))
    sf = models.isoparametric_surface_element_galerkin(tr, spaces.fem_iso_line_quad_brick, *seeder.gauss_legendre_nd(1, 2 * order), tangent_contributions=False)
    st = base_static(**{"assembling mode": ("user element", "user element"), "model": (el, sf), "solver": "bicgstab"})
    if precond == "multigrid":
        st = dict(st, **{"solver": "cg", "type of preconditioner": "multigrid"})
    mask = np.repeat((np.abs(coords[:, 0]) < 1e-9)[:, None], 2, axis=1)
    q0 = 4.0
    settings = {"connectivity": (elems, line), "node coordinates": coords, "dirichlet dofs": mask,
                "dirichlet conditions": np.zeros(mask.shape), "youngs modulus": Em, "poisson ratio": nu, "load multiplier": q0}
    if precond == "multigrid":
        settings["b200 multigrid"] = {"n_elements": (n, n)}

    def mult(s, m):
        s["load multiplier"] = m * q0
        return s
    t = time.perf_counter()
    out = solver.adaptive_load_stepping(np.zeros(mask.shape), settings, st, mult, False, None, newton_tol=1e-8, tol=1e-10)
    dofs = out[0]
    tip = dofs[np.argmax(coords[:, 0] + coords[:, 1])]
    return {"config": "Cook's membrane %dx%d Q%d neo-Hooke, adaptive load stepping, %s" % (n, n, order, "multigrid-PCG" if precond == "multigrid" else "Jacobi-BiCGSTAB"),
            "total_s": time.perf_counter() - t, "multiplier": float(out[1]), "tip_displacement": tip.tolist(),
            "dofs": int(mask.size), **solver.last_stats}


def heat(n, steps):
    coords, elems = mesher.structured_mesh((n, n, n), UNIT_CUBE, "brick")
    face = mesher.boundary_faces((n, n, n), 0, 1)
    gp3, gp2 = seeder.gauss_legendre_nd(3, 2), seeder.gauss_legendre_nd(2, 2)
    cond = models.isoparametric_domain_element_galerkin(models.poisson_weak(lambda x, s: 1.0), spaces.fem_iso_line_quad_brick, *gp3)
    cap = models.isoparametric_domain_element_galerkin(models.forward_backward_euler_weak(lambda x, s: 0.1), spaces.fem_iso_line_quad_brick, *gp3)
    flux = models.isoparametric_surface_element_galerkin(models.neumann_weak(lambda x: -1.0e3), spaces.fem_iso_line_quad_brick, *gp2, tangent_contributions=False)
    st = base_static(**{"assembling mode": ("user element",) * 3, "model": (cond, cap, flux), "solver type": "linear"})
    mask = (np.abs(coords[:, 0]) < 1e-9)[:, None]
    dofs = np.zeros(mask.shape)
    settings = {"connectivity": (elems, elems, face), "node coordinates": coords, "dirichlet dofs": mask,
                "dirichlet conditions": np.zeros(mask.shape), "time increment": 50.0 / 250.0, "dofs n": dofs}
    times = []
    for k in range(steps):
        settings["dofs n"] = dofs
        t = time.perf_counter()
        dofs = dofs + solver.solver(dofs, settings, st, tol=1e-10)[0]             # maze_backward_euler.py:359-374
        times.append(time.perf_counter() - t)
    return {"config": "transient heat %d^3 hex8, backward Euler, %d steps (plan reused)" % (n, steps),
            "first_step_s": times[0], "later_step_ms": 1e3 * float(np.mean(times[1:])) if steps > 1 else None,
            "theta_sum": float(dofs.sum()), "theta_max": float(dofs.max()), **solver.last_stats}


def poisson(n, precond="jacobi", pre=0, post=0, coarsest=0, ratio=0, coarsest_ratio=0, levels=0):
    """python tools/run_config.py poisson 256 multigrid [pre post coarsest ratio coarsest_ratio levels]"""
    import bench
    settings, st, _ = bench.build_problem(n, 0, 1)
    st = dict(st, **{"type of preconditioner": precond})
    if precond == "multigrid":
        settings["b200 multigrid"] = {"n_elements": (n, n, n), "pre": pre, "post": post, "coarsest": coarsest,
                                      "ratio": float(ratio), "coarsest ratio": float(coarsest_ratio)}
        if levels:
            settings["b200 multigrid"]["levels"] = levels
    dofs = np.zeros((settings["node coordinates"].shape[0], 1))
    t = time.perf_counter()
    solver.solver(dofs, settings, st, tol=1e-8)
    t_first = time.perf_counter() - t
    solver.solver(dofs, settings, st, tol=1e-8)
    t = time.perf_counter()
    sol, info = solver.solver(dofs, settings, st, tol=1e-8)
    return {"config": "3D Poisson %d^3 hex8, Newton + %s-PCG 1e-8" % (n, precond), "e2e_ms": (time.perf_counter() - t) * 1e3,
            "first_call_s": t_first, "mg": settings.get("b200 multigrid"),
            "newton": list(map(float, info)), "sum": float(sol.sum()), **solver.last_stats}


def _dist():
    """(rank, world) -- under torchrun the NCCL communicator of the library is initialised (one rank per GPU)."""
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not getattr(_dist, "done", False):
        from autopdex_b200 import backend
        backend.set_device(int(os.environ.get("LOCAL_RANK", rank)))
        from torch.distributed import TCPStore  # plumbing only: ships the NCCL id between ranks
        store = TCPStore(os.environ.get("MASTER_ADDR", "127.0.0.1"), int(os.environ.get("MASTER_PORT", "29500")) + 1,
                         world, rank == 0)
        if rank == 0:
            store.set("id", backend.comm_unique_id())
        backend.comm_init(bytes(store.get("id")), rank, world)
        _dist.done = True
    return rank, world


def neohooke(n, load_steps=2, traction=-1.0, partition="slab", precond="jacobi"):
    """BASELINE config 5 (precond = 'multigrid': CG preconditioned by the geometric multigrid hierarchy instead of the
    Jacobi-BiCGSTAB of the config; slab partitions only).  Under torchrun (WORLD_SIZE > 1) the brick is split into slabs (or RCB parts) and every rank
    solves its part through the public API (settings['b200 partition']): halo exchange + all-reduces inside the Krylov
    loop; times are the max over ranks, the checksums are sums over the owned nodes."""
    from autopdex_b200 import backend
    rank, world = _dist()
    coords, elems = mesher.structured_mesh((n, n, n), UNIT_CUBE, "brick")
    face = mesher.boundary_faces((n, n, n), 0, 1)
    weak = models.hyperelastic_steady_state_weak(models.neo_hooke, lambda x, s: 100.0, lambda x, s: 0.3, "3d")
    el = models.isoparametric_domain_element_galerkin(weak, spaces.fem_iso_line_quad_brick, *seeder.gauss_legendre_nd(3, 2))
    tr = models.neumann_weak(lambda x, s: np.asarray([0.0, 0.0, s["load multiplier"]]))
    sf = models.isoparametric_surface_element_galerkin(tr, spaces.fem_iso_line_quad_brick, *seeder.gauss_legendre_nd(2, 2), tangent_contributions=False)
    st = base_static(**{"assembling mode": ("user element", "user element"), "model": (el, sf), "solver": "bicgstab"})
    if precond == "multigrid":
        st = dict(st, **{"solver": "cg", "type of preconditioner": "multigrid"})
    mask = np.repeat((np.abs(coords[:, 0]) < 1e-9)[:, None], 3, axis=1)
    settings = {"connectivity": (elems, face), "node coordinates": coords, "dirichlet dofs": mask,
                "dirichlet conditions": np.zeros(mask.shape), "load multiplier": 0.0}
    if precond == "multigrid":
        settings["b200 multigrid"] = {"n_elements": (n, n, n)}
    own = slice(0, coords.shape[0])
    if world > 1:
        pt = (mesher.rcb_partition(coords, (elems, face), rank, world) if partition == "rcb"
              else mesher.slab_partition_mesh(coords, (elems, face), (n, n, n), rank, world))
        nodes = pt["nodes"]
        settings.update({"connectivity": tuple(e.astype(np.int32) for e in pt["elements"]), "node coordinates": coords[nodes],
                         "dirichlet dofs": mask[nodes], "dirichlet conditions": np.zeros((nodes.size, 3)),
                         "b200 partition": pt["b200 partition"]})
        own = slice(pt["b200 partition"]["owned_node_begin"], pt["b200 partition"]["owned_node_end"])
    dofs = np.zeros(settings["dirichlet dofs"].shape)
    hist = []
    t0 = time.perf_counter()
    for k in range(1, load_steps + 1):
        settings["load multiplier"] = traction * k / load_steps
        t = time.perf_counter()
        dofs, info = solver.solver(dofs, settings, st, newton_tol=1e-8, tol=1e-8)
        ls = solver.last_stats
        tm = backend.comm_allreduce_host([time.perf_counter() - t, ls["assembly_tangent_ms"], ls["assembly_residual_ms"],
                                          ls["krylov_ms"], ls["total_ms"]], "max")
        hist.append({"newton_steps": int(info[0]), "res_norm": float(info[1]), "diverged": bool(info[2]), "step_s": float(tm[0]),
                     "assembly_tangent_ms": float(tm[1]), "assembly_residual_ms": float(tm[2]), "krylov_ms": float(tm[3]),
                     "krylov_iters": int(ls["krylov_iters"]), "total_ms": float(tm[4]),
                     "ms_per_krylov_iteration": float(tm[3]) / max(int(ls["krylov_iters"]), 1),
                     "ms_per_newton_step": float(tm[4]) / max(int(info[0]), 1)})
    d = np.asarray(dofs)[own]
    sums = backend.comm_allreduce_host([float((d * d).sum()), float(d[:, 2].sum())])
    out = {"config": "3D neo-Hooke %d^3 hex8 (%d dofs), Newton + %s 1e-8, %d load steps, %d GPU(s), %s partition"
                     % (n, mask.size, "multigrid-PCG" if precond == "multigrid" else "Jacobi-BiCGSTAB", load_steps, world,
                        partition if world > 1 else "no"),
           "n_gpus": world, "total_s": time.perf_counter() - t0, "load_steps": hist, "u_dot_u": float(sums[0]), "sum_uz": float(sums[1])}
    if world > 1:
        solver.clear_plan_cache()
        backend.comm_destroy()
    return out if rank == 0 else None


if __name__ == "__main__":
    name, args = sys.argv[1], [int(a) if a.lstrip("-").isdigit() else a for a in sys.argv[2:]]
    res = {"readme": readme, "cook": cook, "heat": heat, "poisson": poisson, "neohooke": neohooke}[name](*args)
    if res is not None:
        print(json.dumps(res), flush=True)
