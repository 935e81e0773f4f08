"""Multi-GPU parity worker (launched by torchrun, one rank per GPU).

    torchrun --nproc-per-node N tests/multi_gpu_worker.py [poisson|neohooke] [M] [cg|bicgstab] [slab|rcb]

poisson   M^3 hex8 Poisson problem, nf = 1 (BASELINE config 4 family)
neohooke  M^3 hex8 neo-Hooke brick clamped at x = 0 with a traction face at x = 1, nf = 3, a domain and a surface set
          (BASELINE config 5 family)
Every rank assembles and solves its slab (or its recursive-coordinate-bisection part) through the public API
(settings['b200 partition']); the owned parts are summed into a global vector with the library's host all-reduce and
compared on rank 0 with the oracle ('scipy'/'lapack' reference path) on the whole mesh.
Exit code 0 = parity within 1e-8 relative L2 and equal Newton step counts.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def neohooke_api_problem(p):
    """settings / static_settings of the public API for problems.neo_hooke_brick (global mesh)."""
    from autopdex_b200 import models, seeder, spaces
    dom, sur = p["sets"]
    t = np.asarray(sur["model"]["traction"], dtype=np.float64)
    Em, nu = dom["model"]["youngs_modulus"], dom["model"]["poisson_ratio"]
    weak = models.hyperelastic_steady_state_weak(models.neo_hooke, lambda x, s: Em, lambda x, s: nu, "3d")
    el = models.isoparametric_domain_element_galerkin(weak, spaces.fem_iso_line_quad_brick, *seeder.gauss_legendre_nd(3, 2))
    tr = models.neumann_weak(lambda x, s: t)
    sf = models.isoparametric_surface_element_galerkin(tr, spaces.fem_iso_line_quad_brick, *seeder.gauss_legendre_nd(2, 2),
                                                       tangent_contributions=False)
    static_settings = {"assembling mode": ("user element", "user element"), "solution structure": ("nodal imposition",) * 2,
                       "model": (el, sf), "solver type": "newton", "solver backend": "b200", "solver": "bicgstab",
                       "type of preconditioner": "jacobi", "verbose": -1}
    return static_settings


def run_case(problem, m, krylov, partition, rank, world, precond="jacobi"):
    import bench
    from autopdex_b200 import backend, mesher, solver
    from tests import problems
    pg = problems.poisson_hex(m) if problem == "poisson" else problems.neo_hooke_brick(m)
    nf = pg["nf"]
    conns = tuple(s["conn"] for s in pg["sets"])
    if problem == "poisson" and partition == "slab" and precond == "jacobi":
        # the bench's memory-lean slab construction (never materialises the global mesh)
        settings, static_settings, _ = bench.build_problem(m, rank, world)
        sp = mesher.slab_partition((m, m, m), rank, world)
        nodes = np.arange(sp["node_lo"], sp["node_hi"])
    else:
        if partition == "rcb":
            pt = mesher.rcb_partition(pg["coords"], conns, rank, world)
        else:
            pt = mesher.slab_partition_mesh(pg["coords"], conns, (m, m, m), rank, world)
        nodes = pt["nodes"]
        if problem == "poisson":
            _, static_settings, _ = bench.build_problem(2, 0, 1)
        else:
            static_settings = neohooke_api_problem(pg)
        settings = {"connectivity": tuple(e.astype(np.int32) for e in pt["elements"]), "node coordinates": pg["coords"][nodes],
                    "dirichlet dofs": pg["mask"][nodes], "dirichlet conditions": np.zeros((nodes.size, nf)),
                    "b200 partition": pt["b200 partition"]}
    static_settings = dict(static_settings, solver=krylov)
    if precond == "multigrid":       # partitioned geometric multigrid hierarchy (slab partitions only)
        static_settings["type of preconditioner"] = "multigrid"
        settings["b200 multigrid"] = {"n_elements": (m, m, m)}
        if os.environ.get("APDX_TEST_MG_LEVELS"):      # experiments: truncate the hierarchy (large coarsest level)
            settings["b200 multigrid"]["levels"] = int(os.environ["APDX_TEST_MG_LEVELS"])
    n_local = settings["node coordinates"].shape[0]
    sol, (steps, res, div) = solver.solver(np.zeros((n_local, nf)), settings, static_settings, tol=1e-12)
    st = dict(solver.last_stats)
    part = settings["b200 partition"]
    glob = np.zeros((pg["coords"].shape[0], nf))
    own = slice(part["owned_node_begin"], part["owned_node_end"])
    glob[nodes[own]] = np.asarray(sol).reshape(n_local, nf)[own]
    glob = backend.comm_allreduce_host(glob.ravel()).reshape(glob.shape)
    ok = True
    if rank == 0:
        from oracle import solve as osolve
        prob = osolve.Problem(pg["sets"], pg["coords"], pg["mask"], pg["values"])
        ref, (rsteps, _, rdiv) = osolve.damped_newton(prob, np.zeros(pg["mask"].shape))
        err = np.linalg.norm(glob.ravel() - ref.ravel()) / np.linalg.norm(ref)
        ok = err < 1e-8 and steps == rsteps and div == rdiv
        print("multi-gpu parity: ranks=%d %s m=%d nf=%d %s%s %s steps=%d/%d res=%.2e krylov_iters=%d rel-L2=%.2e -> %s"
              % (world, problem, m, nf, krylov, "+multigrid" if precond == "multigrid" else "", partition, steps, rsteps, res,
                 int(st["krylov_iters"]), err, "OK" if ok else "FAIL"), flush=True)
    solver.clear_plan_cache()
    return bool(backend.comm_allreduce_host([0.0 if ok else 1.0], "max")[0] == 0.0)


def run_quad_case(n, precond, rank, world):
    """2-D: the README problem (Q1 quads, dict dofs, 'user potential' route) on slabs, Jacobi or multigrid, against the oracle."""
    from autopdex_b200 import backend, mesher, models, seeder, solver, spaces
    from oracle import solve as osolve
    from tests import problems
    p = problems.readme_poisson(n)
    coords, elems = p["coords"], p["sets"][0]["conn"]
    pt = mesher.slab_partition_mesh(coords, (elems,), (n, n), rank, world)
    nodes = pt["nodes"]
    integrand = models.poisson_potential("phi", source_fun=problems.readme_source)
    pot = models.mixed_reference_domain_potential(integrand, {"phi": spaces.fem_iso_line_quad_brick},
                                                  *seeder.gauss_legendre_nd(dimension=2, order=2), "phi")
    static_settings = {"assembling mode": ("user potential",), "solution structure": ("nodal imposition",), "model": (pot,),
                       "solver type": "newton", "solver backend": "b200", "solver": "cg", "type of preconditioner": precond, "verbose": -1}
    settings = {"connectivity": ({"phi": pt["elements"][0].astype(np.int32)},), "dirichlet dofs": {"phi": p["mask"][nodes, 0]},
                "node coordinates": {"phi": coords[nodes]}, "dirichlet conditions": {"phi": np.zeros(nodes.size)},
                "b200 partition": pt["b200 partition"]}
    if precond == "multigrid":
        settings["b200 multigrid"] = {"n_elements": (n, n)}
    sol, (steps, res, div) = solver.solver({"phi": np.zeros(nodes.size)}, settings, static_settings, tol=1e-12)
    its = int(solver.last_stats["krylov_iters"])
    part = settings["b200 partition"]
    own = slice(part["owned_node_begin"], part["owned_node_end"])
    glob = np.zeros(coords.shape[0])
    glob[nodes[own]] = np.asarray(sol["phi"])[own]
    glob = backend.comm_allreduce_host(glob)
    ok = True
    if rank == 0:
        prob = osolve.Problem(p["sets"], p["coords"], p["mask"], p["values"])
        ref, (rsteps, _, rdiv) = osolve.damped_newton(prob, np.zeros(p["mask"].shape))
        err = np.linalg.norm(glob - ref.ravel()) / np.linalg.norm(ref)
        ok = err < 1e-8 and steps == rsteps and div == rdiv
        print("multi-gpu parity: ranks=%d readme quad4 n=%d (2-D, dict dofs) cg+%s slab steps=%d/%d krylov_iters=%d rel-L2=%.2e -> %s"
              % (world, n, precond, steps, rsteps, its, err, "OK" if ok else "FAIL"), flush=True)
    solver.clear_plan_cache()
    return bool(backend.comm_allreduce_host([0.0 if ok else 1.0], "max")[0] == 0.0)


def run_dae_case(m, precond, rank, world, scheme="backward_euler"):
    """dae.TimeSteppingManager (BackwardEuler, or another scheme of tests/test_zz_gpu_r02_dae._reference_steps; transient heat conduction through the 'user residual' route) on slab
    partitions, against a SciPy time loop on the oracle's global mass / stiffness matrices (tests/test_zz_gpu_r02_dae.py)."""
    from autopdex_b200 import backend, dae, mesher, solver
    from tests import test_zz_gpu_r02_dae as T
    dt, n_steps = 0.05, 3
    coords, K, M, F, mask, values, res, gs = T._settings(m)
    elems = gs["connectivity"][0]["theta"]
    pt = mesher.slab_partition_mesh(coords, (elems,), (m, m, m), rank, world)
    nodes = pt["nodes"]
    settings = {"connectivity": ({"theta": pt["elements"][0].astype(np.int32)},), "node coordinates": {"theta": coords[nodes]},
                "dirichlet dofs": {"theta": mask[nodes]}, "dirichlet conditions": {"theta": values[nodes]}, "current time": 0.0,
                "b200 partition": pt["b200 partition"]}
    q0 = 0.3 * np.cos(coords[:, 1])
    integ = T._integrator(scheme)
    ref_steps = T._reference_steps(scheme, K, M, F, mask, values, q0, dt, n_steps)[1] if rank == 0 else None
    static_settings = {"assembling mode": ("user residual",), "model": (res,), "time integrators": {"theta": integ},
                       "solver backend": "b200", "solver": "cg", "type of preconditioner": precond, "verbose": -1}
    if precond == "multigrid":
        settings["b200 multigrid"] = {"n_elements": (m, m, m)}
    mgr = dae.TimeSteppingManager(static_settings, tol=1e-13)
    out = mgr.run({"theta": q0[nodes]}, dt, dt * n_steps, 100, settings)
    part = settings["b200 partition"]
    own = slice(part["owned_node_begin"], part["owned_node_end"])
    glob = np.zeros(coords.shape[0])
    glob[nodes[own]] = np.asarray(out.q["theta"])[own]
    glob = backend.comm_allreduce_host(glob)
    ok = True
    if rank == 0:
        ref = ref_steps[-1]
        err = np.linalg.norm(glob - ref) / np.linalg.norm(ref)
        ok = err < 1e-9 and out.num_accepted == n_steps and all(it == 1 for it in out.newton_iterations)
        print("multi-gpu parity: ranks=%d dae %s heat m=%d cg+%s slab steps=%d rel-L2=%.2e -> %s"
              % (world, scheme.replace("_", "-"), m, precond, out.num_accepted, err, "OK" if ok else "FAIL"), flush=True)
    solver.clear_plan_cache()
    return bool(backend.comm_allreduce_host([0.0 if ok else 1.0], "max")[0] == 0.0)


def main():
    """argv: [poisson|neohooke] [M] [cg|bicgstab] [slab|rcb] [jacobi|multigrid]   one case
             matrix [Mp] [Mn]                                   {poisson Mp, neohooke Mn} x {cg, bicgstab} x {slab, rcb}
             mgmatrix [Mp] [Mn]                                 {poisson Mp, neohooke Mn}, multigrid-preconditioned CG on slabs"""
    from autopdex_b200 import backend
    argv = sys.argv[1:]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    backend.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    from torch.distributed import TCPStore
    store = TCPStore(os.environ.get("MASTER_ADDR", "127.0.0.1"), int(os.environ.get("MASTER_PORT", "29500")) + 1,
                     world, rank == 0)
    if rank == 0:
        store.set("id", backend.comm_unique_id())
    backend.comm_init(bytes(store.get("id")), rank, world)
    if argv and argv[0] == "matrix":
        mp = int(argv[1]) if len(argv) > 1 else 20
        mn = int(argv[2]) if len(argv) > 2 else 12
        mn = max(mn, 2 * world)      # the brick is clamped at x = 0: every slab needs a free node plane of its own
        mp = max(mp, 2 * world)      # all faces are Dirichlet faces: the first and last slab need an interior plane
        cases = [(pb, m, k, pt) for pb, m in (("poisson", mp), ("neohooke", mn)) for k in ("cg", "bicgstab")
                 for pt in ("slab", "rcb")]
    elif argv and argv[0] == "dae":       # dae [M]: the time-stepping manager on slabs, Jacobi and multigrid
        m = max(int(argv[1]) if len(argv) > 1 else 8, 4 * world)
        cases = [("dae", m, "jacobi"), ("dae", m, "multigrid"), ("dae", m, "jacobi", "dirk2"), ("dae", m, "multigrid", "am2")]
    elif argv and argv[0] == "mgmatrix":
        mp = int(argv[1]) if len(argv) > 1 else 32
        mn = int(argv[2]) if len(argv) > 2 else 16
        cases = [("poisson", mp, "cg", "slab", "multigrid"), ("neohooke", mn, "cg", "slab", "multigrid"),
                 ("quad", max(mp, 4 * world), "multigrid")]
    else:
        problem = "poisson"
        if argv and not argv[0].isdigit():
            problem, argv = argv[0], argv[1:]
        cases = [(problem, int(argv[0]) if len(argv) > 0 else 20, argv[1] if len(argv) > 1 else "cg",
                  argv[2] if len(argv) > 2 else "slab", argv[3] if len(argv) > 3 else "jacobi")]
    ok = True
    import signal

    def _watchdog(signum, frame):   # a rank that raised leaves the others inside a collective: bound every case
        print("multi-gpu parity: rank %d: case timed out (a peer probably failed) -> FAIL" % rank, flush=True)
        os._exit(3)
    signal.signal(signal.SIGALRM, _watchdog)
    for c in cases:
        signal.alarm(int(os.environ.get("APDX_CASE_TIMEOUT", "150")))
        try:
            ok = (run_dae_case(c[1], c[2], rank, world, *c[3:]) if c[0] == "dae" else
                  run_quad_case(c[1], c[2], rank, world) if c[0] == "quad" else run_case(*c[:4], rank, world, *c[4:])) and ok
        except Exception as e:   # keep the remaining cases running; every rank raises alike (collective set-up errors)
            ok = False
            print("multi-gpu parity: ranks=%d %s -> ERROR on rank %d: %s" % (world, " ".join(map(str, c)), rank, e),
                  flush=True)
            from autopdex_b200 import solver
            solver.clear_plan_cache()
    signal.alarm(0)
    backend.comm_destroy()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
