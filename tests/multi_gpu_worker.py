"""Multi-GPU parity worker (launched by torchrun, one rank per GPU).

    torchrun --nproc-per-node N tests/multi_gpu_worker.py [M] [cg|bicgstab] [slab|rcb]

Every rank assembles and solves its slab (or its recursive-coordinate-bisection part) of the M^3 hex8 Poisson problem
through the public API (settings['b200 partition']); the owned parts are summed into a global vector with the library's
host all-reduce and compared on rank 0 with the oracle ('scipy'/'lapack' reference path) on the whole mesh.
Exit code 0 = parity within 1e-8 relative L2 and equal Newton step counts.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import bench
    from autopdex_b200 import backend, solver
    m = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    krylov = sys.argv[2] if len(sys.argv) > 2 else "cg"
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    backend.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    from torch.distributed import TCPStore
    store = TCPStore(os.environ.get("MASTER_ADDR", "127.0.0.1"), int(os.environ.get("MASTER_PORT", "29500")) + 1,
                     world, rank == 0)
    if rank == 0:
        store.set("id", backend.comm_unique_id())
    backend.comm_init(bytes(store.get("id")), rank, world)
    partition = sys.argv[3] if len(sys.argv) > 3 else "slab"
    from autopdex_b200 import mesher
    if partition == "rcb":
        # general partition: recursive coordinate bisection of the whole mesh, halo through neighbour lists
        # (apdx_plan_set_partition_lists); every rank builds the global mesh and keeps its part
        from tests import problems
        pg = problems.poisson_hex(m)
        _, static_settings, _ = bench.build_problem(2, 0, 1)
        pt = mesher.rcb_partition(pg["coords"], (pg["sets"][0]["conn"],), rank, world)
        nodes = pt["nodes"]
        settings = {"connectivity": (pt["elements"][0].astype(np.int32),), "node coordinates": pg["coords"][nodes],
                    "dirichlet dofs": pg["mask"][nodes], "dirichlet conditions": np.zeros((nodes.size, 1)),
                    "b200 partition": pt["b200 partition"]}
    else:
        settings, static_settings, _ = bench.build_problem(m, rank, world)
    static_settings = dict(static_settings, solver=krylov)
    n_local = settings["node coordinates"].shape[0]
    sol, (steps, res, div) = solver.solver(np.zeros((n_local, 1)), settings, static_settings, tol=1e-12)
    part = settings["b200 partition"]
    glob = np.zeros((m + 1) ** 3)
    own = slice(part["owned_node_begin"], part["owned_node_end"])
    if partition == "rcb":
        glob[nodes[own]] = sol[own, 0]
    else:
        sp = mesher.slab_partition((m, m, m), rank, world)
        glob[sp["owned_node_lo"]:sp["owned_node_hi"]] = sol[own, 0]
    glob = backend.comm_allreduce_host(glob)
    ok = True
    if rank == 0:
        from oracle import solve as osolve
        from tests import problems
        p = problems.poisson_hex(m)
        prob = osolve.Problem(p["sets"], p["coords"], p["mask"], p["values"])
        ref, (rsteps, _, rdiv) = osolve.damped_newton(prob, np.zeros(p["mask"].shape))
        err = np.linalg.norm(glob - ref.ravel()) / np.linalg.norm(ref)
        ok = err < 1e-8 and steps == rsteps and div == rdiv
        print("multi-gpu parity: ranks=%d m=%d %s %s steps=%d/%d res=%.2e rel-L2=%.2e -> %s"
              % (world, m, krylov, partition, steps, rsteps, res, err, "OK" if ok else "FAIL"))
    solver.clear_plan_cache()
    backend.comm_destroy()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
