"""Host-side logic of the multi-GPU path on CPU: world_size-2 `gloo` processes build their slab-local meshes
(bench.local_poisson_mesh / mesher.slab_partition) and the ORACLE assembles each local mesh; the owned rows of the
local tangents and residuals, mapped back to global ids, must reproduce the global assembly exactly ("ghost
element" redundancy), ghost planes must coincide with the neighbour's owned planes, and the owned
ranges must tile the node set."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, m, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    import bench
    from oracle import assemble as oasm
    from oracle import quadrature as oquad
    dist.init_process_group("gloo", rank=rank, world_size=world)
    part, coords, elems, mask, owned_elems = bench.local_poisson_mesh(m, rank, world)
    st = dict(kind="domain", etype="hex8", conn=elems.astype(np.int64), nf=1, gp=oquad.gauss_legendre_nd(3, 2),
              model=dict(name="poisson_weak", coefficient=1.0, source=1.0))
    rng = np.random.default_rng(0)
    glob_dofs = rng.uniform(-1, 1, ((m + 1) ** 3, 1))
    dofs = glob_dofs[part["node_lo"]:part["node_hi"]]
    R, data = oasm.assemble([st], coords, dofs, {})
    rows, cols = oasm.coo_indices([st])
    n_loc = coords.shape[0]
    K = oasm.scipy_assembling(data, rows, cols, n_loc)
    o0, o1 = part["owned_node_lo"] - part["node_lo"], part["owned_node_hi"] - part["node_lo"]
    Kown = K[o0:o1].tocoo()
    # owned rows in GLOBAL numbering
    trip = np.stack([Kown.row + part["owned_node_lo"], Kown.col + part["node_lo"], Kown.data], axis=1)
    payload = dict(rank=rank, part=part, trip=trip, R=R[o0:o1], n_owned_elems=owned_elems,
                   ghost_lo=coords[:o0], ghost_hi=coords[o1:], owned_first=coords[o0:o0 + part["per_plane"]],
                   owned_last=coords[o1 - part["per_plane"]:o1], mask_own=mask[o0:o1])
    gathered = [None] * world
    dist.all_gather_object(gathered, payload)
    # halo consistency through a real exchange: send my first/last owned plane ids to the neighbours
    t = torch.tensor([float(part["owned_node_lo"]), float(part["owned_node_hi"])])
    allt = [torch.zeros(2) for _ in range(world)]
    dist.all_gather(allt, t)
    if rank == 0:
        q.put((gathered, [a.tolist() for a in allt]))
    dist.barrier()
    dist.destroy_process_group()


def test_slab_partition_world_size_2_matches_global_assembly():
    import torch.multiprocessing as mp
    from oracle import assemble as oasm
    from tests import problems
    m, world, port = 6, 2, 29631
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, m, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered, ranges = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # owned ranges tile the nodes
    assert ranges[0][0] == 0 and ranges[-1][1] == (m + 1) ** 3 and ranges[0][1] == ranges[1][0]
    # ghost planes coincide with the neighbour's owned boundary planes
    assert np.array_equal(gathered[0]["ghost_hi"], gathered[1]["owned_first"])
    assert np.array_equal(gathered[1]["ghost_lo"], gathered[0]["owned_last"])
    # global assembly on one process
    p = problems.poisson_hex(m)
    rng = np.random.default_rng(0)
    glob_dofs = rng.uniform(-1, 1, ((m + 1) ** 3, 1))
    R, data = oasm.assemble(p["sets"], p["coords"], glob_dofs, {})
    rows, cols = oasm.coo_indices(p["sets"])
    n = glob_dofs.size
    K = oasm.scipy_assembling(data, rows, cols, n)
    import scipy.sparse as sp
    trip = np.concatenate([g["trip"] for g in gathered])
    Kp = sp.csr_matrix(sp.coo_matrix((trip[:, 2], (trip[:, 0].astype(int), trip[:, 1].astype(int))), shape=(n, n)))
    Kp.sort_indices()
    assert np.array_equal(Kp.indptr, K.indptr) and np.array_equal(Kp.indices, K.indices)
    assert np.abs(Kp.data - K.data).max() <= 1e-14 * np.abs(K.data).max()
    Rp = np.concatenate([g["R"] for g in gathered])
    assert np.abs(Rp - R).max() <= 1e-14 * np.abs(R).max()
    assert np.array_equal(np.concatenate([g["mask_own"] for g in gathered]), p["mask"][:, 0])


# ---- nf = 3 (BASELINE config 5 family): domain + surface set through mesher.slab_partition_mesh -------------------
def _worker_nf3(rank, world, port, m, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    from autopdex_b200 import mesher
    from oracle import assemble as oasm
    from tests import problems
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = problems.neo_hooke_brick(m)
    nf = p["nf"]
    pt = mesher.slab_partition_mesh(p["coords"], tuple(s["conn"] for s in p["sets"]), (m, m, m), rank, world)
    nodes, bp = pt["nodes"], pt["b200 partition"]
    sets = [dict(s, conn=c) for s, c in zip(p["sets"], pt["elements"])]
    rng = np.random.default_rng(0)
    glob_dofs = rng.uniform(-0.02, 0.02, (p["coords"].shape[0], nf))
    R, data = oasm.assemble(sets, p["coords"][nodes], glob_dofs[nodes], {})
    rows, cols = oasm.coo_indices(sets)
    K = oasm.scipy_assembling(data, rows, cols, nodes.size * nf)
    o0, o1 = bp["owned_node_begin"] * nf, bp["owned_node_end"] * nf
    gdof = (nodes[:, None] * nf + np.arange(nf)).ravel()
    Kown = K[o0:o1].tocoo()
    trip = np.stack([gdof[o0:o1][Kown.row], gdof[Kown.col], Kown.data], axis=1)
    # a real halo exchange of the owned interface planes (what ncclSend/ncclRecv does on the device): the ghost
    # dofs received from the neighbour must equal the neighbour's owned values
    x = np.zeros(nodes.size * nf)
    x[o0:o1] = glob_dofs.ravel()[gdof[o0:o1]]
    plane = pt["slab"]["per_plane"] * nf
    reqs = []
    xt = torch.from_numpy(x)
    lo, hi = bp["rank_lo"], bp["rank_hi"]
    if lo >= 0:
        reqs.append(dist.isend(xt[o0:o0 + plane].clone(), lo))
        reqs.append(dist.irecv(xt[:o0], lo))
    if hi >= 0:
        reqs.append(dist.isend(xt[o1 - plane:o1].clone(), hi))
        reqs.append(dist.irecv(xt[o1:], hi))
    for r in reqs:
        r.wait()
    halo_ok = bool(np.array_equal(x, glob_dofs.ravel()[gdof]))
    gathered = [None] * world
    dist.all_gather_object(gathered, dict(rank=rank, trip=trip, R=R.ravel()[o0:o1], gdof=gdof[o0:o1], halo_ok=halo_ok))
    if rank == 0:
        q.put(gathered)
    dist.barrier()
    dist.destroy_process_group()


def test_slab_partition_nf3_world_size_2_matches_global_assembly():
    import scipy.sparse as sp
    import torch.multiprocessing as mp
    from oracle import assemble as oasm
    from tests import problems
    m, world, port = 4, 2, 29671
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_nf3, args=(r, world, port, m, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(g["halo_ok"] for g in gathered)
    p = problems.neo_hooke_brick(m)
    n = p["coords"].shape[0] * p["nf"]
    assert np.array_equal(np.concatenate([g["gdof"] for g in gathered]), np.arange(n))      # owned dofs tile the system
    rng = np.random.default_rng(0)
    glob_dofs = rng.uniform(-0.02, 0.02, (p["coords"].shape[0], p["nf"]))
    R, data = oasm.assemble(p["sets"], p["coords"], glob_dofs, {})
    rows, cols = oasm.coo_indices(p["sets"])
    K = oasm.scipy_assembling(data, rows, cols, n)
    trip = np.concatenate([g["trip"] for g in gathered])
    Kp = sp.csr_matrix(sp.coo_matrix((trip[:, 2], (trip[:, 0].astype(int), trip[:, 1].astype(int))), shape=(n, n)))
    Kp.sort_indices()
    assert np.array_equal(Kp.indptr, K.indptr) and np.array_equal(Kp.indices, K.indices)
    assert np.abs(Kp.data - K.data).max() <= 1e-13 * np.abs(K.data).max()
    Rp = np.concatenate([g["R"] for g in gathered])
    assert np.abs(Rp - R.ravel()).max() <= 1e-13 * np.abs(R).max()
