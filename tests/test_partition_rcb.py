"""Host logic of the general (recursive-coordinate-bisection) partition on CPU: mesher.rcb_partition against the ORACLE.
For every part the owned rows of the locally assembled tangent / residual, mapped back to global ids, must reproduce
the global assembly ("ghost element" redundancy); the owned sets must tile the nodes; and the halo lists
(send_nodes of the owner == ghost block of the receiver, in the receiver's order) must deliver exactly the owners'
values.  One world_size-2 `gloo` run performs the exchange through torch.distributed send/recv; the other cases run
all parts in one process."""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

from autopdex_b200 import mesher
from oracle import assemble as oasm
from tests import problems

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _problem(name):
    if name == "poisson_hex":
        return problems.poisson_hex(5, distort=0.2)
    if name == "neo_hooke_brick":                      # nf = 3, domain + surface set
        return problems.neo_hooke_brick(3)
    if name == "elasticity_quad9":                     # nf = 2, Q2
        return problems.elasticity_quad(4, "plain strain", "quad9")
    if name == "shuffled_quad":                        # node numbering without any structure
        p = problems.readme_poisson(7)
        perm = np.random.default_rng(1).permutation(p["coords"].shape[0])
        inv = np.argsort(perm)
        p = dict(p)
        p["coords"] = p["coords"][perm]
        p["mask"], p["values"] = p["mask"][perm], p["values"][perm]
        p["sets"] = [dict(s, conn=inv[s["conn"]]) for s in p["sets"]]
        return p
    raise KeyError(name)


def _local_sets(p, part):
    out = []
    for st, conn, ids in zip(p["sets"], part["elements"], part["element_ids"]):
        st = dict(st, conn=conn)
        model = dict(st["model"])
        for k, v in model.items():                     # per-element parameter arrays follow their elements
            if isinstance(v, np.ndarray) and v.ndim >= 1 and v.shape[0] == p["sets"][len(out)]["conn"].shape[0] and v.ndim > 1:
                model[k] = v[ids]
        st["model"] = model
        out.append(st)
    return out


def _global_assembly(p, dofs):
    R, data = oasm.assemble(p["sets"], p["coords"], dofs, {})
    rows, cols = oasm.coo_indices(p["sets"])
    return R, oasm.scipy_assembling(data, rows, cols, dofs.size)


def _owned_rows_in_global_ids(p, part, dofs_global):
    nf = p["nf"]
    nodes = part["nodes"]
    sets = _local_sets(p, part)
    R, data = oasm.assemble(sets, p["coords"][nodes], dofs_global[nodes], {})
    rows, cols = oasm.coo_indices(sets)
    n_loc = nodes.size * nf
    K = oasm.scipy_assembling(data, rows, cols, n_loc)
    n_own = part["n_owned"] * nf
    gdof = (nodes[:, None] * nf + np.arange(nf)).ravel()
    Kown = K[:n_own].tocoo()
    return gdof[:n_own], np.stack([gdof[Kown.row], gdof[Kown.col], Kown.data], axis=1), R.ravel()[:n_own]


@pytest.mark.parametrize("name,nranks", [("poisson_hex", 2), ("poisson_hex", 3), ("poisson_hex", 8), ("neo_hooke_brick", 4),
                                         ("elasticity_quad9", 4), ("shuffled_quad", 5)])
def test_rcb_parts_reproduce_the_global_assembly_and_halo(name, nranks):
    p = _problem(name)
    nf, n_nodes = p["nf"], p["coords"].shape[0]
    conns = tuple(s["conn"] for s in p["sets"])
    rng = np.random.default_rng(0)
    dofs = rng.uniform(-0.02, 0.02, (n_nodes, nf))
    Rg, Kg = _global_assembly(p, dofs)
    owner = mesher.rcb_owner(p["coords"], nranks)
    counts = np.bincount(owner, minlength=nranks)
    assert counts.sum() == n_nodes and counts.max() - counts.min() <= 1 + n_nodes % 2 + nranks      # balanced leaves
    parts = [mesher.rcb_partition(p["coords"], conns, r, nranks) for r in range(nranks)]
    assert all(np.array_equal(pt["owner"], owner) for pt in parts)                                 # deterministic
    # owned node sets tile the mesh
    allowned = np.concatenate([pt["nodes"][:pt["n_owned"]] for pt in parts])
    assert np.array_equal(np.sort(allowned), np.arange(n_nodes))
    # owned rows == global rows
    trips, Rp = [], np.zeros(n_nodes * nf)
    for pt in parts:
        gd, trip, R = _owned_rows_in_global_ids(p, pt, dofs)
        trips.append(trip)
        Rp[gd] = R
    trip = np.concatenate(trips)
    n = n_nodes * nf
    Kp = sp.csr_matrix(sp.coo_matrix((trip[:, 2], (trip[:, 0].astype(int), trip[:, 1].astype(int))), shape=(n, n)))
    Kp.sort_indices()
    assert np.array_equal(Kp.indptr, Kg.indptr) and np.array_equal(Kp.indices, Kg.indices)
    assert np.abs(Kp.data - Kg.data).max() <= 1e-13 * np.abs(Kg.data).max()
    assert np.abs(Rp - Rg.ravel()).max() <= 1e-13 * max(np.abs(Rg).max(), 1e-300)
    # halo lists: what q packs for r is r's ghost block owned by q, entry by entry
    x = rng.standard_normal(n_nodes)
    for r, pt in enumerate(parts):
        assert pt["b200 partition"]["owned_node_end"] == pt["n_owned"]
        covered = pt["n_owned"]
        for q, (b, e) in zip(pt["neighbours"], pt["recv_node_ranges"]):
            assert b == covered and e > b                     # ghost blocks are contiguous, in neighbour order
            covered = e
            assert (owner[pt["nodes"][b:e]] == q).all()
            other = parts[q]
            assert r in other["neighbours"]                   # neighbourhood is symmetric
            send = other["send_nodes"][other["neighbours"].index(r)]
            assert (send < other["n_owned"]).all()
            packed = x[other["nodes"]][send]                  # k_halo_pack on rank q
            assert np.array_equal(packed, x[pt["nodes"][b:e]])
        assert covered == pt["nodes"].size


def _gloo_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    from autopdex_b200 import mesher as m2
    from tests import problems as pr
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = pr.poisson_hex(5, distort=0.2)
    pt = m2.rcb_partition(p["coords"], (p["sets"][0]["conn"],), rank, world)
    xg = np.random.default_rng(7).standard_normal(p["coords"].shape[0])
    x = np.full(pt["nodes"].size, np.nan)
    x[:pt["n_owned"]] = xg[pt["nodes"][:pt["n_owned"]]]          # every rank knows its owned values only
    reqs, bufs = [], []
    for nb, send, (b, e) in zip(pt["neighbours"], pt["send_nodes"], pt["recv_node_ranges"]):
        out = torch.from_numpy(np.ascontiguousarray(x[send]))    # pack
        buf = torch.empty(e - b, dtype=torch.float64)
        reqs += [dist.isend(out, nb), dist.irecv(buf, nb)]
        bufs.append((b, e, buf, out))
    for rq in reqs:
        rq.wait()
    for b, e, buf, _ in bufs:
        x[b:e] = buf.numpy()
    ok = bool(np.array_equal(x, xg[pt["nodes"]]))
    res = [None] * world
    dist.all_gather_object(res, ok)
    if rank == 0:
        q.put(res)
    dist.barrier()
    dist.destroy_process_group()


def test_rcb_halo_exchange_world_size_2_gloo():
    import torch.multiprocessing as mp
    world, port = 2, 29641
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [True, True]


def _reduced_halo_lists(part, mask_local, nf):
    """Python restatement of the index bookkeeping of apdx_plan_set_partition_lists (csrc/api.cu): local dof lists ->
    lists in the reduced (Dirichlet-free) numbering.  Returns (f1, send_idx per neighbour, (recv_begin, count))."""
    m = np.asarray(mask_local).ravel()
    fid = np.where(m, -1, np.cumsum(~m) - 1)                       # free_id: reduced index or -1
    n_free = int((~m).sum())

    def lower(dof):
        nxt = np.flatnonzero(~m[dof:])
        return int(fid[dof + nxt[0]]) if nxt.size else n_free
    dofs_of = lambda nodes: (np.asarray(nodes)[:, None] * nf + np.arange(nf)).ravel()
    send = [fid[dofs_of(v)][fid[dofs_of(v)] >= 0] for v in part["send_nodes"]]
    recv = [(lower(b * nf), lower(e * nf) - lower(b * nf)) for b, e in part["recv_node_ranges"]]
    return lower(part["n_owned"] * nf), send, recv


@pytest.mark.parametrize("name,nranks", [("neo_hooke_brick", 4), ("poisson_hex", 3)])
def test_rcb_halo_in_reduced_numbering_with_dirichlet_dofs(name, nranks):
    """With Dirichlet dofs removed on both sides (the device works in the reduced numbering) the packed send lists
    still line up with the receivers' ghost blocks, and the message lengths agree pairwise (what comm_halo_setup_lists
    cross-checks at set-up)."""
    p = _problem(name)
    nf = p["nf"]
    conns = tuple(s["conn"] for s in p["sets"])
    mask = np.asarray(p["mask"]).reshape(-1, nf).copy()
    mask[::7, 0] = True                                            # a few more constrained dofs, partially constrained nodes
    parts = [mesher.rcb_partition(p["coords"], conns, r, nranks) for r in range(nranks)]
    xg = np.random.default_rng(3).standard_normal(mask.shape)      # a global dof field
    red = []
    for pt in parts:
        ml = mask[pt["nodes"]]
        f1, send, recv = _reduced_halo_lists(pt, ml, nf)
        xl = xg[pt["nodes"]].ravel()[~ml.ravel()]                  # local reduced vector [owned | ghosts]
        assert f1 == int((~ml[:pt["n_owned"]]).sum())
        red.append((send, recv, xl))
    for r, pt in enumerate(parts):
        send_r, recv_r, x_r = red[r]
        for i, q in enumerate(pt["neighbours"]):
            send_q, _, x_q = red[q]
            j = parts[q]["neighbours"].index(r)
            b, cnt = recv_r[i]
            assert cnt == send_q[j].size                           # pairwise message lengths agree
            assert np.array_equal(x_q[send_q[j]], x_r[b:b + cnt])  # k_halo_pack on q == ghost block on r
