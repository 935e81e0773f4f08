"""Host half of the partitioned multigrid hierarchy (autopdex_b200/multigrid.py, slab partitions): the local transfer
operators of every slab must be the owned rows of the global ones, and need nothing beyond the slab's own planes.
(The device half runs on emulated ranks in tests/test_emu_cuda_source.py and on hardware through tests/multi_gpu_worker.py.)"""
import numpy as np
import pytest
import scipy.sparse as sp

from autopdex_b200 import mesher, multigrid as mg


def _reduced(free, lo, hi):
    """reduced ids of the free dofs inside the full-dof range [lo, hi)"""
    red = np.cumsum(free.ravel()) - 1
    sel = free.ravel()[lo:hi]
    return red[lo:hi][sel]


@pytest.mark.parametrize("shape,nf", [((8, 4, 4), 1), ((12, 4), 2), ((16, 4, 6), 3)])
def test_slab_transfer_operators_are_the_owned_rows_of_the_global_ones(shape, nf):
    shape_c = tuple(n // 2 for n in shape)
    rng = np.random.default_rng(1)
    free_f = rng.random((mg.node_count(shape), nf)) > 0.15
    free_c = free_f[mg.fine_node_ids(shape)]
    (pp, pi, pd), _ = mg.prolongation(shape, nf, free_f, free_c)
    Pg = sp.csr_matrix((pd, pi, pp), shape=(free_f.sum(), free_c.sum())).toarray()
    per_f = int(np.prod([n + 1 for n in shape[1:]])) * nf
    per_c = int(np.prod([n + 1 for n in shape_c[1:]])) * nf
    checked = 0
    for world in (2, 3, 4):
        for r in range(world):
            part = mesher.slab_partition(shape, r, world)
            planes = (part["plane_lo"], part["plane_hi"], part["owned_plane_lo"], part["owned_plane_hi"])
            lv = mg.slab_levels(shape, planes, min_owned=1)
            if len(lv) < 2:
                continue
            g0, g1, p0, p1 = planes
            G0, G1, P0, P1 = lv[1][1]
            assert (P0, P1) == ((p0 + 1) // 2, (p1 + 1) // 2) and G0 == max(P0 - 1, 0) and G1 == min(P1 + 1, shape_c[0] + 1)
            ff, fc = free_f.ravel()[g0 * per_f:g1 * per_f].reshape(-1, nf), free_c.ravel()[G0 * per_c:G1 * per_c].reshape(-1, nf)
            (a, b, c), (ra, rb, rc) = mg.prolongation(shape, nf, ff, fc, slab=((g0, g1), (G0, G1)))
            Pl = sp.csr_matrix((c, b, a), shape=(ff.sum(), fc.sum())).toarray()
            Rl = sp.csr_matrix((rc, rb, ra), shape=(fc.sum(), ff.sum())).toarray()
            rows_g, cols_g = _reduced(free_f, p0 * per_f, p1 * per_f), _reduced(free_c, G0 * per_c, G1 * per_c)
            rows_l = _reduced(ff, (p0 - g0) * per_f, (p1 - g0) * per_f)
            # prolongation: owned fine rows, local coarse columns only
            assert np.array_equal(Pg[rows_g][:, cols_g], Pl[rows_l])
            assert np.abs(np.delete(Pg[rows_g], cols_g, axis=1)).max(initial=0) == 0
            # restriction R = P^T: owned coarse rows need the local fine planes (owned + one ghost plane) and nothing else
            crow_g, fcol_g = _reduced(free_c, P0 * per_c, P1 * per_c), _reduced(free_f, g0 * per_f, g1 * per_f)
            crow_l = _reduced(fc, (P0 - G0) * per_c, (P1 - G0) * per_c)
            assert np.abs(np.delete(Pg.T[crow_g], fcol_g, axis=1)).max(initial=0) == 0
            assert np.array_equal(Rl[crow_l], Pg.T[crow_g][:, fcol_g])
            checked += 1
    assert checked >= 4


def test_slab_levels_end_where_a_rank_would_own_fewer_than_two_planes():
    # 256^3 on 8 ranks: 257 planes -> 32-33 per rank -> 16 -> 8 -> 4 -> 2 -> (1: stop): five levels, coarsest 16^3
    lens = []
    for r in range(8):
        part = mesher.slab_partition((256, 256, 256), r, 8)
        lv = mg.slab_levels((256, 256, 256), (part["plane_lo"], part["plane_hi"], part["owned_plane_lo"], part["owned_plane_hi"]))
        lens.append(len(lv))
        for (shape, (G0, G1, P0, P1)) in lv:
            assert P1 - P0 >= 2 and G0 <= P0 and P1 <= G1 <= shape[0] + 1
    assert min(lens) == 5
    # owned coarse planes of neighbouring ranks tile the coarse mesh
    for level in range(1, 5):
        owned = []
        for r in range(8):
            part = mesher.slab_partition((256, 256, 256), r, 8)
            lv = mg.slab_levels((256, 256, 256), (part["plane_lo"], part["plane_hi"], part["owned_plane_lo"], part["owned_plane_hi"]))
            owned.append(lv[level][1][2:])
        assert owned[0][0] == 0 and owned[-1][1] == (256 >> level) + 1
        assert all(owned[i][1] == owned[i + 1][0] for i in range(7))


def test_injection_map_points_ghost_planes_at_local_fine_planes():
    shape = (8, 4, 4)
    part = mesher.slab_partition(shape, 1, 2)            # owned fine planes [4, 9), ghost plane 3
    planes_f = (part["plane_lo"], part["plane_hi"], part["owned_plane_lo"], part["owned_plane_hi"])
    planes_c = mg.slab_levels(shape, planes_f)[1][1]      # owned coarse planes [2, 5), ghost coarse plane 1 = fine plane 2: not local
    assert planes_f == (3, 9, 4, 9) and planes_c == (1, 5, 2, 5)
    ids = mg.fine_node_ids_slab(shape, planes_f, planes_c)
    per_f, per_c = 25, 9
    assert ids.min() >= 0 and ids.max() < (planes_f[1] - planes_f[0]) * per_f and ids.size == 4 * per_c
    assert np.array_equal(ids[per_c:2 * per_c] // per_f, np.full(per_c, 1))   # coarse plane 2 = fine plane 4 = local fine plane 1
    assert np.array_equal(ids[:per_c] // per_f, np.zeros(per_c))              # the non-local ghost plane is clamped (exchanged afterwards)
    assert np.array_equal(ids[per_c:2 * per_c] % per_f, [0, 2, 4, 10, 12, 14, 20, 22, 24])   # every other node inside the plane
