"""CPU-side tests: the C-ABI library loads and exports every declared symbol, host logic of the
reference-facing layer (settings validation / rejection, model recognition, callable evaluation,
mesher, quadrature, shape tables) and the loud failure without a GPU.  No compute calls."""
import ctypes
import os
import re

import numpy as np
import pytest

from autopdex_b200 import _lib, backend, mesher, models, seeder, solver, spaces, utility
from oracle import mesher as omesh
from oracle import quadrature as oquad
from oracle import shapes as oshapes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    txt = open(os.path.join(ROOT, "include", "apdx_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(apdx_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    names = _header_functions()
    assert len(names) >= 30
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "libapdx_b200.so does not export %s" % n
    assert sorted(_lib.SIGNATURES) == names          # the ctypes table mirrors the header one to one
    assert _lib.load().apdx_abi_version() == 1


def test_no_gpu_fails_loudly():
    if _lib.device_count() > 0:
        pytest.skip("a GPU is present")
    conn = np.array([[0, 1, 2, 3]])
    st = backend.SetSpec("domain", "poisson_weak", conn, family="quad_brick", gp=seeder.gauss_legendre_nd(2, 2))
    with pytest.raises(RuntimeError, match="no CPU path"):
        backend.Plan(2, 4, 1, [st], None)


# ---- settings validation: reject at read time, never fall back ------------------------------------
def _static(**over):
    weak = models.poisson_weak()
    elem = models.isoparametric_domain_element_galerkin(weak, spaces.fem_iso_line_quad_brick,
                                                        *seeder.gauss_legendre_nd(2, 2))
    s = {"assembling mode": ("user element",), "solution structure": ("nodal imposition",), "model": (elem,),
         "solver type": "newton", "solver backend": "b200", "solver": "cg", "type of preconditioner": "jacobi",
         "verbose": -1}
    s.update(over)
    return s


def test_validate_accepts_builtin_configuration():
    cfg = solver.validate(_static())
    assert cfg.nodal_imposition and cfg.jacobi and cfg.krylov == "cg"


@pytest.mark.parametrize("over,msg", [
    ({"solver type": "minimize"}, "solver type"),
    ({"solver type": "diagonal linear"}, "solver type"),
    ({"solver": "gmres"}, "'solver'"),
    ({"solver": "lapack"}, "'solver'"),
    ({"type of preconditioner": "ilu"}, "preconditioner"),
    ({"assembling mode": ("dense",)}, "assembling mode"),
    ({"solution structure": ("first order set",)}, "solution structure"),
    ({"known sparsity pattern": "diagonal"}, "sparsity"),
    ({"solver backend": "scipy"}, "b200"),
])
def test_validate_rejects_unsupported(over, msg):
    with pytest.raises(ValueError, match=msg):
        solver.validate(_static(**over))


def test_user_written_models_are_rejected():
    def my_integrand(x_int, ansatz_fun, settings, static_settings, elem_number, set):
        return 0.0
    with pytest.raises(ValueError, match="user-written integrand"):
        models.mixed_reference_domain_potential(my_integrand, {"phi": spaces.fem_iso_line_quad_brick},
                                                *seeder.gauss_legendre_nd(2, 2), "phi")
    with pytest.raises(ValueError, match="not supported"):
        solver.validate(_static(model=(lambda *a: 0.0,)))
    with pytest.raises(ValueError, match="strain energy"):
        models.hyperelastic_steady_state_weak(lambda F, p: 0.0, lambda x: 1.0, lambda x: 0.3, "3d")
    with pytest.raises(ValueError, match="sparse"):
        solver.validate(_static(**{"assembling mode": ("sparse",), "model": (models.poisson_weak(),),
                                   "variational scheme": ("least square pde loss",), "solution space": ("mls",),
                                   "shape function mode": "direct"}))


def test_reference_closures_are_recognised_by_qualname():
    """The reference's models are anonymous closures (models.py); the backend identifies them by
    __qualname__ and closure cells without calling them.  Mimic their shape here."""
    def fem_iso_line_quad_brick(x, xI, fI, settings, overwrite_diff, n_dim):
        raise AssertionError("must not be called")

    def neo_hooke(F, param):
        raise AssertionError("must not be called")

    def hyperelastic_steady_state_weak(strain_energy_fun, youngs_mod_fun, poisson_ratio_fun, mode, volume_load_fun=None):
        def pde_fun(x, ansatz, test_ansatz, settings, static_settings, int_point_number, set):
            return strain_energy_fun, youngs_mod_fun, poisson_ratio_fun, mode, volume_load_fun
        return pde_fun

    def isoparametric_domain_element_galerkin(weak_form_fun, ansatz_fun, ref_int_coor, ref_int_weights, initial_config=True):
        def user_element(fI, xI, elem_number, settings, static_settings, mode, set):
            return weak_form_fun, ansatz_fun, ref_int_coor, ref_int_weights, initial_config
        return user_element

    E = lambda x, settings: settings["youngs modulus"]
    nu = lambda x, settings: settings["poisson ratio"]
    ref_elem = isoparametric_domain_element_galerkin(hyperelastic_steady_state_weak(neo_hooke, E, nu, "plain strain"),
                                                     fem_iso_line_quad_brick, *seeder.gauss_legendre_nd(2, 4))
    m = models.recognise(ref_elem)
    assert isinstance(m, models.ElementModel) and m.kind == "domain" and m.family == "quad_brick"
    assert m.weak.name == "neo_hooke" and m.weak.mode == "plain strain"
    assert m.weak.funs["youngs_modulus"] is E and len(m.gp[1]) == 9
    cfg = solver.validate(_static(model=(ref_elem,), solver="bicgstab"))
    assert cfg.sets[0][1].weak.name == "neo_hooke"

    def linear_elastic_strain_energy(F, param):        # the reference's small-strain energy (models.py:1167-1185)
        raise AssertionError("must not be called")
    lin = models.recognise(isoparametric_domain_element_galerkin(
        hyperelastic_steady_state_weak(linear_elastic_strain_energy, E, nu, "plain strain"), fem_iso_line_quad_brick,
        *seeder.gauss_legendre_nd(2, 2)))
    assert lin.weak.name == "linear_elasticity" and lin.weak.mode == "lame"       # the isotropic tensor, not linear_elasticity_weak's

    def isochoric_neo_hooke(F, mu):
        raise AssertionError("must not be called")
    with pytest.raises(ValueError, match="strain energy"):
        models.recognise(isoparametric_domain_element_galerkin(hyperelastic_steady_state_weak(isochoric_neo_hooke, E, nu, "3d"),
                                                               fem_iso_line_quad_brick, *seeder.gauss_legendre_nd(3, 2)))


def test_callable_evaluation_reference_vs_physical_points():
    pts = np.random.default_rng(0).uniform(size=(100, 2))
    f = lambda x: 20 * (np.sin(10 * np.sum(x * x, axis=-1)))
    v = solver._eval_points(f, pts, {}, 1)
    assert v.shape == (100,) and np.allclose(v[7], f(pts[7]))
    g = lambda x: 20 * np.sin(10 * x @ x)            # NOT vectorisable: falls back to the per-point loop
    assert np.allclose(solver._eval_points(g, pts, {}, 1), [g(p) for p in pts])
    t = lambda x, settings: np.asarray([0.0, settings["load multiplier"]])
    assert np.array_equal(solver._eval_points(t, pts[:3], {"load multiplier": 4.0}, 2), [[0, 4.0]] * 3)
    assert solver._eval_points(lambda x: 1.0, pts, {}, 1).shape == ()
    with pytest.raises(ValueError, match="vectorized"):
        solver._eval_points(g, pts, {}, 1, vectorized=True)


def test_callables_that_reduce_over_the_point_axis_are_evaluated_per_point():
    """The reference calls coefficient callables per point: a callable that reduces over the batch axis (prod / sum /
    norm) or indexes it must not be taken for a constant or a point-wise batch result."""
    pts = np.random.default_rng(2).uniform(0.1, 0.9, (200, 2))
    for f in (lambda x: np.prod(np.sin(np.pi * x)), lambda x: np.sum(x * x), lambda x: np.linalg.norm(x)):
        v = solver._eval_points(f, pts, {}, 1)
        assert v.shape == (200,) and np.allclose(v, [f(p) for p in pts], rtol=1e-14)
    h = lambda x: x[0] * np.array([1.0, 0.0])          # dim == ncomp: batch call returns the value at point 0, shape (2,)
    v = solver._eval_points(h, pts, {}, 2)
    assert v.shape == (200, 2) and np.allclose(v, [h(p) for p in pts])
    k = lambda x: np.cumsum(np.atleast_2d(x), axis=0)[..., 0].squeeze() if np.ndim(x) > 1 else x[0]   # position-dependent
    v = solver._eval_points(k, pts, {}, 1)
    assert np.allclose(v, pts[:, 0])
    assert solver._eval_points(lambda x, s: np.array([0.0, s["q"]]), pts, {"q": 3.0}, 2).shape == (2,)   # true constant


# ---- host restatements agree with the oracle's independent ones -------------------------------------
@pytest.mark.parametrize("family,dim,nen,name", [
    ("quad_brick", 1, 2, "line2"), ("quad_brick", 1, 3, "line3"), ("quad_brick", 2, 4, "quad4"),
    ("quad_brick", 2, 9, "quad9"), ("quad_brick", 3, 8, "hex8"), ("quad_brick", 3, 27, "hex27"),
    ("tri_tet", 2, 3, "tri3"), ("tri_tet", 2, 6, "tri6"), ("tri_tet", 3, 4, "tet4"), ("tri_tet", 3, 10, "tet10")])
def test_shape_tables_match_oracle(family, dim, nen, name):
    xi = np.random.default_rng(1).uniform(0.0, 0.3, (9, dim))
    N, dN = spaces.shape_tables(family, nen, dim, xi)
    No, dNo = oshapes.shape_tables(name, xi)
    assert np.abs(N - No).max() < 1e-14 and np.abs(dN - dNo).max() < 1e-13
    assert np.allclose(N.sum(axis=1), 1.0) and np.abs(dN.sum(axis=1)).max() < 1e-13
    nodes = np.asarray(oshapes.REF_NODES[name], dtype=float)
    assert np.allclose(spaces.shape_tables(family, nen, dim, nodes)[0], np.eye(nen), atol=1e-14)


def test_quadrature_and_mesher_match_oracle():
    for d in (1, 2, 3):
        for o in (1, 2, 3, 4, 6):
            a, b = seeder.gauss_legendre_nd(d, o), oquad.gauss_legendre_nd(d, o)
            assert np.abs(a[0] - b[0]).max() < 1e-15 and np.abs(a[1] - b[1]).max() < 1e-15
            assert np.isclose(a[1].sum(), 2.0 ** d)
    cube = [[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1.2], [1, 1, 1], [0, 1, 1]]
    quad = [[0, 0], [2, 0], [2.5, 1.5], [0, 1]]
    for args in [((3, 4), quad, "quad"), ((3, 4), quad, "tri"), ((2, 3, 4), cube, "brick"), ((2, 3, 2), cube, "tet")]:
        a, b = mesher.structured_mesh(*args), omesh.structured_mesh(*args)
        assert np.array_equal(a[1], b[1]) and np.abs(a[0] - b[0]).max() < 1e-15
    c, e = mesher.structured_mesh((2, 2, 2), cube, "brick")
    a, b = mesher.elevate_mesh_order(c, e), omesh.elevate_bricks(c, e)
    assert np.array_equal(a[1], b[1]) and np.allclose(a[0], b[0]) and a[1].shape[1] == 27
    # reference quirk: mesher quads are clockwise w.r.t. quad4 -> det J < 0
    c, e = mesher.structured_mesh((2, 2), [[0, 0], [1, 0], [1, 1], [0, 1]], "quad")
    _, dN = spaces.shape_tables("quad_brick", 4, 2, np.zeros((1, 2)))
    J = np.einsum("ad,ak->dk", c[e[0]], dN[0])
    assert np.linalg.det(J) < 0


def test_utility_and_slab_partition():
    d = {"phi": np.arange(6.0).reshape(3, 2)}
    flat = utility.dict_flatten(d)
    assert np.array_equal(utility.reshape_as(flat, d)["phi"], d["phi"])
    assert utility.dof_select(np.array([True, False]), np.array([True, True])).shape == (2, 2)
    m, world = 9, 4
    owned = []
    for r in range(world):
        p = mesher.slab_partition((m, m, m), r, world)
        owned.append((p["owned_node_lo"], p["owned_node_hi"]))
        assert p["node_lo"] <= p["owned_node_lo"] < p["owned_node_hi"] <= p["node_hi"]
        assert (p["rank_lo"] == -1) == (r == 0) and (p["rank_hi"] == -1) == (r == world - 1)
    assert owned[0][0] == 0 and owned[-1][1] == (m + 1) ** 3
    assert all(owned[i][1] == owned[i + 1][0] for i in range(world - 1))


def test_host_result_blocks_are_recycled_only_after_the_last_view_died():
    """backend.host_result: large device->host results land in recycled (page-locked on a GPU box) blocks; a block must
    not be reused while any array derived from the result is alive."""
    import gc
    from autopdex_b200 import backend
    n = 1 << 18
    a = backend.host_result(n)
    a[:] = 3.0
    ptr = a.ctypes.data
    view = a.reshape(-1, 1)[5:9]
    del a
    gc.collect()
    b = backend.host_result(n)
    assert b.ctypes.data != ptr and np.all(view == 3.0)
    del view
    gc.collect()
    c = backend.host_result(n)
    assert c.ctypes.data == ptr
    assert backend.host_result(16).base is None      # small results are plain arrays


def _read_legacy_vtk(path):
    """Minimal reader of the binary legacy files utility.write_vtk produces (test helper)."""
    raw = open(path, "rb").read()
    out, pos = {}, 0

    def line():
        nonlocal pos
        end = raw.index(b"\n", pos)
        s = raw[pos:end].decode()
        pos = end + 1
        return s

    def block(dtype, count):
        nonlocal pos
        a = np.frombuffer(raw, dtype=dtype, count=count, offset=pos)
        pos += a.nbytes + 1
        return a
    assert line().startswith("# vtk DataFile") and line() and line() == "BINARY" and line() == "DATASET UNSTRUCTURED_GRID"
    n = int(line().split()[1])
    out["points"] = block(">f8", 3 * n).reshape(n, 3)
    ne, tot = map(int, line().split()[1:])
    out["cells"] = block(">i4", tot).reshape(ne, -1)
    assert int(line().split()[1]) == ne
    out["types"] = block(">i4", ne)
    while pos < len(raw):
        w = line().split()
        if w[0] == "POINT_DATA":
            continue
        if w[0] == "VECTORS":
            out[w[1]] = block(">f8", 3 * n).reshape(n, 3)
        elif w[0] == "SCALARS":
            assert line() == "LOOKUP_TABLE default"
            out[w[1]] = block(">f8", n)
    return out


def test_write_vtk_round_trip(tmp_path):
    from autopdex_b200 import mesher, utility
    coords, elems = mesher.structured_mesh((2, 3, 2), [[0., 0., 0.], [1., 0., 0.], [1., 1., 0.], [0., 1., 0.], [0., 0., 1.],
                                                       [1., 0., 1.], [1., 1., 1.], [0., 1., 1.]], "brick")
    rng = np.random.default_rng(0)
    u, theta, s6 = rng.normal(size=(coords.shape[0], 3)), rng.normal(size=coords.shape[0]), rng.normal(size=(coords.shape[0], 6))
    utility.write_vtk(tmp_path / "brick.vtk", coords, elems, {"u": u, "theta": theta, "stress": s6})
    got = _read_legacy_vtk(tmp_path / "brick.vtk")
    assert np.array_equal(got["points"], coords) and np.array_equal(got["cells"][:, 1:], elems)
    assert (got["cells"][:, 0] == 8).all() and (got["types"] == 12).all()
    assert np.array_equal(got["u"], u) and np.array_equal(got["theta"], theta) and np.array_equal(got["stress_4"], s6[:, 4])
    # 2-D mesh, quad9 after elevation; surface set of a 3-D problem; rejections
    c2, e2 = mesher.structured_mesh((2, 2), [[0., 0.], [1., 0.], [1., 1.], [0., 1.]], "quad")
    c9, e9 = mesher.elevate_quads(c2, e2)
    utility.write_vtk(tmp_path / "q9.vtk", c9, e9, {"u": np.zeros((c9.shape[0], 2))})
    got = _read_legacy_vtk(tmp_path / "q9.vtk")
    assert (got["types"] == 28).all() and got["points"][:, 2].max() == 0.0 and got["u"].shape == (c9.shape[0], 3)
    utility.write_vtk(tmp_path / "face.vtk", coords, mesher.boundary_faces((2, 3, 2), 0, 1), cell_dim=2)
    assert (_read_legacy_vtk(tmp_path / "face.vtk")["types"] == 9).all()
    with pytest.raises(ValueError):
        utility.write_vtk(tmp_path / "bad.vtk", coords, elems[:, :5])
    with pytest.raises(ValueError):
        utility.write_vtk(tmp_path / "bad.vtk", coords, elems + coords.shape[0])


def test_jax_ffi_shim_compiles_and_binds_only_declared_entry_points():
    """csrc/xla/apdx_b200_xla.cc (the jax.ffi handlers of INTEGRATION.md) cannot be built against jaxlib here; a
    syntax-only stand-in of xla/ffi/api/ffi.h keeps it compiling and checks every handler against its Bind() chain.
    jax_ffi.py must refuse to import without JAX instead of providing another path."""
    import importlib
    import shutil
    import subprocess
    src = os.path.join(ROOT, "autopdex_b200", "csrc", "xla", "apdx_b200_xla.cc")
    cuda_inc = "/usr/local/cuda/include"
    if shutil.which("g++") is None or not os.path.exists(os.path.join(cuda_inc, "cuda_runtime_api.h")):
        pytest.skip("g++ / CUDA headers not available")
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-I", os.path.join(ROOT, "tests", "stubs"),
                        "-I", os.path.join(ROOT, "include"), "-I", cuda_inc, src], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    text = open(src).read()
    declared = set(_header_functions())
    import re
    used = set(re.findall(r"\b(apdx_[a-z0-9_]+)\(", text)) - {"apdx_krylov_opts"}
    used = {u for u in used if not u.endswith("_ffi")}
    assert used and used <= declared, used - declared
    try:
        import jax  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError, match="jax.ffi"):
            importlib.import_module("autopdex_b200.jax_ffi")


def test_csr_diagonal_reads_the_diagonal_of_a_canonical_csr():
    import scipy.sparse as sp
    from autopdex_b200 import assembler
    rng = np.random.default_rng(2)
    A = sp.random(40, 40, density=0.15, random_state=3, format="csr")
    A = A + sp.diags(rng.standard_normal(40) * (rng.uniform(size=40) > 0.3))     # some diagonal entries are absent
    A = sp.csr_matrix(A)
    A.sort_indices()
    got = assembler.csr_diagonal(assembler.CSR(A.data, A.indices.astype(np.int64), A.indptr.astype(np.int64), A.shape))
    assert np.array_equal(got, A.diagonal())


def test_get_indices_array_and_dict_block_order():
    """assembler._get_indices (assembler.py:41-141): array dofs against the oracle's COO order; dict dofs (two fields,
    different dofs per node and different connectivities) against a literal transcription of the reference loop."""
    from autopdex_b200 import assembler
    from oracle import assemble as oasm
    rng = np.random.default_rng(5)
    conn = rng.integers(0, 11, size=(7, 4))
    idx = assembler._get_indices(conn, np.zeros((11, 2)))
    r, c = oasm.coo_indices([{"conn": conn, "nf": 2}])
    assert idx.dtype == np.int64 and np.array_equal(idx[:, 0], r) and np.array_equal(idx[:, 1], c)
    idx1 = assembler._get_indices(conn, np.zeros(11))
    r, c = oasm.coo_indices([{"conn": conn, "nf": 1}])
    assert np.array_equal(idx1[:, 0], r) and np.array_equal(idx1[:, 1], c)
    # dict: fields u (11 nodes x 2) and p (5 nodes, scalar) with their own connectivities
    dofs = {"u": np.zeros((11, 2)), "p": np.zeros(5)}
    cn = {"u": conn, "p": rng.integers(0, 5, size=(7, 3))}
    got = assembler._get_indices(cn, dofs)
    off = {"u": 0, "p": 22}
    nfd = {"u": 2, "p": 1}
    want = []
    for fi in dofs:                      # assembler.py:79-80
        for fj in dofs:
            for e in range(7):           # vmap over elements, assembler.py:116-117
                gi = (off[fi] + cn[fi][e][:, None] * nfd[fi] + np.arange(nfd[fi])).ravel()
                gj = (off[fj] + cn[fj][e][:, None] * nfd[fj] + np.arange(nfd[fj])).ravel()
                want.append(np.stack([np.repeat(gi, gj.size), np.tile(gj, gi.size)], axis=-1))
    assert np.array_equal(got, np.concatenate(want))


def test_multigrid_transfer_operators_match_kron_interpolation():
    """multigrid.prolongation (row-by-row construction on the free dofs) against kron(P1x, P1y[, P1z]) x I_nf restricted
    to the free rows / columns; coarse meshes are the structured meshes of the halved element counts."""
    import scipy.sparse as sp
    from autopdex_b200 import mesher, multigrid as mg

    def p1(n):
        P = np.zeros((n + 1, n // 2 + 1))
        for a in range(n + 1):
            if a % 2 == 0:
                P[a, a // 2] = 1.0
            else:
                P[a, a // 2] = P[a, a // 2 + 1] = 0.5
        return sp.csr_matrix(P)

    assert mg.level_shapes((16, 8, 8)) == [(16, 8, 8), (8, 4, 4), (4, 2, 2)]
    assert mg.level_shapes((16, 16), levels=2) == [(16, 16), (8, 8)]
    quad = [[0., 0.], [2., 0.], [2.5, 1.], [0., 1.]]
    cube = [[0., 0., 0.], [1., 0., 0.], [1., 1., 0.], [0., 1., 0.], [0., 0., 1.], [1., 0., 1.], [1., 1., 1.], [0., 1., 1.]]
    for shape, nf in (((4, 6), 1), ((4, 2, 6), 3), ((8, 4, 4), 1), ((6, 4), 2)):
        Pk = p1(shape[0])
        for n in shape[1:]:
            Pk = sp.kron(Pk, p1(n))
        Pk = sp.kron(Pk, sp.identity(nf)).tocsr()
        rng = np.random.default_rng(0)
        ff = rng.random((mg.node_count(shape), nf)) > 0.2
        fc = ff[mg.fine_node_ids(shape)]
        (ip, ix, dt), (rp, ri, rd) = mg.prolongation(shape, nf, ff, fc)
        P = sp.csr_matrix((dt, ix, ip), shape=(ff.sum(), fc.sum()))
        assert abs(P - Pk[ff.ravel()][:, fc.ravel()]).max() == 0.0
        assert abs(sp.csr_matrix((rd, ri, rp), shape=(fc.sum(), ff.sum())) - P.T).max() == 0.0
        assert ip.dtype == np.int32 and ix.dtype == np.int32
        verts, et = (quad, "quad") if len(shape) == 2 else (cube, "brick")
        c, e = mesher.structured_mesh(shape, verts, et)
        assert np.array_equal(e, mg.structured_connectivity(shape))
        cc, _ = mesher.structured_mesh(tuple(n // 2 for n in shape), verts, et)
        assert np.allclose(cc, c[mg.fine_node_ids(shape)])      # injected coordinates = the coarse structured mesh
