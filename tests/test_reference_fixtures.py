"""Oracle and host code against OUTPUTS OF THE REFERENCE ITSELF.

tests/golden/reference_fixtures.npz was produced by tests/golden/make_reference_fixtures.py, which executes the
unmodified AutoPDEx modules on a NumPy stand-in for JAX (tests/golden/fakejax.py; AD replaced by numerical
differentiation: tangents ~1e-9, everything else ~1e-14).  The reference tree does not travel to the GPU box,
the fixture file does.
"""
import os

import numpy as np
import pytest

from autopdex_b200 import mesher, seeder, spaces
from oracle import assemble as oasm
from oracle import mesher as omesh
from oracle import quadrature as oquad
from oracle import shapes as oshapes
from oracle import solve as osolve
from tests import problems

FIX = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_fixtures.npz"))
TANGENT_RTOL = 2e-7        # numerical AD in the generator
EXACT_RTOL = 1e-12


def rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(b).max(), 1e-300)


# ---- tables: quadrature, shape functions, meshes, COO index order ------------------------------------------
def test_gauss_rules_match_reference_tables():
    for d in (1, 2, 3):
        for o in (1, 2, 3, 4, 5, 6):
            for impl in (seeder.gauss_legendre_nd, oquad.gauss_legendre_nd):
                x, w = impl(d, o)
                assert np.abs(x - FIX["gauss_%d_%d_x" % (d, o)]).max() < 4e-15
                assert np.abs(w - FIX["gauss_%d_%d_w" % (d, o)]).max() < 4e-15


def test_simplex_rules_low_order_match_reference():
    # the simplex rules are the reference's tables themselves, point order included: exact equality with the
    # reference-run fixtures (orders 1 and 2 of both families are in the fixture file; all ten orders are checked for
    # their polynomial exactness below)
    for fam, fun in (("tri", seeder.int_pts_ref_tri), ("tet", seeder.int_pts_ref_tet)):
        for order in (1, 2):
            x, w = fun(order)
            assert np.array_equal(x, FIX["%s_rule_%d_x" % (fam, order)]) and np.array_equal(w, FIX["%s_rule_%d_w" % (fam, order)])
    import math
    for order in range(1, 11):
        x, w = seeder.int_pts_ref_tri(order)
        for a in range(order + 1):
            for b in range(order + 1 - a):       # integral of x^a y^b over the reference triangle = a! b! / (a+b+2)!
                exact = math.factorial(a) * math.factorial(b) / math.factorial(a + b + 2)
                assert abs(np.sum(w * x[:, 0] ** a * x[:, 1] ** b) - exact) < 2e-13, (order, a, b)
        x, w = seeder.int_pts_ref_tet(order)
        for a in range(order + 1):
            for b in range(order + 1 - a):
                for c in range(order + 1 - a - b):
                    exact = math.factorial(a) * math.factorial(b) * math.factorial(c) / math.factorial(a + b + c + 3)
                    assert abs(np.sum(w * x[:, 0] ** a * x[:, 1] ** b * x[:, 2] ** c) - exact) < 2e-13, (order, a, b, c)
    with pytest.raises(ValueError):
        seeder.int_pts_ref_tri(11)


def test_meshes_match_reference_mesher():
    cube = [[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1.2], [1, 1, 1], [0, 1, 1]]
    quad = [[0, 0], [2, 0], [2.5, 1.5], [0, 1]]
    for name, args in (("quad", ((3, 4), quad, "quad")), ("tri", ((3, 4), quad, "tri")),
                       ("brick", ((2, 3, 4), cube, "brick")), ("tet", ((2, 3, 2), cube, "tet"))):
        for impl in (mesher.structured_mesh, omesh.structured_mesh):
            c, e = impl(*args)
            assert np.array_equal(e, FIX["mesh_%s_elems" % name]) and np.abs(c - FIX["mesh_%s_coords" % name]).max() < 1e-15
    c, e = mesher.structured_mesh((2, 2, 2), cube, "brick")
    for impl in (mesher.elevate_mesh_order, omesh.elevate_bricks):
        c2, e2 = impl(c, e)
        assert np.array_equal(e2, FIX["mesh_hex27_elems"]) and np.allclose(c2, FIX["mesh_hex27_coords"], atol=1e-15)
    c, e = mesher.structured_mesh((3, 2), quad, "tri")
    for impl in (mesher.elevate_mesh_order, omesh.elevate_triangles):
        c2, e2 = impl(c, e)
        assert np.array_equal(e2, FIX["mesh_tri6_elems"]) and np.allclose(c2, FIX["mesh_tri6_coords"], atol=1e-15)
    x, w, n, conn = seeder.int_pts_in_tri_mesh(c, e, 2)
    assert np.array_equal(conn, FIX["intpts_tri_conn"]) and np.allclose(w, FIX["intpts_tri_w"], rtol=1e-13)
    # same points per element, the rule's internal point order may differ
    assert np.allclose(np.sort(x.reshape(-1, 3, 2), axis=1), np.sort(FIX["intpts_tri_x"].reshape(-1, 3, 2), axis=1))


def test_coo_index_order_matches_reference_get_indices():
    conn = np.array([[0, 3, 4, 1], [1, 4, 5, 2]])
    rows, cols = oasm.coo_indices([dict(conn=conn, nf=2)])
    assert np.array_equal(np.stack([rows, cols], axis=1), FIX["indices_nf2"])


def test_get_indices_two_fields_matches_reference_run():
    """a7: the product's host-side _get_indices against the unmodified reference function run on dict dofs with two
    fields (u: 8 nodes x 2, p: 4 nodes scalar; tests/golden/make_reference_fixtures.py:case_indices_dict)."""
    from autopdex_b200 import assembler
    cu = np.array([[0, 3, 4, 1], [1, 4, 5, 2], [3, 6, 7, 4]])
    cp = np.array([[0, 1, 2], [1, 2, 3], [2, 3, 0]])
    got = assembler._get_indices({"u": cu, "p": cp}, {"u": np.zeros((8, 2)), "p": np.zeros(4)})
    assert got.dtype == np.int64 and np.array_equal(got, FIX["indices_dict_up"])
    # array dofs too
    conn = np.array([[0, 3, 4, 1], [1, 4, 5, 2]])
    assert np.array_equal(assembler._get_indices(conn, np.zeros((6, 2))), FIX["indices_nf2"])


# ---- assembled residuals / tangents / solutions produced by the reference's assembler and solver ---------------
def _check_against_reference(p, tag, settings=None, check_sol=False):
    dofs = FIX[tag + "_dofs"].reshape(p["mask"].shape)
    R, data = oasm.assemble(p["sets"], p["coords"], dofs, settings or {})
    rows, cols = oasm.coo_indices(p["sets"])
    assert np.array_equal(rows, FIX[tag + "_K_rows"]) and np.array_equal(cols, FIX[tag + "_K_cols"])
    assert rel(R, FIX[tag + "_R"].ravel()) < EXACT_RTOL
    assert rel(data, FIX[tag + "_K_data"]) < TANGENT_RTOL
    if check_sol:
        prob = osolve.Problem(p["sets"], p["coords"], p["mask"], p["values"])
        sol, (it, rn, div) = osolve.damped_newton(prob, np.zeros(p["mask"].shape))
        assert it == int(FIX[tag + "_infos"][0]) and div == bool(FIX[tag + "_infos"][2])
        assert rel(sol.ravel(), FIX[tag + "_sol"].ravel()) < 1e-9


@pytest.mark.parametrize("n", [3, 5])
def test_readme_potential_against_reference_run(n):
    p = problems.readme_poisson(n)
    tag = "readme%d" % n
    assert np.array_equal(p["mask"][:, 0], FIX[tag + "_mask"])        # geometry.psdf_polygon: corners are NOT Dirichlet
    _check_against_reference(p, tag, check_sol=True)
    if n == 5:
        assert np.isclose(FIX[tag + "_sol"].sum(), 1.9066412530282952, rtol=1e-13)   # the run reproduces the reference's golden


def test_newton_loop_semantics_against_reference_damped_newton():
    def run(norms, newton_tol=1e-8, maxiter=30):
        k = {"i": 0}

        def residual(d):
            r = norms[min(k["i"], len(norms) - 1)]
            k["i"] += 1
            return np.array([r, 0.0])
        _, (it, rn, div) = osolve.damped_newton(None, np.zeros(2), newton_tol, maxiter, 1.0, lin_solve_fun=lambda d: np.zeros(2),
                                                residual_fun=residual, free=np.array([True, True]))
        return it, rn, div
    for name, args in (("newton_converge", ([1.0, 1e-3, 1e-9],)), ("newton_diverge", ([1.0, 0.5, 0.4, 8.0, 1e-9],)),
                       ("newton_early_jump_ok", ([1.0, 50.0, 1e-9],)), ("newton_nan", ([1.0, float("nan"), 1e-9],)),
                       ("newton_maxiter", ([1.0] * 10, 1e-8, 3))):
        it, rn, div = run(*args)
        ref = FIX[name]
        assert it == int(ref[0]) and div == bool(ref[2]), name
        assert (np.isnan(rn) and np.isnan(ref[1])) or np.isclose(rn, ref[1]), name


# ---- 'user element' sets executed by the reference's assembler (numerical AD in the generator) -------------------
def _element_problem(tag):
    from oracle import mesher as om
    from oracle import quadrature as oq
    cube2 = [[0, 0, 0], [1, 0, 0], [1.1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1.2], [1, 1, 1], [0, 1, 1]]
    mat = dict(youngs_modulus=100.0, poisson_ratio=0.3)
    if tag == "cook_q9":
        sets = [dict(kind="domain", etype="quad9", conn=problems.COOK_ELEMS[[2, 4]], nf=2, gp=oq.gauss_legendre_nd(2, 4),
                     model=dict(name="neo_hooke", mode="plain strain", **mat)),
                dict(kind="surface", etype="line3", conn=problems.COOK_SURF[[4, 5]], nf=2, gp=oq.gauss_legendre_nd(1, 4),
                     model=dict(name="neumann", traction=np.array([0.0, 4.0])))]
        return sets, problems.COOK_NODES, 2
    if tag == "hex8_neo":
        c, e = om.structured_mesh((1, 1, 2), cube2, "brick")
        sets = [dict(kind="domain", etype="hex8", conn=e, nf=3, gp=oq.gauss_legendre_nd(3, 2),
                     model=dict(name="neo_hooke", mode="3d", **mat)),
                dict(kind="surface", etype="quad4", conn=np.array([[2, 8, 11, 5]]), nf=3, gp=oq.gauss_legendre_nd(2, 2),
                     model=dict(name="neumann", traction=np.array([0.0, 0.5, -4.0])))]
        return sets, c, 3
    if tag in ("linel_plain_strain", "linel_plain_stress"):
        c, e = om.structured_mesh((2, 1), [[0, 0], [2, 0], [2.5, 1.5], [0, 1]], "quad")
        mode = tag[6:].replace("_", " ")
        return [dict(kind="domain", etype="quad4", conn=e, nf=2, gp=oq.gauss_legendre_nd(2, 2),
                     model=dict(name="linear_elasticity", mode=mode, body_load=np.array([0.3, -1.0]), **mat))], c, 2
    if tag == "linel_3d":
        c, e = om.structured_mesh((1, 1, 2), cube2, "brick")
        return [dict(kind="domain", etype="hex8", conn=e[:1], nf=3, gp=oq.gauss_legendre_nd(3, 2),
                     model=dict(name="linear_elasticity", mode="3d", body_load=np.array([0.3, -1.0, 0.2]), **mat))], c, 3
    if tag == "tri6_neo":
        c, e = om.structured_mesh((1, 1), [[0, 0], [2, 0], [2.3, 1.0], [0, 1]], "tri")
        c, e = om.elevate_triangles(c, e)
        gp = (FIX["tri_rule_2_x"], FIX["tri_rule_2_w"])
        return [dict(kind="domain", etype="tri6", conn=e, nf=2, gp=gp,
                     model=dict(name="neo_hooke", mode="plain strain", **mat))], c, 2
    # ---- session 3 (generator: case_elements_more) ----
    tets = (np.array([[0., 0., 0.], [1., 0., 0.1], [0.1, 1., 0.], [0., 0.2, 1.], [1.1, 1.2, 0.9]]), np.array([[0, 1, 2, 3], [1, 2, 3, 4]]))
    if tag == "quad4_neo_line2":
        c, e = om.structured_mesh((2, 2), [[0., 0.], [48., 44.], [48., 60.], [0., 44.]], "quad")
        sets = [dict(kind="domain", etype="quad4", conn=e, nf=2, gp=oq.gauss_legendre_nd(2, 2),
                     model=dict(name="neo_hooke", mode="plain strain", **mat)),
                dict(kind="surface", etype="line2", conn=np.array([[6, 7], [7, 8]]), nf=2, gp=oq.gauss_legendre_nd(1, 2),
                     model=dict(name="neumann", traction=np.array([0.0, 4.0])))]
        return sets, c, 2
    if tag == "tet4_neo_tri3":
        sets = [dict(kind="domain", etype="tet4", conn=tets[1], nf=3, gp=(FIX["tet_rule_2_x"], FIX["tet_rule_2_w"]),
                     model=dict(name="neo_hooke", mode="3d", **mat)),
                dict(kind="surface", etype="tri3", conn=np.array([[1, 2, 4]]), nf=3, gp=(FIX["tri_rule_2_x"], FIX["tri_rule_2_w"]),
                     model=dict(name="neumann", traction=np.array([0.2, 0.5, -4.0])))]
        return sets, tets[0], 3
    if tag == "tri3_linel":
        c, e = om.structured_mesh((2, 1), [[0, 0], [2, 0], [2.5, 1.5], [0, 1]], "tri")
        return [dict(kind="domain", etype="tri3", conn=e, nf=2, gp=(FIX["tri_rule_2_x"], FIX["tri_rule_2_w"]),
                     model=dict(name="linear_elasticity", mode="plain stress", body_load=np.array([0.3, -1.0]), **mat))], c, 2
    if tag in ("hyperlin_plain_strain", "hyperlin_3d"):      # generator: case_hyper_linear
        if tag == "hyperlin_3d":
            c, e = om.structured_mesh((1, 1, 2), cube2, "brick")
            return [dict(kind="domain", etype="hex8", conn=e[:1], nf=3, gp=oq.gauss_legendre_nd(3, 2),
                         model=dict(name="linear_elasticity", mode="lame", **mat))], c, 3
        c, e = om.structured_mesh((2, 1), [[0, 0], [2, 0], [2.5, 1.5], [0, 1]], "quad")
        return [dict(kind="domain", etype="quad4", conn=e, nf=2, gp=oq.gauss_legendre_nd(2, 2),
                     model=dict(name="linear_elasticity", mode="lame", **mat))], c, 2
    raise KeyError(tag)


ELEMENT_TAGS = ["cook_q9", "hex8_neo", "linel_plain_strain", "linel_plain_stress", "linel_3d", "tri6_neo"]
ELEMENT_TAGS_MORE = ["quad4_neo_line2", "tet4_neo_tri3", "tri3_linel"]     # session 3: GPU half in test_zz_gpu_first_run.py
ELEMENT_TAGS_HYPERLIN = ["hyperlin_plain_strain", "hyperlin_3d"]           # GPU half in test_zz_gpu_r02_hyper_linear.py


@pytest.mark.parametrize("tag", ELEMENT_TAGS + ELEMENT_TAGS_MORE + ELEMENT_TAGS_HYPERLIN)
def test_user_elements_against_reference_run(tag):
    if tag + "_R" not in FIX:
        pytest.skip("fixture %s not generated yet" % tag)
    sets, coords, nf = _element_problem(tag)
    dofs = FIX[tag + "_dofs"]
    R, data = oasm.assemble(sets, coords, dofs, {})
    rows, cols = oasm.coo_indices(sets)
    assert np.array_equal(rows, FIX[tag + "_K_rows"]) and np.array_equal(cols, FIX[tag + "_K_cols"])
    assert rel(R, FIX[tag + "_R"].ravel()) < 1e-11
    assert rel(data, FIX[tag + "_K_data"]) < TANGENT_RTOL


# ---- README-style 'user potential' with dict dofs on hex8 / hex27 / tet4 / tri3 / tri6 (session 3): the route on which the
# reference itself runs the scalar Poisson problem on isoparametric elements (generator: case_potential3d)
POTENTIAL_TAGS = ["pot_hex8", "pot_hex27", "pot_tet4", "pot_tri3", "pot_tri6", "pot_quad9", "pot_tet10"]


def _potential_source(x):
    shift = np.array([0.3, -0.2, 0.5][:x.shape[-1]])
    return 3.0 * np.sin(2.0 * np.sum(x * x, axis=-1)) - np.cos(np.sum((x - shift) ** 2, axis=-1))


def potential_problem(tag):
    from oracle import elements as oel
    coords, elems = FIX[tag + "_coords"], FIX[tag + "_elems"]
    st = dict(kind="domain", etype=tag[4:], conn=elems, nf=1, gp=(FIX[tag + "_gp_x"], FIX[tag + "_gp_w"]),
              model=dict(name="poisson_potential"))
    st["model"]["source"] = _potential_source(oel.gauss_point_coordinates(st, coords))
    mask = np.zeros((coords.shape[0], 1), dtype=bool)
    return dict(sets=[st], coords=coords, mask=mask, values=np.zeros(mask.shape), nf=1)


@pytest.mark.parametrize("tag", POTENTIAL_TAGS)
def test_user_potential_3d_and_simplex_against_reference_run(tag):
    if tag + "_R" not in FIX:
        pytest.skip("fixture %s not generated yet" % tag)
    p = potential_problem(tag)
    dofs = FIX[tag + "_dofs"][:, None]
    R, data = oasm.assemble(p["sets"], p["coords"], dofs, {})
    rows, cols = oasm.coo_indices(p["sets"])
    assert np.array_equal(rows, FIX[tag + "_K_rows"]) and np.array_equal(cols, FIX[tag + "_K_cols"])
    assert rel(R, FIX[tag + "_R"].ravel()) < 1e-11
    assert rel(data, FIX[tag + "_K_data"]) < TANGENT_RTOL


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ELEMENT_TAGS + ["readme3", "readme5"])
def test_gpu_against_reference_run(tag):
    """The CUDA path against the reference's own outputs (no oracle in between)."""
    if tag + "_R" not in FIX:
        pytest.skip("fixture %s not generated yet" % tag)
    from autopdex_b200 import backend
    from tests import gpu_util
    if tag.startswith("readme"):
        p = problems.readme_poisson(int(tag[6:]))
        sets, coords, nf, mask = p["sets"], p["coords"], 1, p["mask"]
    else:
        sets, coords, nf = _element_problem(tag)
        mask = np.zeros((coords.shape[0], nf), dtype=bool)
        mask[0] = True                       # any mask: the full-CSR values do not depend on it
    p = dict(sets=sets, coords=coords, mask=mask, values=np.zeros(mask.shape), nf=nf)
    plan = gpu_util.make_plan(p)
    dofs = FIX[tag + "_dofs"].reshape(mask.shape)
    d, r = backend.DeviceArray.from_host(dofs), backend.DeviceArray(mask.size)
    plan.assemble(d, True, r)
    n = mask.size
    import scipy.sparse as sp
    ref = sp.csr_matrix(sp.coo_matrix((FIX[tag + "_K_data"], (FIX[tag + "_K_rows"], FIX[tag + "_K_cols"])), shape=(n, n)))
    ref.sort_indices()
    indptr, indices = plan.csr(False)
    assert np.array_equal(indptr, ref.indptr) and np.array_equal(indices, ref.indices)
    assert rel(plan.values(False), ref.data) < TANGENT_RTOL
    assert rel(r.download(), FIX[tag + "_R"].ravel()) < 1e-11
    plan.destroy()


def test_sparse_mode_compiled_against_reference_run():
    """'sparse' mode chain of the reference (assembler.py:874-1035, variational_schemes.py:185-252, solution_structures.py:
    242-248) for conduction + Euler capacity + surface inflow, and one 'linear' solver step (mixed return vector)."""
    p = problems.heat_sparse(2, 2, 1)
    dofs_n = FIX["sparse_dofs_n"]
    assert np.array_equal(dofs_n, p["settings"]["dofs n"])
    R, data = oasm.assemble(p["sets"], p["coords"], dofs_n, p["settings"])
    rows, cols = oasm.coo_indices(p["sets"])
    assert np.array_equal(rows, FIX["sparse_K_rows"]) and np.array_equal(cols, FIX["sparse_K_cols"])
    assert rel(R, FIX["sparse_R"].ravel()) < 1e-11
    assert rel(data, FIX["sparse_K_data"]) < TANGENT_RTOL
    prob = osolve.Problem(p["sets"], p["coords"], p["mask"], p["values"], p["settings"])
    delta = osolve.solve_linear(prob, dofs_n)
    assert rel(delta, FIX["sparse_delta"]) < 1e-9


@pytest.mark.parametrize("name", ["tri3", "tri6", "tet4"])
def test_fem_ini_simplex_direct_tables_against_reference(name):
    """Physical-space P1/P2 shape values and gradients of spaces.fem_ini_simplex (the 'direct' shape function mode)."""
    if "simplex_%s_N" % name not in FIX:
        pytest.skip("fixture not generated")
    xI, x = FIX["simplex_%s_xI" % name], FIX["simplex_%s_x" % name]
    for impl in (spaces.simplex_physical_tables, oshapes.simplex_physical_tables):
        N, dN = impl(x[None, :], xI[None, :, :])
        assert np.abs(N[0] - FIX["simplex_%s_N" % name]).max() < 1e-10
        assert np.abs(dN[0] - FIX["simplex_%s_dN" % name]).max() < 1e-6 * max(1.0, np.abs(dN).max())


_DAE_TAGS = (["be"] + ["bdf%d" % k for k in range(1, 7)] + ["am%d" % k for k in range(1, 7)] + ["dirk%d" % k for k in (1, 2, 3)])


@pytest.mark.parametrize("tag", _DAE_TAGS)
def test_time_integrator_rules_against_reference_integrators(tag):
    """autopdex_b200.dae integrators against the reference's own classes (autopdex/dae.py:288-318, 420-481, 537-587,
    707-766 executed by tests/golden/make_reference_fixtures.py `dae_rules`): for every stage the value / first derivative
    the residual is evaluated at (`_rule`) must be value = x + d, q_t = a value + b with this package's (a, b, d), and the
    end-of-step update (`_update`) must agree."""
    from autopdex_b200 import dae
    integ = (dae.BackwardEuler() if tag == "be" else dae.BackwardDiffFormula(int(tag[3:])) if tag.startswith("bdf")
             else dae.AdamsMoulton(int(tag[2:])) if tag.startswith("am") else dae.DiagonallyImplicitRungeKutta(int(tag[4:])))
    pre = "dae_%s_" % tag
    q_n, q_t_n, stages, dt = FIX[pre + "q_n"], FIX[pre + "q_t_n"], FIX[pre + "stages"], float(FIX[pre + "dt"])
    assert (integ.num_steps, integ.num_stages) == (q_n.shape[0], stages.shape[0])
    assert np.allclose(integ.stage_positions, FIX[pre + "positions"], rtol=1e-15) and integ.order == int(FIX[pre + "order"])
    scale = np.abs(FIX[pre + "q_t"]).max()
    for s in range(integ.num_stages):
        a, b, d = integ.stage_rule(s, dt, stages, q_n, q_t_n)
        value = stages[s] + d
        assert np.abs(value - FIX[pre + "value"][s]).max() < 1e-13 * max(1.0, np.abs(value).max()), (tag, s)
        assert np.abs(a * value + b - FIX[pre + "q_t"][s]).max() < 1e-13 * scale, (tag, s)
    q_n1, q_t_n1 = integ.update(stages, q_n, q_t_n, dt)
    assert np.abs(q_n1 - FIX[pre + "q_n1"]).max() < 1e-13 * max(1.0, np.abs(q_n1).max())
    assert np.abs(q_t_n1 - FIX[pre + "q_t_n1"].reshape(q_t_n1.shape)).max() < 1e-13 * scale


def test_save_policies_and_step_size_controllers_against_reference_classes():
    """autopdex_b200.dae save policies / step-size controllers against the reference's own classes (autopdex/dae.py:1186-1311,
    1474-1573; fixture case `dae_control`, same input sequences as tests/golden/make_reference_fixtures.py)."""
    from autopdex_b200 import dae
    times = [0.0, 0.2, 0.35, 0.5, 0.7, 0.9, 1.0]
    seq = [(True, 2, 0.1), (True, 9, 0.2), (False, 4, 0.25), (True, 1, 0.29), (False, 3, 0.02), (False, 3, 0.011)]
    q0 = np.arange(4.0)
    for tag, pol, max_steps in (("equi3", dae.SaveEquidistantPolicy(num_points=3, tol=1e-6), 10),
                                ("equi_default", dae.SaveEquidistantPolicy(), 5), ("all", dae.SaveAllPolicy(), 10),
                                ("all_clipped", dae.SaveAllPolicy(), 4)):
        st = pol.initialize({"a": q0}, 1.0, max_steps, {"u": np.asarray(0.0)})
        for t in times:
            st = pol.save_step(st, t, {"a": q0 + t}, {"u": np.asarray(2 * t)})
        h = pol.finalize(st)
        for got, name in ((h.t, "t"), (h.q["a"], "q"), (h.user["u"], "u")):
            ref = FIX["dae_save_%s_%s" % (tag, name)]
            assert got.shape == ref.shape and np.allclose(got, ref, rtol=1e-15, atol=0, equal_nan=True), (tag, name)
    for tag, ctrl in (("root", dae.RootIterationController(target_niters=6, gamma=0.5, max_step_size=0.3, min_step_size=0.01)),
                      ("const", dae.ConstantStepSizeController())):
        st, rows = ctrl.initialize(), []
        for conv, its, dt in seq:
            st = ctrl.compute_scaler(st, conv, its, dt)
            st = ctrl.check_accept(st, conv, -1)
            rows.append([st["step_scaler"], st["dt"], float(st["accept"]), float(st["interrupt"])])
        assert np.allclose(np.array(rows), FIX["dae_ctrl_%s" % tag], rtol=1e-14, atol=0), tag


@pytest.mark.parametrize("tag", ["smooth", "halving", "stall", "slow"])
def test_adaptive_load_stepping_control_flow_against_reference_run(tag, monkeypatch):
    """solver.adaptive_load_stepping (rollback on divergence, halving, growth by the Newton count, capping at
    max_multiplier / max_increment, stop below min_increment: solver.py:296-379) against the reference's own function
    driven by the same scripted Newton solver (fixture case `load_stepping`): identical sequences of tried multipliers and
    the same final carry."""
    from autopdex_b200 import solver
    script = {"smooth": [(6, False)],
              "halving": [(5, False), (30, True), (30, True), (4, False), (9, False), (30, True), (3, False)],
              "stall": [(30, True)], "slow": [(12, False)]}[tag]
    calls = []

    def fake_solver(dofs, settings, static_settings, newton_tol=1e-8, **kwargs):
        steps, div = script[min(len(calls), len(script) - 1)]
        calls.append(float(settings["load multiplier"]))
        return dofs + 1.0, (steps, 1e-12, div)
    monkeypatch.setattr(solver, "solver", fake_solver)
    st = {"solver type": "newton", "solver backend": "b200", "solver": "cg", "verbose": -1, "assembling mode": ("user element",),
          "model": (None,)}
    monkeypatch.setattr(solver, "_Config", lambda s: type("C", (), {"solver_type": "newton", "verbose": -1})())
    carry = solver.adaptive_load_stepping(np.zeros(3), {"load multiplier": 0.0}, st, path_dependent=True, implicit_diff_mode=None)
    tried, final = FIX["loadstep_%s_tried" % tag], FIX["loadstep_%s_final" % tag]
    assert len(calls) == len(tried) and np.allclose(calls, tried, rtol=1e-14, atol=0)
    assert np.allclose([carry[0][0], carry[1], carry[2]], final, rtol=1e-13, atol=1e-300)


def test_dict_utilities_against_reference_run():
    """autopdex_b200.utility.dict_flatten / reshape_as / dict_zeros_like / dof_select against the reference's functions
    (utility.py:82-199, 448-462; fixture case `utility`): key order, shapes and values identical."""
    from autopdex_b200 import utility
    rng = np.random.default_rng(21)
    d = {"u": rng.normal(size=(5, 2)), "p": rng.normal(size=(5,)), "T": rng.normal(size=(3, 1))}
    flat = rng.normal(size=5 * 2 + 5 + 3)
    nodes = rng.random(6) < 0.5
    assert np.array_equal(utility.dict_flatten(d), FIX["util_flat"])
    back, zeros = utility.reshape_as(flat, d), utility.dict_zeros_like(d)
    assert list(back.keys()) == list(d.keys())
    for k in d:
        assert back[k].shape == FIX["util_reshape_" + k].shape and np.array_equal(back[k], FIX["util_reshape_" + k])
        assert zeros[k].shape == FIX["util_zeros_" + k].shape and not zeros[k].any()
    assert np.array_equal(utility.dof_select(nodes, [True, False, True]), FIX["util_dofsel_list"])
    assert np.array_equal(utility.dof_select(nodes, True), FIX["util_dofsel_bool"])
