"""`'solver backend': 'b200'` behind the reference entry point
solver.solver(dofs, settings, static_settings, **kwargs)  (autopdex/solver.py:41-137).

Same call signature and return contract as the reference:
  'newton' / 'damped newton'  -> (dofs_like, (n_steps, res_norm, diverged))        solver.py:948
  'linear'                    -> (dofs_like, None), the MIXED vector of solve_linear  solver.py:648-659
Assembly, nodal Dirichlet imposition, the Newton loop and the Krylov solve run on the device
inside libapdx_b200.so; this module reads the two settings dicts, rejects what the backend does
not support (ValueError, never a CPU path), evaluates the Python coefficient callables on the
host once per call, and owns the plan cache (pattern + index maps are built once per mesh).

Keyword arguments: newton_tol, maxiter (Newton, as solver.solve_newton solver.py:741-743),
damping_coefficient ('damped newton'); tol / atol / krylov_maxiter (Krylov, as the **kwargs of
jax.scipy.sparse.linalg.cg in linear_solve_jax, solver.py:1116).  The reference's solve_newton
silently drops Krylov kwargs (solver.py:766-768); here they are honoured.
"""
from collections import OrderedDict
from collections.abc import Mapping
from inspect import signature

import numpy as np

from . import backend, models, spaces, utility

_SUPPORTED_SOLVER_TYPES = ("newton", "damped newton", "linear")
_PLAN_CACHE = OrderedDict()
_PLAN_CACHE_SIZE = 4
last_stats = {}


# ---- reading static_settings ---------------------------------------------------------------------
def _get(static_settings, key, default=None, required=False):
    try:
        return static_settings[key]
    except KeyError:
        if required:
            raise ValueError("b200 backend: static_settings[%r] is required" % key)
        return default


def _per_set(value, n_sets, key):
    if isinstance(value, (tuple, list)):
        if len(value) != n_sets:
            raise ValueError("b200 backend: static_settings[%r] must have one entry per domain" % key)
        return list(value)
    return [value] * n_sets


class _Config:
    """Validated view of static_settings for the b200 path (settings-read-time rejection)."""

    def __init__(self, static_settings):
        if _get(static_settings, "solver backend") != "b200":
            raise ValueError("autopdex_b200.solver handles 'solver backend': 'b200' only, got %r"
                             % (_get(static_settings, "solver backend"),))
        self.solver_type = _get(static_settings, "solver type", required=True)
        if self.solver_type not in _SUPPORTED_SOLVER_TYPES:
            raise ValueError("b200 backend: 'solver type' %r not supported (supported: %s)"
                             % (self.solver_type, ", ".join(_SUPPORTED_SOLVER_TYPES)))
        self.krylov = _get(static_settings, "solver", "cg")
        if self.krylov not in ("cg", "bicgstab"):
            raise ValueError("b200 backend: 'solver' must be 'cg' or 'bicgstab' (Jacobi-preconditioned Krylov), got %r"
                             % (self.krylov,))
        pc = _get(static_settings, "type of preconditioner", None)
        if pc not in (None, "jacobi", "none", "multigrid"):
            raise ValueError("b200 backend: 'type of preconditioner' %r not supported (jacobi, multigrid, none)" % (pc,))
        self.jacobi = pc == "jacobi"
        # geometric multigrid V-cycle on a hierarchy derived from settings['b200 multigrid'] (multigrid.py), the
        # device counterpart of the reference's pyamg / PETSc preconditioners (solver.py:1399-1491, 1224-1333)
        self.multigrid = pc == "multigrid"
        if self.multigrid and self.krylov != "cg":
            raise ValueError("b200 backend: 'type of preconditioner': 'multigrid' needs 'solver': 'cg' (symmetric V-cycle)")
        self.precond = "multigrid" if self.multigrid else ("jacobi" if self.jacobi else "none")
        self.verbose = _get(static_settings, "verbose", 0)
        modes = _get(static_settings, "assembling mode", required=True)
        self.n_sets = len(modes)
        structure = _per_set(_get(static_settings, "solution structure", "off"), self.n_sets, "solution structure")
        for s in structure:
            if s not in ("nodal imposition", "off"):
                raise ValueError("b200 backend: 'solution structure' %r not supported (nodal imposition, off)" % (s,))
        # solver.py:584: `"nodal imposition" in static_settings["solution structure"]`
        self.nodal_imposition = "nodal imposition" in structure
        model_list = _get(static_settings, "model", required=True)
        if len(model_list) != self.n_sets:
            raise ValueError("b200 backend: one model per domain expected")
        if _get(static_settings, "known sparsity pattern", "none") != "none":
            raise ValueError("b200 backend: 'known sparsity pattern' is not supported")
        self.shape_mode = _get(static_settings, "shape function mode", None)
        # device sets: (route, model, domain) -- `domain` indexes settings['connectivity'] etc.; a transient 'user
        # residual' domain (dae.TimeSteppingManager) is served by two device sets on the same connectivity
        self.sets = []
        self.transient = False
        for i, mode in enumerate(modes):
            m = models.recognise(model_list[i])
            if mode == "user residual" and isinstance(m, models.TimeElementModel):
                if not getattr(self, "_dae", False) and not static_settings.get("time integrators"):
                    raise ValueError("b200 backend: domain %d: a time-dependent 'user residual' is driven by "
                                     "autopdex_b200.dae.TimeSteppingManager ('time integrators' in static_settings)" % i)
                self.sets.append(("element", m.steady, i))
                self.sets.append(("element", m.capacity, i))
                self.transient = True
            elif mode in ("user element", "user potential", "user residual"):
                if not isinstance(m, models.ElementModel):
                    raise ValueError("b200 backend: domain %d: assembling mode %r needs an isoparametric element model" % (i, mode))
                route = getattr(m, "route", "user potential" if m.weak.name == "poisson_potential" else "user element")
                if route != mode:
                    raise ValueError("b200 backend: domain %d: model %r does not match assembling mode %r" % (i, m.weak, mode))
                self.sets.append(("element", m, i))
            elif mode == "sparse":
                if not isinstance(m, models.WeakForm):
                    raise ValueError("b200 backend: domain %d: 'sparse' mode needs a weak form" % i)
                scheme = _per_set(_get(static_settings, "variational scheme", required=True), self.n_sets, "variational scheme")[i]
                space = _per_set(_get(static_settings, "solution space", required=True), self.n_sets, "solution space")[i]
                if scheme != "weak form galerkin" or space != "fem simplex":
                    raise ValueError("b200 backend: domain %d: 'sparse' mode supports 'weak form galerkin' on 'fem simplex' "
                                     "only, got %r / %r" % (i, scheme, space))
                if self.shape_mode not in ("direct", "compiled"):
                    raise ValueError("b200 backend: 'shape function mode' must be 'direct' or 'compiled'")
                self.sets.append(("sparse", m, i))
            else:
                raise ValueError("b200 backend: assembling mode %r of domain %d is not supported "
                                 "(user element, user potential, user residual, sparse)" % (mode, i))
        self.model_objects = tuple(model_list)     # their id() is part of the cache key: kept alive with the plan
        self.key = (self.solver_type, self.krylov, self.precond, self.nodal_imposition, self.shape_mode,
                    tuple(id(model_list[i]) for i in range(self.n_sets)), tuple(modes))


    def coarse(self, kept):
        """Configuration of a coarse multigrid level: the kept (domain) sets, Jacobi inside (the level is driven by the
        V-cycle of the finest plan)."""
        import copy
        c = copy.copy(self)
        c.sets = [(self.sets[i][0], self.sets[i][1], j) for j, i in enumerate(kept)]   # coarse settings: one entry per kept set
        c.n_sets = len(kept)
        c.multigrid, c.jacobi, c.precond = False, True, "jacobi"
        return c


def validate(static_settings):
    """Reject unsupported static_settings (raises ValueError); returns the parsed configuration."""
    return _Config(static_settings)


# ---- host evaluation of coefficient callables ---------------------------------------------------------
def _call(fun, x, settings):
    if len(signature(fun).parameters) == 1:
        return fun(x)
    return fun(x, settings)


def _eval_points(fun, pts, settings, ncomp, vectorized=None):
    """Values of `fun` at pts (n, dim) -> (n,) or (n, ncomp); constants stay constants."""
    if fun is None:
        return None
    if not callable(fun):
        return np.asarray(fun, dtype=np.float64)
    n = pts.shape[0]
    want = (n,) if ncomp == 1 else (n, ncomp)
    if vectorized or (vectorized is None and n > 64):
        try:
            v = np.asarray(_call(fun, pts, settings), dtype=np.float64)
        except Exception:
            v = None
        if v is not None:
            # The reference calls these per point; a batched call is only trusted if it reproduces per-point calls at
            # three probe points (a callable that reduces over the point axis, e.g. np.prod(np.sin(x)), must not turn
            # into a constant) and, for point-wise results, a second, shuffled batch (a callable that indexes the
            # batch axis, e.g. x[0] * e_1, gives position-dependent values).
            probe = [0, n // 2, n - 1]
            same = lambda a, b: np.shape(a) == np.shape(b) and np.allclose(a, b, rtol=1e-13, atol=1e-300)
            one = [np.asarray(_call(fun, pts[i], settings), dtype=np.float64) for i in probe]
            if v.shape == (() if ncomp == 1 else (ncomp,)):
                if all(same(o, v) for o in one):
                    return v                               # constant
            elif v.shape == want:
                ok = all(same(one[j], v[i]) for j, i in enumerate(probe))
                if ok:
                    idx = np.random.default_rng(n).permutation(n)[:min(n, 16)]
                    try:
                        v2 = np.asarray(_call(fun, pts[idx], settings), dtype=np.float64)
                        ok = same(v2, v[idx])
                    except Exception:
                        ok = False
                if ok:
                    return v
        if vectorized:
            raise ValueError("b200 backend: %r was declared vectorized but did not map an (n, dim) array to %s" % (fun, want))
        if n > 2_000_000:
            raise ValueError("b200 backend: coefficient callable %r is not vectorisable over %d points" % (fun, n))
    out = np.empty(want)
    for i in range(n):
        out[i] = np.asarray(_call(fun, pts[i], settings), dtype=np.float64)
    return out


_NCOMP_NF = ("body_load", "traction")


class _State:
    """Plan + device buffers for one (mesh, model configuration)."""

    def __init__(self, cfg, dofs, settings):
        self.cfg = cfg
        self.dict_key = None
        # Multi-field dict dofs (assembler.py:61-121, utility.dict_flatten utility.py:104-128): the global numbering is
        # field-major -- all dofs of the first key, then the second, ... -- so with the SAME number of dofs per node in
        # every field the problem is a single-field problem on the concatenated node set (node k of field f becomes
        # node node_offset[f] + k).  A domain whose (recognised, single-field) model acts on field f gets its connectivity
        # shifted accordingly; the structural entries the reference emits for the field pairs the integrand does not
        # couple (explicit zero blocks of jacfwd, assembler.py:79-117) come from one pattern-only device set per domain
        # with the element's nodes of ALL fields, which makes the CSR pattern identical to the reference's.
        self.fields = None
        if isinstance(dofs, Mapping):
            keys = list(dofs.keys())
            if len(keys) == 1:
                self.dict_key = keys[0]
            elif len(keys) == 0:
                raise ValueError("b200 backend: empty dofs dict")
            else:
                self.fields = keys
                shapes = [np.shape(dofs[k]) for k in keys]
                nfs = {1 if len(sh) == 1 else sh[-1] for sh in shapes}
                if len(nfs) != 1 or len({len(sh) for sh in shapes}) != 1:
                    raise ValueError("b200 backend: multi-field dict dofs need the same number of dofs per node in every "
                                     "field (got shapes %s)" % (dict(zip(keys, shapes)),))
                self.field_nodes = [sh[0] for sh in shapes]
                self.node_off = dict(zip(keys, np.concatenate([[0], np.cumsum(self.field_nodes)[:-1]]).astype(np.int64)))
                if settings.get("b200 partition") or cfg.multigrid:
                    raise ValueError("b200 backend: multi-field dict dofs run on one GPU with the Jacobi preconditioner")
        d0 = self._unwrap(dofs)
        self.dofs_ndim = d0.ndim
        coords = np.asarray(self._unwrap(settings["node coordinates"]), dtype=np.float64)
        self.n_nodes, self.dim = coords.shape
        self.nf = 1 if d0.ndim == 1 else d0.shape[-1]
        if d0.size != self.n_nodes * self.nf:
            raise ValueError("b200 backend: dofs and node coordinates disagree on the number of nodes")
        if self.fields is None:
            self.conn_refs = [self._unwrap(settings["connectivity"][dom]) for _, _, dom in cfg.sets]   # per DEVICE set
        else:
            self.conn_refs = [self._field_connectivity(settings["connectivity"][dom], m, route, dom) for route, m, dom in cfg.sets]
        specs = []
        for i, (route, m, dom) in enumerate(cfg.sets):
            conn = np.asarray(self.conn_refs[i])
            if route == "element":
                specs.append(backend.SetSpec(m.kind, m.weak.name, conn, family=m.family, gp=m.gp, mode=m.weak.mode))
            else:
                specs.append(backend.SetSpec("intpoint", m.name, conn, mode=m.mode))
        self.n_model_sets = len(specs)
        if self.fields is not None:      # one pattern-only set per domain: the element's nodes of every field
            for dom in sorted({d for _, _, d in cfg.sets}):
                c = settings["connectivity"][dom]
                allc = np.concatenate([np.asarray(c[k], dtype=np.int64) + self.node_off[k] for k in self.fields], axis=1)
                self.conn_refs.append(allc)
                specs.append(backend.SetSpec("domain", "pattern_only", allc))
        mask = None
        if cfg.nodal_imposition:
            mask = np.asarray(self._unwrap(settings["dirichlet dofs"])).astype(bool).ravel()
        self.mask = mask
        self.keepalive = []                                     # filled by _state_for: objects behind the cache key
        self.plan = backend.Plan(self.dim, self.n_nodes, self.nf, specs, mask)
        # multi-GPU: this process holds one slab (owned nodes + ghost planes, local ids in global order) or one part of
        # a general partition (mesher.rcb_partition: owned nodes first, then the ghosts grouped by owning rank)
        self.partition = settings.get("b200 partition")
        if self.partition:
            pt = self.partition
            if "neighbours" in pt:
                nf = self.nf
                dofs_of = lambda nodes: (np.asarray(nodes, dtype=np.int64)[:, None] * nf + np.arange(nf)).ravel()
                if pt.get("owned_node_begin", 0) != 0:
                    raise ValueError("b200 backend: a list partition numbers its owned nodes first")
                self.plan.set_partition_lists(pt["owned_node_end"] * nf, pt["neighbours"],
                                              [dofs_of(v) for v in pt["send_nodes"]],
                                              [(b * nf, e * nf) for b, e in pt["recv_node_ranges"]])
            else:
                self.plan.set_partition(pt["owned_node_begin"] * self.nf, pt["owned_node_end"] * self.nf,
                                        pt.get("rank_lo", -1), pt.get("rank_hi", -1))
        n = self.plan.n_dofs
        self.dofs_d = backend.DeviceArray(n)
        self.vals_d = backend.DeviceArray(n)
        self.out_d = backend.DeviceArray(n)
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self.coarse_state = None
        if cfg.multigrid:
            self._build_hierarchy(settings)

    # -- multigrid hierarchy (multigrid.py): a chain of coarse _State objects below this one -----------------------
    def _wrap(self, arr):
        if self.fields is not None:              # split the concatenated node axis back into the fields
            arr = np.asarray(arr)
            bounds = np.cumsum(self.field_nodes)[:-1]
            return dict(zip(self.fields, np.split(arr, bounds, axis=0)))
        return {self.dict_key: arr} if self.dict_key is not None else arr

    def _mg_kinds(self):
        kinds = []
        for route, m, dom in self.cfg.sets:
            if route != "element":
                raise ValueError("b200 backend: the multigrid preconditioner supports isoparametric element sets only")
            kinds.append((m.kind, dom))
        return kinds

    def _build_hierarchy(self, settings):
        from . import multigrid
        opt = settings.get("b200 multigrid")
        if not isinstance(opt, Mapping) or "n_elements" not in opt:
            raise ValueError("b200 backend: 'type of preconditioner': 'multigrid' needs settings['b200 multigrid'] = "
                             "{'n_elements': (nx, ny[, nz])} (the structured mesh the hierarchy is derived from)")
        self._mg_slab = None
        if self.partition:
            # Partitioned hierarchy (slab partitions): every level is split like the finest one -- the rank that owns fine
            # node plane 2I owns coarse plane I -- and 'n_elements' is the GLOBAL mesh.  The number of levels is agreed
            # between the ranks (every rank must own free dofs on every level).
            pt = self.partition
            if "neighbours" in pt or "planes" not in pt:
                raise ValueError("b200 backend: the multigrid preconditioner on several GPUs needs a slab partition with "
                                 "settings['b200 partition']['planes'] = (plane_lo, plane_hi, owned_plane_lo, owned_plane_hi) "
                                 "(mesher.slab_partition_mesh provides it)")
            planes = tuple(int(v) for v in pt["planes"])
            per_plane = int(np.prod([n + 1 for n in opt["n_elements"][1:]]))
            if (planes[1] - planes[0]) * per_plane != self.n_nodes:
                raise ValueError("b200 backend: settings['b200 multigrid']['n_elements'] = %s and the slab planes %s do not "
                                 "match the %d local nodes" % (tuple(opt["n_elements"]), planes, self.n_nodes))
            levels = multigrid.slab_levels(opt["n_elements"], planes, opt.get("levels"))
            if not opt.get("_levels_agreed"):
                n_lev = int(-backend.comm_allreduce_host([-float(len(levels))], "max")[0])     # minimum over the ranks
                levels = levels[:n_lev]
            shapes = [lv[0] for lv in levels]
            if len(levels) > 1:
                self._mg_slab = dict(planes_f=levels[0][1], planes_c=levels[1][1], rank_lo=pt.get("rank_lo", -1),
                                     rank_hi=pt.get("rank_hi", -1))
        else:
            shapes = multigrid.level_shapes(opt["n_elements"], opt.get("levels"))
            if multigrid.node_count(shapes[0]) != self.n_nodes:
                raise ValueError("b200 backend: settings['b200 multigrid']['n_elements'] = %s does not match the %d nodes"
                                 % (shapes[0], self.n_nodes))
        self.mg_shape = shapes[0]
        if len(shapes) > 1:
            self._mg_cache = {}
            cs, kept, fine_nodes = multigrid.coarse_level_settings(settings, shapes[0], self._mg_kinds(), self._unwrap, self._wrap,
                                                                   self._mg_cache, self._mg_slab)
            cs["b200 multigrid"] = dict(opt, n_elements=shapes[1], levels=len(shapes) - 1, _levels_agreed=True)
            ccfg = self.cfg.coarse(kept)
            ccfg.multigrid = True                                    # recurse: the coarse state builds its own coarse level
            n_c = np.asarray(self._unwrap(cs["node coordinates"])).shape[0]
            d_c = self._wrap(np.zeros((n_c, self.nf)) if self.dofs_ndim > 1 else np.zeros(n_c))
            self.coarse_state = _State(ccfg, d_c, cs)
            # transfer operators (interpolation reduced to the free dofs of the two plans, its transpose, the injection
            # map) are built on the device from the plans' own Dirichlet maps: multigrid.prolongation is the host
            # statement of the same construction (22 s of NumPy at 256^3), kept as the checker of the tests
            dims_f, dims_c = [n + 1 for n in shapes[0]], [n + 1 for n in shapes[1]]
            off_f = off_c = 0
            if self._mg_slab is not None:
                (off_f, g1), (off_c, G1) = self._mg_slab["planes_f"][:2], self._mg_slab["planes_c"][:2]
                dims_f[0], dims_c[0] = g1 - off_f, G1 - off_c
            self.plan.set_coarse_structured(self.coarse_state.plan, dims_f, dims_c, off_f, off_c)
            self.mg_kept = kept
        coarsest, coarsest_ratio = opt.get("coarsest", 0), opt.get("coarsest ratio", 0.0)
        nc = max(shapes[-1])
        if self.partition and not coarsest and not coarsest_ratio and nc > 4:
            # A partitioned hierarchy ends where a rank would own fewer than two planes, so its coarsest level can be
            # large (16^3 for 256^3 on 8 GPUs): widen the Chebyshev interval of the coarsest solve to the spectrum of that
            # mesh (lambda_max / lambda_min of D^-1 A grows like n^2) and raise the degree with its square root.
            coarsest_ratio = max(40.0, 0.6 * nc * nc)
            coarsest = max(12, int(np.ceil(1.3 * np.sqrt(coarsest_ratio))))
        self.plan.set_multigrid(opt.get("pre", 0), opt.get("post", 0), coarsest, opt.get("ratio", 0.0), coarsest_ratio)

    def update_coarse_fields(self, settings):
        """Per-call upload of the coarse levels' fields (injected coordinates, coefficient values at the coarse Gauss
        points, 'dofs n')."""
        if self.coarse_state is None:
            return
        from . import multigrid
        cs, _, _ = multigrid.coarse_level_settings(settings, self.mg_shape, self._mg_kinds(), self._unwrap, self._wrap,
                                                   self._mg_cache, self._mg_slab)
        self.coarse_state.update_fields(cs)
        self.h2d_bytes += self.coarse_state.h2d_bytes

    def destroy(self):
        if self.coarse_state is not None:
            self.coarse_state.destroy()
            self.coarse_state = None
        self.plan.destroy()

    def _unwrap(self, x):
        if isinstance(x, Mapping):
            if self.fields is not None:          # multi-field: concatenate over the node axis in key order
                if list(x.keys()) != self.fields:
                    raise ValueError("b200 backend: dict-valued settings must have the fields %s of the dofs dict, in that "
                                     "order (got %s)" % (self.fields, list(x.keys())))
                parts = [np.asarray(x[k]) for k in self.fields]
                for k, a, n in zip(self.fields, parts, self.field_nodes):
                    if a.shape[0] != n:
                        raise ValueError("b200 backend: field %r has %d nodes in the dofs but %d here" % (k, n, a.shape[0]))
                return np.concatenate(parts, axis=0)
            if self.dict_key is None or list(x.keys()) != [self.dict_key]:
                raise ValueError("b200 backend: dict-valued settings must have exactly the field of the dofs dict")
            return np.asarray(x[self.dict_key])
        return np.asarray(x)

    def _field_connectivity(self, conn, m, route, dom):
        """Connectivity of the device set of a multi-field domain: the nodes of the field the model acts on, shifted
        into the concatenated node numbering."""
        if route != "element" or getattr(m, "field", None) is None:
            raise ValueError("b200 backend: domain %d: multi-field dict dofs need models built on a named field "
                             "(mixed_reference_domain_potential / _residual with a tagged integrand)" % dom)
        if not isinstance(conn, Mapping) or list(conn.keys()) != self.fields:
            raise ValueError("b200 backend: domain %d: settings['connectivity'] must be a dict with the fields %s" % (dom, self.fields))
        if m.field not in self.fields:
            raise ValueError("b200 backend: domain %d: the model acts on field %r, the dofs have %s" % (dom, m.field, self.fields))
        return np.asarray(conn[m.field], dtype=np.int64) + self.node_off[m.field]

    # -- per-call upload of everything that lives in `settings` --------------------------------------
    def update_fields(self, settings):
        cfg, plan = self.cfg, self.plan
        self.h2d_bytes = 0
        coords = np.ascontiguousarray(self._unwrap(settings["node coordinates"]), dtype=np.float64)
        plan.set_coords(coords)
        self.h2d_bytes += coords.nbytes
        for i, (route, m, dom) in enumerate(cfg.sets):
            weak = m.weak if route == "element" else m
            conn = np.asarray(self.conn_refs[i])
            if route == "element":
                xi = m.gp[0].reshape(len(m.gp[1]), -1)
                if m.physical_x:
                    N, _ = spaces.shape_tables(m.family, conn.shape[1], xi.shape[1], xi)
                    pts = np.einsum("ga,nad->ngd", N, coords[conn]).reshape(-1, self.dim)
                else:
                    pts = xi                                        # reference coordinates (models.py:1671-1679)
                n_gp = xi.shape[0]
            else:
                n_gp = 1
                x_int = np.asarray(settings["integration coordinates"][dom], dtype=np.float64)
                if cfg.shape_mode == "compiled":
                    pts = np.zeros((1, self.dim))                   # variational_schemes.py:214-215
                else:
                    pts = x_int
                self._upload_intpoint_tables(i, dom, settings, conn, coords, x_int)
            for name, fun in weak.funs.items():
                ncomp = self.nf if name in _NCOMP_NF else 1
                v = _eval_points(fun, pts, settings, ncomp, getattr(weak, "vectorized", None))
                if v is None:
                    continue
                if route == "element" and m.physical_x and v.ndim >= 1 and v.shape[0] == pts.shape[0]:
                    v = v.reshape((conn.shape[0], n_gp) + v.shape[1:])
                plan.set_param(i, name, v)
                self.h2d_bytes += np.asarray(v).nbytes
            if weak.name == "capacity":
                plan.set_time_increment(float(settings["time increment"]))
                dn = np.asarray(settings["dofs n"], dtype=np.float64)
                plan.set_dofs_n(dn)
                self.h2d_bytes += dn.nbytes
        self.update_coarse_fields(settings)

    def _upload_intpoint_tables(self, i, dom, settings, conn, coords, x_int):
        w = np.asarray(settings["integration weights"][dom], dtype=np.float64)
        if self.cfg.shape_mode == "compiled":
            f, df = settings["compiled shape functions"][dom][:2]
            N, dN = np.asarray(f, dtype=np.float64), np.asarray(df, dtype=np.float64)
        else:
            N, dN = spaces.simplex_physical_tables(x_int, coords[conn])
        self.plan.set_intpoint_tables(i, N.reshape(conn.shape), dN.reshape(conn.shape + (self.dim,)), w)
        self.h2d_bytes += N.nbytes + dN.nbytes + w.nbytes


def _fingerprint(arr, full_bytes=1 << 20):
    """Cheap content fingerprint of an index / mask array for the plan-cache key: shape, dtype and a hash of the bytes
    (all of them up to 1 MiB, else ~64 Ki evenly strided items plus both ends).  id() alone can be reused after garbage
    collection and does not see in-place edits."""
    a = np.asarray(arr)
    flat = a.reshape(-1)
    if flat.nbytes <= full_bytes:
        sample = np.ascontiguousarray(flat)
    else:
        step = max(1, flat.size // 65536)
        sample = np.ascontiguousarray(np.concatenate([flat[::step], flat[:1024], flat[-1024:]]))
    return (a.shape, a.dtype.str, hash(sample.tobytes()))


def _state_for(cfg, dofs, settings):
    conns = settings["connectivity"]
    ids, keep = [], []
    for c in conns:
        for arr in (c.values() if isinstance(c, Mapping) else (c,)):
            ids.append((id(arr),) + _fingerprint(arr))
            keep.append(arr)
    dd = settings.get("dirichlet dofs") if cfg.nodal_imposition else None
    mask_key = None
    if dd is not None:
        parts = list(dd.values()) if isinstance(dd, Mapping) else [dd]
        # the plan bakes the Dirichlet mask: identity + content
        mask_key = tuple((id(p),) + _fingerprint(p) for p in parts)
        keep.extend(parts)
    d0 = next(iter(dofs.values())) if isinstance(dofs, Mapping) else dofs
    mg = settings.get("b200 multigrid") if cfg.multigrid else None
    mg_key = None if mg is None else tuple(sorted((k, tuple(v) if isinstance(v, (tuple, list)) else v) for k, v in mg.items()))
    key = (cfg.key, tuple(ids), mask_key, tuple(np.shape(d0)), mg_key)
    st = _PLAN_CACHE.get(key)
    if st is None:
        st = _State(cfg, dofs, settings)
        # keep every object whose id() is part of the key alive as long as the plan (a freed object's id can be reused)
        st.keepalive = keep + [m for m in (settings.get("b200 partition"),) if m is not None] + list(cfg.model_objects)
        _PLAN_CACHE[key] = st
        while len(_PLAN_CACHE) > _PLAN_CACHE_SIZE:
            _, old = _PLAN_CACHE.popitem(last=False)
            old.destroy()
    else:
        _PLAN_CACHE.move_to_end(key)
    return st


def clear_plan_cache():
    while _PLAN_CACHE:
        _, old = _PLAN_CACHE.popitem()
        old.destroy()


def _warn_unconverged(stats, where):
    """The reference's 'scipy' backend cannot return an unconverged solve silently (direct solver); the Krylov loop can
    (maxiter, breakdown), so say so: warnings.warn, and the flag stays in last_stats['krylov_converged']."""
    if not stats.get("krylov_converged", True):
        import warnings
        warnings.warn("b200 backend: the last Krylov solve of %s() ended without meeting its tolerance "
                      "(relative residual %.3e after %d iterations in total): raise krylov_maxiter or check the system"
                      % (where, stats.get("krylov_relres", float("nan")), int(stats.get("krylov_iters", 0))), RuntimeWarning,
                      stacklevel=3)


# ---- the entry point --------------------------------------------------------------------------------------
def solver(dofs, settings, static_settings, newton_tol=1e-8, maxiter=30, damping_coefficient=None,
           tol=1e-10, atol=0.0, krylov_maxiter=0, **kwargs):
    """Drop-in for autopdex.solver.solver with 'solver backend': 'b200' (solver.py:41-137)."""
    global last_stats
    cfg = _Config(static_settings)
    if kwargs:
        raise ValueError("b200 backend: unknown keyword arguments %s" % sorted(kwargs))
    st = _state_for(cfg, dofs, settings)
    st.update_fields(settings)
    plan = st.plan
    d0 = np.ascontiguousarray(st._unwrap(dofs), dtype=np.float64)
    st.dofs_d.upload(d0.ravel())
    st.h2d_bytes += d0.nbytes
    vals_d = None
    if cfg.nodal_imposition:
        dv = np.ascontiguousarray(st._unwrap(settings["dirichlet conditions"]), dtype=np.float64)
        st.vals_d.upload(dv.ravel())
        st.h2d_bytes += dv.nbytes
        vals_d = st.vals_d
    opts = backend.KrylovOptions(cfg.krylov, rtol=tol, atol=atol, maxiter=krylov_maxiter, jacobi=cfg.precond)

    def wrap(flat):
        return st._wrap(flat.reshape(d0.shape))

    if cfg.solver_type == "linear":
        plan.linear_step(opts, st.dofs_d, vals_d, st.out_d)
        sol = st.out_d.download()
        infos = None
    else:
        damping = 1.0 if cfg.solver_type == "newton" else (0.8 if damping_coefficient is None else damping_coefficient)
        n_it, res, div = plan.newton(opts, st.dofs_d, vals_d, newton_tol, maxiter, damping)
        sol = st.dofs_d.download()
        infos = (n_it, res, div)
        if cfg.verbose > 0:                      # solver.py:906-912, one line per iteration (printed after the device loop)
            for i, r in enumerate(plan.newton_history()):
                print("Residual after Newton iteration %d: %s" % (i + 1, r))
                if cfg.verbose > 1:
                    print("")
        if div and n_it > maxiter:
            print("Warning: Newton scheme could not converge!")
    st.d2h_bytes = sol.nbytes
    last_stats = dict(plan.stats(), h2d_bytes=st.h2d_bytes, d2h_bytes=st.d2h_bytes)
    _warn_unconverged(last_stats, "solver")
    return wrap(sol), infos


def tangent_solve(dofs, rhs, settings, static_settings, transpose=False, tol=1e-10, atol=0.0, krylov_maxiter=0):
    """Linear solve with the tangent at `dofs` for sensitivities: what the reference's implicit
    differentiation asks of a solver backend, `solve_fun(mat, rhs, free_dofs_flat)` and `solve_fun(mat.T, ...)` with
    `mat = assemble_tangent(sol)` (implicit_diff.py:139-183, 225-234, 274-304; backend dispatch solver.py:232-283).

    `dofs` is the state the tangent is taken at (Dirichlet values are imposed as in the Newton step), `rhs` a dof-shaped
    array (dict dofs: dict).  Returns the dof-shaped solution with zeros on the Dirichlet dofs, i.e.
    `mask_op(zeros, free_dofs_flat, u_f, 'set')`.  `transpose=True` is the adjoint solve `A^T u = v` of `_root_vjp`; the
    in-scope tangents are symmetric, so both run the same device solve.  The plan of `solver()` is reused."""
    global last_stats
    cfg = _Config(static_settings)
    st = _state_for(cfg, dofs, settings)
    st.update_fields(settings)
    plan = st.plan
    d0 = np.array(st._unwrap(dofs), dtype=np.float64)
    if cfg.nodal_imposition:
        dv = np.asarray(st._unwrap(settings["dirichlet conditions"]), dtype=np.float64).reshape(d0.shape)
        mask = np.asarray(st._unwrap(settings["dirichlet dofs"]), dtype=bool).reshape(d0.shape)
        d0[mask] = dv[mask]                                    # solver.py:586-604
    r0 = np.ascontiguousarray(st._unwrap(rhs), dtype=np.float64)
    if r0.shape != d0.shape:
        raise ValueError("tangent_solve: rhs has shape %s, the dofs %s" % (r0.shape, d0.shape))
    st.dofs_d.upload(d0.ravel())
    st.vals_d.upload(r0.ravel())                               # staging buffer of the same size
    st.h2d_bytes += d0.nbytes + r0.nbytes
    opts = backend.KrylovOptions(cfg.krylov, rtol=tol, atol=atol, maxiter=krylov_maxiter, jacobi=cfg.precond)
    plan.tangent_solve(opts, st.dofs_d, st.vals_d, st.out_d, transpose=transpose)
    out = st.out_d.download().reshape(d0.shape)
    st.d2h_bytes = out.nbytes
    last_stats = dict(plan.stats(), h2d_bytes=st.h2d_bytes, d2h_bytes=st.d2h_bytes)
    _warn_unconverged(last_stats, "tangent_solve")
    return st._wrap(out)


def adaptive_load_stepping(dofs, settings, static_settings,
                           multiplier_settings=lambda settings, multiplier: (settings.update({"load multiplier": multiplier}), settings)[1],
                           path_dependent=True, implicit_diff_mode=None, max_multiplier=1.0, min_increment=0.01,
                           max_increment=1.0, init_increment=0.2, max_load_steps=1000, target_num_newton_iter=7,
                           newton_tol=1e-10, **kwargs):
    """autopdex.solver.adaptive_load_stepping (solver.py:155-457) around the b200 Newton solve.
    Returns the reference's carry (dofs, multiplier, increment, load_step, settings, res_norm)."""
    if implicit_diff_mode is not None:
        raise ValueError("b200 backend: implicit differentiation is not supported (implicit_diff_mode must be None)")
    cfg = _Config(static_settings)
    if cfg.solver_type not in ("newton", "damped newton"):
        raise ValueError("adaptive_load_stepping works only with solver types 'newton' and 'damped newton'")
    settings = dict(settings)
    multiplier, increment, res_norm, load_step = 0.0, init_increment, 0.0, 0.0
    while multiplier < max_multiplier and increment > min_increment:      # solver.py:296-300
        multiplier += increment                                           # :316
        if cfg.verbose > -1:
            print("Multiplier: %s" % multiplier)
        settings = multiplier_settings(settings, multiplier)              # :323
        new, (steps, res_norm, diverged) = solver(dofs, settings, static_settings, newton_tol=newton_tol, **kwargs)
        if diverged:                                                      # :356-371
            multiplier -= increment
            increment *= 0.5
        else:
            increment *= 1 + 0.5 * (target_num_newton_iter - steps) / target_num_newton_iter
            dofs = new
        increment = min(increment, max_increment)                         # :363
        if multiplier + increment > max_multiplier:                       # :374-378
            increment = max_multiplier - multiplier
    if multiplier < max_multiplier - min_increment and increment < min_increment:
        print("Adaptive load stepping could not converge; increment size below min_increment!")
    return dofs, multiplier, increment, load_step, settings, res_norm
