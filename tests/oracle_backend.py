"""CPU stand-in for autopdex_b200.backend.Plan / DeviceArray that answers with the ORACLE (test infrastructure).

Purpose: the host layer of the product (settings parsing, device-set expansion, coefficient evaluation, hierarchy
construction, the dae time loop, BCOO/index bookkeeping) is Python and can be exercised without a GPU -- every call the
host makes on a plan is recorded or answered by oracle/ on the same data, so `-m "not gpu"` tests check the host logic
end to end (the CUDA kernels behind the real Plan are checked against the same oracle by the `-m gpu` tests).
Never imported by the product."""
import numpy as np

from oracle import assemble as oasm
from oracle import solve as osolve

_ETYPE = {("quad_brick", 2): "line2", ("quad_brick", 3): "line3", ("quad_brick", 4): "quad4", ("quad_brick", 9): "quad9",
          ("quad_brick", 8): "hex8", ("quad_brick", 27): "hex27", ("tri_tet", 3): "tri3", ("tri_tet", 6): "tri6",
          ("tri_tet", 4): "tet4", ("tri_tet", 10): "tet10"}


class FakeDeviceArray:
    def __init__(self, n, dtype=np.float64):
        self.n, self.a = int(n), np.zeros(int(n), dtype=dtype)
        self.ptr = id(self)

    @classmethod
    def from_host(cls, arr, dtype=np.float64):
        out = cls(np.size(arr), dtype)
        out.upload(arr)
        return out

    def upload(self, arr):
        arr = np.asarray(arr, dtype=self.a.dtype).ravel()
        if arr.size != self.n:
            raise ValueError("size mismatch")
        self.a[:] = arr

    def download(self, out=None):
        return self.a.copy()

    def zero(self):
        self.a[:] = 0

    def free(self):
        pass


class FakePlan:
    instances = []

    def __init__(self, dim, n_nodes, nf, sets, dirichlet_mask=None):
        self.dim, self.n_nodes, self.nf, self.sets = int(dim), int(n_nodes), int(nf), list(sets)
        self.n_dofs = self.n_nodes * self.nf
        self.mask = (np.zeros(self.n_dofs, bool) if dirichlet_mask is None else np.asarray(dirichlet_mask, bool).ravel())
        assert self.mask.size == self.n_dofs
        self.n_free = int((~self.mask).sum())
        self.f0, self.f1 = 0, self.n_free
        for st in self.sets:
            assert st.conn.ndim == 2 and st.conn.min() >= 0 and st.conn.max() < self.n_nodes, "connectivity out of range"
            if st.kind != "intpoint":
                st.n_gp = 1 if st.model == "pattern_only" else len(st.gp[1])
        self.n_coo = sum(st.conn.shape[0] * (st.conn.shape[1] * self.nf) ** 2 for st in self.sets)
        self.params = [dict(st.params) for st in self.sets]
        self.coords, self.dt, self.dofs_n = None, None, None
        self.coarse, self.mg_options, self.destroyed, self.history = None, None, False, []
        self.calls = []
        self.device_bytes = 0
        FakePlan.instances.append(self)

    # ---- fields ----
    def set_coords(self, coords):
        c = np.asarray(coords, dtype=np.float64)
        assert c.shape == (self.n_nodes, self.dim)
        self.coords = c.copy()

    def set_param(self, iset, name, value):
        st = self.sets[iset]
        ncomp = self.nf if name in ("body_load", "traction") else 1
        v = np.asarray(value, dtype=np.float64)
        n_rows, n_gp = st.conn.shape[0], (1 if st.kind == "intpoint" else st.n_gp)
        ok = (v.size == ncomp and v.ndim <= 1) or (v.shape in ((n_gp,), (n_gp, ncomp)) and v.size == n_gp * ncomp) \
            or v.size == n_rows * n_gp * ncomp
        if not ok:
            raise ValueError("parameter %r of set %d has shape %s" % (name, iset, v.shape))
        if v.size == n_rows * n_gp * ncomp and v.size != ncomp:
            v = v.reshape((n_rows, n_gp) + ((ncomp,) if ncomp > 1 else ()))
        self.params[iset][name] = v.copy()

    def set_time_increment(self, dt):
        self.dt = float(dt)

    def set_dofs_n(self, dofs_n):
        self.dofs_n = np.asarray(dofs_n, dtype=np.float64).reshape(self.n_nodes, self.nf).copy()

    def set_intpoint_tables(self, iset, N, dNdx, w):
        self.sets[iset].tables = (np.asarray(N), np.asarray(dNdx), np.asarray(w))

    # ---- oracle problem of the current fields ----
    def _problem(self, values=None):
        sets = []
        for i, st in enumerate(self.sets):
            model = dict(name=st.model, mode=st.mode, **{k: v for k, v in self.params[i].items() if v is not None})
            if st.model == "pattern_only":
                sets.append(dict(kind="domain", etype=None, conn=st.conn, nf=self.nf, gp=None, model=model))
            elif st.kind == "intpoint":
                N, dN, w = st.tables
                sets.append(dict(kind="intpoint", conn=st.conn, nf=self.nf, N=N, dNdx=dN, w=w, model=model))
            else:
                sets.append(dict(kind=st.kind, etype=_ETYPE[(st.family, st.conn.shape[1])], conn=st.conn, nf=self.nf,
                                 gp=st.gp, model=model))
        settings = {}
        if self.dt is not None:
            settings["time increment"] = self.dt
        if self.dofs_n is not None:
            settings["dofs n"] = self.dofs_n
        shape = (self.n_nodes, self.nf)
        vals = np.zeros(shape) if values is None else np.asarray(values).reshape(shape)
        return osolve.Problem(sets, self.coords, self.mask.reshape(shape), vals, settings)

    # ---- solves ----
    def newton(self, opts, dofs_d, dirichlet_values_d, newton_tol=1e-8, maxiter=30, damping=1.0):
        prob = self._problem(None if dirichlet_values_d is None else dirichlet_values_d.a)
        self.history = []
        sol, (it, rn, div) = osolve.damped_newton(prob, dofs_d.a.reshape(self.n_nodes, self.nf), newton_tol, maxiter, damping,
                                                  nodal_imposition=dirichlet_values_d is not None, history=self.history)
        dofs_d.a[:] = sol.ravel()
        self.calls.append(("newton", opts.c.jacobi, opts.c.method))
        return it, rn, div

    def linear_step(self, opts, dofs_d, dirichlet_values_d, delta_d):
        prob = self._problem(None if dirichlet_values_d is None else dirichlet_values_d.a)
        delta_d.a[:] = osolve.solve_linear(prob, dofs_d.a.reshape(self.n_nodes, self.nf),
                                           nodal_imposition=dirichlet_values_d is not None).ravel()
        self.calls.append(("linear_step", opts.c.jacobi, opts.c.method))
        return 1

    def assemble(self, dofs_d, want_tangent=True, residual_d=None):
        prob = self._problem()
        R, data = oasm.assemble(prob.sets, prob.coords, dofs_d.a.reshape(self.n_nodes, self.nf), prob.settings, want_tangent)
        if residual_d is not None:
            residual_d.a[:] = R
        if want_tangent:
            self._coo = data
            rows, cols = prob.coo()
            self._csr = oasm.scipy_assembling(data, rows, cols, self.n_dofs)

    def coo_values(self, offset=0, count=None):
        return self._coo[offset:(None if count is None else offset + count)].copy()

    def csr(self, reduced=False):
        m = self._csr if not reduced else self._reduced()
        return m.indptr.astype(np.int64), m.indices.astype(np.int64)

    def values(self, reduced=False):
        return (self._csr if not reduced else self._reduced()).data.copy()

    def _reduced(self):
        free = ~self.mask
        m = self._csr[:, free][free]
        m.sort_indices()
        return m

    # ---- multigrid / bookkeeping ----
    def set_coarse(self, coarse, P, R, inject):
        assert np.asarray(P[0]).size == self.n_free + 1 and np.asarray(R[0]).size == coarse.n_free + 1
        assert np.asarray(inject).size == coarse.n_dofs and np.asarray(P[0])[-1] == np.asarray(R[0])[-1]
        assert np.asarray(P[1]).max() < coarse.n_free and np.asarray(R[1]).max() < self.n_free
        assert np.asarray(inject).max() < self.n_dofs
        self.coarse, self.transfer = coarse, (P, R, np.asarray(inject))

    def set_coarse_structured(self, coarse, dims_f, dims_c, plane_off_f=0, plane_off_c=0):
        """Stand-in of apdx_plan_set_coarse_structured: the host statement of the construction (multigrid.prolongation)."""
        from autopdex_b200 import multigrid
        dims_f, dims_c = [int(v) for v in dims_f], [int(v) for v in dims_c]
        assert int(np.prod(dims_f)) == self.n_nodes and int(np.prod(dims_c)) == coarse.n_nodes
        shape_f = tuple(2 * (n - 1) for n in dims_c)           # global element counts; the slowest direction is never read
        free_f = ~np.asarray(self.mask, dtype=bool).reshape(self.n_nodes, -1)
        free_c = ~np.asarray(coarse.mask, dtype=bool).reshape(coarse.n_nodes, -1)
        nf = free_f.shape[1]
        slab = ((plane_off_f, plane_off_f + dims_f[0]), (plane_off_c, plane_off_c + dims_c[0]))
        P, R = multigrid.prolongation(shape_f, nf, free_f, free_c, slab)
        i0 = np.clip(2 * (np.arange(dims_c[0]) + plane_off_c) - plane_off_f, 0, dims_f[0] - 1)
        grids = np.meshgrid(i0, *[np.arange(0, n, 2) for n in dims_f[1:]], indexing="ij")
        strides = np.cumprod([1] + dims_f[::-1])[::-1][1:]
        nodes = sum(g.ravel().astype(np.int64) * int(st) for g, st in zip(grids, strides))
        self.set_coarse(coarse, P, R, (nodes[:, None] * nf + np.arange(nf)).ravel())

    def set_multigrid(self, pre=0, post=0, coarsest=0, ratio=0.0, coarsest_ratio=0.0):
        self.mg_options = (pre, post, coarsest, ratio, coarsest_ratio)

    def newton_history(self):
        return np.asarray(self.history)

    def stats(self):
        return {"assembly_tangent_ms": 0.0, "assembly_residual_ms": 0.0, "krylov_ms": 0.0, "krylov_iters": 0.0,
                "spmv_launches": 0.0, "total_ms": 0.0, "kernel_launches": 0.0, "sell_bytes": 0.0, "krylov_relres": 0.0,
                "krylov_converged": True}

    def device_bytes_now(self):
        return 0

    def destroy(self):
        self.destroyed = True


def install(monkeypatch):
    """Route the product's host layer to the oracle-backed stand-ins for the duration of one test."""
    from autopdex_b200 import backend, solver
    FakePlan.instances = []
    monkeypatch.setattr(backend, "Plan", FakePlan)
    monkeypatch.setattr(backend, "DeviceArray", FakeDeviceArray)
    solver._PLAN_CACHE.clear()
    return FakePlan
