"""Thin object layer over the C ABI: device arrays and the assembly/solve plan.

Everything numerical happens inside libapdx_b200.so; this module marshals NumPy arrays
into the C structs of include/apdx_b200.h.
"""
import ctypes as C
import weakref

import numpy as np

from . import _lib, spaces


# ---- page-locking of host buffers that are uploaded repeatedly ----------------------------------------
_SEEN, _PINNED = {}, {}
_PIN_MIN_BYTES = 1 << 20


def _unpin(ptr):
    if _PINNED.pop(ptr, None):
        try:
            _lib.load().apdx_host_unregister(C.c_void_p(ptr))
        except Exception:
            pass


def pin_if_repeated(arr):
    """Settings arrays are usually the same NumPy buffers call after call (coordinates, Dirichlet values, the
    initial guess).  The second time a large buffer is uploaded it is page-locked in place (cudaHostRegister), so
    later host->device copies are DMA transfers instead of staged pageable copies; a finalizer unregisters it."""
    if arr.nbytes < _PIN_MIN_BYTES or not arr.flags.c_contiguous:
        return
    ptr = arr.ctypes.data
    key = (ptr, arr.nbytes)
    if ptr in _PINNED:
        return
    _SEEN[key] = _SEEN.get(key, 0) + 1
    if _SEEN[key] == 2:
        base = arr
        while isinstance(base.base, np.ndarray):
            base = base.base
        if _lib.load().apdx_host_register(C.c_void_p(ptr), arr.nbytes) == 0:
            _PINNED[ptr] = True
            try:
                weakref.finalize(base, _unpin, ptr)
            except TypeError:
                pass
        if len(_SEEN) > 4096:
            _SEEN.clear()


# ---- recycled, page-locked host blocks for large results ------------------------------------------------
# A device->host copy into a fresh np.empty pays the page faults of the new pages and a staged pageable copy
# (~10 GB/s); into a block that was page-locked once it is one DMA transfer.  Results are handed out as NumPy arrays
# over such a block; the block goes back to the pool only when the LAST array referring to it is gone: every view
# NumPy derives from the result has the small _Lease object as its .base, and the lease's finalizer recycles the block.
_POOL = {}
_POOL_MAX_BLOCKS = 4


class _Lease:
    def __init__(self, mem, n, dtype):
        self._mem = mem
        self.__array_interface__ = {"data": (mem.ctypes.data, False), "shape": (n,), "typestr": np.dtype(dtype).str,
                                    "version": 3}


def _recycle(nbytes, mem):
    blocks = _POOL.setdefault(nbytes, [])
    if len(blocks) < _POOL_MAX_BLOCKS:
        blocks.append(mem)
    else:
        _unpin(mem.ctypes.data)


def host_result(n, dtype=np.float64):
    """An uninitialised host array of n items for a device->host copy (page-locked and recycled when large)."""
    dtype = np.dtype(dtype)
    nbytes = int(n) * dtype.itemsize
    if nbytes < _PIN_MIN_BYTES:
        return np.empty(n, dtype=dtype)
    blocks = _POOL.get(nbytes)
    if blocks:
        mem = blocks.pop()
    else:
        mem = np.empty(nbytes, dtype=np.uint8)
        ptr = mem.ctypes.data
        if ptr % dtype.itemsize == 0 and _lib.load().apdx_host_register(C.c_void_p(ptr), nbytes) == 0:
            _PINNED[ptr] = True
    lease = _Lease(mem, int(n), dtype)
    weakref.finalize(lease, _recycle, nbytes, mem)
    return np.asarray(lease)


class Stream:
    """A CUDA stream for callers without a CUDA binding of their own (apdx_stream_*)."""

    def __init__(self):
        self.ptr = C.c_void_p()
        _lib.check(_lib.load().apdx_stream_create(C.byref(self.ptr)))

    def synchronize(self):
        _lib.check(_lib.load().apdx_stream_synchronize(self.ptr))

    def destroy(self):
        if self.ptr:
            _lib.load().apdx_stream_destroy(self.ptr)
            self.ptr = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class DeviceArray:
    """A caller-owned FP64 (or raw byte) buffer in HBM."""

    def __init__(self, n, dtype=np.float64):
        self.dtype = np.dtype(dtype)
        self.n = int(n)
        self.ptr = C.c_void_p()
        _lib.check(_lib.load().apdx_malloc(C.byref(self.ptr), max(self.n, 1) * self.dtype.itemsize))

    @classmethod
    def from_host(cls, arr, dtype=np.float64):
        arr = np.ascontiguousarray(arr, dtype=dtype)
        out = cls(arr.size, dtype)
        out.upload(arr)
        return out

    def upload(self, arr):
        arr = np.ascontiguousarray(arr, dtype=self.dtype)
        if arr.size != self.n:
            raise ValueError("size mismatch: buffer %d, array %d" % (self.n, arr.size))
        pin_if_repeated(arr)
        _lib.check(_lib.load().apdx_memcpy_h2d(self.ptr, arr.ctypes.data_as(C.c_void_p), arr.nbytes))

    def download(self, out=None):
        if out is None:
            out = host_result(self.n, self.dtype)
        _lib.check(_lib.load().apdx_memcpy_d2h(out.ctypes.data_as(C.c_void_p), self.ptr, self.n * self.dtype.itemsize))
        return out

    def zero(self):
        _lib.check(_lib.load().apdx_memset(self.ptr, 0, self.n * self.dtype.itemsize))

    def free(self):
        if self.ptr:
            _lib.load().apdx_free(self.ptr)
            self.ptr = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class SetSpec:
    """One connectivity set with its recognised closed-form model.

    kind    'domain' | 'surface' | 'intpoint'
    model   key of _lib.APDX_MODEL
    family  'quad_brick' | 'tri_tet'   (domain / surface)
    conn    (n_rows, nen) integer array
    gp      (xi, w) reference Gauss rule (domain / surface)
    params  {name: array}  name in _lib.APDX_PARAM; shapes () | (ncomp,) const,
            (n_gp,) | (n_gp, ncomp) per Gauss point, (n_rows, n_gp[, ncomp]) per element
    tables  (N, dNdx, w) per integration point (intpoint)
    """

    def __init__(self, kind, model, conn, family=None, gp=None, mode=None, params=None, tables=None):
        self.kind, self.model, self.mode, self.family = kind, model, mode, family
        self.conn = np.ascontiguousarray(conn)
        if self.conn.dtype not in (np.int32, np.int64):
            self.conn = self.conn.astype(np.int64)
        self.gp = gp
        self.params = dict(params or {})
        self.tables = tables


class KrylovOptions:
    def __init__(self, method="cg", rtol=1e-8, atol=0.0, maxiter=0, jacobi=True, check_every=0):
        """jacobi: True / False, or the preconditioner by name: 'jacobi', 'none', 'multigrid' (APDX_PRECOND_*)."""
        if method not in _lib.APDX_KRYLOV:
            raise ValueError("'solver' must be 'cg' or 'bicgstab' for the b200 backend, got %r" % (method,))
        pc = {"none": 0, "jacobi": 1, "multigrid": 2}[jacobi] if isinstance(jacobi, str) else (1 if jacobi else 0)
        self.c = _lib.KrylovOpts(_lib.APDX_KRYLOV[method], int(maxiter), float(rtol), float(atol), pc, int(check_every))


class Plan:
    """Pattern + index maps + device work space for one mesh / model configuration."""

    def __init__(self, dim, n_nodes, nf, sets, dirichlet_mask=None):
        lib = _lib.load()
        if _lib.device_count() < 1:
            raise RuntimeError("autopdex_b200: no CUDA device visible; the b200 backend has no CPU path")
        self.dim, self.n_nodes, self.nf = int(dim), int(n_nodes), int(nf)
        self.sets = list(sets)
        descs = (_lib.SetDesc * len(self.sets))()
        keep = []
        for i, st in enumerate(self.sets):
            d = descs[i]
            d.kind = _lib.APDX_KIND[st.kind]
            d.model = _lib.APDX_MODEL[st.model]
            d.mode = _lib.APDX_MODE[st.mode]
            d.n_rows, d.nen = st.conn.shape
            d.conn_itemsize = st.conn.dtype.itemsize
            d.conn_h = st.conn.ctypes.data_as(C.c_void_p)
            if st.kind == "intpoint":
                d.n_gp, d.dim_ref = 1, dim
            elif st.model == "pattern_only":
                d.n_gp, d.dim_ref = 1, dim          # no arithmetic, no shape tables
                st.n_gp = 1
            else:
                xi, w = st.gp
                dr = dim if st.kind == "domain" else dim - 1
                xi = np.asarray(xi, dtype=np.float64).reshape(-1, dr)
                N, dN = spaces.shape_tables(st.family, d.nen, dr, xi)
                N, dN = np.ascontiguousarray(N), np.ascontiguousarray(dN)
                w = np.ascontiguousarray(w, dtype=np.float64)
                keep += [N, dN, w]
                d.n_gp, d.dim_ref = xi.shape[0], dr
                d.shape_n_h = N.ctypes.data_as(C.c_void_p)
                d.shape_dn_h = dN.ctypes.data_as(C.c_void_p)
                d.gp_w_h = w.ctypes.data_as(C.c_void_p)
                st.n_gp = xi.shape[0]
        mask_p = None
        if dirichlet_mask is not None:
            m = np.ascontiguousarray(np.asarray(dirichlet_mask).ravel(), dtype=np.uint8)
            if m.size != self.n_nodes * self.nf:
                raise ValueError("'dirichlet dofs' has %d entries, expected %d" % (m.size, self.n_nodes * self.nf))
            keep.append(m)
            mask_p = m.ctypes.data_as(C.c_void_p)
        self.h = C.c_void_p()
        _lib.check(lib.apdx_plan_create(C.byref(self.h), self.dim, self.n_nodes, self.nf, len(self.sets), descs, mask_p))
        q = (C.c_int64 * 8)()
        _lib.check(lib.apdx_plan_query(self.h, q))
        (self.n_dofs, self.n_free, self.nnz, self.nnz_reduced, self.n_coo, self.f0, self.f1, self.device_bytes) = list(q)
        for i, st in enumerate(self.sets):
            for name, val in st.params.items():
                self.set_param(i, name, val)
            if st.kind == "intpoint" and st.tables is not None:
                self.set_intpoint_tables(i, *st.tables)

    # -- life cycle ---------------------------------------------------------------------
    def destroy(self):
        if getattr(self, "h", None):
            _lib.load().apdx_plan_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    # -- pattern ------------------------------------------------------------------------
    def csr(self, reduced=False):
        n, nnz = (self.n_free, self.nnz_reduced) if reduced else (self.n_dofs, self.nnz)
        indptr = np.empty(n + 1, dtype=np.int64)
        indices = np.empty(max(nnz, 1), dtype=np.int64)
        _lib.check(_lib.load().apdx_plan_get_csr(self.h, int(reduced), indptr.ctypes.data_as(C.c_void_p),
                                                 indices.ctypes.data_as(C.c_void_p)))
        return indptr, indices[:nnz]

    def elem_map(self, offset=0, count=None):
        count = self.n_coo - offset if count is None else count
        pos = np.empty(count, dtype=np.int64)
        _lib.check(_lib.load().apdx_plan_get_elem_map(self.h, offset, count, pos.ctypes.data_as(C.c_void_p)))
        return pos

    # -- run-time fields ------------------------------------------------------------------
    def set_coords(self, coords):
        c = np.ascontiguousarray(coords, dtype=np.float64)
        if c.shape != (self.n_nodes, self.dim):
            raise ValueError("'node coordinates' must have shape (%d, %d)" % (self.n_nodes, self.dim))
        pin_if_repeated(c)
        _lib.check(_lib.load().apdx_set_coords(self.h, c.ctypes.data_as(C.c_void_p)))

    def set_param(self, iset, name, value):
        st = self.sets[iset]
        ncomp = self.nf if name in ("body_load", "traction") else 1
        v = np.asarray(value, dtype=np.float64)
        n_rows = st.conn.shape[0]
        n_gp = 1 if st.kind == "intpoint" else st.n_gp
        if v.size == ncomp and v.ndim <= 1:
            layout = "const"
        elif v.shape in ((n_gp,), (n_gp, ncomp)) and v.size == n_gp * ncomp:
            layout = "per_gp"
        elif v.size == n_rows * n_gp * ncomp:
            layout = "per_row_gp"
        else:
            raise ValueError("parameter %r of set %d has shape %s; expected (%d,), (%d,[%d]) or (%d,%d,[%d])"
                             % (name, iset, v.shape, ncomp, n_gp, ncomp, n_rows, n_gp, ncomp))
        v = np.ascontiguousarray(v.ravel())
        _lib.check(_lib.load().apdx_set_param(self.h, iset, _lib.APDX_PARAM[name], _lib.APDX_LAYOUT[layout], ncomp,
                                              v.ctypes.data_as(C.c_void_p)))

    def set_intpoint_tables(self, iset, N, dNdx, w):
        N = np.ascontiguousarray(N, dtype=np.float64)
        dNdx = np.ascontiguousarray(dNdx, dtype=np.float64)
        w = np.ascontiguousarray(w, dtype=np.float64)
        _lib.check(_lib.load().apdx_set_intpoint_tables(self.h, iset, N.ctypes.data_as(C.c_void_p),
                                                        dNdx.ctypes.data_as(C.c_void_p), w.ctypes.data_as(C.c_void_p)))

    def set_time_increment(self, dt):
        _lib.check(_lib.load().apdx_set_time_increment(self.h, float(dt)))

    def set_dofs_n(self, dofs_n):
        v = np.ascontiguousarray(np.asarray(dofs_n, dtype=np.float64).ravel())
        _lib.check(_lib.load().apdx_set_dofs_n(self.h, v.ctypes.data_as(C.c_void_p)))

    # -- assembly / solve -----------------------------------------------------------------
    def assemble(self, dofs_d, want_tangent=True, residual_d=None):
        rp = residual_d.ptr if residual_d is not None else None
        _lib.check(_lib.load().apdx_assemble(self.h, dofs_d.ptr, int(want_tangent), rp))

    def values(self, reduced=False):
        nnz = self.nnz_reduced if reduced else self.nnz
        out = np.empty(max(nnz, 1), dtype=np.float64)
        _lib.check(_lib.load().apdx_get_values(self.h, int(reduced), out.ctypes.data_as(C.c_void_p)))
        return out[:nnz]

    def newton_history(self):
        """Residual norms after every iteration of the last newton() call."""
        n = C.c_int32(0)
        buf = np.zeros(64)
        _lib.check(_lib.load().apdx_plan_newton_history(self.h, buf.ctypes.data_as(C.c_void_p), buf.size, C.byref(n)))
        return buf[:min(n.value, buf.size)].copy()

    def device_bytes_now(self):
        """Bytes held by ALL plans of this process right now (apdx_plan_query out[7])."""
        q = (C.c_int64 * 8)()
        _lib.check(_lib.load().apdx_plan_query(self.h, q))
        return int(q[7])

    def set_coarse(self, coarse, P, R, inject):
        """Link `coarse` (a Plan of the same model on the next-coarser mesh) below this plan: P, R = CSR triples
        (indptr int32, indices int32, data) over the reduced dofs, inject = fine full dof of every coarse full dof."""
        arrs = [np.ascontiguousarray(P[0], dtype=np.int32), np.ascontiguousarray(P[1], dtype=np.int32),
                np.ascontiguousarray(P[2], dtype=np.float64), np.ascontiguousarray(R[0], dtype=np.int32),
                np.ascontiguousarray(R[1], dtype=np.int32), np.ascontiguousarray(R[2], dtype=np.float64),
                np.ascontiguousarray(inject, dtype=np.int64)]
        if arrs[0].size != self.n_free + 1 or arrs[3].size != coarse.n_free + 1 or arrs[6].size != coarse.n_dofs:
            raise ValueError("set_coarse: transfer operators do not match the two plans")
        _lib.check(_lib.load().apdx_plan_set_coarse(self.h, coarse.h, *[a.ctypes.data_as(C.c_void_p) for a in arrs]))
        self.coarse = coarse

    def set_coarse_structured(self, coarse, dims_f, dims_c, plane_off_f=0, plane_off_c=0):
        """Link `coarse` below this plan with the transfer operators of a structured hierarchy built on the device
        (coarse node (I, J, K) = fine node (2I, 2J, 2K)); dims_*: local node counts per direction, slowest first;
        plane_off_*: global index of local plane 0 (slab partitions)."""
        df = np.ascontiguousarray(dims_f, dtype=np.int64)
        dc = np.ascontiguousarray(dims_c, dtype=np.int64)
        if df.size != dc.size or df.size not in (2, 3):
            raise ValueError("set_coarse_structured: dims_f / dims_c must have 2 or 3 entries")
        _lib.check(_lib.load().apdx_plan_set_coarse_structured(self.h, coarse.h, int(df.size), df.ctypes.data_as(C.c_void_p),
                                                               dc.ctypes.data_as(C.c_void_p), int(plane_off_f), int(plane_off_c)))
        self.coarse = coarse

    def get_transfer(self, which):
        """(indptr, indices, data) of the linked P (which = 0) or R (which = 1) and the injection map, copied to the host."""
        n_rows, nnz = C.c_int64(0), C.c_int64(0)
        lib = _lib.load()
        _lib.check(lib.apdx_plan_get_transfer(self.h, int(which), C.byref(n_rows), C.byref(nnz), None, None, None, None))
        ptr = np.empty(n_rows.value + 1, dtype=np.int32)
        idx = np.empty(max(nnz.value, 1), dtype=np.int32)
        val = np.empty(max(nnz.value, 1), dtype=np.float64)
        inj = np.empty(self.coarse.n_dofs, dtype=np.int32)
        _lib.check(lib.apdx_plan_get_transfer(self.h, int(which), None, None, ptr.ctypes.data_as(C.c_void_p),
                                              idx.ctypes.data_as(C.c_void_p), val.ctypes.data_as(C.c_void_p),
                                              inj.ctypes.data_as(C.c_void_p)))
        return (ptr, idx[:nnz.value], val[:nnz.value]), inj

    def set_multigrid(self, pre=0, post=0, coarsest=0, ratio=0.0, coarsest_ratio=0.0):
        _lib.check(_lib.load().apdx_plan_set_multigrid(self.h, int(pre), int(post), int(coarsest), float(ratio),
                                                       float(coarsest_ratio)))

    def coo_values(self, offset=0, count=None):
        """Element-tangent entries of the last tangent assembly in the reference's COO order, duplicates not summed
        (the `data` of the BCOO that assembler.assemble_tangent returns, assembler.py:749-777)."""
        count = self.n_coo - offset if count is None else count
        out = np.empty(max(count, 1), dtype=np.float64)
        _lib.check(_lib.load().apdx_get_coo_values(self.h, int(offset), int(count), out.ctypes.data_as(C.c_void_p)))
        return out[:count]

    def spmv(self, x_d, y_d):
        _lib.check(_lib.load().apdx_spmv(self.h, x_d.ptr, y_d.ptr))

    def comm_info(self):
        """How this partitioned plan communicates (apdx_plan_comm_info)."""
        q = (C.c_double * 4)()
        _lib.check(_lib.load().apdx_plan_comm_info(self.h, q))
        return {"allreduce": "mailbox" if q[0] else "nccl", "halo": "inbox" if q[1] else "nccl",
                "halo_inbox_us": q[2], "halo_nccl_us": q[3]}

    # -- caller-owned streams: the plan enqueues on `stream` (a Stream, a raw cudaStream_t, or None = its own)
    def set_stream(self, stream):
        raw = stream.ptr if isinstance(stream, Stream) else stream
        _lib.check(_lib.load().apdx_plan_set_stream(self.h, raw))

    def assemble_async(self, dofs_d, want_tangent=True, residual_d=None):
        rp = residual_d.ptr if residual_d is not None else None
        _lib.check(_lib.load().apdx_assemble_async(self.h, dofs_d.ptr, int(want_tangent), rp))

    def spmv_async(self, x_d, y_d):
        _lib.check(_lib.load().apdx_spmv_async(self.h, x_d.ptr, y_d.ptr))

    def krylov(self, opts, rhs_d, x_d):
        it, rr = C.c_int32(0), C.c_double(0.0)
        _lib.check(_lib.load().apdx_krylov(self.h, C.byref(opts.c), rhs_d.ptr, x_d.ptr, C.byref(it), C.byref(rr)))
        return it.value, rr.value

    def linear_step(self, opts, dofs_d, dirichlet_values_d, delta_d):
        it = C.c_int32(0)
        dv = dirichlet_values_d.ptr if dirichlet_values_d is not None else None
        _lib.check(_lib.load().apdx_linear_step(self.h, C.byref(opts.c), dofs_d.ptr, dv, delta_d.ptr, C.byref(it)))
        return it.value

    def tangent_solve(self, opts, dofs_d, rhs_d, out_d, transpose=False):
        """out[free] = K(dofs)^-1 rhs[free] (K^-T with transpose; the in-scope tangents are symmetric), zeros on Dirichlet
        dofs.  dofs_d=None reuses the tangent of the last assembly (implicit_diff.py:225-234)."""
        it = C.c_int32(0)
        _lib.check(_lib.load().apdx_tangent_solve(self.h, C.byref(opts.c), dofs_d.ptr if dofs_d is not None else None,
                                                  rhs_d.ptr, int(bool(transpose)), out_d.ptr, C.byref(it)))
        return it.value

    def newton(self, opts, dofs_d, dirichlet_values_d, newton_tol=1e-8, maxiter=30, damping=1.0):
        it, rn, dv = C.c_int32(0), C.c_double(0.0), C.c_int32(0)
        dvals = dirichlet_values_d.ptr if dirichlet_values_d is not None else None
        _lib.check(_lib.load().apdx_newton(self.h, C.byref(opts.c), dofs_d.ptr, dvals, float(newton_tol), int(maxiter),
                                           float(damping), C.byref(it), C.byref(rn), C.byref(dv)))
        return it.value, rn.value, bool(dv.value)

    def time_spmv(self, reps=20):
        ms = C.c_double(0.0)
        _lib.check(_lib.load().apdx_time_spmv(self.h, int(reps), C.byref(ms)))
        return ms.value

    def stats(self):
        out = (C.c_double * 8)()
        _lib.check(_lib.load().apdx_plan_stats(self.h, out))
        keys = ("assembly_tangent_ms", "assembly_residual_ms", "krylov_ms", "krylov_iters", "spmv_launches",
                "total_ms", "kernel_launches", "sell_bytes")
        d = dict(zip(keys, list(out)[:8]))
        rr, ok = C.c_double(0.0), C.c_int32(1)
        _lib.check(_lib.load().apdx_plan_last_krylov(self.h, C.byref(rr), C.byref(ok)))
        d["krylov_relres"], d["krylov_converged"] = rr.value, bool(ok.value)
        return d

    def sell_info(self):
        out = (C.c_int64 * 6)()
        _lib.check(_lib.load().apdx_plan_sell_info(self.h, out))
        return dict(zip(("slices", "stored_values", "index_ints", "mirrored_entries", "symmetric", "dofs_per_node"), list(out)))

    def set_partition(self, owned_dof_begin, owned_dof_end, rank_lo=-1, rank_hi=-1):
        _lib.check(_lib.load().apdx_plan_set_partition(self.h, int(owned_dof_begin), int(owned_dof_end), int(rank_lo),
                                                       int(rank_hi)))
        q = (C.c_int64 * 8)()
        _lib.check(_lib.load().apdx_plan_query(self.h, q))
        self.f0, self.f1 = q[5], q[6]

    def set_partition_lists(self, owned_dof_end, neighbour_ranks, send_dofs, recv_dof_ranges):
        """General partition (mesher.rcb_partition): local dofs are [owned | ghosts grouped by owner]; send_dofs[i] are
        the local dof ids neighbour i ghosts (in its ghost order), recv_dof_ranges[i] = (begin, end) of its block here."""
        nn = len(neighbour_ranks)
        if len(send_dofs) != nn or len(recv_dof_ranges) != nn:
            raise ValueError("one send list and one ghost range per neighbour expected")
        ranks = np.ascontiguousarray(neighbour_ranks, dtype=np.int32)
        lists = [np.ascontiguousarray(v, dtype=np.int64).ravel() for v in send_dofs]
        ptr = np.zeros(nn + 1, dtype=np.int64)
        ptr[1:] = np.cumsum([v.size for v in lists])
        flat = np.concatenate(lists) if nn and ptr[-1] else np.zeros(1, dtype=np.int64)
        rb = np.ascontiguousarray([r[0] for r in recv_dof_ranges], dtype=np.int64)
        re = np.ascontiguousarray([r[1] for r in recv_dof_ranges], dtype=np.int64)
        vp = lambda a: a.ctypes.data_as(C.c_void_p) if a.size else None
        _lib.check(_lib.load().apdx_plan_set_partition_lists(self.h, int(owned_dof_end), nn, vp(ranks), vp(ptr), vp(flat),
                                                             vp(rb), vp(re)))
        q = (C.c_int64 * 8)()
        _lib.check(_lib.load().apdx_plan_query(self.h, q))
        self.f0, self.f1 = q[5], q[6]


# ---- multi-GPU plumbing (one process per GPU) ---------------------------------------------------------
def comm_unique_id():
    buf = (C.c_uint8 * 128)()
    _lib.check(_lib.load().apdx_comm_unique_id(buf))
    return bytes(buf)


def comm_init(unique_id, rank, nranks):
    buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
    _lib.check(_lib.load().apdx_comm_init(buf, int(rank), int(nranks)))


def comm_destroy():
    _lib.check(_lib.load().apdx_comm_destroy())


def comm_allreduce_host(values, op="sum"):
    v = np.ascontiguousarray(values, dtype=np.float64).copy()
    _lib.check(_lib.load().apdx_comm_allreduce_host(v.ctypes.data_as(C.c_void_p), v.size, 1 if op == "max" else 0))
    return v


def comm_exchange_planes(arr, lo_count, hi_count, rank_lo, rank_hi):
    """Slab neighbours swap the boundary blocks of a host array laid out [ghost_lo | owned | ghost_hi] (flattened; counts
    in items): returns a float64 copy whose ghost blocks hold the neighbours' adjacent owned blocks
    (apdx_comm_exchange_planes).  Without a communicator the array comes back unchanged."""
    v = np.ascontiguousarray(arr, dtype=np.float64)
    d = DeviceArray.from_host(v.ravel())
    _lib.check(_lib.load().apdx_comm_exchange_planes(d.ptr, v.size, int(lo_count), int(hi_count), int(rank_lo), int(rank_hi)))
    out = np.array(d.download()).reshape(v.shape)
    d.free()
    return out


def measure_fp64_peak():
    tf = C.c_double(0.0)
    _lib.check(_lib.load().apdx_measure_fp64_peak(C.byref(tf)))
    return tf.value


def set_device(index):
    _lib.check(_lib.load().apdx_set_device(int(index)))
