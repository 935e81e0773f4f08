/* apdx_b200.h -- C ABI of libapdx_b200.so, the B200 (sm_100a) backend for the AutoPDEx
 * hot path "sparse residual/tangent assembly -> Newton linear solve".
 *
 * This is the drop-in boundary: the entry points below are what a
 * `'solver backend': 'b200'` branch inside AutoPDEx's solver.solver
 * (autopdex/solver.py:41-137, backend read at solver.py:576) binds through jax.ffi /
 * ctypes.  INTEGRATION.md shows the reference-side stub.  Each entry point cites the
 * reference function it replaces (paths relative to the AutoPDEx v1.1.4 tree).
 *
 * Conventions
 *   - every function returns APDX_OK (0) or a negative error code; the message of the
 *     last error of the calling thread is available from apdx_last_error();
 *   - pointers suffixed _h are HOST pointers, pointers suffixed _d are DEVICE pointers
 *     owned by the caller (allocate them with apdx_malloc or any CUDA allocator);
 *   - all reals are FP64, dof/flat vectors use the reference's global numbering
 *     gid = node*nf + comp (assembler.py:130, utility.dict_flatten utility.py:104-128);
 *   - a plan is not thread-safe; the library is re-entrant across plans;
 *   - there is NO CPU fallback: every compute entry point needs a CUDA device.
 */
#ifndef APDX_B200_H
#define APDX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define APDX_ABI_VERSION 1

enum {
  APDX_OK = 0,
  APDX_ERR_INVALID = -1,     /* bad argument / inconsistent description */
  APDX_ERR_CUDA = -2,        /* CUDA runtime error (message has the detail) */
  APDX_ERR_UNSUPPORTED = -3, /* model / element outside the supported table: rejected, never emulated */
  APDX_ERR_NOMEM = -4,
  APDX_ERR_NCCL = -5,
  APDX_ERR_STATE = -6        /* call order violated (e.g. solve before assemble) */
};

/* kind of a connectivity set (static_settings['assembling mode'][set]) */
enum {
  APDX_SET_DOMAIN = 0,   /* isoparametric domain element: 'user element' / 'user potential'
                            (models.py:1616-1720, 1188-1269) */
  APDX_SET_SURFACE = 1,  /* isoparametric surface element (models.py:1723-1850) */
  APDX_SET_INTPOINT = 2  /* 'sparse' mode: one row per integration point (assembler.py:874-1035) */
};

/* closed-form models */
enum {
  APDX_MODEL_POISSON_POTENTIAL = 0, /* Pi = 1/2 c |grad phi|^2 - f phi  (README integrand) */
  APDX_MODEL_POISSON_WEAK = 1,      /* models.poisson_weak, models.py:96-134 */
  APDX_MODEL_LINEAR_ELASTICITY = 2, /* models.linear_elasticity_weak, models.py:510-635 */
  APDX_MODEL_NEO_HOOKE = 3,         /* hyperelastic_steady_state_weak + neo_hooke, models.py:917-1000,1122-1146 */
  APDX_MODEL_NEUMANN = 4,           /* models.neumann_weak, models.py:744-779 */
  APDX_MODEL_CAPACITY = 5,          /* models.forward_backward_euler_weak, models.py:1946-2010 */
  APDX_MODEL_PATTERN_ONLY = 6       /* structural entries only (values and residual are zero): the all-pairs block an element
                                       of a MULTI-FIELD dict-dof problem emits for the field pairs its integrand does not
                                       couple (assembler._get_indices, assembler.py:79-117: every (field_i, field_j) block
                                       is part of the BCOO, jacfwd of an uncoupled pair is an explicit zero block).  nen up
                                       to 64, no shape tables, APDX_SET_DOMAIN. */
};

enum { APDX_MODE_NONE = 0, APDX_MODE_PLAIN_STRAIN = 1, APDX_MODE_PLAIN_STRESS = 2, APDX_MODE_3D = 3,
       /* linear elasticity with the isotropic tensor lam 1x1 + 2 mu I_sym in the mesh's dimension: the first Piola-Kirchhoff
        * stress of models.linear_elastic_strain_energy inside hyperelastic_steady_state_weak (models.py:917-1000, 1167-1185);
        * in 2-D the customary plane-strain matrix, which differs from APDX_MODE_PLAIN_STRAIN = the matrix of
        * linear_elasticity_weak with its doubled shear entry (models.py:570-577) */
       APDX_MODE_LAME = 4 };

/* run-time parameters of a set (values of the Python coefficient callables, evaluated by the host) */
enum {
  APDX_PARAM_COEFFICIENT = 0, /* c        (1 comp)  */
  APDX_PARAM_SOURCE = 1,      /* f        (1 comp)  */
  APDX_PARAM_YOUNGS = 2,      /* E        (1 comp)  */
  APDX_PARAM_POISSON_RATIO = 3, /* nu     (1 comp)  */
  APDX_PARAM_BODY_LOAD = 4,   /* b        (nf comps) */
  APDX_PARAM_TRACTION = 5,    /* t        (nf comps) */
  APDX_PARAM_COUNT = 6
};
enum {
  APDX_LAYOUT_CONST = 0,       /* [ncomp]                      */
  APDX_LAYOUT_PER_GP = 1,      /* [n_gp][ncomp]   (same for every element) */
  APDX_LAYOUT_PER_ROW_GP = 2   /* [n_rows][n_gp][ncomp]        */
};

enum { APDX_KRYLOV_CG = 0, APDX_KRYLOV_BICGSTAB = 1 };
enum { APDX_PRECOND_NONE = 0, APDX_PRECOND_JACOBI = 1, APDX_PRECOND_MULTIGRID = 2 };

typedef struct apdx_plan apdx_plan;

/* One entry of settings['connectivity'] together with the recognised model of
 * static_settings['model'][set].  All pointers are host pointers, read during
 * apdx_plan_create only. */
typedef struct {
  int32_t kind;          /* APDX_SET_* */
  int32_t model;         /* APDX_MODEL_* */
  int32_t mode;          /* APDX_MODE_* (elasticity / neo-Hooke) */
  int32_t nen;           /* nodes per row of conn */
  int32_t n_gp;          /* Gauss points per element (1 for APDX_SET_INTPOINT) */
  int32_t dim_ref;       /* reference dimension of the element (dim, or dim-1 for surfaces) */
  int32_t conn_itemsize; /* 4 (int32) or 8 (int64, the reference's index dtype) */
  int32_t reserved;
  int64_t n_rows;        /* elements, or integration points for APDX_SET_INTPOINT */
  const void *conn_h;    /* [n_rows][nen] node ids */
  const double *shape_n_h;  /* [n_gp][nen] shape values at the Gauss points (domain/surface) */
  const double *shape_dn_h; /* [n_gp][nen][dim_ref] reference gradients (domain/surface) */
  const double *gp_w_h;     /* [n_gp] reference weights (domain/surface) */
} apdx_set_desc;

typedef struct {
  int32_t method;    /* APDX_KRYLOV_* (static_settings['solver']: 'cg' | 'bicgstab') */
  int32_t maxiter;   /* kwargs maxiter of linear_solve_jax, solver.py:1116 */
  double rtol;       /* ||r|| <= max(rtol*||b||, atol), as jax.scipy.sparse.linalg.cg */
  double atol;
  int32_t jacobi;    /* APDX_PRECOND_*: 1 = 'type of preconditioner': 'jacobi' (solver.py:1093-1099), 0 = none,
                        2 = multigrid V-cycle over the hierarchy linked with apdx_plan_set_coarse (method must be CG) */
  int32_t check_every; /* iterations between host convergence polls (0 = default) */
} apdx_krylov_opts;

/* ---- library / device ------------------------------------------------------------- */
int apdx_abi_version(void);
const char *apdx_last_error(void);
int apdx_device_count(int *count);
int apdx_set_device(int device);
int apdx_malloc(void **ptr_d, size_t bytes);
int apdx_free(void *ptr_d);
int apdx_host_alloc(void **ptr_h, size_t bytes);   /* pinned */
int apdx_host_free(void *ptr_h);
/* page-lock an existing host buffer (cudaHostRegister) so that repeated uploads of the same settings array run as
 * true DMA; the Python layer does this for buffers it sees twice */
int apdx_host_register(void *ptr_h, size_t bytes);
int apdx_host_unregister(void *ptr_h);
int apdx_memcpy_h2d(void *dst_d, const void *src_h, size_t bytes);
int apdx_memcpy_d2h(void *dst_h, const void *src_d, size_t bytes);
int apdx_memset(void *dst_d, int value, size_t bytes);
int apdx_synchronize(void);
int apdx_mem_info(size_t *free_bytes, size_t *total_bytes);

/* ---- plan: pattern, element->CSR map, Dirichlet maps -------------------------------- *
 * Replaces assembler._get_indices (assembler.py:47-141) + solver.scipy_assembling
 * (solver.py:1180-1222): the COO (row,col) stream of all sets is sorted/uniqued ON DEVICE
 * once; the result is the full CSR pattern, the reduced pattern csr[:,free][free], the
 * element-local -> CSR position map, and the gather lists of the deterministic scatter.
 * dirichlet_mask_h: [n_nodes*nf] bytes (settings['dirichlet dofs'] flattened) or NULL.   */
int apdx_plan_create(apdx_plan **plan, int32_t dim, int64_t n_nodes, int32_t nf, int32_t n_sets,
                     const apdx_set_desc *sets, const uint8_t *dirichlet_mask_h);
int apdx_plan_destroy(apdx_plan *plan);
/* out[0]=n_dofs out[1]=n_free out[2]=nnz_full out[3]=nnz_reduced out[4]=n_coo (nse)
 * out[5]=owned_free_begin out[6]=owned_free_end out[7]=device bytes held by the plan */
int apdx_plan_query(const apdx_plan *plan, int64_t out[8]);
/* CSR pattern as int64 (the reference index dtype): indptr [n+1], indices [nnz] */
int apdx_plan_get_csr(const apdx_plan *plan, int reduced, int64_t *indptr_h, int64_t *indices_h);
/* pos[k] = index in full-CSR data of COO entry k, k in [offset, offset+count) */
int apdx_plan_get_elem_map(const apdx_plan *plan, int64_t offset, int64_t count, int64_t *pos_h);

/* ---- run-time fields (everything that lives in the traced `settings` dict) ----------- */
int apdx_set_coords(apdx_plan *plan, const double *coords_h);            /* [n_nodes][dim] */
int apdx_set_param(apdx_plan *plan, int32_t set, int32_t param, int32_t layout, int32_t ncomp,
                   const double *values_h);
/* per-row tables of an APDX_SET_INTPOINT set: N [n_rows][nen], dNdx [n_rows][nen][dim], w [n_rows]
 * (settings['compiled shape functions'][set], settings['integration weights'][set]) */
int apdx_set_intpoint_tables(apdx_plan *plan, int32_t set, const double *n_h, const double *dndx_h,
                             const double *w_h);
int apdx_set_time_increment(apdx_plan *plan, double dt);                 /* settings['time increment'] */
int apdx_set_dofs_n(apdx_plan *plan, const double *dofs_n_h);            /* settings['dofs n'] */

/* ---- assembly ------------------------------------------------------------------------ *
 * Replaces assembler.assemble_residual / assemble_tangent (assembler.py:587-637,682-777)
 * followed by the duplicate summation of solver.scipy_assembling.  residual_d [n_dofs].
 * With want_tangent the summed values are kept inside the plan (full and reduced CSR).    */
int apdx_assemble(apdx_plan *plan, const double *dofs_d, int want_tangent, double *residual_d);
int apdx_get_values(const apdx_plan *plan, int reduced, double *values_h); /* [nnz] after assemble */
/* The BCOO wire format of assembler.assemble_tangent (autopdex/assembler.py:749-777, data = the flattened element
 * tangents of assembler.py:1383, duplicates NOT summed): values of the element-local pairs [offset, offset+count) of
 * the last tangent assembly in the reference's COO order (set-major, element-major, local row, local column). */
int apdx_get_coo_values(apdx_plan *plan, int64_t offset, int64_t count, double *values_h);

/* ---- linear algebra on the assembled reduced system ---------------------------------- */
int apdx_spmv(apdx_plan *plan, const double *x_d, double *y_d);           /* [n_free] each */
/* Jacobi-preconditioned CG / BiCGSTAB on the reduced system (device analogue of
 * solver.linear_solve_jax, solver.py:1093-1126).  x_d is the initial guess and the result. */
int apdx_krylov(apdx_plan *plan, const apdx_krylov_opts *opts, const double *rhs_d, double *x_d,
                int32_t *iters, double *relres);

/* ---- the Newton hot loop ------------------------------------------------------------- *
 * apdx_linear_step = solver.solve_linear with nodal imposition (solver.py:586-656):
 * delta_d gets the MIXED vector (free entries = Newton increment, Dirichlet entries = imposed
 * values).  apdx_newton = solver.damped_newton (solver.py:837-948) with identical loop
 * semantics; dirichlet_values_d [n_dofs] (only masked entries are read).                  */
int apdx_linear_step(apdx_plan *plan, const apdx_krylov_opts *opts, const double *dofs_d,
                     const double *dirichlet_values_d, double *delta_d, int32_t *krylov_iters);
int apdx_newton(apdx_plan *plan, const apdx_krylov_opts *opts, double *dofs_d,
                const double *dirichlet_values_d, double newton_tol, int32_t maxiter, double damping,
                int32_t *iters, double *res_norm, int32_t *diverged);
/* Linear solve with the tangent for sensitivities (row N3 of SURVEY.md 8f): the device analogue of
 * `solve_fun(mat, rhs, free_dofs_flat)` / `solve_fun(mat.T, ...)` inside implicit_diff._root_vjp / _root_jvp
 * (implicit_diff.py:139-183, 225-234, 274-304) with mat = tangent at dofs_d:
 *   out[free] = K[free][:, free]^-1 (or ^-T) rhs[free],  out[dirichlet] = 0   (utility.mask_op(zeros, free, u_f, 'set')).
 * dofs_d == NULL reuses the tangent of the plan's last assembly.  transpose != 0 asks for the transposed solve: every
 * in-scope tangent is symmetric (that is what the symmetric sliced-ELL storage relies on), so it is the same solve.   */
int apdx_tangent_solve(apdx_plan *plan, const apdx_krylov_opts *opts, const double *dofs_d, const double *rhs_d,
                       int transpose, double *out_d, int32_t *krylov_iters);
/* ---- multigrid preconditioner (row N4 of SURVEY.md 8f) --------------------------------------------------------------
 * The reference's stronger-than-Jacobi preconditioners are algebraic multigrid through pyamg (autopdex/solver.py:1399-1491)
 * and PETSc's pc types (solver.py:1224-1333), both on the host.  Here: a geometric V-cycle on a hierarchy of plans of the
 * same model on coarser meshes (re-discretised coarse operators, assembled by the same element kernels at the injected
 * state), Chebyshev-accelerated Jacobi smoothing, used as the preconditioner of CG (APDX_PRECOND_MULTIGRID).
 * apdx_plan_set_coarse links `coarse` as the next-coarser level of `fine`:
 *   P (prolongation, CSR over the REDUCED dofs: n_free(fine) rows x n_free(coarse) columns), R = P^T (CSR, n_free(coarse)
 *   rows), inject[d] = the fine full dof id that coincides with coarse full dof d (state transfer, n_dofs(coarse) entries).
 * The coarse plan then runs on the fine plan's stream; it must outlive the fine plan's solves and is not owned by it.
 * apdx_plan_set_multigrid: Chebyshev degree of the pre- and post-smoother, of the coarsest-level solve, and the ratios
 * lambda_max / lambda_min of the smoothing and coarsest intervals (values <= 0 keep the defaults 2, 2, 12, 3, 40).
 * Multi-GPU: every level is a slab-partitioned plan (apdx_plan_set_partition on the fine AND the coarse plans, the rank that
 * owns fine node plane 2I owns coarse plane I); P, R and inject are then given in the LOCAL numberings of the two slabs
 * (inject entries of a coarse ghost plane whose fine plane is not local may point at any local dof: the neighbour's values
 * replace them).  Halo exchanges per product and per transfer, all-reduced dot products (csrc/multigrid.cu).            */
int apdx_plan_set_coarse(apdx_plan *fine, apdx_plan *coarse, const int32_t *p_indptr_h, const int32_t *p_indices_h,
                         const double *p_data_h, const int32_t *r_indptr_h, const int32_t *r_indices_h,
                         const double *r_data_h, const int64_t *inject_h);
/* The same link for STRUCTURED hierarchies (coarse node (I, J, K) = fine node (2I, 2J, 2K): the meshes of
 * mesher.structured_mesh 'quad' / 'brick'), with P, R and inject built on the device from the two plans' own Dirichlet
 * maps -- no host construction, no upload (256^3: 22 s of NumPy + 3.4 GB of transfers otherwise).  dims_f / dims_c [dim]:
 * node counts per direction of the two LOCAL meshes, slowest-varying direction first; plane_off_f / plane_off_c: global
 * index of local plane 0 along that direction (slab partitions; 0 on one GPU).  Every other direction must satisfy
 * dims_f = 2 (dims_c - 1) + 1.  The operators equal those of the host construction bit for bit (tests).                */
int apdx_plan_set_coarse_structured(apdx_plan *fine, apdx_plan *coarse, int32_t dim, const int64_t *dims_f,
                                    const int64_t *dims_c, int64_t plane_off_f, int64_t plane_off_c);
/* test / inspection hook: copies of the linked transfer operators (which = 0: P, 1: R) into host arrays; pass NULL
 * pointers to query the sizes only (n_rows, nnz).                                                                       */
int apdx_plan_get_transfer(const apdx_plan *fine, int32_t which, int64_t *n_rows, int64_t *nnz, int32_t *indptr_h,
                           int32_t *indices_h, double *data_h, int32_t *inject_h);
int apdx_plan_set_multigrid(apdx_plan *plan, int32_t pre_degree, int32_t post_degree, int32_t coarsest_degree,
                            double smoother_ratio, double coarsest_ratio);
/* timings (ms, CUDA events) and counters of the last apdx_newton / apdx_linear_step:
 * out[0]=assembly(tangent+residual) out[1]=assembly(residual only) out[2]=krylov
 * out[3]=krylov iterations out[4]=spmv launches out[5]=total out[6]=kernel launches
 * out[7]=bytes of the sliced-ELL matrix the SpMV streams (values + compressed indices)    */
int apdx_plan_stats(const apdx_plan *plan, double out[8]);
/* residual norms after every iteration of the last apdx_newton: what solver.damped_newton prints with verbose > 0
 * ("Residual after Newton iteration {i}: {res}", autopdex/solver.py:906-909).  count = iterations run; at most
 * `capacity` values are written. */
int apdx_plan_newton_history(const apdx_plan *plan, double *res_norms, int32_t capacity, int32_t *count);
/* outcome of the plan's LAST Krylov solve (inside apdx_newton: of the last Newton step): relative residual
 * ||b - A x|| / ||b|| of the recurrence and whether the stopping rule of jax.scipy.sparse.linalg.cg / bicgstab
 * (||r|| <= max(rtol ||b||, atol), solver.py:1116-1126) was met -- 0 when the loop ended on maxiter or broke down.
 * The reference's jax solvers return info = None (no signal at all); the host layer warns. */
int apdx_plan_last_krylov(const apdx_plan *plan, double *relres, int32_t *converged);

/* layout of the sliced-ELL copy of the reduced matrix (built by the first assembly that needs it):
 * out[0]=slices out[1]=stored values (incl. padding) out[2]=index ints (offsets, mirror tables, explicit columns)
 * out[3]=entries read from their transposed position instead of being stored (symmetric storage; 0 with
 * APDX_SELL_SYM=0) out[4]=1 if symmetric storage is enabled out[5]=dofs per node the slices interleave */
int apdx_plan_sell_info(const apdx_plan *plan, int64_t out[6]);

/* ---- caller-owned streams ------------------------------------------------------------------------------
 * A plan enqueues all of its work on ONE stream: its own (default) or, after apdx_plan_set_stream, the caller's
 * (`cuda_stream` is a cudaStream_t; NULL goes back to the plan's own).  The entry points above finish with a host
 * synchronisation of that stream (the Newton / Krylov loops need their convergence scalars on the host anyway);
 * apdx_assemble_async and apdx_spmv_async are the same operations enqueued only: results are valid once the stream has
 * reached that point, no timing statistics are recorded.  An XLA custom call (csrc/xla) hands over XLA's stream this way
 * instead of blocking it.  apdx_stream_*: a stream for callers without a CUDA binding of their own (ctypes).            */
int apdx_plan_set_stream(apdx_plan *plan, void *cuda_stream);
int apdx_assemble_async(apdx_plan *plan, const double *dofs_d, int want_tangent, double *residual_d);
int apdx_spmv_async(apdx_plan *plan, const double *x_d, double *y_d);
int apdx_stream_create(void **cuda_stream);
int apdx_stream_synchronize(void *cuda_stream);
int apdx_stream_destroy(void *cuda_stream);

/* average device time (ms, CUDA events on the plan's stream) of `reps` launches of the SpMV kernel
 * exactly as the CG loop launches it (fused p.Ap dot product included) */
int apdx_time_spmv(apdx_plan *plan, int32_t reps, double *ms_avg);

/* FP64 FMA peak of the current device (TFLOP/s), measured with a register-resident DFMA kernel and CUDA events:
 * the denominator of the assembly kernels' FP64 roofline (MEASURED_PEAKS.json only holds bf16 and HBM figures) */
int apdx_measure_fp64_peak(double *tflops);

/* ---- multi-GPU (one process per GPU, slab partition) -------------------- *
 * Each rank builds a plan of its LOCAL mesh (owned nodes plus one ghost plane per side,
 * local ids in global order).  owned dofs are the contiguous range [begin,end) of local
 * dof ids; rank_lo/rank_hi are the neighbour ranks (-1: none).  Halo planes travel with
 * ncclSend/ncclRecv, dot products and the Newton norm with ncclAllReduce.                 */
int apdx_comm_unique_id(uint8_t id_out[128]);
int apdx_comm_init(const uint8_t id[128], int32_t rank, int32_t nranks);
int apdx_comm_destroy(void);
/* host-side collective for scalars (timings, counters): op 0 = sum, 1 = max; no-op without a communicator */
int apdx_comm_allreduce_host(double *inout_h, int32_t count, int32_t op);
int apdx_plan_set_partition(apdx_plan *plan, int64_t owned_dof_begin, int64_t owned_dof_end,
                            int32_t rank_lo, int32_t rank_hi);
/* How a partitioned plan communicates (decided when the partition is set): out[0] = 1 if the dot-product all-reduces go
 * through the peer-memory mailboxes (else ncclAllReduce), out[1] = 1 if the halo exchange goes through the peer inboxes
 * (else ncclSend/ncclRecv), out[2], out[3] = microseconds per halo exchange through the inboxes / through NCCL measured at
 * set-up (maximum over the ranks; the faster one is kept unless APDX_HALO=nccl|inbox says otherwise; 0 = not measured). */
int apdx_plan_comm_info(const apdx_plan *plan, double out[4]);
/* Slab neighbours swap boundary blocks of a device vector of n doubles laid out [ghost_lo | owned | ghost_hi]: the first
 * lo_count owned entries go to rank_lo, whose answer fills [0, lo_count); the last hi_count owned entries go to rank_hi,
 * whose answer fills [n - hi_count, n).  Collective between neighbours (both sides pass the size of one node plane);
 * used for the node fields of a partitioned multigrid hierarchy (coordinates, masks, injected state).  Synchronous. */
int apdx_comm_exchange_planes(double *buf_d, int64_t n, int64_t lo_count, int64_t hi_count, int32_t rank_lo, int32_t rank_hi);
/* General partition (recursive coordinate bisection, unstructured meshes): the local mesh numbers the owned nodes
 * first ([0, owned_dof_end) in dof units) and then the ghost nodes grouped by owning rank.  For neighbour i
 * (neighbour_rank[i]) send_dof_h[send_ptr_h[i] .. send_ptr_h[i+1]) lists the LOCAL dof ids (owned, ascending in the
 * order the neighbour stores them as ghosts) whose values it needs, and the ghost dofs it owns are the local dof range
 * [recv_dof_begin_h[i], recv_dof_end_h[i]).  Dirichlet dofs are dropped from both sides (the masks of a shared node must
 * agree across ranks).  The halo exchange is then a pack kernel + one grouped ncclSend/ncclRecv per neighbour.        */
int apdx_plan_set_partition_lists(apdx_plan *plan, int64_t owned_dof_end, int32_t n_neighbours,
                                  const int32_t *neighbour_rank, const int64_t *send_ptr_h, const int64_t *send_dof_h,
                                  const int64_t *recv_dof_begin_h, const int64_t *recv_dof_end_h);

#ifdef __cplusplus
}
#endif
#endif /* APDX_B200_H */
