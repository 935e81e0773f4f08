// Element residual / tangent kernels (closed-form FP64 on the FMA pipes) and the
// deterministic gather-reduce scatter.
//
// Device replacement of the vmap'ed AD element routines of the reference:
//   user_element_assemble_* (assembler.py:1310-1389) over
//   models.isoparametric_domain_element_galerkin / _surface_element_galerkin (models.py:1616-1850),
//   user_potential_assemble_* (assembler.py:1064-1149) over mixed_reference_domain_potential
//   (models.py:1188-1269), sparse_assemble_* (assembler.py:874-1035) over
//   variational_schemes.weak_form_galerkin (variational_schemes.py:185-252),
// with the weak forms of models.py:96-134 (poisson_weak), :510-635 (linear_elasticity_weak),
// :744-779 (neumann_weak), :917-1000 + :1122-1146 (neo-Hooke), :1946-2010 (Euler capacity term),
// and the scatter-add of assembler._get_residual (assembler.py:331-436) / the duplicate summing of
// solver.scipy_assembling (solver.py:1207-1211) replaced by a fixed-order segmented reduction
// through the sort-built gather lists (no float atomics).
//
// Generic kernel: a CTA stages a batch of elements in shared memory --
//   phase A  gather nodal coordinates / dofs through the connectivity,
//   phase B  one thread per (element, Gauss point): Jacobian, signed det, physical gradients
//            G_a = dN_a J^-1 (kept in shared memory) and the constitutive point data,
//   phase C  one thread per (element, node pair): nf x nf tangent block summed over Gauss points,
//   phase D  one thread per (element, node): residual entries.
#include "elements.cuh"

namespace apdx {

template <int DIM, int NF>
__global__ void __launch_bounds__(256) k_elements(ElemArgs A) {
  extern __shared__ double sm[];
  constexpr int NPT = (NF == 1) ? NPT_SCALAR : NPT_VECTOR;
  const int nen = A.nen, n_gp = A.n_gp, dr = A.dim_ref, epb = A.epb;
  const int ndof = nen * NF;
  const bool tabulated = (A.kind != APDX_SET_INTPOINT);

  // shared layout
  double *tN = sm;                                   // [n_gp][nen]
  double *tdN = tN + n_gp * nen;                     // [n_gp][nen][dr]
  double *tW = tdN + n_gp * nen * dr;                // [n_gp]
  double *slot0 = tW + n_gp;
  const int per_slot = nen * (DIM + NF + 1) + n_gp * (nen * DIM + NPT) + (tabulated ? 0 : nen);
  auto sX = [&](int s) { return slot0 + (size_t)s * per_slot; };          // [nen][DIM]
  auto sU = [&](int s) { return sX(s) + nen * DIM; };                     // [nen][NF]
  auto sUN = [&](int s) { return sU(s) + nen * NF; };                     // [nen]
  auto sG = [&](int s) { return sUN(s) + nen; };                          // [n_gp][nen][DIM]
  auto sPT = [&](int s) { return sG(s) + n_gp * nen * DIM; };             // [n_gp][NPT]
  auto sNrow = [&](int s) { return sPT(s) + n_gp * NPT; };                // [nen] (intpoint only)

  if (tabulated) {
    for (int i = threadIdx.x; i < n_gp * nen; i += blockDim.x) tN[i] = A.shape_n[i];
    for (int i = threadIdx.x; i < n_gp * nen * dr; i += blockDim.x) tdN[i] = A.shape_dn[i];
    for (int i = threadIdx.x; i < n_gp; i += blockDim.x) tW[i] = A.gp_w[i];
  }

  for (int64_t base = (int64_t)blockIdx.x * epb; base < A.n_rows; base += (int64_t)gridDim.x * epb) {
    const int cnt = (int)min((int64_t)epb, A.n_rows - base);
    __syncthreads();
    // ---- phase A: gather ---------------------------------------------------------------
    for (int i = threadIdx.x; i < cnt * nen; i += blockDim.x) {
      int s = i / nen, a = i - s * nen;
      int64_t row = base + s;
      int32_t node = A.conn[row * nen + a];
      if (tabulated) {
#pragma unroll
        for (int d = 0; d < DIM; ++d) sX(s)[a * DIM + d] = A.coords[(int64_t)node * DIM + d];
      } else {
        sNrow(s)[a] = A.ip_n[row * nen + a];
#pragma unroll
        for (int d = 0; d < DIM; ++d) sG(s)[a * DIM + d] = A.ip_dndx[(row * nen + a) * DIM + d];
      }
#pragma unroll
      for (int c = 0; c < NF; ++c) sU(s)[a * NF + c] = A.dofs[(int64_t)node * NF + c];
      if (NF == 1) sUN(s)[a] = (A.model == APDX_MODEL_CAPACITY) ? A.dofs_n[node] : 0.0;
    }
    __syncthreads();
    // ---- phase B: one thread per (element, Gauss point) ----------------------------------
    for (int i = threadIdx.x; i < cnt * n_gp; i += blockDim.x) {
      int s = i / n_gp, g = i - s * n_gp;
      int64_t row = base + s;
      const double *X = sX(s), *U = sU(s);
      double *G = sG(s) + (size_t)g * nen * DIM;
      double *PT = sPT(s) + g * NPT;
      const double *Ng = tabulated ? (tN + g * nen) : sNrow(s);
      double w;
      if (A.kind == APDX_SET_DOMAIN) {
        double J[DIM][DIM], Ji[DIM][DIM];
#pragma unroll
        for (int d = 0; d < DIM; ++d)
#pragma unroll
          for (int k = 0; k < DIM; ++k) J[d][k] = 0.0;
        for (int a = 0; a < nen; ++a)
#pragma unroll
          for (int d = 0; d < DIM; ++d)
#pragma unroll
            for (int k = 0; k < DIM; ++k) J[d][k] += X[a * DIM + d] * tdN[(g * nen + a) * DIM + k];
        double det = inv_det<DIM>(J, Ji);
        w = tW[g] * det;  // SIGNED det J (models.py:1691-1694)
        for (int a = 0; a < nen; ++a)
#pragma unroll
          for (int d = 0; d < DIM; ++d) {
            double v = 0.0;
#pragma unroll
            for (int k = 0; k < DIM; ++k) v += tdN[(g * nen + a) * DIM + k] * Ji[k][d];
            G[a * DIM + d] = v;
          }
      } else if (A.kind == APDX_SET_SURFACE) {
        // scaling ||dX/dxi|| (2-D) or ||dX/dxi1 x dX/dxi2|| (3-D)  (models.py:1806-1821)
        double t1[DIM], t2[DIM];
#pragma unroll
        for (int d = 0; d < DIM; ++d) t1[d] = t2[d] = 0.0;
        for (int a = 0; a < nen; ++a)
#pragma unroll
          for (int d = 0; d < DIM; ++d) {
            t1[d] += X[a * DIM + d] * tdN[(g * nen + a) * dr + 0];
            if (DIM == 3) t2[d] += X[a * DIM + d] * tdN[(g * nen + a) * dr + 1];
          }
        double sc;
        if constexpr (DIM == 2) {
          sc = sqrt(t1[0] * t1[0] + t1[1] * t1[1]);
        } else {
          double cx = t1[1] * t2[2] - t1[2] * t2[1];
          double cy = t1[2] * t2[0] - t1[0] * t2[2];
          double cz = t1[0] * t2[1] - t1[1] * t2[0];
          sc = sqrt(cx * cx + cy * cy + cz * cz);
        }
        w = tW[g] * sc;
        for (int a = 0; a < nen * DIM; ++a) G[a] = 0.0;
      } else {
        w = A.ip_w[row];  // physical weight of the integration point (seeder.py:3473-3488)
      }

      if constexpr (NF == 1) {
        double gu[DIM], ug = 0.0, ung = 0.0;
#pragma unroll
        for (int d = 0; d < DIM; ++d) gu[d] = 0.0;
        for (int a = 0; a < nen; ++a) {
          double ua = U[a];
          ug += Ng[a] * ua;
          ung += Ng[a] * sUN(s)[a];
#pragma unroll
          for (int d = 0; d < DIM; ++d) gu[d] += G[a * DIM + d] * ua;
        }
        double kc = 0.0, km = 0.0, sf = 0.0;
        if (A.model == APDX_MODEL_POISSON_POTENTIAL || A.model == APDX_MODEL_POISSON_WEAK) {
          double c = par_get(A.par[APDX_PARAM_COEFFICIENT], row, g, 0, 1.0);
          double f = par_get(A.par[APDX_PARAM_SOURCE], row, g, 0, 0.0);
          double sgn = (A.model == APDX_MODEL_POISSON_WEAK) ? -1.0 : 1.0;
          kc = sgn * w * c;
          sf = sgn * w * f;
        } else if (A.model == APDX_MODEL_CAPACITY) {
          double c = par_get(A.par[APDX_PARAM_COEFFICIENT], row, g, 0, 1.0);
          km = -w * c * A.inv_dt;
        } else {  // Neumann, scalar field: -dtheta * q
          sf = w * par_get(A.par[APDX_PARAM_TRACTION], row, g, 0, 0.0);
        }
        PT[0] = kc; PT[1] = km; PT[2] = sf;
#pragma unroll
        for (int d = 0; d < DIM; ++d) PT[3 + d] = gu[d];
        PT[6] = ug - ung;
      } else {
        double H[DIM][DIM];
#pragma unroll
        for (int i2 = 0; i2 < DIM; ++i2)
#pragma unroll
          for (int j2 = 0; j2 < DIM; ++j2) H[i2][j2] = 0.0;
        for (int a = 0; a < nen; ++a)
#pragma unroll
          for (int i2 = 0; i2 < DIM; ++i2)
#pragma unroll
            for (int j2 = 0; j2 < DIM; ++j2) H[i2][j2] += U[a * NF + i2] * G[a * DIM + j2];
        double A0 = 0, A1 = 0, A2 = 0, A3 = 0;
        double T[DIM][DIM], P[DIM][DIM], sb[DIM];
#pragma unroll
        for (int i2 = 0; i2 < DIM; ++i2) {
          sb[i2] = 0.0;
#pragma unroll
          for (int j2 = 0; j2 < DIM; ++j2) { T[i2][j2] = (i2 == j2) ? 1.0 : 0.0; P[i2][j2] = 0.0; }
        }
        if (A.model == APDX_MODEL_NEUMANN) {
#pragma unroll
          for (int i2 = 0; i2 < DIM; ++i2) sb[i2] = w * par_get(A.par[APDX_PARAM_TRACTION], row, g, i2, 0.0);
        } else {
          double Em = par_get(A.par[APDX_PARAM_YOUNGS], row, g, 0, 0.0);
          double nu = par_get(A.par[APDX_PARAM_POISSON_RATIO], row, g, 0, 0.0);
          if (A.par[APDX_PARAM_BODY_LOAD].p) {
#pragma unroll
            for (int i2 = 0; i2 < DIM; ++i2) sb[i2] = w * par_get(A.par[APDX_PARAM_BODY_LOAD], row, g, i2, 0.0);
          }
          if (A.model == APDX_MODEL_LINEAR_ELASTICITY) {
            double c12, c33, cd;
            lin_el_constants(A.mode, Em, nu, c12, c33, cd);
            double tr = 0.0;
#pragma unroll
            for (int i2 = 0; i2 < DIM; ++i2) tr += H[i2][i2];
#pragma unroll
            for (int i2 = 0; i2 < DIM; ++i2)
#pragma unroll
              for (int j2 = 0; j2 < DIM; ++j2)
                P[i2][j2] = w * (c33 * (H[i2][j2] + H[j2][i2]) + ((i2 == j2) ? (c12 * tr + cd * H[i2][i2]) : 0.0));
            A0 = w * c33; A1 = w * c33; A2 = w * c12; A3 = w * cd;
          } else {  // neo-Hooke: P = mu F - c1 F^-T, c1 = mu - lam/2 (J^2-1), c2 = lam J^2
            double lam = Em * nu / ((1.0 + nu) * (1.0 - 2.0 * nu));  // models.py:955-956
            double mu = Em / (2.0 * (1.0 + nu));
            double F[DIM][DIM], Fi[DIM][DIM];
#pragma unroll
            for (int i2 = 0; i2 < DIM; ++i2)
#pragma unroll
              for (int j2 = 0; j2 < DIM; ++j2) F[i2][j2] = H[i2][j2] + ((i2 == j2) ? 1.0 : 0.0);
            double Jd = inv_det<DIM>(F, Fi);  // plain strain: F33 = 1, in-plane block suffices
            double c1 = mu - 0.5 * lam * (Jd * Jd - 1.0);
            double c2 = lam * Jd * Jd;
#pragma unroll
            for (int i2 = 0; i2 < DIM; ++i2)
#pragma unroll
              for (int j2 = 0; j2 < DIM; ++j2) {
                T[i2][j2] = Fi[j2][i2];  // F^-T
                P[i2][j2] = w * (mu * F[i2][j2] - c1 * Fi[j2][i2]);
              }
            A0 = w * mu; A1 = w * c1; A2 = w * c2; A3 = 0.0;
          }
        }
        PT[0] = A0; PT[1] = A1; PT[2] = A2; PT[3] = A3;
#pragma unroll
        for (int i2 = 0; i2 < DIM; ++i2) {
#pragma unroll
          for (int j2 = 0; j2 < DIM; ++j2) {
            PT[4 + i2 * DIM + j2] = T[i2][j2];
            PT[13 + i2 * DIM + j2] = P[i2][j2];
          }
          PT[22 + i2] = sb[i2];
        }
      }
    }
    __syncthreads();
    // ---- phase C: tangent blocks, one thread per (element, node pair a <= b) --------------------
    // Every model of this kernel has a symmetric element tangent (K[(a,i),(b,k)] = K[(b,k),(a,i)]: potential Hessians,
    // Galerkin forms with symmetric constitutive tensors, the capacity matrix), so only the nen (nen + 1) / 2 upper node
    // pairs are integrated (36 of 64 for hex8: phase C is ~97 % of the kernel's flops) and only the upper triangle of
    // the element matrix is stored (300 of 576 entries for hex8 with three dofs per node); the scatter reads entry
    // (J, I) from the slot of (I, J), so the assembled matrix is symmetric to the last bit.
    if (A.want_tangent && A.model != APDX_MODEL_NEUMANN) {
      const int pairs = nen * (nen + 1) / 2;
      for (int i = threadIdx.x; i < cnt * pairs; i += blockDim.x) {
        int s = i / pairs, ab = i - s * pairs;
        int a = 0;
        while (ab >= nen - a) { ab -= nen - a; ++a; }   // row a of the upper triangle holds the pairs (a, a .. nen-1)
        const int b = a + ab;
        int64_t row = base + s;
        const double *Ga0 = sG(s), *PT0 = sPT(s);
        double blk[NF][NF];
#pragma unroll
        for (int i2 = 0; i2 < NF; ++i2)
#pragma unroll
          for (int k2 = 0; k2 < NF; ++k2) blk[i2][k2] = 0.0;
        for (int g = 0; g < n_gp; ++g) {
          const double *Ga = Ga0 + ((size_t)g * nen + a) * DIM;
          const double *Gb = Ga0 + ((size_t)g * nen + b) * DIM;
          const double *PT = PT0 + g * NPT;
          double gg = 0.0;
#pragma unroll
          for (int d = 0; d < DIM; ++d) gg += Ga[d] * Gb[d];
          if constexpr (NF == 1) {
            const double *Ng = tabulated ? (tN + g * nen) : sNrow(s);
            blk[0][0] += PT[0] * gg + PT[1] * Ng[a] * Ng[b];
          } else {
            double ga[DIM], gb[DIM];
#pragma unroll
            for (int i2 = 0; i2 < DIM; ++i2) {
              double va = 0.0, vb = 0.0;
#pragma unroll
              for (int j2 = 0; j2 < DIM; ++j2) {
                va += PT[4 + i2 * DIM + j2] * Ga[j2];
                vb += PT[4 + i2 * DIM + j2] * Gb[j2];
              }
              ga[i2] = va; gb[i2] = vb;
            }
#pragma unroll
            for (int i2 = 0; i2 < DIM; ++i2)
#pragma unroll
              for (int k2 = 0; k2 < DIM; ++k2) {
                double v = PT[1] * ga[k2] * gb[i2] + PT[2] * ga[i2] * gb[k2];
                if (i2 == k2) v += PT[0] * gg + PT[3] * ga[i2] * gb[i2];
                blk[i2][k2] += v;
              }
          }
        }
        // the element matrix is stored as its upper triangle, row by row (SetData::tri; pattern.cu: soa_address):
        // slot(I, J) = I ndof - I (I - 1) / 2 + (J - I) for dofs I <= J.  a < b: the whole block lies above the
        // diagonal; a == b: its upper triangle (the lower one would differ by the rounding of (c g_k) g_i against
        // (c g_i) g_k only, the gather lists read it from the transposed slot).
        double *out = A.ke + row * ((int64_t)ndof * (ndof + 1) / 2);
#pragma unroll
        for (int i2 = 0; i2 < NF; ++i2) {
          const int I = a * NF + i2;
          double *orow = out + ((int64_t)I * ndof - (int64_t)I * (I - 1) / 2 - I);   // + J gives slot(I, J)
#pragma unroll
          for (int k2 = 0; k2 < NF; ++k2)
            if (a != b || k2 >= i2) orow[b * NF + k2] = blk[i2][k2];
        }
      }
    }
    // ---- phase D: residual, one thread per (element, a) ----------------------------------------
    for (int i = threadIdx.x; i < cnt * nen; i += blockDim.x) {
      int s = i / nen, a = i - s * nen;
      int64_t row = base + s;
      double r[NF];
#pragma unroll
      for (int c = 0; c < NF; ++c) r[c] = 0.0;
      for (int g = 0; g < n_gp; ++g) {
        const double *Ga = sG(s) + ((size_t)g * nen + a) * DIM;
        const double *PT = sPT(s) + g * NPT;
        const double *Ng = tabulated ? (tN + g * nen) : sNrow(s);
        if constexpr (NF == 1) {
          double gd = 0.0;
#pragma unroll
          for (int d = 0; d < DIM; ++d) gd += Ga[d] * PT[3 + d];
          r[0] += PT[0] * gd + (PT[1] * PT[6] - PT[2]) * Ng[a];
        } else {
#pragma unroll
          for (int i2 = 0; i2 < DIM; ++i2) {
            double v = 0.0;
#pragma unroll
            for (int j2 = 0; j2 < DIM; ++j2) v += PT[13 + i2 * DIM + j2] * Ga[j2];
            r[i2] += v - Ng[a] * PT[22 + i2];
          }
        }
      }
#pragma unroll
      for (int c = 0; c < NF; ++c) A.re[row * ndof + a * NF + c] = r[c];
    }
  }
}

// values[u] = sum of the element-matrix entries listed for CSR entry u, in ascending COO order
// (fixed order => bitwise reproducible; replaces the duplicate summing of solver.py:1207-1211).
__global__ void k_gather_reduce_full(const double *__restrict__ ke, const uint32_t *__restrict__ perm,
                                     const int32_t *__restrict__ seg_ptr, int64_t nnz,
                                     double *__restrict__ vals) {
  int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= nnz) return;
  double acc = 0.0;
  for (int32_t j = seg_ptr[u]; j < seg_ptr[u + 1]; ++j) acc += ke[perm[j]];
  vals[u] = acc;
}
__global__ void k_gather_reduce_red(const double *__restrict__ ke, const uint32_t *__restrict__ perm,
                                    const int32_t *__restrict__ seg_ptr, const int32_t *__restrict__ red2full,
                                    int64_t nnz_red, double *__restrict__ red_vals) {
  int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nnz_red) return;
  int32_t u = red2full[q];
  double acc = 0.0;
  for (int32_t j = seg_ptr[u]; j < seg_ptr[u + 1]; ++j) acc += ke[perm[j]];
  red_vals[q] = acc;
}
__global__ void k_gather_reduce_res(const double *__restrict__ re, const uint32_t *__restrict__ rperm,
                                    const int32_t *__restrict__ rseg_ptr, int64_t n,
                                    double *__restrict__ residual) {
  int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= n) return;
  double acc = 0.0;
  for (int32_t j = rseg_ptr[d]; j < rseg_ptr[d + 1]; ++j) acc += re[rperm[j]];
  residual[d] = acc;
}

static int smem_doubles(const SetData &st, int dim, int nf, int epb) {
  int nen = st.d.nen, n_gp = st.d.n_gp, dr = st.d.dim_ref;
  int npt = (nf == 1) ? NPT_SCALAR : NPT_VECTOR;
  bool tab = st.d.kind != APDX_SET_INTPOINT;
  int per_slot = nen * (dim + nf + 1) + n_gp * (nen * dim + npt) + (tab ? 0 : nen);
  return n_gp * nen + n_gp * nen * dr + n_gp + epb * per_slot;
}

template <int DIM, int NF>
static int launch_one(apdx_plan *pl, SetData &st, const ElemArgs &args_in) {
  ElemArgs args = args_in;
  // elements per block: fill ~96 KB of shared memory, at most 64 slots
  int epb = 64;
  while (epb > 1 && smem_doubles(st, DIM, NF, epb) * 8 > 96 * 1024) epb >>= 1;
  size_t smem = (size_t)smem_doubles(st, DIM, NF, epb) * 8;
  APDX_REQUIRE(smem <= 220 * 1024, APDX_ERR_UNSUPPORTED, "element too large for shared memory staging");
  args.epb = epb;
  APDX_CUDA(cudaFuncSetAttribute(k_elements<DIM, NF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t blocks = (st.d.n_rows + epb - 1) / epb;
  int64_t cap = (int64_t)sm_count() * 16;
  unsigned grid = (unsigned)(blocks < cap ? blocks : cap);
  k_elements<DIM, NF><<<grid, 256, smem, pl->stream>>>(args);
  pl->stats.kernel_launches += 1;
  return APDX_OK;
}

int launch_element_kernels(apdx_plan *pl, const double *dofs_d, bool want_tangent) {
  APDX_REQUIRE(pl->have_coords, APDX_ERR_STATE, "apdx_set_coords must be called before assembling");
  for (auto &st : pl->sets) {
    if (st.d.n_rows == 0) continue;
    if (st.d.model == APDX_MODEL_PATTERN_ONLY) continue;   // its streams were zeroed when the plan was created
    ElemArgs a{};
    a.conn = st.conn.p;
    a.n_rows = st.d.n_rows;
    a.nen = st.d.nen; a.n_gp = st.d.n_gp; a.dim_ref = st.d.dim_ref;
    a.kind = st.d.kind; a.model = st.d.model; a.mode = st.d.mode;
    a.want_tangent = want_tangent ? 1 : 0;
    a.shape_n = st.shape_n.p; a.shape_dn = st.shape_dn.p; a.gp_w = st.gp_w.p;
    a.ip_n = st.ip_n.p; a.ip_dndx = st.ip_dndx.p; a.ip_w = st.ip_w.p;
    a.coords = pl->coords.p; a.dofs = dofs_d; a.dofs_n = pl->dofs_n.p;
    a.inv_dt = 1.0 / pl->time_increment;
    for (int p = 0; p < APDX_PARAM_COUNT; ++p) a.par[p] = st.pview[p];
    a.ke = pl->ke.p + st.ke_offset;
    a.re = pl->re.p + st.res_offset;
    if (st.d.kind == APDX_SET_INTPOINT)
      APDX_REQUIRE(st.ip_n.p && st.ip_w.p, APDX_ERR_STATE, "integration-point tables of a 'sparse' set are missing");
    if (st.d.model == APDX_MODEL_CAPACITY)
      APDX_REQUIRE(pl->dofs_n.p, APDX_ERR_STATE, "settings['dofs n'] not set (apdx_set_dofs_n)");
    if (st.d.model == APDX_MODEL_LINEAR_ELASTICITY || st.d.model == APDX_MODEL_NEO_HOOKE)
      APDX_REQUIRE(st.pview[APDX_PARAM_YOUNGS].p && st.pview[APDX_PARAM_POISSON_RATIO].p, APDX_ERR_STATE,
                   "Young's modulus / Poisson ratio of an elasticity set are missing");
    bool handled = false;
    APDX_CHECK(launch_fast_elements(pl, st, a, &handled));
    if (handled) continue;
    if (pl->dim == 2 && pl->nf == 1) APDX_CHECK((launch_one<2, 1>(pl, st, a)));
    else if (pl->dim == 3 && pl->nf == 1) APDX_CHECK((launch_one<3, 1>(pl, st, a)));
    else if (pl->dim == 2 && pl->nf == 2) APDX_CHECK((launch_one<2, 2>(pl, st, a)));
    else if (pl->dim == 3 && pl->nf == 3) APDX_CHECK((launch_one<3, 3>(pl, st, a)));
    else APDX_REQUIRE(false, APDX_ERR_UNSUPPORTED, "dim=%d nf=%d not supported", pl->dim, pl->nf);
  }
  APDX_CUDA(cudaGetLastError());
  return APDX_OK;
}

// tangent_flags: bit 0 = full CSR values, bit 1 = reduced CSR values
int launch_gather_reduce(apdx_plan *pl, int tangent_flags, double *residual_d) {
  const int B = 256;
  cudaStream_t s = pl->stream;
  if (residual_d) {
    k_gather_reduce_res<<<(unsigned)((pl->n_dofs + B - 1) / B), B, 0, s>>>(pl->re.p, pl->rperm.p, pl->rseg_ptr.p,
                                                                           pl->n_dofs, residual_d);
    pl->stats.kernel_launches += 1;
  }
  if (tangent_flags & 1) {
    if (!pl->vals.p) APDX_CHECK(pl->vals.alloc(pl->nnz > 0 ? pl->nnz : 1));
    k_gather_reduce_full<<<(unsigned)((pl->nnz + B - 1) / B), B, 0, s>>>(pl->ke.p, pl->perm.p, pl->seg_ptr.p,
                                                                         pl->nnz, pl->vals.p);
    pl->stats.kernel_launches += 1;
  }
  if ((tangent_flags & 2) && pl->nnz_red > 0) {
    if (!pl->red_vals.p) APDX_CHECK(pl->red_vals.alloc(pl->nnz_red));
    k_gather_reduce_red<<<(unsigned)((pl->nnz_red + B - 1) / B), B, 0, s>>>(
        pl->ke.p, pl->perm.p, pl->seg_ptr.p, pl->red2full.p, pl->nnz_red, pl->red_vals.p);
    pl->stats.kernel_launches += 1;
  }
  APDX_CUDA(cudaGetLastError());
  return APDX_OK;
}

}  // namespace apdx
