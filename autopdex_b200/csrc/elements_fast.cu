// Register-resident element kernels for small scalar-field domain elements (quad4, hex8, tri3,
// tri6, tet4, quad9 with nf = 1): the flagship path of BASELINE config 4 (3-D Q1 hex Poisson).
//
// One thread per element.  Connectivity is read with vector loads, nodal coordinates / dofs are
// gathered straight into registers, the shape tables live in the kernel-argument constant bank
// (so dN/dxi and the weights are immediate operands of the DFMAs), the symmetric element tangent
// (NEN(NEN+1)/2 accumulators) and the residual stay in registers over the fully unrolled Gauss
// loop, and the results are stored entry-major (ke[upper-triangle slot][element]) so that every global store
// instruction of a warp writes 256 contiguous bytes.  Same closed forms as the generic kernel
// (elements.cu), same reference citations: models.py:96-134 (poisson_weak), the README potential,
// signed w*det J of models.py:1691-1694, 1257-1261.
#include <cstdlib>

#include "elements.cuh"

namespace apdx {

template <int DIM, int NEN, int NGP>
struct FastTab {
  double N[NGP * NEN];
  double dN[NGP * NEN * DIM];
  double w[NGP];
};

template <int DIM, int NEN, int NGP>
struct FastArgs {
  const int32_t *conn;
  const double *coords, *dofs;
  double *ke, *re;
  int64_t n_rows;
  int model;
  ParamView coef, src;
  FastTab<DIM, NEN, NGP> tab;
};

constexpr int FAST_BLOCK = 128;

template <int DIM, int NEN, int NGP, bool TANGENT>
__global__ void __launch_bounds__(FAST_BLOCK) k_elem_scalar_reg(const __grid_constant__ FastArgs<DIM, NEN, NGP> A) {
  constexpr int NSYM = NEN * (NEN + 1) / 2;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = e < A.n_rows;
  const int64_t ec = active ? e : A.n_rows - 1;

  int32_t nd[NEN];
  if constexpr (NEN % 4 == 0) {
    const int4 *cp = reinterpret_cast<const int4 *>(A.conn + ec * NEN);
#pragma unroll
    for (int q = 0; q < NEN / 4; ++q) {
      int4 v = __ldg(cp + q);
      nd[4 * q] = v.x; nd[4 * q + 1] = v.y; nd[4 * q + 2] = v.z; nd[4 * q + 3] = v.w;
    }
  } else {
#pragma unroll
    for (int a = 0; a < NEN; ++a) nd[a] = __ldg(A.conn + ec * NEN + a);
  }
  double X[NEN][DIM], U[NEN];
#pragma unroll
  for (int a = 0; a < NEN; ++a) {
#pragma unroll
    for (int d = 0; d < DIM; ++d) X[a][d] = __ldg(A.coords + (int64_t)nd[a] * DIM + d);
    U[a] = __ldg(A.dofs + nd[a]);
  }
  double K[TANGENT ? NSYM : 1], R[NEN];
#pragma unroll
  for (int i = 0; i < (TANGENT ? NSYM : 1); ++i) K[i] = 0.0;
#pragma unroll
  for (int a = 0; a < NEN; ++a) R[a] = 0.0;
  const double sgn = (A.model == APDX_MODEL_POISSON_WEAK) ? -1.0 : 1.0;

#pragma unroll
  for (int g = 0; g < NGP; ++g) {
    double J[DIM][DIM], Ji[DIM][DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d)
#pragma unroll
      for (int k = 0; k < DIM; ++k) {
        double v = 0.0;
#pragma unroll
        for (int a = 0; a < NEN; ++a) v += X[a][d] * A.tab.dN[(g * NEN + a) * DIM + k];
        J[d][k] = v;
      }
    const double det = inv_det<DIM>(J, Ji);
    const double w = A.tab.w[g] * det;   // signed
    const double c = par_get(A.coef, ec, g, 0, 1.0);
    const double f = par_get(A.src, ec, g, 0, 0.0);
    const double kc = sgn * w * c, sf = sgn * w * f;
    double G[NEN][DIM], gu[DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d) gu[d] = 0.0;
#pragma unroll
    for (int a = 0; a < NEN; ++a)
#pragma unroll
      for (int d = 0; d < DIM; ++d) {
        double v = 0.0;
#pragma unroll
        for (int k = 0; k < DIM; ++k) v += A.tab.dN[(g * NEN + a) * DIM + k] * Ji[k][d];
        G[a][d] = v;
        gu[d] += v * U[a];
      }
#pragma unroll
    for (int a = 0; a < NEN; ++a) {
      double gd = 0.0, ga[DIM];
#pragma unroll
      for (int d = 0; d < DIM; ++d) { ga[d] = kc * G[a][d]; gd += ga[d] * gu[d]; }
      R[a] += gd - sf * A.tab.N[g * NEN + a];
      if constexpr (TANGENT) {
#pragma unroll
        for (int b = a; b < NEN; ++b) {
          double v = 0.0;
#pragma unroll
          for (int d = 0; d < DIM; ++d) v += ga[d] * G[b][d];
          K[a * NEN - a * (a - 1) / 2 + (b - a)] += v;
        }
      }
    }
  }

  // ---- entry-major stores: ke[slot][e], re[i][e]: consecutive threads (elements) write consecutive addresses.  Only
  // the NSYM upper-triangle slots exist: the gather lists address entry (a,b) and (b,a) through the same slot
  // (pattern.cu: k_to_soa), which saves 28 of the 64 stores of a hex8 matrix ----
  if (!active) return;
  if constexpr (TANGENT) {
#pragma unroll
    for (int i = 0; i < NSYM; ++i) A.ke[(int64_t)i * A.n_rows + e] = K[i];
  }
#pragma unroll
  for (int a = 0; a < NEN; ++a) A.re[(int64_t)a * A.n_rows + e] = R[a];
}

template <int DIM, int NEN, int NGP>
static int launch_fast(apdx_plan *pl, SetData &st, const ElemArgs &a) {
  FastArgs<DIM, NEN, NGP> F;
  F.conn = a.conn; F.coords = a.coords; F.dofs = a.dofs; F.ke = a.ke; F.re = a.re;
  F.n_rows = a.n_rows; F.model = a.model;
  F.coef = a.par[APDX_PARAM_COEFFICIENT];
  F.src = a.par[APDX_PARAM_SOURCE];
  for (int i = 0; i < NGP * NEN; ++i) F.tab.N[i] = st.h_shape_n[i];
  for (int i = 0; i < NGP * NEN * DIM; ++i) F.tab.dN[i] = st.h_shape_dn[i];
  for (int i = 0; i < NGP; ++i) F.tab.w[i] = st.h_gp_w[i];
  // (blocks of 32 / 64 threads at the same 8 warps per SM were measured in round 2: 13.37 / 13.38 / 13.38 ms per
  // apdx_assemble pass at P256, profiles/r02a_asm_elem_block_128_32_64.jsonl -- no difference, removed)
  const int block = FAST_BLOCK;
  const unsigned grid = (unsigned)((a.n_rows + block - 1) / block);
  if (a.want_tangent) k_elem_scalar_reg<DIM, NEN, NGP, true><<<grid, block, 0, pl->stream>>>(F);
  else k_elem_scalar_reg<DIM, NEN, NGP, false><<<grid, block, 0, pl->stream>>>(F);
  pl->stats.kernel_launches += 1;
  APDX_CUDA(cudaGetLastError());
  return APDX_OK;
}

// sets served by the register kernels store their element streams entry-major (see pattern.cu: k_to_soa)
bool fast_kernel_applies(int dim, int nf, const apdx_set_desc &d) {
  if (d.kind != APDX_SET_DOMAIN || nf != 1) return false;
  if (d.model != APDX_MODEL_POISSON_WEAK && d.model != APDX_MODEL_POISSON_POTENTIAL) return false;
  const int t[7][3] = {{3, 8, 8}, {2, 4, 4}, {3, 4, 1}, {3, 4, 4}, {2, 3, 1}, {2, 3, 3}, {2, 9, 9}};
  for (auto &c : t)
    if (dim == c[0] && d.nen == c[1] && d.n_gp == c[2]) return true;
  return false;
}

int launch_fast_elements(apdx_plan *pl, SetData &st, const ElemArgs &a, bool *handled) {
  *handled = false;
  if (!st.soa) return APDX_OK;
  const int dim = pl->dim, nen = a.nen, ngp = a.n_gp;
#define APDX_FAST(D, N, G)                                   \
  if (dim == D && nen == N && ngp == G) {                    \
    *handled = true;                                         \
    return launch_fast<D, N, G>(pl, st, a);                  \
  }
  APDX_FAST(3, 8, 8)    // hex8, Gauss order 2 (BASELINE config 4)
  APDX_FAST(2, 4, 4)    // quad4, Gauss order 2 (README)
  APDX_FAST(3, 4, 1)    // tet4, 1 point
  APDX_FAST(3, 4, 4)    // tet4, 4 points
  APDX_FAST(2, 3, 1)    // tri3, 1 point
  APDX_FAST(2, 3, 3)    // tri3, 3 points
  APDX_FAST(2, 9, 9)    // quad9, Gauss order 4
#undef APDX_FAST
  return APDX_OK;
}

}  // namespace apdx
