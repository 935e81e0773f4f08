// Device helpers shared by the element kernels (elements.cu: generic shared-memory kernel,
// elements_fast.cu: register-resident kernels for small scalar elements).
#pragma once
#include "common.cuh"

namespace apdx {

constexpr int NPT_SCALAR = 8;
constexpr int NPT_VECTOR = 26;

struct ElemArgs {
  const int32_t *conn;
  int64_t n_rows;
  int nen, n_gp, dim_ref, kind, model, mode, epb, want_tangent;
  const double *shape_n, *shape_dn, *gp_w;
  const double *ip_n, *ip_dndx, *ip_w;
  const double *coords, *dofs, *dofs_n;
  double inv_dt;
  ParamView par[APDX_PARAM_COUNT];
  double *ke, *re;
};

__device__ __forceinline__ double par_get(const ParamView &v, int64_t row, int g, int c, double dflt) {
  if (!v.p) return dflt;
  return v.p[row * v.s_row + (int64_t)g * v.s_gp + c];
}

// closed-form inverse and determinant (utility.matrix_inv / matrix_det, utility.py:656-817)
template <int DIM>
__device__ __forceinline__ double inv_det(const double (&J)[DIM][DIM], double (&Ji)[DIM][DIM]) {
  if constexpr (DIM == 2) {
    double det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    double id = 1.0 / det;
    Ji[0][0] = J[1][1] * id;
    Ji[0][1] = -J[0][1] * id;
    Ji[1][0] = -J[1][0] * id;
    Ji[1][1] = J[0][0] * id;
    return det;
  } else {
    double a00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
    double a01 = J[0][2] * J[2][1] - J[0][1] * J[2][2];
    double a02 = J[0][1] * J[1][2] - J[0][2] * J[1][1];
    double a10 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
    double a11 = J[0][0] * J[2][2] - J[0][2] * J[2][0];
    double a12 = J[0][2] * J[1][0] - J[0][0] * J[1][2];
    double a20 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
    double a21 = J[0][1] * J[2][0] - J[0][0] * J[2][1];
    double a22 = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    double det = J[0][0] * a00 + J[0][1] * a10 + J[0][2] * a20;
    double id = 1.0 / det;
    Ji[0][0] = a00 * id; Ji[0][1] = a01 * id; Ji[0][2] = a02 * id;
    Ji[1][0] = a10 * id; Ji[1][1] = a11 * id; Ji[1][2] = a12 * id;
    Ji[2][0] = a20 * id; Ji[2][1] = a21 * id; Ji[2][2] = a22 * id;
    return det;
  }
}

// Voigt material constants exactly as models.py:570-601, folded into the three numbers of
//   A_iJkL = c12 d_iJ d_kL + c33 (d_ik d_JL + d_iL d_Jk) + cd d_iJkL,   cd = c11 - c12 - 2 c33
// (cd != 0 only for the reference's plain-strain matrix, whose shear entry is coeff*(1-2nu)).
__device__ __forceinline__ void lin_el_constants(int mode, double Em, double nu, double &c12, double &c33,
                                                 double &cd) {
  double c11;
  if (mode == APDX_MODE_PLAIN_STRAIN) {
    double mu = Em / (2.0 * (1.0 + nu));
    double c1 = 1.0 - 2.0 * nu, c2 = 1.0 - nu;
    double co = 2.0 * mu / c1;
    c11 = co * c2; c12 = co * nu; c33 = co * c1;
  } else if (mode == APDX_MODE_PLAIN_STRESS) {
    double co = Em / (1.0 - nu * nu);
    c11 = co; c12 = co * nu; c33 = co * (1.0 - nu) / 2.0;
  } else {   // APDX_MODE_3D, APDX_MODE_LAME: c11 = lam + 2 mu, c12 = lam, c33 = mu
    double co = Em / (1.0 + nu);
    double c1 = 1.0 - 2.0 * nu;
    c11 = co * (1.0 - nu) / c1; c12 = co * nu / c1; c33 = co * 0.5;
  }
  cd = c11 - c12 - 2.0 * c33;
}


// elements_fast.cu: returns APDX_OK and sets *handled when a specialised kernel took the set
int launch_fast_elements(apdx_plan *pl, SetData &st, const ElemArgs &args, bool *handled);

}  // namespace apdx
