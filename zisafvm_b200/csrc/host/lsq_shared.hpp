// Least-squares matrix row and pseudo-inverse, one source for the host (stencil.cpp) and the device
// (kernels/precompute.cu).  Both sides evaluate exactly these expressions in IEEE double without fused
// multiply-adds (g++ has no FMA target here; precompute.cu is compiled with -fmad=false; division and square
// root are correctly rounded on both), so the device-built weights are bit-identical to the host-built ones.
//
//   lsq_row           one row of assemble_weno_ao_matrix, src/zisa/reconstruction/lsq_solver.cpp:168-403
//   pinv_householder  W = R^{-1} Q^T of A = Q R; the reference solves the normal equations instead
//                     (LDLT(A^T A), lsq_solver.cpp:47,82) -- same minimiser, squared condition number
#pragma once

#if defined(__CUDACC__)
#define ZFVM_HD __host__ __device__ inline
#else
#define ZFVM_HD inline
#endif

#include <cmath>

namespace zfvm {
namespace lsq {

ZFVM_HD int dof(int deg, int n_dims) {
  return n_dims == 2 ? ((deg + 1) * (deg + 2)) / 2 : ((deg + 1) * (deg + 2) * (deg + 3)) / 6;
}
ZFVM_HD int pidx2(int a, int b) {
  const int n = a + b;
  return ((n + 1) * n) / 2 + b;
}
ZFVM_HD int pidx3(int a, int b, int c) { return dof(a + b + c - 1, 3) + pidx2(b, c); }

/// Row of the LSQ matrix for stencil member j of centre cell 0: (x, y, z) = (x_j - x_0) / l_0, lj = l_j / l_0,
/// C0 / Cj the normalised moments of the two cells.  `row` has dof(order - 1) - 1 entries.
ZFVM_HD void lsq_row(double *row, int nd, int order, double x, double y, double z, double lj, const double *C0,
                     const double *Cj) {
  const double x_10 = x, x_01 = y;
  const int i_10 = nd == 2 ? pidx2(1, 0) : pidx3(1, 0, 0), i_01 = nd == 2 ? pidx2(0, 1) : pidx3(0, 1, 0);
  row[i_10 - 1] = x_10;
  row[i_01 - 1] = x_01;
#define ZFVM_IDX(a, b) (nd == 2 ? pidx2(a, b) : pidx3(a, b, 0))
  if (order >= 3) {
    const int i_20 = ZFVM_IDX(2, 0), i_11 = ZFVM_IDX(1, 1), i_02 = ZFVM_IDX(0, 2);
    const double x_20 = x_10 * x_10, x_11 = x_10 * x_01, x_02 = x_01 * x_01;
    const double lj_2 = lj * lj;
    row[i_20 - 1] = x_20 - C0[i_20] + lj_2 * Cj[i_20];
    row[i_11 - 1] = x_11 - C0[i_11] + lj_2 * Cj[i_11];
    row[i_02 - 1] = x_02 - C0[i_02] + lj_2 * Cj[i_02];
    if (order >= 4) {
      const int i_30 = ZFVM_IDX(3, 0), i_21 = ZFVM_IDX(2, 1), i_12 = ZFVM_IDX(1, 2), i_03 = ZFVM_IDX(0, 3);
      const double x_30 = x_20 * x_10, x_21 = x_20 * x_01, x_12 = x_11 * x_01, x_03 = x_02 * x_01;
      const double lj_3 = lj_2 * lj;
      row[i_30 - 1] = x_30 - C0[i_30] + 3.0 * x_10 * lj_2 * Cj[i_20] + lj_3 * Cj[i_30];
      row[i_21 - 1] = x_21 - C0[i_21] + x_01 * lj_2 * Cj[i_20] + 2.0 * x_10 * lj_2 * Cj[i_11] + lj_3 * Cj[i_21];
      row[i_12 - 1] = x_12 - C0[i_12] + x_10 * lj_2 * Cj[i_02] + 2.0 * x_01 * lj_2 * Cj[i_11] + lj_3 * Cj[i_12];
      row[i_03 - 1] = x_03 - C0[i_03] + 3.0 * x_01 * lj_2 * Cj[i_02] + lj_3 * Cj[i_03];
      if (order >= 5) {
        const int i_40 = ZFVM_IDX(4, 0), i_31 = ZFVM_IDX(3, 1), i_22 = ZFVM_IDX(2, 2), i_13 = ZFVM_IDX(1, 3),
                  i_04 = ZFVM_IDX(0, 4);
        const double x_40 = x_30 * x_10, x_31 = x_30 * x_01, x_22 = x_21 * x_01, x_13 = x_12 * x_01, x_04 = x_03 * x_01;
        const double lj_4 = lj_3 * lj;
        row[i_40 - 1] = x_40 - C0[i_40] + 6.0 * x_20 * lj_2 * Cj[i_20] + 4.0 * x_10 * lj_3 * Cj[i_30] + lj_4 * Cj[i_40];
        row[i_31 - 1] = x_31 - C0[i_31] + 3 * x_11 * lj_2 * Cj[i_20] + 3.0 * x_20 * lj_2 * Cj[i_11] +
                        x_01 * lj_3 * Cj[i_30] + 3.0 * x_10 * lj_3 * Cj[i_21] + lj_4 * Cj[i_31];
        row[i_22 - 1] = x_22 - C0[i_22] + x_02 * lj_2 * Cj[i_20] + x_20 * lj_2 * Cj[i_02] + 4 * x_11 * lj_2 * Cj[i_11] +
                        2.0 * x_01 * lj_3 * Cj[i_21] + 2 * x_10 * lj_3 * Cj[i_12] + lj_4 * Cj[i_22];
        row[i_13 - 1] = x_13 - C0[i_13] + 3 * x_11 * lj_2 * Cj[i_02] + 3.0 * x_02 * lj_2 * Cj[i_11] +
                        x_10 * lj_3 * Cj[i_03] + 3.0 * x_01 * lj_3 * Cj[i_12] + lj_4 * Cj[i_13];
        row[i_04 - 1] = x_04 - C0[i_04] + 6.0 * x_02 * lj_2 * Cj[i_02] + 4.0 * x_01 * lj_3 * Cj[i_03] + lj_4 * Cj[i_04];
      }
    }
  }
#undef ZFVM_IDX
  if (nd == 3) {
    row[pidx3(0, 0, 1) - 1] = z;
    if (order >= 3) {
      const int i_002 = pidx3(0, 0, 2), i_101 = pidx3(1, 0, 1), i_011 = pidx3(0, 1, 1);
      const double lj_2 = lj * lj;
      row[i_002 - 1] = z * z - C0[i_002] + lj_2 * Cj[i_002];
      row[i_101 - 1] = x * z - C0[i_101] + lj_2 * Cj[i_101];
      row[i_011 - 1] = y * z - C0[i_011] + lj_2 * Cj[i_011];
      if (order >= 4) {
        const int i_003 = pidx3(0, 0, 3), i_102 = pidx3(1, 0, 2), i_012 = pidx3(0, 1, 2), i_201 = pidx3(2, 0, 1),
                  i_111 = pidx3(1, 1, 1), i_021 = pidx3(0, 2, 1), i_200 = pidx3(2, 0, 0), i_020 = pidx3(0, 2, 0),
                  i_110 = pidx3(1, 1, 0);
        const double lj_3 = lj * lj * lj;
        row[i_003 - 1] = z * z * z - C0[i_003] + 3.0 * z * lj_2 * Cj[i_002] + lj_3 * Cj[i_003];
        row[i_102 - 1] = x * z * z - C0[i_102] + x * lj_2 * Cj[i_002] + 2.0 * z * lj_2 * Cj[i_101] + lj_3 * Cj[i_102];
        row[i_012 - 1] = y * z * z - C0[i_012] + y * lj_2 * Cj[i_002] + 2.0 * z * lj_2 * Cj[i_011] + lj_3 * Cj[i_012];
        row[i_201 - 1] = x * x * z - C0[i_201] + z * lj_2 * Cj[i_200] + 2.0 * x * lj_2 * Cj[i_101] + lj_3 * Cj[i_201];
        row[i_021 - 1] = y * y * z - C0[i_021] + z * lj_2 * Cj[i_020] + 2.0 * y * lj_2 * Cj[i_011] + lj_3 * Cj[i_021];
        row[i_111 - 1] = x * y * z - C0[i_111] + z * lj_2 * Cj[i_110] + y * lj_2 * Cj[i_101] + x * lj_2 * Cj[i_011] +
                         lj_3 * Cj[i_111];
      }
    }
  }
}

/// Plain strided view of a thread's scratch (stride 1 on the host, the number of resident threads on the device, so
/// that a warp's accesses to one element coalesce).
struct Strided {
  double *p;
  long long stride;
  ZFVM_HD double &operator[](int i) const { return p[(long long)i * stride]; }
};

/// W = pinv(A) for a full-column-rank rows x cols matrix (rows >= cols) by Householder QR.
///   R      rows x cols, row-major through the view; holds A on entry.  On exit the upper triangle is the R factor and
///          the strict lower part of column k keeps the reflector v_k (its leading entry is vk[k], |v_k|^2 is vn[k]).
///   y      rows entries of scratch
///   out(i, c, value)   receives W[i][c], 0 <= i < cols, 0 <= c < rows
/// A rank-deficient column (zero norm) is skipped like a reflector of zero length.
template <class View, class Out>
ZFVM_HD void pinv_householder(View R, View vk, View vn, View y, int rows, int cols, Out out) {
  for (int k = 0; k < cols; ++k) {
    double nrm = 0.0;
    for (int r = k; r < rows; ++r) nrm += R[r * cols + k] * R[r * cols + k];
    nrm = sqrt(nrm);
    vn[k] = 0.0;
    vk[k] = 0.0;
    if (nrm == 0.0) continue;
    const double akk = R[k * cols + k];
    const double alpha = (akk > 0 ? -nrm : nrm);
    const double v0 = akk - alpha;  // v_k[k]; v_k[r] = R[r][k] for r > k stays where it is
    double vnn = v0 * v0;
    for (int r = k + 1; r < rows; ++r) vnn += R[r * cols + k] * R[r * cols + k];
    if (vnn == 0.0) continue;
    vk[k] = v0;
    vn[k] = vnn;
    {  // column k itself: only the diagonal entry is ever read again
      double s = v0 * akk;
      for (int r = k + 1; r < rows; ++r) s += R[r * cols + k] * R[r * cols + k];
      s = 2.0 * s / vnn;
      R[k * cols + k] = akk - s * v0;
    }
    for (int c = k + 1; c < cols; ++c) {
      double s = v0 * R[k * cols + c];
      for (int r = k + 1; r < rows; ++r) s += R[r * cols + k] * R[r * cols + c];
      s = 2.0 * s / vnn;
      R[k * cols + c] -= s * v0;
      for (int r = k + 1; r < rows; ++r) R[r * cols + c] -= s * R[r * cols + k];
    }
  }
  // column c of Q^T (the reflectors applied to e_c), then back substitution with R
  for (int c = 0; c < rows; ++c) {
    for (int r = 0; r < rows; ++r) y[r] = (r == c) ? 1.0 : 0.0;
    for (int k = 0; k < cols; ++k) {
      const double vnn = vn[k];
      if (vnn == 0.0) continue;
      const double v0 = vk[k];
      double s = v0 * y[k];
      for (int r = k + 1; r < rows; ++r) s += R[r * cols + k] * y[r];
      s = 2.0 * s / vnn;
      y[k] -= s * v0;
      for (int r = k + 1; r < rows; ++r) y[r] -= s * R[r * cols + k];
    }
    for (int i = cols - 1; i >= 0; --i) {
      double s = y[i];
      for (int j = i + 1; j < cols; ++j) s -= R[i * cols + j] * y[j];
      y[i] = s / R[i * cols + i];
      out(i, c, y[i]);
    }
  }
}

}  // namespace lsq
}  // namespace zfvm
