// gicp_math.cuh — fp64 per-point linear algebra of the GICP tracker, usable on host and device.
//
// The reference delegates these to Eigen (vendored under submodules/fast_gicp/thirdparty/Eigen,
// version 3.3.90): JacobiSVD<Matrix3d> (Eigen/src/SVD/JacobiSVD.h:664-779 with
// Eigen/src/misc/RealSvd2x2.h:19-49 and Eigen/src/Jacobi/Jacobi.h:92-125), Quaterniond(Matrix3d)
// (Eigen/src/Geometry/Quaternion.h:816-853), Quaternion::toRotationMatrix (:592-624), Matrix::inverse().
// Downstream map Gaussians depend on the exact U (column order and signs) that algorithm produces
// (SURVEY §7), so the two-sided Jacobi sweep is restated here step for step, not replaced by a
// different eigen-solver.  This translation unit is compiled with -fmad=false; oracle/gicp_oracle.cpp
// (its independent CPU restatement) with -ffp-contract=off, so both round identically.
#pragma once
#include <cuda_runtime.h>
#include <float.h>
#include <math.h>

namespace gsicp {

#define GM_HD __host__ __device__ __forceinline__

struct Rot2 {  // planar rotation (c, s)
  double c, s;
};

// rows p,q of a 3x3 (row-major a[r][c]) <- [c s; -s c] * rows
GM_HD void rot_rows(double a[3][3], int p, int q, Rot2 j) {
  if (j.c == 1.0 && j.s == 0.0) return;
  for (int i = 0; i < 3; i++) {
    const double x = a[p][i], y = a[q][i];
    a[p][i] = j.c * x + j.s * y;
    a[q][i] = -j.s * x + j.c * y;
  }
}
// columns p,q <- columns * rotation given as (c, s): col_p' = c*col_p + s*col_q ; col_q' = -s*col_p + c*col_q
GM_HD void rot_cols(double a[3][3], int p, int q, Rot2 j) {
  if (j.c == 1.0 && j.s == 0.0) return;
  for (int i = 0; i < 3; i++) {
    const double x = a[i][p], y = a[i][q];
    a[i][p] = j.c * x + j.s * y;
    a[i][q] = -j.s * x + j.c * y;
  }
}

// Two-sided Jacobi SVD of a real 3x3: A = U diag(S) V^T, S sorted descending, U/V as Eigen returns them.
GM_HD void svd3_jacobi(const double A[3][3], double U[3][3], double S[3], double V[3][3]) {
  const double precision = 2.0 * DBL_EPSILON;
  const double tiny = DBL_MIN;
  double scale = 0.0;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) scale = fmax(scale, fabs(A[i][j]));
  if (scale == 0.0) scale = 1.0;
  double W[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      W[i][j] = A[i][j] / scale;
      U[i][j] = (i == j) ? 1.0 : 0.0;
      V[i][j] = (i == j) ? 1.0 : 0.0;
    }
  double max_diag = fmax(fabs(W[0][0]), fmax(fabs(W[1][1]), fabs(W[2][2])));
  bool finished = false;
  int guard = 0;
  while (!finished && guard++ < 100) {
    finished = true;
    for (int p = 1; p < 3; p++) {
      for (int q = 0; q < p; q++) {
        const double threshold = fmax(tiny, precision * max_diag);
        if (fabs(W[p][q]) > threshold || fabs(W[q][p]) > threshold) {
          finished = false;
          // 2x2 real SVD step: symmetrise with rot1, diagonalise with a Jacobi rotation
          double m00 = W[p][p], m01 = W[p][q], m10 = W[q][p], m11 = W[q][q];
          Rot2 rot1;
          const double t = m00 + m11, d = m10 - m01;
          if (fabs(d) < tiny) {
            rot1.s = 0.0;
            rot1.c = 1.0;
          } else {
            const double u = t / d;
            const double tmp = sqrt(1.0 + u * u);
            rot1.s = 1.0 / tmp;
            rot1.c = u / tmp;
          }
          {  // m <- rot1 applied to its rows
            const double a0 = rot1.c * m00 + rot1.s * m10, a1 = rot1.c * m01 + rot1.s * m11;
            const double b0 = -rot1.s * m00 + rot1.c * m10, b1 = -rot1.s * m01 + rot1.c * m11;
            m00 = a0; m01 = a1; m10 = b0; m11 = b1;
          }
          Rot2 jr;
          {
            const double x = m00, y = m01, z = m11;
            const double deno = 2.0 * fabs(y);
            if (deno < tiny) {
              jr.c = 1.0;
              jr.s = 0.0;
            } else {
              const double tau = (x - z) / deno;
              const double w = sqrt(tau * tau + 1.0);
              const double tt = (tau > 0.0) ? 1.0 / (tau + w) : 1.0 / (tau - w);
              const double sign_t = tt > 0.0 ? 1.0 : -1.0;
              const double n = 1.0 / sqrt(tt * tt + 1.0);
              jr.s = -sign_t * (y / fabs(y)) * fabs(tt) * n;
              jr.c = n;
            }
          }
          // j_left = rot1 * transpose(j_right)
          Rot2 jl;
          jl.c = rot1.c * jr.c - rot1.s * (-jr.s);
          jl.s = rot1.c * (-jr.s) + rot1.s * jr.c;

          rot_rows(W, p, q, jl);
          rot_cols(U, p, q, jl);
          const Rot2 jrt = {jr.c, -jr.s};
          rot_cols(W, p, q, jrt);
          rot_cols(V, p, q, jrt);
          max_diag = fmax(max_diag, fmax(fabs(W[p][p]), fabs(W[q][q])));
        }
      }
    }
  }
  for (int i = 0; i < 3; i++) {
    const double a = W[i][i];
    S[i] = fabs(a);
    if (a < 0.0)
      for (int r = 0; r < 3; r++) U[r][i] = -U[r][i];
  }
  for (int i = 0; i < 3; i++) S[i] *= scale;
  for (int i = 0; i < 3; i++) {  // selection sort, descending, first maximum wins
    int pos = i;
    double mx = S[i];
    for (int j = i + 1; j < 3; j++)
      if (S[j] > mx) {
        mx = S[j];
        pos = j;
      }
    if (mx == 0.0) break;
    if (pos != i) {
      const double ts = S[i]; S[i] = S[pos]; S[pos] = ts;
      for (int r = 0; r < 3; r++) {
        const double tu = U[r][i]; U[r][i] = U[r][pos]; U[r][pos] = tu;
        const double tv = V[r][i]; V[r][i] = V[r][pos]; V[r][pos] = tv;
      }
    }
  }
}

// Quaternion (x,y,z,w) from a 3x3 by Shoemake's method, then normalised — what
// `Eigen::Quaterniond q(U); q.normalize()` computes, also when det(U) = -1 (fast_gicp_impl.hpp:639-646).
GM_HD void quat_from_matrix(const double m[3][3], double q[4]) {
  double t = m[0][0] + m[1][1] + m[2][2];
  if (t > 0.0) {
    t = sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (m[2][1] - m[1][2]) * t;
    q[1] = (m[0][2] - m[2][0]) * t;
    q[2] = (m[1][0] - m[0][1]) * t;
  } else {
    int i = 0;
    if (m[1][1] > m[0][0]) i = 1;
    if (m[2][2] > m[i][i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(m[i][i] - m[j][j] - m[k][k] + 1.0);
    q[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (m[k][j] - m[j][k]) * t;
    q[j] = (m[j][i] + m[i][j]) * t;
    q[k] = (m[k][i] + m[i][k]) * t;
  }
  const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}

// Rotation matrix of quaternion (x,y,z,w) (Quaternion.h:592-624).
GM_HD void quat_to_matrix(double x, double y, double z, double w, double R[3][3]) {
  const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0][0] = 1.0 - (tyy + tzz); R[0][1] = txy - twz;         R[0][2] = txz + twy;
  R[1][0] = txy + twz;         R[1][1] = 1.0 - (txx + tzz); R[1][2] = tyz - twx;
  R[2][0] = txz - twy;         R[2][1] = tyz + twx;         R[2][2] = 1.0 - (txx + tyy);
}

// out = A diag(v) B^T
GM_HD void a_diag_bt(const double A[3][3], const double v[3], const double B[3][3], double out[3][3]) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      out[i][j] = ((A[i][0] * v[0]) * B[j][0] + (A[i][1] * v[1]) * B[j][1]) + (A[i][2] * v[2]) * B[j][2];
}

// inverse of a 3x3 by cofactors; returns false when det == 0
GM_HD bool inverse3(const double a[3][3], double inv[3][3]) {
  const double c00 = a[1][1] * a[2][2] - a[1][2] * a[2][1];
  const double c01 = a[1][2] * a[2][0] - a[1][0] * a[2][2];
  const double c02 = a[1][0] * a[2][1] - a[1][1] * a[2][0];
  const double det = (a[0][0] * c00 + a[0][1] * c01) + a[0][2] * c02;
  if (det == 0.0) return false;
  const double id = 1.0 / det;
  inv[0][0] = c00 * id;
  inv[1][0] = c01 * id;
  inv[2][0] = c02 * id;
  inv[0][1] = (a[0][2] * a[2][1] - a[0][1] * a[2][2]) * id;
  inv[1][1] = (a[0][0] * a[2][2] - a[0][2] * a[2][0]) * id;
  inv[2][1] = (a[0][1] * a[2][0] - a[0][0] * a[2][1]) * id;
  inv[0][2] = (a[0][1] * a[1][2] - a[0][2] * a[1][1]) * id;
  inv[1][2] = (a[0][2] * a[1][0] - a[0][0] * a[1][2]) * id;
  inv[2][2] = (a[0][0] * a[1][1] - a[0][1] * a[1][0]) * id;
  return true;
}

// ---------------------------------------------------------------------------------------------------------------
// Levenberg-Marquardt step algebra of LsqRegistration (lsq_registration_impl.hpp:81-90, 125-173), shared by the host
// loop and the device-resident loop (align_lm_kernel): rigid transforms, so3_exp, pivoted 6x6 LDL^T.
// ---------------------------------------------------------------------------------------------------------------
struct Iso {  // rigid transform, double
  double R[3][3], t[3];
};

GM_HD Iso iso_mul(const Iso& a, const Iso& b) {  // a * b
  Iso r;
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) r.R[i][j] = (a.R[i][0] * b.R[0][j] + a.R[i][1] * b.R[1][j]) + a.R[i][2] * b.R[2][j];
    r.t[i] = ((a.R[i][0] * b.t[0] + a.R[i][1] * b.t[1]) + a.R[i][2] * b.t[2]) + a.t[i];
  }
  return r;
}

GM_HD void so3_exp_matrix(const double w[3], double R[3][3]) {  // FG/include/fast_gicp/so3/so3.hpp:58-77
  const double theta_sq = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  double imag, real;
  if (theta_sq < 1e-10) {
    const double theta_quad = theta_sq * theta_sq;
    imag = 0.5 - 1.0 / 48.0 * theta_sq + 1.0 / 3840.0 * theta_quad;
    real = 1.0 - 1.0 / 8.0 * theta_sq + 1.0 / 384.0 * theta_quad;
  } else {
    const double theta = sqrt(theta_sq);
    const double half = 0.5 * theta;
    imag = sin(half) / theta;
    real = cos(half);
  }
  quat_to_matrix(imag * w[0], imag * w[1], imag * w[2], real, R);
}

GM_HD void gm_swap(double& a, double& b) {
  const double t = a;
  a = b;
  b = t;
}

// Pivoted LDL^T of a symmetric 6x6 and solve, following Eigen's LDLT (Eigen/src/Cholesky/LDLT.h:
// unblocked lower factorisation with diagonal pivoting, then P^T L^-T D^-1 L^-1 P b).
GM_HD void ldlt_solve6(const double Hin[6][6], const double rhs[6], double x[6]) {
  // Every loop below has compile-time bounds once the k loop is unrolled and the pivot row is matched against the static
  // candidates p = k+1..5, so on the device the whole factorisation lives in registers (no local-memory arrays with
  // run-time indices): the LM decision is a serial section of align_lm_kernel.  Same operations in the same order as the
  // plain triple loop.
  constexpr int n = 6;
  double A[6][6];
  int tr[6];
#pragma unroll
  for (int i = 0; i < n; i++)
#pragma unroll
    for (int j = 0; j < n; j++) A[i][j] = Hin[i][j];
  bool stop = false;
#pragma unroll
  for (int k = 0; k < n; k++) {
    if (stop) {
      continue;
    }
    int piv = k;
    double big = fabs(A[k][k]);
#pragma unroll
    for (int i = k + 1; i < n; i++)
      if (fabs(A[i][i]) > big) {
        big = fabs(A[i][i]);
        piv = i;
      }
    tr[k] = piv;
#pragma unroll
    for (int p = k + 1; p < n; p++) {
      if (piv == p) {  // symmetric row/column interchange on the lower triangle
#pragma unroll
        for (int c = 0; c < k; c++) gm_swap(A[k][c], A[p][c]);
#pragma unroll
        for (int r = p + 1; r < n; r++) gm_swap(A[r][k], A[r][p]);
        gm_swap(A[k][k], A[p][p]);
#pragma unroll
        for (int i = k + 1; i < p; i++) gm_swap(A[i][k], A[p][i]);
      }
    }
    if (k > 0) {
      double temp[6];
#pragma unroll
      for (int c = 0; c < k; c++) temp[c] = A[c][c] * A[k][c];
      double acc = 0.0;
#pragma unroll
      for (int c = 0; c < k; c++) acc += A[k][c] * temp[c];
      A[k][k] -= acc;
#pragma unroll
      for (int r = k + 1; r < n; r++) {
        double a2 = 0.0;
#pragma unroll
        for (int c = 0; c < k; c++) a2 += A[r][c] * temp[c];
        A[r][k] -= a2;
      }
    }
    const double akk = A[k][k];
    const bool valid = fabs(akk) > 0.0;
    if (k == 0 && !valid) {
#pragma unroll
      for (int j = 0; j < n; j++) tr[j] = j;
      stop = true;
    } else if (valid) {
#pragma unroll
      for (int r = k + 1; r < n; r++) A[r][k] /= akk;
    }
  }
  double y[6];
#pragma unroll
  for (int i = 0; i < n; i++) y[i] = rhs[i];
#pragma unroll
  for (int k = 0; k < n; k++) {                                    // P b
#pragma unroll
    for (int p = k + 1; p < n; p++)
      if (tr[k] == p) gm_swap(y[k], y[p]);
  }
#pragma unroll
  for (int i = 0; i < n; i++)                                      // L^-1
#pragma unroll
    for (int c = 0; c < i; c++) y[i] -= A[i][c] * y[c];
  const double tol = 1.0 / DBL_MAX;
#pragma unroll
  for (int i = 0; i < n; i++) y[i] = (fabs(A[i][i]) > tol) ? y[i] / A[i][i] : 0.0;  // D^-1
#pragma unroll
  for (int i = n - 1; i >= 0; i--)                                 // L^-T
#pragma unroll
    for (int c = i + 1; c < n; c++) y[i] -= A[c][i] * y[c];
#pragma unroll
  for (int k = n - 1; k >= 0; k--) {                               // P^T
#pragma unroll
    for (int p = k + 1; p < n; p++)
      if (tr[k] == p) gm_swap(y[k], y[p]);
  }
#pragma unroll
  for (int i = 0; i < n; i++) x[i] = y[i];
}

GM_HD bool lm_is_converged(double rot_eps, double trans_eps, const Iso& delta) {  // lsq:81-90
  double m = 0.0;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) m = fmax(m, (1.0 / rot_eps) * fabs(delta.R[i][j] - (i == j ? 1.0 : 0.0)));
  double mt = 0.0;
  for (int i = 0; i < 3; i++) mt = fmax(mt, (1.0 / trans_eps) * fabs(delta.t[i]));
  return fmax(m, mt) < 1;
}

// One trial of step_lm (lsq:137-146): solve (H + lambda I) d = -b, delta = (so3_exp(d[0:3]), d[3:6]), xi = delta * x0.
GM_HD void lm_trial(const double H[6][6], const double b[6], double lambda, const Iso& x0, double d[6], Iso& delta, Iso& xi) {
  double A[6][6], nb[6];
  for (int i = 0; i < 6; i++) {
    for (int j = 0; j < 6; j++) A[i][j] = H[i][j] + (i == j ? lambda : 0.0);
    nb[i] = -b[i];
  }
  ldlt_solve6(A, nb, d);
  so3_exp_matrix(d, delta.R);
  delta.t[0] = d[3]; delta.t[1] = d[4]; delta.t[2] = d[5];
  xi = iso_mul(delta, x0);
}

}  // namespace gsicp
