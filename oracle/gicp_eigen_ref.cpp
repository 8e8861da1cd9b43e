/*
 * gicp_eigen_ref.cpp — the per-point expressions of fast_gicp evaluated with the REFERENCE's own vendored
 * Eigen (compiled with -I /root/reference/submodules/fast_gicp/thirdparty/Eigen into
 * oracle/_ref/libref_gicp_eigen.so by oracle/Makefile).  TEST INFRASTRUCTURE: pins the hand-written linear
 * algebra of gicp_oracle.cpp (and through it the CUDA kernels) against Eigen's JacobiSVD, Quaterniond,
 * Matrix4d::inverse, LDLT and the so3_exp formula.  Each function names the reference lines whose Eigen
 * expression it evaluates (fgi = submodules/fast_gicp/include/fast_gicp/gicp/impl/fast_gicp_impl.hpp).
 */
#include <Eigen/Cholesky>
#include <Eigen/Dense>
#include <Eigen/Geometry>
#include <Eigen/SVD>
#include <cmath>

using RowM3 = Eigen::Matrix<double, 3, 3, Eigen::RowMajor>;
using RowM4 = Eigen::Matrix<double, 4, 4, Eigen::RowMajor>;
using RowM6 = Eigen::Matrix<double, 6, 6, Eigen::RowMajor>;

extern "C" {

// fgi:638 JacobiSVD<Matrix3d>(cov, ComputeFullU | ComputeFullV)
void eig_svd3(const double* A9, double* U9, double* S3, double* V9) {
  Eigen::Matrix3d A = Eigen::Map<const RowM3>(A9);
  Eigen::JacobiSVD<Eigen::Matrix3d> svd(A, Eigen::ComputeFullU | Eigen::ComputeFullV);
  Eigen::Map<RowM3> map_U9(U9); map_U9 = svd.matrixU();
  Eigen::Map<RowM3> map_V9(V9); map_V9 = svd.matrixV();
  Eigen::Map<Eigen::Vector3d> map_S3(S3); map_S3 = svd.singularValues();
}

// fgi:639-646 Quaterniond qfrommat(svd.matrixU()); qfrommat.normalize();  -> x,y,z,w
void eig_quat_from_matrix(const double* M9, double* q4) {
  Eigen::Matrix3d M = Eigen::Map<const RowM3>(M9);
  Eigen::Quaterniond q(M);
  q.normalize();
  q4[0] = q.x(); q4[1] = q.y(); q4[2] = q.z(); q4[3] = q.w();
}

// fgi:631-697 from the neighbourhood covariance to (rotation quaternion, scales, regularised covariance)
void eig_cov_pipeline(const double* C9, int clamp, float* q4, float* s3, double* out9) {
  Eigen::Matrix3d cov = Eigen::Map<const RowM3>(C9);
  Eigen::JacobiSVD<Eigen::Matrix3d> svd(cov, Eigen::ComputeFullU | Eigen::ComputeFullV);
  Eigen::Quaterniond q(svd.matrixU());
  q.normalize();
  q4[0] = (float)q.x(); q4[1] = (float)q.y(); q4[2] = (float)q.z(); q4[3] = (float)q.w();
  Eigen::Vector3d scale = svd.singularValues().cwiseSqrt();
  s3[0] = (float)scale.x(); s3[1] = (float)scale.y(); s3[2] = (float)scale.z();
  Eigen::Vector3d values;
  if (svd.singularValues()(1) == 0) {
    values = Eigen::Vector3d(1e-9, 1e-9, 1e-9);
  } else {
    values = svd.singularValues() / svd.singularValues()(1);
    if (clamp) values = values.array().max(1e-3);
  }
  Eigen::Matrix3d r = svd.matrixU() * values.asDiagonal() * svd.matrixV().transpose();
  Eigen::Map<RowM3> map_out9(out9); map_out9 = r;
}

// fgi:864-898 setCovariances: note Quaterniond(w,x,y,z) constructor fed with the stored (x,y,z,w)
void eig_cov_from_qs(const float* rot4, const float* scale3, double* out9) {
  Eigen::Vector3d sv = {(double)scale3[0] * scale3[0], (double)scale3[1] * scale3[1], (double)scale3[2] * scale3[2]};
  if (sv(1) < 1e-3) sv = Eigen::Vector3d(1e-3, 1e-3, 1e-3);
  else sv = sv / sv(1);
  Eigen::Quaterniond q((double)rot4[0], (double)rot4[1], (double)rot4[2], (double)rot4[3]);
  q = q.normalized();
  Eigen::Matrix3d r = q.toRotationMatrix() * sv.asDiagonal() * q.toRotationMatrix().transpose();
  Eigen::Map<RowM3> map_out9(out9); map_out9 = r;
}

// fgi:280-291 RCR = cov_B + T cov_A T^T (4x4), RCR(3,3) = 1, inverse, (3,3) = 0  -> 3x3 block
void eig_mahalanobis(const double* covA9, const double* covB9, const double* T16, double* M9) {
  Eigen::Matrix4d A = Eigen::Matrix4d::Zero(), B = Eigen::Matrix4d::Zero();
  A.block<3, 3>(0, 0) = Eigen::Map<const RowM3>(covA9);
  B.block<3, 3>(0, 0) = Eigen::Map<const RowM3>(covB9);
  Eigen::Matrix4d T = Eigen::Map<const RowM4>(T16);
  Eigen::Matrix4d RCR = B + T * A * T.transpose();
  RCR(3, 3) = 1.0;
  Eigen::Matrix4d M = RCR.inverse();
  M(3, 3) = 0.0;
  Eigen::Map<RowM3> map_M9(M9); map_M9 = M.block<3, 3>(0, 0);
}

// fgi:317-339 one point's contribution to H, b and the error
void eig_linearize_point(const double* T16, const float* a3, const float* b3, const double* M9, double* H36, double* b6,
                         double* err) {
  const Eigen::Matrix4d Tm = Eigen::Map<const RowM4>(T16);
  Eigen::Isometry3d trans;
  trans.matrix() = Tm;
  Eigen::Matrix4d M = Eigen::Matrix4d::Zero();
  M.block<3, 3>(0, 0) = Eigen::Map<const RowM3>(M9);
  const Eigen::Vector4d mean_A(a3[0], a3[1], a3[2], 1.0), mean_B(b3[0], b3[1], b3[2], 1.0);
  const Eigen::Vector4d transed_mean_A = trans * mean_A;
  const Eigen::Vector4d error = mean_B - transed_mean_A;
  *err = error.transpose() * M * error;
  Eigen::Matrix<double, 4, 6> dtdx0 = Eigen::Matrix<double, 4, 6>::Zero();
  Eigen::Matrix3d skew = Eigen::Matrix3d::Zero();
  const Eigen::Vector3d x = transed_mean_A.head<3>();
  skew(0, 1) = -x[2]; skew(0, 2) = x[1]; skew(1, 0) = x[2]; skew(1, 2) = -x[0]; skew(2, 0) = -x[1]; skew(2, 1) = x[0];
  dtdx0.block<3, 3>(0, 0) = skew;
  dtdx0.block<3, 3>(0, 3) = -Eigen::Matrix3d::Identity();
  Eigen::Matrix<double, 6, 6> Hi = dtdx0.transpose() * M * dtdx0;
  Eigen::Matrix<double, 6, 1> bi = dtdx0.transpose() * M * error;
  Eigen::Map<RowM6> map_H36(H36); map_H36 = Hi;
  Eigen::Map<Eigen::Matrix<double, 6, 1>> map_b6(b6); map_b6 = bi;
}

// lsq:135-136 LDLT<Matrix6d>(H).solve(b)
void eig_ldlt_solve6(const double* H36, const double* b6, double* x6) {
  Eigen::Matrix<double, 6, 6> H = Eigen::Map<const RowM6>(H36);
  Eigen::Matrix<double, 6, 1> b = Eigen::Map<const Eigen::Matrix<double, 6, 1>>(b6);
  Eigen::LDLT<Eigen::Matrix<double, 6, 6>> solver(H);
  Eigen::Matrix<double, 6, 1> d = solver.solve(b);
  Eigen::Map<Eigen::Matrix<double, 6, 1>> map_x6(x6); map_x6 = d;
}

// so3.hpp:58-77 followed by toRotationMatrix (lsq:139)
void eig_so3_exp(const double* w3, double* R9) {
  const Eigen::Vector3d omega(w3[0], w3[1], w3[2]);
  const double theta_sq = omega.dot(omega);
  double imag_factor, real_factor;
  if (theta_sq < 1e-10) {
    const double theta_quad = theta_sq * theta_sq;
    imag_factor = 0.5 - 1.0 / 48.0 * theta_sq + 1.0 / 3840.0 * theta_quad;
    real_factor = 1.0 - 1.0 / 8.0 * theta_sq + 1.0 / 384.0 * theta_quad;
  } else {
    const double theta = std::sqrt(theta_sq), half_theta = 0.5 * theta;
    imag_factor = std::sin(half_theta) / theta;
    real_factor = std::cos(half_theta);
  }
  Eigen::Quaterniond q(real_factor, imag_factor * omega.x(), imag_factor * omega.y(), imag_factor * omega.z());
  Eigen::Map<RowM3> map_R9(R9); map_R9 = q.toRotationMatrix();
}

}  // extern "C"
