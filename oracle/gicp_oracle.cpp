/*
 * gicp_oracle.cpp — CPU restatement of the reference GICP tracker (fast_gicp FastGICP + LsqRegistration).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in gs_icp_slam_b200/ may import, link or call this file; only tests/,
 * __graft_entry__.smoke() and bench.py (cpu_baseline leg and `--impl reference`) use it.
 *
 * fast_gicp itself cannot be built in this image (PCL, FLANN and boost are absent), so this
 * file follows the reference source line ranges (FG = submodules/fast_gicp, fgi =
 * FG/include/fast_gicp/gicp/impl/fast_gicp_impl.hpp, lsq = .../lsq_registration_impl.hpp):
 *   Oracle::set_cloud            FG/src/python/main.cpp:37-45,167-168 ; fgi:94-105,120-130
 *   Oracle::covariances          fgi:382-479 (clamp) / 588-706 (source, filter) / 710-825 (target, filter)
 *   Oracle::covs_from_qs         fgi:828-902 (incl. the (w,x,y,z) constructor quirk :890-894)
 *   Oracle::update_correspondences / linearize / compute_error   fgi:242-293 / 296-352 / 355-378
 *   Oracle::step_lm / align      lsq:125-173 / lsq:53-78 + pcl::Registration::align ; is_converged lsq:81-90
 *   so3_exp                      FG/include/fast_gicp/so3/so3.hpp:58-77
 * The third-party arithmetic the reference calls is restated from its published algorithm:
 *   exact k-NN            PCL 1.10 pcl::search::KdTree -> FLANN 1.9.1 KDTreeSingleIndex (exact, L2, sorted
 *                         ascending; versions from docker_folder/Dockerfile:22-26, not pinned by the repo).
 *                         Here: an exact kd-tree; ties on equal distance resolved by LOWER INDEX (FLANN's tie
 *                         order is traversal-dependent and unspecified — "parity unpinned" for ties).
 *   JacobiSVD<Matrix3d>   Eigen 3.3.90 (vendored FG/thirdparty/Eigen): Eigen/src/SVD/JacobiSVD.h:664-779,
 *                         misc/RealSvd2x2.h:19-49, Jacobi/Jacobi.h:92-125
 *   Quaterniond(Matrix3d) Eigen/src/Geometry/Quaternion.h:816-853 ; toRotationMatrix :592-624
 *   LDLT<Matrix6d>        Eigen/src/Cholesky/LDLT.h (unblocked, diagonal pivoting)
 * PINNING: tests/test_gicp_oracle.py checks svd3/quaternion/cov/Mahalanobis/LDLT/so3_exp of this file against
 * the real Eigen headers (oracle/gicp_eigen_ref.cpp -> oracle/_ref/libref_gicp_eigen.so) and the whole
 * align() against the reference's own acceptance fixture (FG/data/251370668.pcd <-> 251371071.pcd with
 * FG/data/relative.txt, bound 0.05 m / 1 deg, FG/src/test/gicp_test.cpp:147-201) through committed golden
 * vectors.  Neighbour indices / covariances / H,b of the real fast_gicp binary are NOT pinned (it cannot
 * run here): "parity unpinned" at that level, stated in DESIGN.md.
 *
 * Deviations of O(machine epsilon), shared with the CUDA implementation so that both agree bit for bit:
 *   - regularised covariances and Mahalanobis matrices are stored as their upper triangle (symmetric);
 *   - sums over neighbours / matrix products are evaluated left to right without fma contraction
 *     (-ffp-contract=off); the CUDA translation unit is compiled with -fmad=false;
 *   - a singular RCR (reference: pseudo-inverse, fgi:283-286) contributes a zero Mahalanobis matrix.
 */
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <numeric>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// ------------------------------------------------------------------------------------------------
// exact kd-tree
// ------------------------------------------------------------------------------------------------
inline float dist2f(const float* a, const float* b) {
  const float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
  return (dx * dx + dy * dy) + dz * dz;
}

struct Cand {
  float d2;
  int idx;
};
inline bool cand_less(const Cand& a, const Cand& b) { return a.d2 < b.d2 || (a.d2 == b.d2 && a.idx < b.idx); }

struct KdTree {
  struct Node {
    int left = -1, right = -1;  // children, or -1 for a leaf
    int begin = 0, end = 0;     // leaf: range in `order`
    int dim = 0;
    float split = 0.f;
    float lo[3], hi[3];         // bounding box of the node's points
  };
  const float* pts = nullptr;
  int n = 0;
  std::vector<int> order;
  std::vector<Node> nodes;

  void build(const float* xyz, int count) {
    pts = xyz;
    n = count;
    order.resize(n);
    std::iota(order.begin(), order.end(), 0);
    nodes.clear();
    nodes.reserve(n / 4 + 16);
    if (n > 0) build_rec(0, n);
  }
  int build_rec(int b, int e) {
    const int id = (int)nodes.size();
    nodes.emplace_back();
    Node nd;
    nd.begin = b;
    nd.end = e;
    for (int d = 0; d < 3; d++) { nd.lo[d] = FLT_MAX; nd.hi[d] = -FLT_MAX; }
    for (int i = b; i < e; i++)
      for (int d = 0; d < 3; d++) {
        const float v = pts[3 * (size_t)order[i] + d];
        nd.lo[d] = std::min(nd.lo[d], v);
        nd.hi[d] = std::max(nd.hi[d], v);
      }
    if (e - b > 12) {
      int dim = 0;
      float ext = nd.hi[0] - nd.lo[0];
      for (int d = 1; d < 3; d++)
        if (nd.hi[d] - nd.lo[d] > ext) { ext = nd.hi[d] - nd.lo[d]; dim = d; }
      if (ext > 0.f) {
        const int mid = (b + e) / 2;
        std::nth_element(order.begin() + b, order.begin() + mid, order.begin() + e, [&](int x, int y) {
          const float vx = pts[3 * (size_t)x + dim], vy = pts[3 * (size_t)y + dim];
          return vx < vy || (vx == vy && x < y);
        });
        nd.dim = dim;
        nd.split = pts[3 * (size_t)order[mid] + dim];
        nodes[id] = nd;
        const int l = build_rec(b, mid);
        const int r = build_rec(mid, e);
        nodes[id].left = l;
        nodes[id].right = r;
        return id;
      }
    }
    nodes[id] = nd;
    return id;
  }
  static float box_dist2(const Node& nd, const float* q) {
    float s = 0.f;
    for (int d = 0; d < 3; d++) {
      float t = 0.f;
      if (q[d] < nd.lo[d]) t = nd.lo[d] - q[d];
      else if (q[d] > nd.hi[d]) t = q[d] - nd.hi[d];
      s += t * t;
    }
    return s;
  }
  // best: sorted ascending by (d2, idx), capacity k
  void search(const float* q, int k, std::vector<Cand>& best) const {
    best.clear();
    if (n == 0 || k <= 0) return;
    search_rec(0, q, k, best);
  }
  void search_rec(int id, const float* q, int k, std::vector<Cand>& best) const {
    const Node& nd = nodes[id];
    if ((int)best.size() == k) {
      // prune only when the box is strictly farther than the current k-th (keeps equal-distance ties reachable);
      // the 1e-6 relative slack covers the rounding of box_dist2 versus dist2f
      const float bd = box_dist2(nd, q);
      if (bd * (1.0f - 1e-5f) > best.back().d2) return;
    }
    if (nd.left < 0) {
      for (int i = nd.begin; i < nd.end; i++) {
        const int idx = order[i];
        Cand c{dist2f(pts + 3 * (size_t)idx, q), idx};
        if ((int)best.size() < k) {
          best.insert(std::upper_bound(best.begin(), best.end(), c, cand_less), c);
        } else if (cand_less(c, best.back())) {
          best.pop_back();
          best.insert(std::upper_bound(best.begin(), best.end(), c, cand_less), c);
        }
      }
      return;
    }
    const bool left_first = q[nd.dim] < nd.split;
    search_rec(left_first ? nd.left : nd.right, q, k, best);
    search_rec(left_first ? nd.right : nd.left, q, k, best);
  }
};

// ------------------------------------------------------------------------------------------------
// 3x3 linear algebra (restating Eigen, see header)
// ------------------------------------------------------------------------------------------------
struct Rot2 { double c, s; };
void rot_rows(double a[3][3], int p, int q, Rot2 j) {
  if (j.c == 1.0 && j.s == 0.0) return;
  for (int i = 0; i < 3; i++) {
    const double x = a[p][i], y = a[q][i];
    a[p][i] = j.c * x + j.s * y;
    a[q][i] = -j.s * x + j.c * y;
  }
}
void rot_cols(double a[3][3], int p, int q, Rot2 j) {
  if (j.c == 1.0 && j.s == 0.0) return;
  for (int i = 0; i < 3; i++) {
    const double x = a[i][p], y = a[i][q];
    a[i][p] = j.c * x + j.s * y;
    a[i][q] = -j.s * x + j.c * y;
  }
}

void svd3(const double A[3][3], double U[3][3], double S[3], double V[3][3]) {
  const double precision = 2.0 * DBL_EPSILON, tiny = DBL_MIN;
  double scale = 0.0;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) scale = std::fmax(scale, std::fabs(A[i][j]));
  if (scale == 0.0) scale = 1.0;
  double W[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      W[i][j] = A[i][j] / scale;
      U[i][j] = V[i][j] = (i == j) ? 1.0 : 0.0;
    }
  double max_diag = std::fmax(std::fabs(W[0][0]), std::fmax(std::fabs(W[1][1]), std::fabs(W[2][2])));
  bool finished = false;
  int guard = 0;
  while (!finished && guard++ < 100) {
    finished = true;
    for (int p = 1; p < 3; p++)
      for (int q = 0; q < p; q++) {
        const double threshold = std::fmax(tiny, precision * max_diag);
        if (std::fabs(W[p][q]) > threshold || std::fabs(W[q][p]) > threshold) {
          finished = false;
          double m00 = W[p][p], m01 = W[p][q], m10 = W[q][p], m11 = W[q][q];
          Rot2 r1;
          const double t = m00 + m11, d = m10 - m01;
          if (std::fabs(d) < tiny) { r1.s = 0.0; r1.c = 1.0; }
          else {
            const double u = t / d, tmp = std::sqrt(1.0 + u * u);
            r1.s = 1.0 / tmp;
            r1.c = u / tmp;
          }
          const double a0 = r1.c * m00 + r1.s * m10, a1 = r1.c * m01 + r1.s * m11;
          const double b1 = -r1.s * m01 + r1.c * m11;
          m00 = a0; m01 = a1; m11 = b1;
          Rot2 jr;
          const double deno = 2.0 * std::fabs(m01);
          if (deno < tiny) { jr.c = 1.0; jr.s = 0.0; }
          else {
            const double tau = (m00 - m11) / deno, w = std::sqrt(tau * tau + 1.0);
            const double tt = (tau > 0.0) ? 1.0 / (tau + w) : 1.0 / (tau - w);
            const double sign_t = tt > 0.0 ? 1.0 : -1.0;
            const double nn = 1.0 / std::sqrt(tt * tt + 1.0);
            jr.s = -sign_t * (m01 / std::fabs(m01)) * std::fabs(tt) * nn;
            jr.c = nn;
          }
          Rot2 jl{r1.c * jr.c - r1.s * (-jr.s), r1.c * (-jr.s) + r1.s * jr.c};
          rot_rows(W, p, q, jl);
          rot_cols(U, p, q, jl);
          const Rot2 jrt{jr.c, -jr.s};
          rot_cols(W, p, q, jrt);
          rot_cols(V, p, q, jrt);
          max_diag = std::fmax(max_diag, std::fmax(std::fabs(W[p][p]), std::fabs(W[q][q])));
        }
      }
  }
  for (int i = 0; i < 3; i++) {
    const double a = W[i][i];
    S[i] = std::fabs(a);
    if (a < 0.0)
      for (int r = 0; r < 3; r++) U[r][i] = -U[r][i];
  }
  for (int i = 0; i < 3; i++) S[i] *= scale;
  for (int i = 0; i < 3; i++) {
    int pos = i;
    double mx = S[i];
    for (int j = i + 1; j < 3; j++)
      if (S[j] > mx) { mx = S[j]; pos = j; }
    if (mx == 0.0) break;
    if (pos != i) {
      std::swap(S[i], S[pos]);
      for (int r = 0; r < 3; r++) { std::swap(U[r][i], U[r][pos]); std::swap(V[r][i], V[r][pos]); }
    }
  }
}

void quat_from_matrix(const double m[3][3], double q[4]) {  // x,y,z,w ; normalised
  double t = m[0][0] + m[1][1] + m[2][2];
  if (t > 0.0) {
    t = std::sqrt(t + 1.0);
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
    t = std::sqrt(m[i][i] - m[j][j] - m[k][k] + 1.0);
    q[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (m[k][j] - m[j][k]) * t;
    q[j] = (m[j][i] + m[i][j]) * t;
    q[k] = (m[k][i] + m[i][k]) * t;
  }
  const double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int i = 0; i < 4; i++) q[i] /= n;
}

void quat_to_matrix(double x, double y, double z, double w, double R[3][3]) {
  const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0][0] = 1.0 - (tyy + tzz); R[0][1] = txy - twz; R[0][2] = txz + twy;
  R[1][0] = txy + twz; R[1][1] = 1.0 - (txx + tzz); R[1][2] = tyz - twx;
  R[2][0] = txz - twy; R[2][1] = tyz + twx; R[2][2] = 1.0 - (txx + tyy);
}

void a_diag_bt(const double A[3][3], const double v[3], const double B[3][3], double out[3][3]) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      out[i][j] = ((A[i][0] * v[0]) * B[j][0] + (A[i][1] * v[1]) * B[j][1]) + (A[i][2] * v[2]) * B[j][2];
}

bool inverse3(const double a[3][3], double inv[3][3]) {
  const double c00 = a[1][1] * a[2][2] - a[1][2] * a[2][1];
  const double c01 = a[1][2] * a[2][0] - a[1][0] * a[2][2];
  const double c02 = a[1][0] * a[2][1] - a[1][1] * a[2][0];
  const double det = (a[0][0] * c00 + a[0][1] * c01) + a[0][2] * c02;
  if (det == 0.0) return false;
  const double id = 1.0 / det;
  inv[0][0] = c00 * id; inv[1][0] = c01 * id; inv[2][0] = c02 * id;
  inv[0][1] = (a[0][2] * a[2][1] - a[0][1] * a[2][2]) * id;
  inv[1][1] = (a[0][0] * a[2][2] - a[0][2] * a[2][0]) * id;
  inv[2][1] = (a[0][1] * a[2][0] - a[0][0] * a[2][1]) * id;
  inv[0][2] = (a[0][1] * a[1][2] - a[0][2] * a[1][1]) * id;
  inv[1][2] = (a[0][2] * a[1][0] - a[0][0] * a[1][2]) * id;
  inv[2][2] = (a[0][0] * a[1][1] - a[0][1] * a[1][0]) * id;
  return true;
}

void ldlt_solve6(const double Hin[6][6], const double rhs[6], double x[6]) {
  const int n = 6;
  double A[6][6];
  int tr[6];
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) A[i][j] = Hin[i][j];
  for (int k = 0; k < n; k++) {
    int piv = k;
    double big = std::fabs(A[k][k]);
    for (int i = k + 1; i < n; i++)
      if (std::fabs(A[i][i]) > big) { big = std::fabs(A[i][i]); piv = i; }
    tr[k] = piv;
    if (piv != k) {
      const int s = n - piv - 1;
      for (int c = 0; c < k; c++) std::swap(A[k][c], A[piv][c]);
      for (int r = 0; r < s; r++) std::swap(A[piv + 1 + r][k], A[piv + 1 + r][piv]);
      std::swap(A[k][k], A[piv][piv]);
      for (int i = k + 1; i < piv; i++) std::swap(A[i][k], A[piv][i]);
    }
    const int rs = n - k - 1;
    if (k > 0) {
      double temp[6];
      for (int c = 0; c < k; c++) temp[c] = A[c][c] * A[k][c];
      double acc = 0.0;
      for (int c = 0; c < k; c++) acc += A[k][c] * temp[c];
      A[k][k] -= acc;
      for (int r = 0; r < rs; r++) {
        double a2 = 0.0;
        for (int c = 0; c < k; c++) a2 += A[k + 1 + r][c] * temp[c];
        A[k + 1 + r][k] -= a2;
      }
    }
    const double akk = A[k][k];
    const bool valid = std::fabs(akk) > 0.0;
    if (k == 0 && !valid) {
      for (int j = 0; j < n; j++) tr[j] = j;
      break;
    }
    if (rs > 0 && valid)
      for (int r = 0; r < rs; r++) A[k + 1 + r][k] /= akk;
  }
  double y[6];
  for (int i = 0; i < n; i++) y[i] = rhs[i];
  for (int k = 0; k < n; k++) std::swap(y[k], y[tr[k]]);
  for (int i = 0; i < n; i++)
    for (int c = 0; c < i; c++) y[i] -= A[i][c] * y[c];
  const double tol = 1.0 / std::numeric_limits<double>::max();
  for (int i = 0; i < n; i++) y[i] = (std::fabs(A[i][i]) > tol) ? y[i] / A[i][i] : 0.0;
  for (int i = n - 1; i >= 0; i--)
    for (int c = i + 1; c < n; c++) y[i] -= A[c][i] * y[c];
  for (int k = n - 1; k >= 0; k--) std::swap(y[k], y[tr[k]]);
  for (int i = 0; i < n; i++) x[i] = y[i];
}

struct Iso { double R[3][3], t[3]; };

void so3_exp_matrix(const double w[3], double R[3][3]) {
  const double theta_sq = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  double imag, real;
  if (theta_sq < 1e-10) {
    const double theta_quad = theta_sq * theta_sq;
    imag = 0.5 - 1.0 / 48.0 * theta_sq + 1.0 / 3840.0 * theta_quad;
    real = 1.0 - 1.0 / 8.0 * theta_sq + 1.0 / 384.0 * theta_quad;
  } else {
    const double theta = std::sqrt(theta_sq), half = 0.5 * theta;
    imag = std::sin(half) / theta;
    real = std::cos(half);
  }
  quat_to_matrix(imag * w[0], imag * w[1], imag * w[2], real, R);
}

Iso iso_mul(const Iso& a, const Iso& b) {
  Iso r;
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) r.R[i][j] = (a.R[i][0] * b.R[0][j] + a.R[i][1] * b.R[1][j]) + a.R[i][2] * b.R[2][j];
    r.t[i] = ((a.R[i][0] * b.t[0] + a.R[i][1] * b.t[1]) + a.R[i][2] * b.t[2]) + a.t[i];
  }
  return r;
}

// ------------------------------------------------------------------------------------------------
// the registration object
// ------------------------------------------------------------------------------------------------
struct Cloud {
  std::vector<float> xyz;  // 3n
  int n = 0;
  KdTree tree;
  std::vector<double> cov;  // 6 per covariance (upper triangle)
  int cov_n = 0;
  std::vector<float> rots, scales;
  std::vector<int> filter;
  bool has_filter = false;
  int num_trackable = 0;
  std::vector<float> z_values;  // fgi:182-186 / 205-209
};

constexpr int kChunk = 128;  // reduction granularity of linearize (= the CUDA block size)

struct Oracle {
  Cloud src, tgt;
  double max_corr = (double)std::numeric_limits<float>::max();
  float knn_max = 0.5f;
  int k = 10, max_iterations = 64, lm_max_iterations = 10;
  double rot_eps = 2e-3, trans_eps = 5e-4, lm_init_lambda_factor = 1e-9, lm_lambda = -1.0;
  bool converged = false;
  int nr_iterations = 0, n_lin = 0, n_err = 0;
  float final_transformation[16];
  double final_hessian[36];
  std::vector<int> corr;
  std::vector<float> sqd;
  std::vector<double> mahal;  // 6 per source point

  void set_cloud(Cloud& c, const double* xyz64, const float* xyz32, int n) {
    c.n = n;
    c.xyz.resize((size_t)n * 3);
    for (size_t i = 0; i < (size_t)n * 3; i++) c.xyz[i] = xyz64 ? (float)xyz64[i] : xyz32[i];
    c.tree.build(c.xyz.data(), n);
    c.cov_n = 0;
    c.cov.clear();
    c.rots.clear();
    c.scales.clear();
  }

  // withz: calculate_covariances_withz (fgi:481-583) = the clamped variant with the exported scales divided by
  // z = max(1, z_value^1.5 * 2)
  int covariances(Cloud& c, bool with_filter, bool clamp, bool withz = false) {
    const int n = c.n;
    if (n == 0) { fprintf(stderr, "no point cloud\n"); return 0; }
    if (withz && (int)c.z_values.size() != n) return -4;
    int slots = n;
    if (with_filter) {
      if (!c.has_filter) {
        c.filter.resize(n);
        for (int i = 0; i < n; i++) c.filter[i] = i + 1;
        c.num_trackable = n;
      } else if ((int)c.filter.size() != n) return -4;
      slots = c.num_trackable;
    }
    c.cov.assign((size_t)slots * 6, 0.0);
    c.rots.assign((size_t)n * 4, 0.f);
    c.scales.assign((size_t)n * 3, 0.f);
    std::vector<float> new_xyz(with_filter ? (size_t)slots * 3 : 0);
    const float* xyz = c.xyz.data();
#pragma omp parallel
    {
      std::vector<Cand> nn;
#pragma omp for schedule(guided, 8)
      for (int i = 0; i < n; i++) {
        c.tree.search(xyz + 3 * (size_t)i, k, nn);
        int reliable = 0;
        for (size_t j = 0; j < nn.size(); j++)
          if (nn[j].d2 < knn_max) ++reliable;  // squared distance against knn_max_distance_ (fgi:620)
        double mean[3] = {0, 0, 0};
        for (int j = 0; j < reliable; j++)
          for (int d = 0; d < 3; d++) mean[d] += (double)xyz[3 * (size_t)nn[j].idx + d];
        for (int d = 0; d < 3; d++) mean[d] /= (double)reliable;
        double C[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
        for (int j = 0; j < reliable; j++) {
          const double dx = (double)xyz[3 * (size_t)nn[j].idx] - mean[0];
          const double dy = (double)xyz[3 * (size_t)nn[j].idx + 1] - mean[1];
          const double dz = (double)xyz[3 * (size_t)nn[j].idx + 2] - mean[2];
          C[0][0] += dx * dx; C[0][1] += dx * dy; C[0][2] += dx * dz;
          C[1][1] += dy * dy; C[1][2] += dy * dz; C[2][2] += dz * dz;
        }
        const double kd = (double)k;  // fgi:635: divides by k_correspondences_
        C[0][0] /= kd; C[0][1] /= kd; C[0][2] /= kd; C[1][1] /= kd; C[1][2] /= kd; C[2][2] /= kd;
        C[1][0] = C[0][1]; C[2][0] = C[0][2]; C[2][1] = C[1][2];
        double U[3][3], S[3], V[3][3], q[4];
        svd3(C, U, S, V);
        quat_from_matrix(U, q);
        for (int d = 0; d < 4; d++) c.rots[4 * (size_t)i + d] = (float)q[d];
        for (int d = 0; d < 3; d++) c.scales[3 * (size_t)i + d] = (float)std::sqrt(S[d]);
        if (withz) {
          const float z = (float)std::max(1., std::pow((double)c.z_values[i], 1.5) * 2.);  // fgi:534
          for (int d = 0; d < 3; d++) c.scales[3 * (size_t)i + d] = (float)std::sqrt(S[d]) / z;
        }
        int slot = i;
        if (with_filter) {
          if (c.filter[i] == 0) continue;
          slot = c.filter[i] - 1;
        }
        double values[3];
        if (S[1] == 0.0) values[0] = values[1] = values[2] = 1e-9;
        else {
          for (int d = 0; d < 3; d++) values[d] = S[d] / S[1];
          if (clamp)
            for (int d = 0; d < 3; d++) values[d] = std::fmax(values[d], 1e-3);
        }
        double Rg[3][3];
        a_diag_bt(U, values, V, Rg);
        double* o = &c.cov[6 * (size_t)slot];
        o[0] = Rg[0][0]; o[1] = Rg[0][1]; o[2] = Rg[0][2]; o[3] = Rg[1][1]; o[4] = Rg[1][2]; o[5] = Rg[2][2];
        if (with_filter)
          for (int d = 0; d < 3; d++) new_xyz[3 * (size_t)slot + d] = xyz[3 * (size_t)i + d];
      }
    }
    c.cov_n = slots;
    if (with_filter) {
      c.xyz.swap(new_xyz);
      c.n = slots;
      c.has_filter = false;
      c.tree.build(c.xyz.data(), c.n);
    }
    return 0;
  }

  void covs_from_qs(Cloud& c, const float* rots, const float* scales, int n) {
    c.rots.assign(rots, rots + (size_t)n * 4);
    c.scales.assign(scales, scales + (size_t)n * 3);
    c.cov.assign((size_t)n * 6, 0.0);
#pragma omp parallel for schedule(guided, 8)
    for (int i = 0; i < n; i++) {
      double sv[3];
      for (int d = 0; d < 3; d++) {
        const double s = (double)scales[3 * (size_t)i + d];
        sv[d] = s * s;
      }
      if (sv[1] < 1e-3) sv[0] = sv[1] = sv[2] = 1e-3;
      else {
        const double m = sv[1];
        for (int d = 0; d < 3; d++) sv[d] = sv[d] / m;
      }
      // Eigen::Quaterniond q(rot[0], rot[1], rot[2], rot[3]) — constructor order is (w, x, y, z)
      double w = (double)rots[4 * (size_t)i], x = (double)rots[4 * (size_t)i + 1], y = (double)rots[4 * (size_t)i + 2],
             z = (double)rots[4 * (size_t)i + 3];
      const double nrm = std::sqrt(x * x + y * y + z * z + w * w);
      if (nrm > 0.0) { x /= nrm; y /= nrm; z /= nrm; w /= nrm; }
      double R[3][3], Cc[3][3];
      quat_to_matrix(x, y, z, w, R);
      a_diag_bt(R, sv, R, Cc);
      double* o = &c.cov[6 * (size_t)i];
      o[0] = Cc[0][0]; o[1] = Cc[0][1]; o[2] = Cc[0][2]; o[3] = Cc[1][1]; o[4] = Cc[1][2]; o[5] = Cc[2][2];
    }
    c.cov_n = n;
  }

  // fgi:242-293 + 296-352.  out28 = 21 (upper H) + 6 (b) + err, reduced per 128-point chunk then in chunk order.
  void linearize(const Iso& T, double out28[28], bool update_corr) {
    const int n = src.n;
    corr.resize(n);
    sqd.resize(n);
    mahal.resize((size_t)n * 6);
    float Rf[3][3], tf[3];
    for (int i = 0; i < 3; i++) {
      for (int j = 0; j < 3; j++) Rf[i][j] = (float)T.R[i][j];
      tf[i] = (float)T.t[i];
    }
    const double thr2 = max_corr * max_corr;
    const int chunks = (n + kChunk - 1) / kChunk;
    std::vector<double> part((size_t)std::max(chunks, 1) * 28, 0.0);
    const float* sx = src.xyz.data();
    const float* tx = tgt.xyz.data();
#pragma omp parallel
    {
      std::vector<Cand> nn;
#pragma omp for schedule(dynamic, 1)
      for (int ch = 0; ch < chunks; ch++) {
        double* acc = &part[(size_t)ch * 28];
        for (int i = ch * kChunk; i < std::min(n, (ch + 1) * kChunk); i++) {
          const float* p = sx + 3 * (size_t)i;
          if (update_corr) {
            float pt[3];
            // Eigen evaluates the packet product trans_f * getVector4fMap() (fgi:260) as (c0*x + c1*y) + (c2*z + c3*1):
            // pinned against the reference build (oracle/_ref/fast_gicp, tests/test_gicp_reference.py)
            for (int d = 0; d < 3; d++) pt[d] = (Rf[d][0] * p[0] + Rf[d][1] * p[1]) + (Rf[d][2] * p[2] + tf[d]);
            tgt.tree.search(pt, 1, nn);
            const float d2 = nn.empty() ? FLT_MAX : nn[0].d2;
            sqd[i] = d2;
            corr[i] = (!nn.empty() && (double)d2 < thr2) ? nn[0].idx : -1;
          }
          const int j = corr[i];
          if (j < 0) continue;
          double M[3][3];
          double* mo = &mahal[6 * (size_t)i];
          if (update_corr) {
            const double* ca = &src.cov[6 * (size_t)i];
            const double* cb = &tgt.cov[6 * (size_t)j];
            const double A[3][3] = {{ca[0], ca[1], ca[2]}, {ca[1], ca[3], ca[4]}, {ca[2], ca[4], ca[5]}};
            const double B[3][3] = {{cb[0], cb[1], cb[2]}, {cb[1], cb[3], cb[4]}, {cb[2], cb[4], cb[5]}};
            double RA[3][3], RCR[3][3];
            for (int r = 0; r < 3; r++)
              for (int c = 0; c < 3; c++) RA[r][c] = (T.R[r][0] * A[0][c] + T.R[r][1] * A[1][c]) + T.R[r][2] * A[2][c];
            for (int r = 0; r < 3; r++)
              for (int c = 0; c < 3; c++)
                RCR[r][c] = B[r][c] + ((RA[r][0] * T.R[c][0] + RA[r][1] * T.R[c][1]) + RA[r][2] * T.R[c][2]);
            if (!inverse3(RCR, M))
              for (int r = 0; r < 3; r++)
                for (int c = 0; c < 3; c++) M[r][c] = 0.0;
            mo[0] = M[0][0]; mo[1] = M[0][1]; mo[2] = M[0][2]; mo[3] = M[1][1]; mo[4] = M[1][2]; mo[5] = M[2][2];
          }
          M[0][0] = mo[0]; M[0][1] = mo[1]; M[0][2] = mo[2]; M[1][0] = mo[1]; M[1][1] = mo[3]; M[1][2] = mo[4];
          M[2][0] = mo[2]; M[2][1] = mo[4]; M[2][2] = mo[5];
          const double ax = p[0], ay = p[1], az = p[2];
          const double qx = ((T.R[0][0] * ax + T.R[0][1] * ay) + T.R[0][2] * az) + T.t[0];
          const double qy = ((T.R[1][0] * ax + T.R[1][1] * ay) + T.R[1][2] * az) + T.t[1];
          const double qz = ((T.R[2][0] * ax + T.R[2][1] * ay) + T.R[2][2] * az) + T.t[2];
          const double e[3] = {(double)tx[3 * (size_t)j] - qx, (double)tx[3 * (size_t)j + 1] - qy, (double)tx[3 * (size_t)j + 2] - qz};
          double Me[3];
          for (int r = 0; r < 3; r++) Me[r] = (M[r][0] * e[0] + M[r][1] * e[1]) + M[r][2] * e[2];
          acc[27] += (e[0] * Me[0] + e[1] * Me[1]) + e[2] * Me[2];
          if (!update_corr) continue;  // compute_error: error only
          const double S[3][3] = {{0.0, -qz, qy}, {qz, 0.0, -qx}, {-qy, qx, 0.0}};  // skew(T p) (so3.hpp:21-31)
          double MS[3][3], H[6][6];
          for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) MS[r][c] = (M[r][0] * S[0][c] + M[r][1] * S[1][c]) + M[r][2] * S[2][c];
          for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) {
              H[r][c] = (S[0][r] * MS[0][c] + S[1][r] * MS[1][c]) + S[2][r] * MS[2][c];
              H[r][3 + c] = -((S[0][r] * M[0][c] + S[1][r] * M[1][c]) + S[2][r] * M[2][c]);
              H[3 + r][3 + c] = M[r][c];
            }
          int o = 0;
          for (int r = 0; r < 6; r++)
            for (int c = r; c < 6; c++) acc[o++] += H[r][c];
          for (int r = 0; r < 3; r++) {
            acc[21 + r] += (S[0][r] * Me[0] + S[1][r] * Me[1]) + S[2][r] * Me[2];
            acc[24 + r] += -Me[r];
          }
        }
      }
    }
    for (int k2 = 0; k2 < 28; k2++) out28[k2] = 0.0;
    for (int ch = 0; ch < chunks; ch++)
      for (int k2 = 0; k2 < 28; k2++) out28[k2] += part[(size_t)ch * 28 + k2];
  }

  bool is_converged(const Iso& d) const {
    double m = 0.0, mt = 0.0;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) m = std::max(m, (1.0 / rot_eps) * std::fabs(d.R[i][j] - (i == j ? 1.0 : 0.0)));
    for (int i = 0; i < 3; i++) mt = std::max(mt, (1.0 / trans_eps) * std::fabs(d.t[i]));
    return std::max(m, mt) < 1;
  }

  int step_lm(Iso& x0, Iso& delta) {
    double r28[28], H[6][6], b[6];
    linearize(x0, r28, true);
    n_lin++;
    int o = 0;
    for (int r = 0; r < 6; r++)
      for (int c = r; c < 6; c++) { H[r][c] = r28[o]; H[c][r] = r28[o]; o++; }
    for (int r = 0; r < 6; r++) b[r] = r28[21 + r];
    const double y0 = r28[27];
    if (lm_lambda < 0.0) {
      double mx = 0.0;
      for (int i = 0; i < 6; i++) mx = std::max(mx, std::fabs(H[i][i]));
      lm_lambda = lm_init_lambda_factor * mx;
    }
    double nu = 2.0;
    for (int it = 0; it < lm_max_iterations; it++) {
      double A[6][6], nb[6], d[6];
      for (int i = 0; i < 6; i++) {
        for (int j = 0; j < 6; j++) A[i][j] = H[i][j] + (i == j ? lm_lambda : 0.0);
        nb[i] = -b[i];
      }
      ldlt_solve6(A, nb, d);
      so3_exp_matrix(d, delta.R);
      delta.t[0] = d[3]; delta.t[1] = d[4]; delta.t[2] = d[5];
      const Iso xi = iso_mul(delta, x0);
      double e28[28];
      linearize(xi, e28, false);
      n_err++;
      const double yi = e28[27];
      double dot = 0.0;
      for (int i = 0; i < 6; i++) dot += d[i] * (lm_lambda * d[i] - b[i]);
      const double rho = (y0 - yi) / dot;
      if (rho < 0) {
        if (is_converged(delta)) return 1;
        lm_lambda = nu * lm_lambda;
        nu = 2 * nu;
        continue;
      }
      x0 = xi;
      lm_lambda = lm_lambda * std::max(1.0 / 3.0, 1 - std::pow(2 * rho - 1, 3));
      for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++) final_hessian[6 * i + j] = H[i][j];
      return 1;
    }
    return 0;
  }

  int align(const float guess[16], float out[16]) {
    if (tgt.n == 0 || src.n == 0) return -4;
    converged = false;
    n_lin = n_err = 0;
    if (src.cov_n != src.n)
      if (int e = covariances(src, true, false)) return e;
    if (tgt.cov_n != tgt.n)
      if (int e = covariances(tgt, false, true)) return e;
    if (src.n == 0) return -4;
    Iso x0;
    for (int i = 0; i < 3; i++) {
      for (int j = 0; j < 3; j++) x0.R[i][j] = (double)guess[4 * i + j];
      x0.t[i] = (double)guess[4 * i + 3];
    }
    lm_lambda = -1.0;
    int iters = 0;
    for (int i = 0; i < max_iterations && !converged; i++) {
      nr_iterations = i;
      iters = i + 1;
      Iso delta;
      const int rc = step_lm(x0, delta);
      if (rc == 0) { fprintf(stderr, "lm not converged!!\n"); break; }
      converged = is_converged(delta);
    }
    for (int i = 0; i < 3; i++) {
      for (int j = 0; j < 3; j++) final_transformation[4 * i + j] = (float)x0.R[i][j];
      final_transformation[4 * i + 3] = (float)x0.t[i];
    }
    final_transformation[12] = final_transformation[13] = final_transformation[14] = 0.f;
    final_transformation[15] = 1.f;
    std::memcpy(out, final_transformation, sizeof(float) * 16);
    return iters;
  }
};

void cov6_to9(const std::vector<double>& c6, int n, double* out) {
  for (int i = 0; i < n; i++) {
    const double* s = &c6[(size_t)i * 6];
    double* o = out + (size_t)i * 9;
    o[0] = s[0]; o[1] = s[1]; o[2] = s[2]; o[3] = s[1]; o[4] = s[3]; o[5] = s[4]; o[6] = s[2]; o[7] = s[4]; o[8] = s[5];
  }
}

}  // namespace

extern "C" {

void* go_create() {
  Oracle* o = new Oracle();
  for (int i = 0; i < 16; i++) o->final_transformation[i] = (i % 5 == 0) ? 1.f : 0.f;
  for (int i = 0; i < 36; i++) o->final_hessian[i] = (i % 7 == 0) ? 1.0 : 0.0;
  return o;
}
void go_destroy(void* h) { delete (Oracle*)h; }
int go_num_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void go_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}
void go_set_max_correspondence_distance(void* h, double d) { ((Oracle*)h)->max_corr = d; }
void go_set_max_knn_distance(void* h, double d) { ((Oracle*)h)->knn_max = (float)d; }
void go_set_correspondence_randomness(void* h, int k) { ((Oracle*)h)->k = k; }
void go_set_max_iterations(void* h, int n) { ((Oracle*)h)->max_iterations = n; }
void go_set_input_source(void* h, const double* xyz, int n) { Oracle* o = (Oracle*)h; o->set_cloud(o->src, xyz, nullptr, n); o->corr.clear(); }
void go_set_input_target(void* h, const double* xyz, int n) { Oracle* o = (Oracle*)h; o->set_cloud(o->tgt, xyz, nullptr, n); }
static void set_filter(Cloud& c, int nt, const int32_t* f, int n) {
  c.num_trackable = nt;
  c.filter.assign(f, f + n);
  c.has_filter = true;
}
void go_set_source_filter(void* h, int nt, const int32_t* f, int n) { set_filter(((Oracle*)h)->src, nt, f, n); }
void go_set_target_filter(void* h, int nt, const int32_t* f, int n) { set_filter(((Oracle*)h)->tgt, nt, f, n); }
int go_calculate_target_covariance_with_filter(void* h) { Oracle* o = (Oracle*)h; return o->covariances(o->tgt, true, false); }
int go_calculate_source_covariance(void* h) { Oracle* o = (Oracle*)h; return o->covariances(o->src, false, true); }
int go_calculate_target_covariance(void* h) { Oracle* o = (Oracle*)h; return o->covariances(o->tgt, false, true); }
int go_calculate_target_covariance_withz(void* h) { Oracle* o = (Oracle*)h; return o->covariances(o->tgt, false, true, true); }
void go_set_source_z_values(void* h, const float* z, int n) { ((Oracle*)h)->src.z_values.assign(z, z + n); }
void go_set_target_z_values(void* h, const float* z, int n) { ((Oracle*)h)->tgt.z_values.assign(z, z + n); }
// swapSourceAndTarget (fgi:66-76): clouds, search structures, covariances, rotations, scales; filters and z values stay
void go_swap_source_and_target(void* h) {
  Oracle* o = (Oracle*)h;
  std::swap(o->src, o->tgt);
  std::swap(o->src.filter, o->tgt.filter);
  std::swap(o->src.has_filter, o->tgt.has_filter);
  std::swap(o->src.num_trackable, o->tgt.num_trackable);
  std::swap(o->src.z_values, o->tgt.z_values);
  o->corr.clear();
  o->sqd.clear();
}
// pcl::Registration::getFitnessScore(max_range): mean squared distance of the transformed source points to their nearest
// target point over the points whose squared distance is <= max_range; DBL_MAX when there is none.
double go_get_fitness_score(void* h, double max_range) {
  Oracle* o = (Oracle*)h;
  const float* T = o->final_transformation;
  double sum = 0.0;
  long nr = 0;
  std::vector<Cand> nn;
  for (int i = 0; i < o->src.n; i++) {
    const float* p = &o->src.xyz[3 * (size_t)i];
    float q[3];
    for (int r = 0; r < 3; r++) q[r] = (T[4 * r] * p[0] + T[4 * r + 1] * p[1]) + (T[4 * r + 2] * p[2] + T[4 * r + 3]);  // pcl::transformPointCloud order
    o->tgt.tree.search(q, 1, nn);
    if (!nn.empty() && (double)nn[0].d2 <= max_range) {
      sum += (double)nn[0].d2;
      nr++;
    }
  }
  return nr > 0 ? sum / (double)nr : std::numeric_limits<double>::max();
}
void go_set_source_covariances_fromqs(void* h, const float* r, const float* s, int n) { Oracle* o = (Oracle*)h; o->covs_from_qs(o->src, r, s, n); }
void go_set_target_covariances_fromqs(void* h, const float* r, const float* s, int n) { Oracle* o = (Oracle*)h; o->covs_from_qs(o->tgt, r, s, n); }
int go_align(void* h, const float* guess, float* out) { return ((Oracle*)h)->align(guess, out); }
int go_has_converged(void* h) { return ((Oracle*)h)->converged; }
void go_get_final_hessian(void* h, double* out) { std::memcpy(out, ((Oracle*)h)->final_hessian, sizeof(double) * 36); }
int go_source_size(void* h) { return ((Oracle*)h)->src.n; }
int go_target_size(void* h) { return ((Oracle*)h)->tgt.n; }
int go_source_rotationsq_size(void* h) { return (int)((Oracle*)h)->src.rots.size(); }
int go_target_rotationsq_size(void* h) { return (int)((Oracle*)h)->tgt.rots.size(); }
int go_source_scales_size(void* h) { return (int)((Oracle*)h)->src.scales.size(); }
int go_target_scales_size(void* h) { return (int)((Oracle*)h)->tgt.scales.size(); }
void go_get_source_rotationsq(void* h, float* o) { auto& v = ((Oracle*)h)->src.rots; std::memcpy(o, v.data(), v.size() * 4); }
void go_get_target_rotationsq(void* h, float* o) { auto& v = ((Oracle*)h)->tgt.rots; std::memcpy(o, v.data(), v.size() * 4); }
void go_get_source_scales(void* h, float* o) { auto& v = ((Oracle*)h)->src.scales; std::memcpy(o, v.data(), v.size() * 4); }
void go_get_target_scales(void* h, float* o) { auto& v = ((Oracle*)h)->tgt.scales; std::memcpy(o, v.data(), v.size() * 4); }
int go_source_cov_size(void* h) { return ((Oracle*)h)->src.cov_n; }
int go_target_cov_size(void* h) { return ((Oracle*)h)->tgt.cov_n; }
void go_get_source_covariances(void* h, double* o) { Oracle* q = (Oracle*)h; cov6_to9(q->src.cov, q->src.cov_n, o); }
void go_get_target_covariances(void* h, double* o) { Oracle* q = (Oracle*)h; cov6_to9(q->tgt.cov, q->tgt.cov_n, o); }
int go_get_source_correspondence(void* h, int32_t* corr, float* sqd) {
  Oracle* o = (Oracle*)h;
  if ((int)o->corr.size() != o->src.n) return -4;
  std::memcpy(corr, o->corr.data(), o->corr.size() * 4);
  std::memcpy(sqd, o->sqd.data(), o->sqd.size() * 4);
  return 0;
}
int go_linearize(void* h, const double* pose16, double* H36, double* b6, double* err) {
  Oracle* o = (Oracle*)h;
  if (o->src.n == 0 || o->tgt.n == 0 || o->src.cov_n != o->src.n || o->tgt.cov_n != o->tgt.n) return -4;
  Iso x;
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) x.R[i][j] = pose16[4 * i + j];
    x.t[i] = pose16[4 * i + 3];
  }
  double r28[28];
  o->linearize(x, r28, true);
  int k = 0;
  for (int r = 0; r < 6; r++)
    for (int c = r; c < 6; c++) { H36[6 * r + c] = r28[k]; H36[6 * c + r] = r28[k]; k++; }
  for (int r = 0; r < 6; r++) b6[r] = r28[21 + r];
  *err = r28[27];
  return 0;
}
int go_compute_error(void* h, const double* pose16, double* err) {
  Oracle* o = (Oracle*)h;
  if ((int)o->corr.size() != o->src.n) return -4;
  Iso x;
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) x.R[i][j] = pose16[4 * i + j];
    x.t[i] = pose16[4 * i + 3];
  }
  double r28[28];
  o->linearize(x, r28, false);
  *err = r28[27];
  return 0;
}
void go_last_counts(void* h, int* n_lin, int* n_err) { *n_lin = ((Oracle*)h)->n_lin; *n_err = ((Oracle*)h)->n_err; }

/* component access for the pinning tests against Eigen (tests/test_gicp_oracle.py) */
void go_svd3(const double* A9, double* U9, double* S3, double* V9) {
  double A[3][3], U[3][3], V[3][3];
  std::memcpy(A, A9, sizeof(A));
  svd3(A, U, S3, V);
  std::memcpy(U9, U, sizeof(U));
  std::memcpy(V9, V, sizeof(V));
}
void go_quat_from_matrix(const double* M9, double* q4) {
  double M[3][3];
  std::memcpy(M, M9, sizeof(M));
  quat_from_matrix(M, q4);
}
void go_quat_to_matrix(const double* q4, double* R9) {
  double R[3][3];
  quat_to_matrix(q4[0], q4[1], q4[2], q4[3], R);
  std::memcpy(R9, R, sizeof(R));
}
int go_inverse3(const double* A9, double* I9) {
  double A[3][3], I[3][3];
  std::memcpy(A, A9, sizeof(A));
  const bool ok = inverse3(A, I);
  std::memcpy(I9, I, sizeof(I));
  return ok;
}
void go_ldlt_solve6(const double* H36, const double* b6, double* x6) {
  double H[6][6];
  std::memcpy(H, H36, sizeof(H));
  ldlt_solve6(H, b6, x6);
}
void go_so3_exp(const double* w3, double* R9) {
  double R[3][3];
  so3_exp_matrix(w3, R);
  std::memcpy(R9, R, sizeof(R));
}
/* exact k-NN of every point of a cloud in itself (sorted by (d2, idx)) — pins the CUDA grid search */
void go_knn(const float* xyz, int n, int k, int32_t* idx_out, float* d2_out) {
  KdTree t;
  t.build(xyz, n);
#pragma omp parallel
  {
    std::vector<Cand> nn;
#pragma omp for schedule(guided, 8)
    for (int i = 0; i < n; i++) {
      t.search(xyz + 3 * (size_t)i, k, nn);
      for (int j = 0; j < k; j++) {
        idx_out[(size_t)i * k + j] = j < (int)nn.size() ? nn[j].idx : -1;
        d2_out[(size_t)i * k + j] = j < (int)nn.size() ? nn[j].d2 : FLT_MAX;
      }
    }
  }
}

}  // extern "C"
