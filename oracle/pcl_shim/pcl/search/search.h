// oracle/pcl_shim — see README.md.  TEST INFRASTRUCTURE (own code, nothing copied from PCL / FLANN).
//
// pcl::search::Search<PointT> (abstract) and pcl::search::KdTree<PointT>: EXACT k-nearest-neighbour search with results
// in ascending (squared distance, point index) order.  PCL delegates to FLANN's KDTreeSingleIndex (exact, sorted,
// L2_Simple); the only freedom an exact search has is the order of equidistant points, which this shim fixes to
// "lower index first" (SURVEY.md §8c).  Squared distances: ((dx*dx + dy*dy) + dz*dz) in float, no contraction
// (compile with -ffp-contract=off or without FMA targets), like FLANN's L2_Simple for float points.
#pragma once
#include <algorithm>
#include <cmath>
#include <limits>
#include <memory>
#include <numeric>
#include <vector>
#include <pcl/point_cloud.h>

namespace pcl {
using Indices = std::vector<int>;
using IndicesPtr = shared_ptr<Indices>;
using IndicesConstPtr = shared_ptr<const Indices>;

namespace search {

template <typename PointT>
class Search {
public:
  using PointCloud = pcl::PointCloud<PointT>;
  using PointCloudPtr = typename PointCloud::Ptr;
  using PointCloudConstPtr = typename PointCloud::ConstPtr;
  using Ptr = shared_ptr<Search<PointT>>;
  using ConstPtr = shared_ptr<const Search<PointT>>;

  virtual ~Search() {}
  virtual void setInputCloud(const PointCloudConstPtr& cloud, const IndicesConstPtr& indices = IndicesConstPtr()) {
    input_ = cloud;
    indices_ = indices;
  }
  virtual PointCloudConstPtr getInputCloud() const { return input_; }
  virtual int nearestKSearch(const PointT& point, int k, std::vector<int>& k_indices,
                             std::vector<float>& k_sqr_distances) const = 0;

protected:
  PointCloudConstPtr input_;
  IndicesConstPtr indices_;
};

template <typename PointT>
class KdTree : public Search<PointT> {
public:
  using PointCloudConstPtr = typename Search<PointT>::PointCloudConstPtr;
  using Ptr = shared_ptr<KdTree<PointT>>;
  using ConstPtr = shared_ptr<const KdTree<PointT>>;

  explicit KdTree(bool sorted = true) { (void)sorted; }

  void setInputCloud(const PointCloudConstPtr& cloud, const IndicesConstPtr& indices = IndicesConstPtr()) override {
    Search<PointT>::setInputCloud(cloud, indices);
    build();
  }

  int nearestKSearch(const PointT& point, int k, std::vector<int>& k_indices,
                     std::vector<float>& k_sqr_distances) const override {
    const int n = static_cast<int>(pts_.size() / 3);
    if (k > n) k = n;
    k_indices.resize(k);
    k_sqr_distances.resize(k);
    if (k <= 0) return 0;
    const float q[3] = {point.x, point.y, point.z};
    if (k == 1) {
      Best1 b;
      search1(0, q, b);
      k_indices[0] = b.id;
      k_sqr_distances[0] = b.d2;
      return 1;
    }
    // bounded sorted list (k is small: 10 in the SLAM)
    int found = 0;
    float* d = k_sqr_distances.data();
    int* id = k_indices.data();
    searchk(0, q, k, d, id, found);
    return found;
  }

private:
  struct Node {
    float lo[3], hi[3];  // bounding box of the points below this node
    int left = -1, right = -1;
    int begin = 0, end = 0;  // leaf: range in order_
  };
  struct Best1 {
    float d2 = std::numeric_limits<float>::infinity();
    int id = std::numeric_limits<int>::max();
  };
  static constexpr int kLeaf = 12;

  std::vector<float> pts_;   // xyz in ORIGINAL order
  std::vector<int> order_;   // point ids, grouped by leaf
  std::vector<float> leaf_;  // xyz in leaf order (contiguous scans)
  std::vector<Node> nodes_;

  static inline float dist2(const float* a, const float* b) {
    const float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
    return (dx * dx + dy * dy) + dz * dz;
  }
  // lower bound of dist2(q, p) for every p in the box; float rounding is monotone, so the bound holds for the
  // rounded distances as well
  static inline float box_dist2(const Node& nd, const float* q) {
    float g[3];
    for (int a = 0; a < 3; a++) {
      const float below = nd.lo[a] - q[a], above = q[a] - nd.hi[a];
      g[a] = below > 0.f ? below : (above > 0.f ? above : 0.f);
    }
    return (g[0] * g[0] + g[1] * g[1]) + g[2] * g[2];
  }

  void build() {
    pts_.clear();
    order_.clear();
    nodes_.clear();
    leaf_.clear();
    if (!this->input_) return;
    const auto& cloud = *this->input_;
    const int n = static_cast<int>(cloud.size());
    pts_.resize(3 * static_cast<std::size_t>(n));
    for (int i = 0; i < n; i++) {
      pts_[3 * i] = cloud.points[i].x;
      pts_[3 * i + 1] = cloud.points[i].y;
      pts_[3 * i + 2] = cloud.points[i].z;
    }
    order_.resize(n);
    std::iota(order_.begin(), order_.end(), 0);
    if (n == 0) return;
    nodes_.reserve(2 * (n / kLeaf + 2));
    build_rec(0, n);
    leaf_.resize(3 * static_cast<std::size_t>(n));
    for (int i = 0; i < n; i++) {
      leaf_[3 * i] = pts_[3 * order_[i]];
      leaf_[3 * i + 1] = pts_[3 * order_[i] + 1];
      leaf_[3 * i + 2] = pts_[3 * order_[i] + 2];
    }
  }

  int build_rec(int begin, int end) {
    const int me = static_cast<int>(nodes_.size());
    nodes_.emplace_back();
    Node nd;
    for (int a = 0; a < 3; a++) {
      nd.lo[a] = std::numeric_limits<float>::infinity();
      nd.hi[a] = -std::numeric_limits<float>::infinity();
    }
    for (int i = begin; i < end; i++)
      for (int a = 0; a < 3; a++) {
        const float v = pts_[3 * order_[i] + a];
        // NaN coordinates never tighten the box: such points are unreachable except through an unbounded search
        if (v < nd.lo[a]) nd.lo[a] = v;
        if (v > nd.hi[a]) nd.hi[a] = v;
      }
    nd.begin = begin;
    nd.end = end;
    int axis = 0;
    float ext = nd.hi[0] - nd.lo[0];
    for (int a = 1; a < 3; a++)
      if (nd.hi[a] - nd.lo[a] > ext) {
        ext = nd.hi[a] - nd.lo[a];
        axis = a;
      }
    if (end - begin > kLeaf && ext > 0.f) {
      const int mid = (begin + end) / 2;
      std::nth_element(order_.begin() + begin, order_.begin() + mid, order_.begin() + end, [&](int a, int b) {
        const float va = pts_[3 * a + axis], vb = pts_[3 * b + axis];
        return va < vb || (va == vb && a < b);
      });
      const int l = build_rec(begin, mid);
      const int r = build_rec(mid, end);
      nd.left = l;
      nd.right = r;
    }
    nodes_[me] = nd;
    return me;
  }

  void search1(int ni, const float* q, Best1& b) const {
    const Node& nd = nodes_[ni];
    if (nd.left < 0) {
      for (int i = nd.begin; i < nd.end; i++) {
        const float d2 = dist2(&leaf_[3 * i], q);
        const int id = order_[i];
        if (d2 < b.d2 || (d2 == b.d2 && id < b.id)) {
          b.d2 = d2;
          b.id = id;
        }
      }
      return;
    }
    const float dl = box_dist2(nodes_[nd.left], q), dr = box_dist2(nodes_[nd.right], q);
    const int first = dl <= dr ? nd.left : nd.right, second = dl <= dr ? nd.right : nd.left;
    const float df = dl <= dr ? dl : dr, ds = dl <= dr ? dr : dl;
    if (!(df > b.d2)) search1(first, q, b);
    if (!(ds > b.d2)) search1(second, q, b);
  }

  static inline void insert(int k, float* d, int* id, int& found, float d2, int pid) {
    if (found == k) {
      if (d2 > d[k - 1] || (d2 == d[k - 1] && pid > id[k - 1])) return;
    }
    int pos = found < k ? found : k - 1;
    while (pos > 0 && (d[pos - 1] > d2 || (d[pos - 1] == d2 && id[pos - 1] > pid))) {
      d[pos] = d[pos - 1];
      id[pos] = id[pos - 1];
      pos--;
    }
    d[pos] = d2;
    id[pos] = pid;
    if (found < k) found++;
  }

  void searchk(int ni, const float* q, int k, float* d, int* id, int& found) const {
    const Node& nd = nodes_[ni];
    if (nd.left < 0) {
      for (int i = nd.begin; i < nd.end; i++) insert(k, d, id, found, dist2(&leaf_[3 * i], q), order_[i]);
      return;
    }
    const float dl = box_dist2(nodes_[nd.left], q), dr = box_dist2(nodes_[nd.right], q);
    const int first = dl <= dr ? nd.left : nd.right, second = dl <= dr ? nd.right : nd.left;
    const float df = dl <= dr ? dl : dr, ds = dl <= dr ? dr : dl;
    if (found < k || !(df > d[k - 1])) searchk(first, q, k, d, id, found);
    if (found < k || !(ds > d[k - 1])) searchk(second, q, k, d, id, found);
  }
};

}  // namespace search
}  // namespace pcl
