// oracle/pcl_shim — see README.md.  TEST INFRASTRUCTURE (own code, nothing copied from PCL).
#pragma once
#include <Eigen/Core>
#include <Eigen/StdVector>
#include <memory>
#include <vector>
#include <pcl/pcl_macros.h>

namespace pcl {

struct PCLHeader {
  unsigned int seq = 0;
  unsigned long long stamp = 0;
  std::string frame_id;
};

template <typename PointT>
class PointCloud {
public:
  using PointType = PointT;
  using VectorType = std::vector<PointT, Eigen::aligned_allocator<PointT>>;
  using Ptr = shared_ptr<PointCloud<PointT>>;
  using ConstPtr = shared_ptr<const PointCloud<PointT>>;
  using iterator = typename VectorType::iterator;
  using const_iterator = typename VectorType::const_iterator;

  PCLHeader header;
  VectorType points;
  unsigned int width = 0, height = 0;
  bool is_dense = true;

  std::size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  void resize(std::size_t n) {
    points.resize(n);
    if (width * height != n) {
      width = static_cast<unsigned int>(n);
      height = 1;
    }
  }
  void clear() {
    points.clear();
    width = height = 0;
  }
  void push_back(const PointT& p) {
    points.push_back(p);
    width = static_cast<unsigned int>(points.size());
    height = 1;
  }
  const PointT& at(std::size_t n) const { return points.at(n); }
  PointT& at(std::size_t n) { return points.at(n); }
  const PointT& operator[](std::size_t n) const { return points[n]; }
  PointT& operator[](std::size_t n) { return points[n]; }
  iterator begin() { return points.begin(); }
  iterator end() { return points.end(); }
  const_iterator begin() const { return points.begin(); }
  const_iterator end() const { return points.end(); }
  Ptr makeShared() const { return Ptr(new PointCloud<PointT>(*this)); }
};

}  // namespace pcl
