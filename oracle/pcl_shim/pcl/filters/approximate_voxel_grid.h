// oracle/pcl_shim — see README.md.  TEST INFRASTRUCTURE (own code).  Voxel-centroid down-sampling with the interface of
// pcl::ApproximateVoxelGrid used by main.cpp:50-55,82-92 (`downsample`, `align_points`; never called by the SLAM).
// NOT bit-compatible with PCL's hashed approximation.
#pragma once
#include <cmath>
#include <cstdint>
#include <unordered_map>
#include <pcl/point_cloud.h>

namespace pcl {
template <typename PointT>
class ApproximateVoxelGrid {
public:
  void setLeafSize(float lx, float ly, float lz) { leaf_[0] = lx; leaf_[1] = ly; leaf_[2] = lz; }
  void setInputCloud(const typename PointCloud<PointT>::ConstPtr& cloud) { input_ = cloud; }
  void filter(PointCloud<PointT>& out) {
    struct Acc { double s[3] = {0, 0, 0}; int n = 0; std::size_t order = 0; };
    std::unordered_map<std::uint64_t, Acc> cells;
    std::vector<std::uint64_t> keys;
    for (const auto& p : input_->points) {
      std::uint64_t key = 0;
      const float c[3] = {p.x, p.y, p.z};
      for (int a = 0; a < 3; a++) {
        const std::int64_t v = static_cast<std::int64_t>(std::floor(c[a] / leaf_[a])) & 0x1fffff;
        key = (key << 21) | static_cast<std::uint64_t>(v);
      }
      auto it = cells.find(key);
      if (it == cells.end()) {
        it = cells.emplace(key, Acc()).first;
        keys.push_back(key);
      }
      for (int a = 0; a < 3; a++) it->second.s[a] += c[a];
      it->second.n++;
    }
    out.clear();
    for (std::uint64_t k : keys) {
      const Acc& a = cells[k];
      PointT q;
      q.x = static_cast<float>(a.s[0] / a.n);
      q.y = static_cast<float>(a.s[1] / a.n);
      q.z = static_cast<float>(a.s[2] / a.n);
      out.push_back(q);
    }
  }

private:
  float leaf_[3] = {1.f, 1.f, 1.f};
  typename PointCloud<PointT>::ConstPtr input_;
};
}  // namespace pcl
