// oracle/pcl_shim — boost::hash_combine (fast_vgicp_voxel.hpp:50-52; the voxel map is not on the SLAM's path).
#pragma once
#include <cstddef>
#include <functional>
namespace boost {
template <typename T>
inline void hash_combine(std::size_t& seed, const T& v) {
  seed ^= std::hash<T>()(v) + 0x9e3779b97f4a7c15ull + (seed << 6) + (seed >> 2);
}
}  // namespace boost
