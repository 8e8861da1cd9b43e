// oracle/pcl_shim — see README.md.  TEST INFRASTRUCTURE (own code, nothing copied from PCL).
#pragma once
#include <Eigen/Core>
#include <pcl/point_cloud.h>

namespace pcl {
// cloud_out = transform * cloud_in (rigid / affine 4x4, homogeneous coordinate kept at 1)
template <typename PointT, typename Scalar>
void transformPointCloud(const PointCloud<PointT>& cloud_in, PointCloud<PointT>& cloud_out,
                         const Eigen::Matrix<Scalar, 4, 4>& transform) {
  if (&cloud_in != &cloud_out) {
    cloud_out.header = cloud_in.header;
    cloud_out.is_dense = cloud_in.is_dense;
    cloud_out.points.assign(cloud_in.points.begin(), cloud_in.points.end());
    cloud_out.width = cloud_in.width;
    cloud_out.height = cloud_in.height;
  }
  const Eigen::Matrix<float, 4, 4> tf = transform.template cast<float>();
  for (std::size_t i = 0; i < cloud_out.points.size(); ++i) {
    const Eigen::Vector4f p(cloud_in.points[i].x, cloud_in.points[i].y, cloud_in.points[i].z, 1.0f);
    const Eigen::Vector4f q = tf * p;
    cloud_out.points[i].x = q[0];
    cloud_out.points[i].y = q[1];
    cloud_out.points[i].z = q[2];
  }
}
}  // namespace pcl
