// oracle/pcl_shim — see README.md.  TEST INFRASTRUCTURE (own code, nothing copied from PCL).
// Point types with PCL's memory layout: 16-byte aligned float[4] with data[3] = 1 (homogeneous coordinate).
#pragma once
#include <Eigen/Core>
#include <pcl/pcl_macros.h>

namespace pcl {

#define PCL_SHIM_POINT4D                                                                                      \
  union EIGEN_ALIGN16 {                                                                                       \
    float data[4];                                                                                            \
    struct {                                                                                                  \
      float x, y, z;                                                                                          \
    };                                                                                                        \
  };                                                                                                          \
  inline Eigen::Map<Eigen::Vector3f> getVector3fMap() { return Eigen::Map<Eigen::Vector3f>(data); }         \
  inline const Eigen::Map<const Eigen::Vector3f> getVector3fMap() const {                                     \
    return Eigen::Map<const Eigen::Vector3f>(data);                                                           \
  }                                                                                                           \
  inline Eigen::Map<Eigen::Vector4f, Eigen::Aligned> getVector4fMap() {                                      \
    return Eigen::Map<Eigen::Vector4f, Eigen::Aligned>(data);                                                 \
  }                                                                                                           \
  inline const Eigen::Map<const Eigen::Vector4f, Eigen::Aligned> getVector4fMap() const {                     \
    return Eigen::Map<const Eigen::Vector4f, Eigen::Aligned>(data);                                           \
  }

struct EIGEN_ALIGN16 PointXYZ {
  PCL_SHIM_POINT4D
  PointXYZ() : PointXYZ(0.f, 0.f, 0.f) {}
  PointXYZ(float x_, float y_, float z_) {
    x = x_; y = y_; z = z_;
    data[3] = 1.0f;
  }
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW
};

struct EIGEN_ALIGN16 PointXYZI {
  PCL_SHIM_POINT4D
  union {
    struct {
      float intensity;
    };
    float data_c[4];
  };
  PointXYZI() {
    x = y = z = 0.f;
    data[3] = 1.0f;
    intensity = 0.f;
  }
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW
};

struct EIGEN_ALIGN16 PointNormal {
  PCL_SHIM_POINT4D
  union EIGEN_ALIGN16 {
    float data_n[4];
    float normal[3];
    struct {
      float normal_x, normal_y, normal_z;
    };
  };
  union {
    struct {
      float curvature;
    };
    float data_c[4];
  };
  PointNormal() {
    x = y = z = 0.f;
    data[3] = 1.0f;
    normal_x = normal_y = normal_z = data_n[3] = 0.f;
    curvature = 0.f;
  }
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW
};

template <typename PointT>
inline bool isFinite(const PointT& p) {
  return std::isfinite(p.x) && std::isfinite(p.y) && std::isfinite(p.z);
}

}  // namespace pcl
