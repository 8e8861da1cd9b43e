// oracle/pcl_shim — see README.md.  TEST INFRASTRUCTURE (own code, nothing copied from PCL).
//
// pcl::Registration<PointSource, PointTarget, Scalar>: the members and control flow fast_gicp relies on
// (lsq_registration.hpp:36-43, fast_gicp.hpp:41-45, main.cpp:166-179): setInputSource/Target, align() — identity reset,
// output = copy of the input, computeTransformation(output, guess) —, initCompute() building the base class's own target
// kd-tree once per new target cloud (PCL does that in addition to fast_gicp's search_target_), getFitnessScore().
#pragma once
#include <Eigen/Core>
#include <iostream>
#include <limits>
#include <string>
#include <pcl/common/transforms.h>
#include <pcl/point_cloud.h>
#include <pcl/search/kdtree.h>

namespace pcl {

template <typename PointSource, typename PointTarget, typename Scalar = float>
class Registration {
public:
  using Matrix4 = Eigen::Matrix<Scalar, 4, 4>;
  using PointCloudSource = pcl::PointCloud<PointSource>;
  using PointCloudSourcePtr = typename PointCloudSource::Ptr;
  using PointCloudSourceConstPtr = typename PointCloudSource::ConstPtr;
  using PointCloudTarget = pcl::PointCloud<PointTarget>;
  using PointCloudTargetPtr = typename PointCloudTarget::Ptr;
  using PointCloudTargetConstPtr = typename PointCloudTarget::ConstPtr;
  using KdTree = pcl::search::KdTree<PointTarget>;
  using KdTreePtr = typename KdTree::Ptr;
  using Ptr = shared_ptr<Registration<PointSource, PointTarget, Scalar>>;
  using ConstPtr = shared_ptr<const Registration<PointSource, PointTarget, Scalar>>;

  Registration()
      : tree_(new KdTree),
        nr_iterations_(0),
        max_iterations_(10),
        final_transformation_(Matrix4::Identity()),
        transformation_(Matrix4::Identity()),
        previous_transformation_(Matrix4::Identity()),
        transformation_epsilon_(0.0),
        transformation_rotation_epsilon_(0.0),
        euclidean_fitness_epsilon_(-std::numeric_limits<double>::max()),
        corr_dist_threshold_(std::sqrt(std::numeric_limits<double>::max())),
        converged_(false),
        target_cloud_updated_(true),
        source_cloud_updated_(true),
        force_no_recompute_(false) {}
  virtual ~Registration() {}

  virtual void setInputSource(const PointCloudSourceConstPtr& cloud) {
    source_cloud_updated_ = true;
    input_ = cloud;
  }
  inline PointCloudSourceConstPtr const getInputSource() { return input_; }
  virtual void setInputTarget(const PointCloudTargetConstPtr& cloud) {
    if (!cloud || cloud->points.empty()) {
      std::cerr << "[pcl::" << reg_name_ << "::setInputTarget] Invalid or empty point cloud dataset given!" << std::endl;
      return;
    }
    target_ = cloud;
    target_cloud_updated_ = true;
  }
  inline PointCloudTargetConstPtr const getInputTarget() { return target_; }

  inline Matrix4 getFinalTransformation() { return final_transformation_; }
  inline Matrix4 getLastIncrementalTransformation() { return transformation_; }
  inline void setMaximumIterations(int nr_iterations) { max_iterations_ = nr_iterations; }
  inline int getMaximumIterations() { return max_iterations_; }
  inline void setMaxCorrespondenceDistance(double distance_threshold) { corr_dist_threshold_ = distance_threshold; }
  inline double getMaxCorrespondenceDistance() { return corr_dist_threshold_; }
  inline void setTransformationEpsilon(double epsilon) { transformation_epsilon_ = epsilon; }
  inline double getTransformationEpsilon() { return transformation_epsilon_; }
  inline void setEuclideanFitnessEpsilon(double epsilon) { euclidean_fitness_epsilon_ = epsilon; }
  inline bool hasConverged() const { return converged_; }
  inline const std::string& getClassName() const { return reg_name_; }

  inline double getFitnessScore(double max_range = std::numeric_limits<double>::max()) {
    double fitness_score = 0.0;
    PointCloudSource input_transformed;
    transformPointCloud(*input_, input_transformed, final_transformation_);
    std::vector<int> nn_indices(1);
    std::vector<float> nn_dists(1);
    int nr = 0;
    for (std::size_t i = 0; i < input_transformed.points.size(); ++i) {
      PointTarget q;
      q.x = input_transformed.points[i].x;
      q.y = input_transformed.points[i].y;
      q.z = input_transformed.points[i].z;
      tree_->nearestKSearch(q, 1, nn_indices, nn_dists);
      if (nn_dists[0] <= max_range) {
        fitness_score += nn_dists[0];
        nr++;
      }
    }
    if (nr > 0) return fitness_score / nr;
    return std::numeric_limits<double>::max();
  }

  inline void align(PointCloudSource& output) { align(output, Matrix4::Identity()); }

  inline void align(PointCloudSource& output, const Matrix4& guess) {
    if (!initCompute()) return;
    // output starts as a copy of the input with the homogeneous coordinate set
    output.header = input_->header;
    output.points.assign(input_->points.begin(), input_->points.end());
    output.width = static_cast<unsigned int>(output.points.size());
    output.height = 1;
    output.is_dense = input_->is_dense;
    converged_ = false;
    final_transformation_ = transformation_ = previous_transformation_ = Matrix4::Identity();
    for (std::size_t i = 0; i < output.points.size(); ++i) output.points[i].data[3] = 1.0f;
    computeTransformation(output, guess);
  }

protected:
  bool initCompute() {
    if (!target_) {
      std::cerr << "[pcl::registration::" << reg_name_ << "::compute] No input target dataset was given!" << std::endl;
      return false;
    }
    if (target_cloud_updated_ && !force_no_recompute_) {
      tree_->setInputCloud(target_);
      target_cloud_updated_ = false;
    }
    if (!input_) {
      std::cerr << "[pcl::" << reg_name_ << "::compute] No input source dataset was given!" << std::endl;
      return false;
    }
    return true;
  }

  virtual void computeTransformation(PointCloudSource& output, const Matrix4& guess) = 0;

  std::string reg_name_;
  KdTreePtr tree_;
  int nr_iterations_;
  int max_iterations_;
  PointCloudSourceConstPtr input_;
  PointCloudTargetConstPtr target_;
  Matrix4 final_transformation_, transformation_, previous_transformation_;
  double transformation_epsilon_, transformation_rotation_epsilon_, euclidean_fitness_epsilon_;
  double corr_dist_threshold_;
  bool converged_;
  bool target_cloud_updated_, source_cloud_updated_, force_no_recompute_;

public:
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW
};

}  // namespace pcl
