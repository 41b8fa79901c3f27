// picp_solver.h — pr::PICPSolver with the reference's interface (src/picp_solver.h:21-59). The state
// (points, packed correspondence stream, pose, H, b) lives in HBM behind a vo_picp handle.
#pragma once
#include <memory>

#include "camera.h"

namespace pr {

class PICPSolver {
 public:
  PICPSolver();  // damping 1, kernel threshold 1000, like src/picp_solver.cpp:8-15

  // Copies the camera and UPLOADS the points (the reference keeps raw pointers, src/picp_solver.cpp:21-22,
  // which dangle in exec/icp_test.cpp:81-85; copying is the safe reading of that contract).
  void init(const Camera& camera, const Vector3fVector& world_points, const Vector2fVector& image_points);

  float kernelThreshold() const { return kernel_threshold_; }
  void setKernelThreshold(float t) { kernel_threshold_ = t; }
  const Camera& camera() const;  // pose is read back from the device on demand
  float chiInliers() const { return chi_inliers_; }
  float chiOutliers() const { return chi_outliers_; }
  int numInliers() const { return num_inliers_; }

  // One Gauss-Newton round (src/picp_solver.cpp:93-105). The correspondences (first: measurement,
  // second: model) are uploaded and gathered only when they differ from the previous call's.
  bool oneRound(const IntPairVector& correspondences, bool keep_outliers);

 private:
  struct Handle;
  std::shared_ptr<Handle> h_;  // shared: the reference copies solvers by value (src/cam.cpp:34)
  mutable Camera camera_;
  mutable bool pose_stale_;  // device pose is ahead of camera_
  float kernel_threshold_;
  float damping_;
  int min_num_inliers_;
  float chi_inliers_, chi_outliers_;
  int num_inliers_;
  const void* corr_ptr_;
  size_t corr_size_;
  uint64_t corr_hash_;
};

}  // namespace pr
