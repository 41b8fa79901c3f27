// picp_solver.h — drop-in for the reference's src/picp_solver.h:21-59 (pr::PICPSolver): the same public interface;
// the state (points, correspondences, pose, H, b) lives in HBM behind a vo_picp handle of libvo_b200.so.
#pragma once
#include <memory>

#include "camera.h"
#include "defs.h"

namespace pr {

class PICPSolver {
 public:
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW;
  PICPSolver();
  // copies the camera and UPLOADS the points (the reference keeps raw pointers, src/picp_solver.cpp:21-22, which
  // dangle in exec/icp_test.cpp:81-85: the temporaries of extract_V3fV / extract_V2fV die with the statement)
  void init(const Camera& camera, const Vector3fVector& world_points, const Vector2fVector& image_points);
  inline float kernelThreshold() const { return _kernel_thereshold; }
  inline void setKernelThreshold(float kernel_threshold) { _kernel_thereshold = kernel_threshold; }
  const Camera& camera() const;
  const float chiInliers() const { return _chi_inliers; }
  const float chiOutliers() const { return _chi_outliers; }
  const int numInliers() const { return _num_inliers; }
  bool oneRound(const IntPairVector& correspondences, bool keep_outliers);

 protected:
  struct Impl;
  std::shared_ptr<Impl> _impl;  // shared: the reference copies solvers by value (src/cam.cpp:33-34)
  mutable Camera _camera;
  float _kernel_thereshold;
  float _damping;
  int _min_num_inliers;
  float _chi_inliers;
  float _chi_outliers;
  int _num_inliers;
};

}  // namespace pr
