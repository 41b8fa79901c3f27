// cam.h — drop-in for the reference's src/cam.h:18-158 (class Cam): the same public interface over libvo_b200.so.
#pragma once
#include <Eigen/Core>
#include <memory>
#include <opencv2/core/eigen.hpp>

#include "camera.h"
#include "data_point.h"
#include "defs.h"
#include "my_utilities.h"
#include "picp_solver.h"

class Cam {
 public:
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW
  Cam();  // K = [180 0 320; 0 180 240; 0 0 1], 640x480, the camera-to-image rotation (src/cam.cpp:10-35)
  // cv::findEssentialMat(RANSAC) + cv::recoverPose, restated on the GPU (src/cam.cpp:37-91); mask: n x 1 CV_8U
  void computeEssentialAndRecoverPose(const std::vector<std::pair<Data_Point, Data_Point>>& matches, cv::Mat& mask);
  // cv::triangulatePoints + convertPointsFromHomogeneous on the GPU (src/cam.cpp:94-140); appends to points3D
  void triangulatePoints(const Eigen::Isometry3f& T1, const Eigen::Isometry3f& T2,
                         std::vector<std::pair<Data_Point, Data_Point>>& matches, std::vector<World_Point>& points3D);
  Eigen::Matrix3f getEigenCamera();
  cv::Mat getRotationMatrix() const { return R_; }
  cv::Mat getTranslationVector() const { return t_; }
  int getHeight() const;
  int getWidth() const;
  void initOneRound(std::vector<World_Point> world_points, std::vector<Data_Point> img_points);  // src/cam.cpp:178-189
  void oneRound(pr::IntPairVector correspondences);                                              // src/cam.cpp:191-224
  Eigen::Isometry3f getPose();
  void setPose(Eigen::Isometry3f pose);
  Eigen::Isometry3f cameraToImage();

 private:
  struct Impl;
  std::shared_ptr<Impl> impl_;
  cv::Mat R_, t_;  // 3x3 / 3x1 CV_64F as recoverPose returns them
};
