// cam.h — class Cam with the reference's interface (src/cam.h:18-158): hard-coded intrinsics, the
// essential-matrix initialisation, triangulation and the embedded PICP driver, over the C-ABI.
#pragma once
#include "camera.h"
#include "data_point.h"
#include "my_utilities.h"
#include "picp_solver.h"

class Cam {
 public:
  Cam();  // K = [180 0 320; 0 180 240; 0 0 1], 640x480, cameraToImage rotation (src/cam.cpp:10-35)

  // src/cam.cpp:37-91. `mask` receives recoverPose's inlier mask (0/255), one byte per match.
  // Stores [R|t]^-1 as the PICP camera pose, so getPose() returns camera-2-in-world.
  void computeEssentialAndRecoverPose(const std::vector<std::pair<Data_Point, Data_Point>>& matches,
                                      std::vector<uint8_t>& mask);

  // src/cam.cpp:94-140: appends one World_Point per match (descriptor / ids of the FIRST view's point).
  void triangulatePoints(const vo::Iso3f& T1, const vo::Iso3f& T2, std::vector<std::pair<Data_Point, Data_Point>>& matches,
                         std::vector<World_Point>& points3D);

  vo::Mat3f getEigenCamera() const { return K_; }
  int getHeight() const { return height_; }
  int getWidth() const { return width_; }
  const double* getRotationMatrix() const { return R_; }     // 3x3 row-major, CV_64F in the reference
  const double* getTranslationVector() const { return t_; }  // 3

  void initOneRound(const std::vector<World_Point>& world_points, const std::vector<Data_Point>& img_points);  // :178-189
  void oneRound(const pr::IntPairVector& correspondences);                                                     // :191-224
  vo::Iso3f getPose() const { return picp_cam_.worldInCameraPose(); }
  void setPose(const vo::Iso3f& pose) { picp_cam_.setWorldInCameraPose(pose); }
  vo::Iso3f cameraToImage() const { return camera_to_image_; }

 private:
  vo::Mat3f K_;
  vo::Iso3f camera_to_image_;
  float z_near_, z_far_;
  int width_, height_;
  double R_[9], t_[3];
  pr::Vector3fVector world_points_picp_;
  pr::Vector2fVector image_points_picp_;
  pr::Camera picp_cam_;
  pr::PICPSolver picp_solver_;
};
