// camera.h — drop-in for the reference's src/camera.h:13-51 (pr::Camera): same constructor and members; the batch
// projection runs on the GPU (vo_project_points), the single-point one stays inline host arithmetic in the
// reference's float32 order (compile with -ffp-contract=off).
#pragma once
#include "defs.h"

namespace pr {

class Camera {
 public:
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW;
  Camera(int rows = 100, int cols = 100, const Eigen::Matrix3f& camera_matrix = Eigen::Matrix3f::Identity(),
         const Eigen::Isometry3f& world_in_camera_pose = Eigen::Isometry3f::Identity())
      : _rows(rows), _cols(cols), _camera_matrix(camera_matrix), _world_in_camera_pose(world_in_camera_pose) {}

  inline bool projectPoint(Eigen::Vector2f& image_point, const Eigen::Vector3f& world_point) {
    Eigen::Vector3f camera_point = _world_in_camera_pose * world_point;
    if (camera_point.z() <= 0) return false;
    Eigen::Vector3f projected_point = _camera_matrix * camera_point;
    const float iz = (float)(1. / projected_point.z());
    image_point = Eigen::Vector2f(projected_point.x() * iz, projected_point.y() * iz);
    if (image_point.x() < 0 || image_point.x() > _cols - 1) return false;
    if (image_point.y() < 0 || image_point.y() > _rows - 1) return false;
    return true;
  }

  int projectPoints(Vector2fVector& image_points, const Vector3fVector& world_points, bool keep_indices = false);

  inline const Eigen::Isometry3f& worldInCameraPose() const { return _world_in_camera_pose; }
  inline void setWorldInCameraPose(const Eigen::Isometry3f& pose) { _world_in_camera_pose = pose; }
  inline const Eigen::Matrix3f& cameraMatrix() const { return _camera_matrix; }
  int rows() const { return _rows; }
  int cols() const { return _cols; }

 protected:
  int _rows;
  int _cols;
  Eigen::Matrix3f _camera_matrix;
  Eigen::Isometry3f _world_in_camera_pose;
};

}  // namespace pr
