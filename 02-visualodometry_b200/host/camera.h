// camera.h — pr::Camera with the reference's interface (src/camera.h:13-51), over the C-ABI.
#pragma once
#include "vo_math.h"

namespace pr {

class Camera {
 public:
  Camera(int rows = 100, int cols = 100, const vo::Mat3f& camera_matrix = vo::Mat3f::Identity(),
         const vo::Iso3f& world_in_camera_pose = vo::Iso3f::Identity())
      : rows_(rows), cols_(cols), K_(camera_matrix), pose_(world_in_camera_pose) {}

  // One point is host arithmetic in the reference too (inline, src/camera.h:24-36); evaluated here
  // in the same float32 order (this header must be compiled with -ffp-contract=off).
  bool projectPoint(vo::Vec2f& image_point, const vo::Vec3f& world_point) const;

  // Batch projection on the GPU (src/camera.cpp:14-35): keep_indices => one output per input with
  // (-1,-1) for invalid points, else compacted. Returns the number of points inside the image.
  int projectPoints(Vector2fVector& image_points, const Vector3fVector& world_points, bool keep_indices = false) const;

  const vo::Iso3f& worldInCameraPose() const { return pose_; }
  void setWorldInCameraPose(const vo::Iso3f& pose) { pose_ = pose; }
  const vo::Mat3f& cameraMatrix() const { return K_; }
  int rows() const { return rows_; }
  int cols() const { return cols_; }

 private:
  int rows_, cols_;
  vo::Mat3f K_;
  vo::Iso3f pose_;
};

}  // namespace pr
