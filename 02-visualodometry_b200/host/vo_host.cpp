// vo_host.cpp — process context, pr::Camera and pr::PICPSolver of the host mirror.
// Compiled with -ffp-contract=off: projectPoint below must round exactly like the reference's inline.
#include <cstdlib>
#include <stdexcept>
#include <string>

#include "camera.h"
#include "picp_solver.h"
#include <iostream>

namespace vo {

void check(int status, const char* what) {
  if (status == VO_OK) return;
  throw std::runtime_error(std::string(what) + ": " + vo_status_str(status) + " (" +
                           vo_last_error(default_ctx()) + ")");
}

vo_ctx* default_ctx() {
  static vo_ctx* ctx = [] {
    const char* dev = std::getenv("VO_B200_DEVICE");
    vo_ctx* c = nullptr;
    const int st = vo_ctx_create(dev ? std::atoi(dev) : 0, nullptr, &c);
    if (st != VO_OK)
      throw std::runtime_error(std::string("vo_ctx_create: ") + vo_status_str(st) +
                               " - no usable CUDA device (this library has no CPU fallback)");
    return c;
  }();
  return ctx;
}

}  // namespace vo

namespace pr {

// reference: src/camera.h:24-36
bool Camera::projectPoint(vo::Vec2f& image_point, const vo::Vec3f& world_point) const {
  const vo::Vec3f c = pose_ * world_point;
  if (c.z() <= 0) return false;
  float q[3];
  for (int i = 0; i < 3; ++i) {
    const float x0 = K_(i, 0) * c[0], x1 = K_(i, 1) * c[1], x2 = K_(i, 2) * c[2];
    q[i] = x0 + (x1 + x2);
  }
  const float iz = (float)(1. / (double)q[2]);
  image_point = vo::Vec2f(q[0] * iz, q[1] * iz);
  if (image_point.x() < 0 || image_point.x() > (float)(cols_ - 1)) return false;
  if (image_point.y() < 0 || image_point.y() > (float)(rows_ - 1)) return false;
  return true;
}

int Camera::projectPoints(Vector2fVector& image_points, const Vector3fVector& world_points, bool keep_indices) const {
  image_points.resize(world_points.size());
  int64_t n_out = 0, n_inside = 0;
  vo::check(vo_project_points(vo::default_ctx(), K_.data(), rows_, cols_, pose_.data(),
                              world_points.empty() ? nullptr : world_points[0].v, (int64_t)world_points.size(),
                              keep_indices ? 1 : 0, image_points.empty() ? nullptr : image_points[0].v, &n_out, &n_inside),
            "vo_project_points");
  image_points.resize((size_t)n_out);
  return (int)n_inside;
}

struct PICPSolver::Handle {
  vo_picp* p = nullptr;
  Handle() { vo::check(vo_picp_create(vo::default_ctx(), &p), "vo_picp_create"); }
  ~Handle() { vo_picp_destroy(p); }
  Handle(const Handle&) = delete;
  Handle& operator=(const Handle&) = delete;
};

PICPSolver::PICPSolver()
    : pose_stale_(false), kernel_threshold_(1000.f), damping_(1.f), min_num_inliers_(0), chi_inliers_(0.f),
      chi_outliers_(0.f), num_inliers_(0), corr_ptr_(nullptr), corr_size_(0), corr_hash_(0) {}

void PICPSolver::init(const Camera& camera, const Vector3fVector& world_points, const Vector2fVector& image_points) {
  // The reference copies solvers by value (src/cam.cpp:33-34); copies of this class share the device state until one
  // of them is initialised again, at which point it gets a device state of its own.
  if (!h_ || h_.use_count() > 1) h_ = std::make_shared<Handle>();
  camera_ = camera;
  pose_stale_ = false;
  vo::check(vo_picp_set_camera(h_->p, camera.cameraMatrix().data(), camera.rows(), camera.cols(),
                               camera.worldInCameraPose().data()),
            "vo_picp_set_camera");
  vo::check(vo_picp_set_points(h_->p, world_points.empty() ? nullptr : world_points[0].v, (int64_t)world_points.size(),
                               image_points.empty() ? nullptr : image_points[0].v, (int64_t)image_points.size()),
            "vo_picp_set_points");
  corr_ptr_ = nullptr;  // a new point set invalidates the gathered correspondence stream
  corr_size_ = 0;
  corr_hash_ = 0;
}

const Camera& PICPSolver::camera() const {
  if (pose_stale_ && h_) {
    vo::Iso3f T;
    vo::check(vo_picp_get_pose(h_->p, T.m), "vo_picp_get_pose");
    camera_.setWorldInCameraPose(T);
    pose_stale_ = false;
  }
  return camera_;
}

// Fingerprint of a correspondence vector: the whole of a small one, a bounded sample of a large one (head, tail and
// 4096 evenly spaced pairs) - at 10 M correspondences hashing all 80 MB on one core would cost more per round than the
// Gauss-Newton round itself.  A caller that rewrites a few pairs in the middle of a large vector in place, at the same
// address and size, must call init() again (the reference's drivers build a fresh vector per frame).
static uint64_t corr_fingerprint(const void* data, size_t n_pairs) {
  const uint64_t* p = (const uint64_t*)data;  // one IntPair = 8 bytes
  uint64_t h = 1469598103934665603ull;
  auto mix = [&](uint64_t v) { h = (h ^ v) * 1099511628211ull; h ^= h >> 29; };
  const size_t full = 8192;
  if (n_pairs <= full) {
    for (size_t i = 0; i < n_pairs; ++i) mix(p[i]);
    return h;
  }
  for (size_t i = 0; i < 2048; ++i) mix(p[i]);
  for (size_t i = n_pairs - 2048; i < n_pairs; ++i) mix(p[i]);
  const size_t step = n_pairs / 4096;
  for (size_t i = 0; i < n_pairs; i += step) mix(p[i]);
  return h;
}

bool PICPSolver::oneRound(const IntPairVector& correspondences, bool keep_outliers) {
  if (!h_) throw std::runtime_error("PICPSolver::oneRound before init");
  // the caller hands the same vector every iteration (exec/icp_test.cpp:94-107): upload + gather once
  const void* ptr = correspondences.empty() ? nullptr : (const void*)&correspondences[0];
  const uint64_t hash = corr_fingerprint(ptr, correspondences.size());
  if (ptr != corr_ptr_ || correspondences.size() != corr_size_ || hash != corr_hash_ || corr_ptr_ == nullptr) {
    vo::check(vo_picp_set_correspondences(h_->p, (const int32_t*)ptr, (int64_t)correspondences.size()),
              "vo_picp_set_correspondences");
    corr_ptr_ = ptr ? ptr : (const void*)this;
    corr_size_ = correspondences.size();
    corr_hash_ = hash;
  }
  vo_picp_stats st;
  vo::check(vo_picp_one_round(h_->p, kernel_threshold_, damping_, keep_outliers ? 1 : 0, &st), "vo_picp_one_round");
  chi_inliers_ = st.chi_inliers;
  chi_outliers_ = st.chi_outliers;
  num_inliers_ = st.num_inliers;
  pose_stale_ = true;
  if (num_inliers_ < min_num_inliers_) {  // dead in the reference too (min is 0, src/picp_solver.cpp:97-100)
    std::cerr << "too few inliers, skipping" << std::endl;
    return false;
  }
  return true;
}

}  // namespace pr
