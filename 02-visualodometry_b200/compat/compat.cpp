// compat.cpp — bodies of the replacement headers in compat/src/ (the drop-in for the reference's cam.cpp camera.cpp
// picp_solver.cpp my_utilities.cpp, CMakeLists.txt:25-59): thin adapters from the Eigen / OpenCV typed interface of
// the reference to the host mirror in ../host/ (which talks to libvo_b200.so through the C-ABI).  The host mirror
// defines classes with the reference's own names (Cam, pr::Camera, pr::PICPSolver, Data_Point, ...) on dependency-
// free types; to link both into one program its sources are compiled here inside the namespace `vohost`.
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <limits>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/vo_b200.h"

namespace vohost {
#include "../host/vo_math.h"
#include "../host/vo_records.h"
#include "../host/camera.h"
#include "../host/picp_solver.h"
#include "../host/my_utilities.h"
#include "../host/cam.h"
#include "../host/vo_host.cpp"
#include "../host/my_utilities.cpp"
#include "../host/cam.cpp"
}  // namespace vohost

#include "src/cam.h"
#include "src/my_utilities.h"

namespace {

vohost::vo::Iso3f to_host(const Eigen::Isometry3f& T) {
  vohost::vo::Iso3f r;
  T.to12(r.m);
  return r;
}
Eigen::Isometry3f from_host(const vohost::vo::Iso3f& T) { return Eigen::Isometry3f::from12(T.m); }
vohost::vo::Mat3f to_host(const Eigen::Matrix3f& K) {
  vohost::vo::Mat3f r;
  for (int i = 0; i < 9; ++i) r.m[i] = K.m[i];
  return r;
}
Eigen::Matrix3f from_host(const vohost::vo::Mat3f& K) {
  Eigen::Matrix3f r;
  for (int i = 0; i < 9; ++i) r.m[i] = K.m[i];
  return r;
}
vohost::vo::Descriptor to_host(const Eigen::VectorXf& d) { return vohost::vo::Descriptor(d.data(), d.data() + d.size()); }
Eigen::VectorXf from_host(const vohost::vo::Descriptor& d) {
  Eigen::VectorXf r((int)d.size());
  for (size_t i = 0; i < d.size(); ++i) r[(int)i] = d[i];
  return r;
}
vohost::Data_Point to_host(const Data_Point& p) {
  return vohost::Data_Point(p.id_meas, p.id_real, vohost::vo::Point2f(p.coordinates.x, p.coordinates.y), to_host(p.descriptor));
}
Data_Point from_host(const vohost::Data_Point& p) {
  return Data_Point(p.id_meas, p.id_real, cv::Point2f(p.coordinates.x, p.coordinates.y), from_host(p.descriptor));
}
vohost::World_Point to_host(const World_Point& p) {
  return vohost::World_Point(vohost::vo::Point3f(p.coordinates.x, p.coordinates.y, p.coordinates.z), to_host(p.descriptor),
                             p.id_meas, p.id_real);
}
World_Point from_host(const vohost::World_Point& p) {
  return World_Point(cv::Point3f(p.coordinates.x, p.coordinates.y, p.coordinates.z), from_host(p.descriptor), p.id_meas,
                     p.id_real);
}
template <class A, class B>
std::vector<B> map_vec(const std::vector<A>& v, B (*f)(const A&)) {
  std::vector<B> r;
  r.reserve(v.size());
  for (const A& a : v) r.push_back(f(a));
  return r;
}
std::vector<vohost::Data_Point> to_host(const std::vector<Data_Point>& v) {
  return map_vec<Data_Point, vohost::Data_Point>(v, to_host);
}
std::vector<vohost::World_Point> to_host(const std::vector<World_Point>& v) {
  return map_vec<World_Point, vohost::World_Point>(v, to_host);
}
std::vector<std::pair<vohost::Data_Point, vohost::Data_Point>> to_host(const std::vector<std::pair<Data_Point, Data_Point>>& v) {
  std::vector<std::pair<vohost::Data_Point, vohost::Data_Point>> r;
  r.reserve(v.size());
  for (const auto& p : v) r.emplace_back(to_host(p.first), to_host(p.second));
  return r;
}
vohost::pr::Vector3fVector to_host(const pr::Vector3fVector& v) {
  vohost::pr::Vector3fVector r(v.size());
  if (!v.empty()) std::memcpy(r.data(), v.data(), v.size() * 12);
  return r;
}
vohost::pr::Vector2fVector to_host(const pr::Vector2fVector& v) {
  vohost::pr::Vector2fVector r(v.size());
  if (!v.empty()) std::memcpy(r.data(), v.data(), v.size() * 8);
  return r;
}
vohost::pr::Camera to_host(const pr::Camera& c) {
  return vohost::pr::Camera(c.rows(), c.cols(), to_host(c.cameraMatrix()), to_host(c.worldInCameraPose()));
}

}  // namespace

// ---------------------------------------------------------------------------------------------- my_utilities
namespace vo_compat {
void match_rows(const float* descA, long long n1, const float* descB, long long n2, int dim, const int* idA,
                const int* idB, std::vector<int>& pairs, long long stats[2]) {
  pairs.assign((size_t)2 * n1, 0);
  int64_t n = 0, st[2] = {0, 0};
  vohost::vo::check(vo_match(vohost::vo::default_ctx(), descA, n1, descB, n2, dim, DISTANCE_THRESHOLD, RATIO_THRESHOLD, idA,
                             idB, 0, n1, pairs.data(), n1, &n, st),
                    "vo_match");
  pairs.resize((size_t)2 * n);
  stats[0] = st[0];
  stats[1] = st[1];
}
}  // namespace vo_compat

void extract_coordinates_from_matches(std::vector<std::pair<Data_Point, Data_Point>> matches,
                                      std::vector<cv::Point2f>& matches1, std::vector<cv::Point2f>& matches2) {
  for (const auto& m : matches) {
    matches1.push_back(m.first.coordinates);
    matches2.push_back(m.second.coordinates);
  }
}
Vector2fVector extract_V2fV(const std::vector<Data_Point>& points) {
  Vector2fVector r;
  r.reserve(points.size());
  for (const auto& p : points) r.push_back(Eigen::Vector2f(p.coordinates.x, p.coordinates.y));
  return r;
}
Vector3fVector extract_V3fV(const std::vector<World_Point>& points) {
  Vector3fVector r;
  r.reserve(points.size());
  for (const auto& p : points) r.push_back(Eigen::Vector3f(p.coordinates.x, p.coordinates.y, p.coordinates.z));
  return r;
}
std::vector<std::string> split(const std::string& str, const std::string& delimiter) { return vohost::split(str, delimiter); }
static Measurement from_host(const vohost::Measurement& m) {
  Measurement r;
  r.seq = m.seq;
  r.gt_pose = Eigen::Vector3f(m.gt_pose[0], m.gt_pose[1], m.gt_pose[2]);
  r.odometry_pose = Eigen::Vector3f(m.odometry_pose[0], m.odometry_pose[1], m.odometry_pose[2]);
  for (const auto& p : m.data_points) r.data_points.push_back(from_host(p));
  return r;
}
Measurement extract_measurement(const std::string& filename) { return from_host(vohost::extract_measurement(filename)); }
std::vector<Measurement> extract_measurements(const std::string& filename, int n_meas) {
  std::vector<Measurement> r;
  for (const auto& m : vohost::extract_measurements(filename, n_meas)) r.push_back(from_host(m));
  return r;
}
std::vector<Measurement> load_and_initialize_data(const std::string& path, int num_measurements) {
  std::vector<Measurement> r;
  for (const auto& m : vohost::load_and_initialize_data(path, num_measurements)) r.push_back(from_host(m));
  return r;
}
std::vector<World_Point> load_world_points(const std::string& filename) {
  std::vector<World_Point> r;
  for (const auto& p : vohost::load_world_points(filename)) r.push_back(from_host(p));
  return r;
}
Eigen::Isometry3f oneRound(Eigen::Isometry3f last_pose_estimate, pr::Camera& pr_cam,
                           const pr::Vector3fVector& world_points, const pr::Vector2fVector& image_points,
                           const pr::IntPairVector& correspondences) {
  vohost::pr::Camera hc = to_host(pr_cam);
  const vohost::vo::Iso3f out =
      vohost::oneRound(to_host(last_pose_estimate), hc, to_host(world_points), to_host(image_points), correspondences);
  if (correspondences.size() >= 10) pr_cam.setWorldInCameraPose(last_pose_estimate);  // src/my_utilities.cpp:277
  return from_host(out);
}
Eigen::Isometry3f augment_pose(const Eigen::Vector3f& pose) {
  return from_host(vohost::augment_pose(vohost::vo::Vec3f(pose[0], pose[1], pose[2])));
}
float compute_scale(const std::vector<Eigen::Vector3f>& a, const std::vector<Eigen::Vector3f>& b) {
  std::vector<vohost::vo::Vec3f> ha, hb;
  for (const auto& p : a) ha.emplace_back(p[0], p[1], p[2]);
  for (const auto& p : b) hb.emplace_back(p[0], p[1], p[2]);
  return vohost::compute_scale(ha, hb);
}
void create_plot(const std::vector<Eigen::Isometry3f>&, const std::vector<Eigen::Isometry3f>&, const std::string& title) {
  std::cout << "create_plot(\"" << title << "\"): plotting skipped (headless drop-in; the window of the reference blocks on waitKey)"
            << std::endl;
}
float computeRotationError(const Eigen::Matrix3f& R_err) { return vohost::computeRotationError(to_host(R_err)); }
std::vector<std::pair<Data_Point, Data_Point>> add_new_world_points(
    std::vector<std::pair<Data_Point, World_Point>> img_world_matches, std::vector<std::pair<Data_Point, Data_Point>> img_matches) {
  // src/my_utilities.cpp:413-434: keep the image<->image matches whose second point is not already in the map
  std::vector<int32_t> matched(img_world_matches.size()), cand(img_matches.size());
  for (size_t i = 0; i < matched.size(); ++i) matched[i] = img_world_matches[i].first.id_meas;
  for (size_t i = 0; i < cand.size(); ++i) cand[i] = img_matches[i].second.id_meas;
  std::vector<uint8_t> keep(cand.size() ? cand.size() : 1);
  int64_t n_keep = 0;
  vohost::vo::check(vo_anti_join(vohost::vo::default_ctx(), matched.data(), (int64_t)matched.size(), cand.data(),
                                 (int64_t)cand.size(), keep.data(), &n_keep),
                    "vo_anti_join");
  std::vector<std::pair<Data_Point, Data_Point>> r;
  for (size_t i = 0; i < cand.size(); ++i)
    if (keep[i]) r.push_back(img_matches[i]);
  return r;
}
int check_world_points_sanity(const std::vector<World_Point>& world_points) {
  return vohost::check_world_points_sanity(to_host(world_points));
}
Eigen::Affine3f alignTrajectories(const std::vector<Eigen::Isometry3f>& poses, const std::vector<Eigen::Isometry3f>& gt_poses) {
  // Eigen::umeyama(P, Q, with_scaling) (src/my_utilities.cpp:459-478).  The one thing a caller reads from the result
  // is the scale, linear().col(0).norm() (exec/icp_test.cpp:164): the returned transform carries that scale on a
  // rotation-free linear part.
  std::vector<vohost::vo::Iso3f> hp, hg;
  for (const auto& p : poses) hp.push_back(to_host(p));
  for (const auto& p : gt_poses) hg.push_back(to_host(p));
  const float s = vohost::alignTrajectoriesScale(hp, hg);
  Eigen::Affine3f T;
  T.linear() = Eigen::Matrix3f::Identity() * s;
  return T;
}

// ---------------------------------------------------------------------------------------------- pr::Camera
int pr::Camera::projectPoints(Vector2fVector& image_points, const Vector3fVector& world_points, bool keep_indices) {
  vohost::pr::Vector2fVector out;
  const int n = to_host(*this).projectPoints(out, to_host(world_points), keep_indices);
  image_points.resize(out.size());
  if (!out.empty()) std::memcpy(image_points.data(), out.data(), out.size() * 8);
  return n;
}

// ---------------------------------------------------------------------------------------------- pr::PICPSolver
struct pr::PICPSolver::Impl {
  vohost::pr::PICPSolver s;
};
pr::PICPSolver::PICPSolver()
    : _kernel_thereshold(1000.f), _damping(1.f), _min_num_inliers(0), _chi_inliers(0), _chi_outliers(0), _num_inliers(0) {}
void pr::PICPSolver::init(const Camera& camera, const Vector3fVector& world_points, const Vector2fVector& image_points) {
  _impl = std::make_shared<Impl>();
  _camera = camera;
  _impl->s.init(to_host(camera), to_host(world_points), to_host(image_points));
}
const pr::Camera& pr::PICPSolver::camera() const {
  if (_impl) _camera.setWorldInCameraPose(from_host(_impl->s.camera().worldInCameraPose()));
  return _camera;
}
bool pr::PICPSolver::oneRound(const IntPairVector& correspondences, bool keep_outliers) {
  if (!_impl) throw std::runtime_error("PICPSolver::oneRound before init");
  _impl->s.setKernelThreshold(_kernel_thereshold);
  const bool ok = _impl->s.oneRound(correspondences, keep_outliers);
  _chi_inliers = _impl->s.chiInliers();
  _chi_outliers = _impl->s.chiOutliers();
  _num_inliers = _impl->s.numInliers();
  return ok;
}

// ---------------------------------------------------------------------------------------------- Cam
struct Cam::Impl {
  vohost::Cam cam;
};
Cam::Cam() : impl_(std::make_shared<Impl>()) {}
void Cam::computeEssentialAndRecoverPose(const std::vector<std::pair<Data_Point, Data_Point>>& matches, cv::Mat& mask) {
  std::vector<uint8_t> m;
  impl_->cam.computeEssentialAndRecoverPose(to_host(matches), m);
  mask.create((int)m.size(), 1, CV_8U);
  for (size_t i = 0; i < m.size(); ++i) mask.at<unsigned char>((int)i) = m[i];
  R_.create(3, 3, CV_64F);
  t_.create(3, 1, CV_64F);
  for (int i = 0; i < 9; ++i) R_.at<double>(i / 3, i % 3) = impl_->cam.getRotationMatrix()[i];
  for (int i = 0; i < 3; ++i) t_.at<double>(i) = impl_->cam.getTranslationVector()[i];
}
void Cam::triangulatePoints(const Eigen::Isometry3f& T1, const Eigen::Isometry3f& T2,
                            std::vector<std::pair<Data_Point, Data_Point>>& matches, std::vector<World_Point>& points3D) {
  auto hm = to_host(matches);
  std::vector<vohost::World_Point> out;
  impl_->cam.triangulatePoints(to_host(T1), to_host(T2), hm, out);
  for (const auto& p : out) points3D.push_back(from_host(p));
}
Eigen::Matrix3f Cam::getEigenCamera() { return from_host(impl_->cam.getEigenCamera()); }
int Cam::getHeight() const { return impl_->cam.getHeight(); }
int Cam::getWidth() const { return impl_->cam.getWidth(); }
void Cam::initOneRound(std::vector<World_Point> world_points, std::vector<Data_Point> img_points) {
  impl_->cam.initOneRound(to_host(world_points), to_host(img_points));
}
void Cam::oneRound(pr::IntPairVector correspondences) { impl_->cam.oneRound(correspondences); }
Eigen::Isometry3f Cam::getPose() { return from_host(impl_->cam.getPose()); }
void Cam::setPose(Eigen::Isometry3f pose) { impl_->cam.setPose(to_host(pose)); }
Eigen::Isometry3f Cam::cameraToImage() { return from_host(impl_->cam.cameraToImage()); }
