// cam.cpp — class Cam of the host mirror (reference: src/cam.cpp).
#include "cam.h"

#include <cstdlib>
#include <cstring>
#include <iostream>

Cam::Cam() : z_near_(0.f), z_far_(5.f), width_(640), height_(480) {
  K_(0, 0) = 180.f; K_(0, 2) = 320.f;
  K_(1, 1) = 180.f; K_(1, 2) = 240.f;
  K_(2, 2) = 1.f;
  vo::Mat3f m;  // camera axes -> image axes (src/cam.cpp:18-27), zero translation
  m(0, 2) = 1.f;
  m(1, 0) = -1.f;
  m(2, 1) = -1.f;
  camera_to_image_.setLinear(m);
  std::memset(R_, 0, sizeof(R_));
  std::memset(t_, 0, sizeof(t_));
  picp_cam_ = pr::Camera(height_, width_, K_, vo::Iso3f::Identity());
  picp_solver_ = pr::PICPSolver();
}

void Cam::computeEssentialAndRecoverPose(const std::vector<std::pair<Data_Point, Data_Point>>& matches,
                                         std::vector<uint8_t>& mask) {
  std::vector<vo::Point2f> p1, p2;
  extract_coordinates_from_matches(matches, p1, p2);
  mask.assign(matches.size(), 0);
  double E[9];
  int good = 0;
  const int st = vo_essential_recover(vo::default_ctx(), K_.data(), p1.empty() ? nullptr : &p1[0].x,
                                      p2.empty() ? nullptr : &p2[0].x, (int64_t)matches.size(), E, R_, t_,
                                      mask.empty() ? nullptr : mask.data(), &good);
  if (st != VO_OK) {  // the reference exits when findEssentialMat returns an empty matrix (src/cam.cpp:56-59)
    std::cerr << "Essential matrix computation failed!" << std::endl;
    exit(EXIT_FAILURE);
  }
  vo::Iso3f T;  // cv2eigen: CV_64F -> float
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) T(r, c) = (float)R_[3 * r + c];
    T(r, 3) = (float)t_[r];
  }
  picp_cam_.setWorldInCameraPose(T.inverse());
}

void Cam::triangulatePoints(const vo::Iso3f& T1, const vo::Iso3f& T2,
                            std::vector<std::pair<Data_Point, Data_Point>>& matches, std::vector<World_Point>& points3D) {
  std::vector<vo::Point2f> p1, p2;
  extract_coordinates_from_matches(matches, p1, p2);
  if (p1.empty() || p2.empty()) {
    std::cout << "Skipping triangulation: not enough points." << std::endl;
    return;
  }
  std::vector<float> xyz(3 * p1.size());
  vo::check(vo_triangulate(vo::default_ctx(), K_.data(), T1.data(), T2.data(), &p1[0].x, &p2[0].x, (int64_t)p1.size(),
                           xyz.data()),
            "vo_triangulate");
  std::cout << "Number of triangulated world points before checking duplicates: " << p1.size() << std::endl;
  for (size_t i = 0; i < p1.size(); ++i) {
    const Data_Point& src = matches[i].first;
    points3D.emplace_back(vo::Point3f(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]), src.descriptor, src.id_meas, src.id_real);
  }
}

void Cam::initOneRound(const std::vector<World_Point>& world_points, const std::vector<Data_Point>& img_points) {
  world_points_picp_ = extract_V3fV(world_points);
  image_points_picp_ = extract_V2fV(img_points);
  picp_solver_.init(picp_cam_, world_points_picp_, image_points_picp_);
  picp_solver_.setKernelThreshold(1000.0f);
  if (!picp_cam_.worldInCameraPose().isApprox(picp_solver_.camera().worldInCameraPose())) {
    std::cerr << "Cam::initOneRound failed: camera poses are different!" << std::endl;
    exit(EXIT_FAILURE);
  }
}

void Cam::oneRound(const pr::IntPairVector& correspondences) {
  for (int i = 0; i < 5; ++i) picp_solver_.oneRound(correspondences, false);  // exactly five rounds (src/cam.cpp:214-216)
  picp_cam_ = picp_solver_.camera();
  std::cout << "PICP inliers: " << picp_solver_.numInliers() << "/" << correspondences.size() << std::endl;
}
