// my_utilities.h — the live part of the reference's src/my_utilities.{h,cpp}: data loading, the
// match_points<> template, frame-loop glue and the evaluation tail, on the dependency-free types of
// vo_math.h. Dead declarations of the reference header and the OpenCV plotting
// (create_plot) are not reproduced.
#pragma once
#include <iostream>
#include <limits>
#include <string>
#include <utility>
#include <vector>

#include "data_point.h"
#include "picp_solver.h"

struct Measurement {
  int seq = 0;
  vo::Vec3f gt_pose;        // x y theta
  vo::Vec3f odometry_pose;  // x y theta
  std::vector<Data_Point> data_points;
};

// src/my_utilities.h:44-47
inline const float DISTANCE_THRESHOLD = 0.2f;
inline const float FRAMES_DISTANCE_THRESHOLD = 0.1f;
inline const float RATIO_THRESHOLD = 0.8f;
inline const int PICP_RUNS = 10;

// Brute-force descriptor matching with ratio test (src/my_utilities.h:70-120). Appends to `matches`
// and `correspondences` (never clears), ascending in the index of points1, and prints the reference's
// summary line. For image<->world matching pass the image points first.
template <typename PointType1, typename PointType2>
void match_points(const std::vector<PointType1>& points1, const std::vector<PointType2>& points2,
                  std::vector<std::pair<PointType1, PointType2>>& matches, pr::IntPairVector& correspondences) {
  long long possible = 0, correct = 0;
  const size_t before = matches.size();
  if (!points1.empty() && !points2.empty()) {
    const int dim = (int)points1[0].descriptor.size();
    std::vector<float> dA, dB;
    std::vector<int32_t> iA, iB;
    vo::gather_descriptors(points1, dA);  // vo_records.h: row-major float[N][D] + the id_real column
    vo::gather_descriptors(points2, dB);
    vo::gather_real_ids(points1, iA);
    vo::gather_real_ids(points2, iB);
    std::vector<int32_t> pairs(2 * points1.size());
    int64_t n = 0, stats[2] = {0, 0};
    vo::check(vo_match(vo::default_ctx(), dA.data(), (int64_t)points1.size(), dB.data(), (int64_t)points2.size(), dim,
                       DISTANCE_THRESHOLD, RATIO_THRESHOLD, iA.data(), iB.data(), 0, (int64_t)points1.size(),
                       pairs.data(), (int64_t)points1.size(), &n, stats),
              "vo_match");
    for (int64_t k = 0; k < n; ++k) {
      const int i = pairs[2 * k], j = pairs[2 * k + 1];
      matches.push_back(std::make_pair(points1[i], points2[j]));
      correspondences.push_back(pr::IntPair(i, j));
    }
    possible = stats[0];
    correct = stats[1];
  }
  (void)before;
  std::cout << "Matches: Out of " << possible << " possible matches, found " << matches.size() << ", of which "
            << correct << " are correct" << std::endl;
}

// ---- data loading (src/my_utilities.cpp:20-182)
std::vector<std::string> split(const std::string& str, const std::string& delimiter);
Measurement extract_measurement(const std::string& filename);
std::vector<Measurement> extract_measurements(const std::string& filename, int n_meas);
std::vector<Measurement> load_and_initialize_data(const std::string& path, int num_measurements);
std::vector<World_Point> load_world_points(const std::string& filename);

// ---- gathers (src/my_utilities.cpp:185-223)
void extract_coordinates_from_matches(const std::vector<std::pair<Data_Point, Data_Point>>& matches,
                                      std::vector<vo::Point2f>& matches1, std::vector<vo::Point2f>& matches2);
pr::Vector2fVector extract_V2fV(const std::vector<Data_Point>& points);
pr::Vector3fVector extract_V3fV(const std::vector<World_Point>& points);

// ---- frame-loop glue
std::vector<std::pair<Data_Point, Data_Point>> add_new_world_points(
    const std::vector<std::pair<Data_Point, World_Point>>& img_world_matches,
    const std::vector<std::pair<Data_Point, Data_Point>>& img_matches);  // src/my_utilities.cpp:413-434
vo::Iso3f oneRound(vo::Iso3f last_pose_estimate, pr::Camera& pr_cam, const pr::Vector3fVector& world_points,
                   const pr::Vector2fVector& image_points,
                   const pr::IntPairVector& correspondences);  // src/my_utilities.cpp:263-315
int check_world_points_sanity(const std::vector<World_Point>& world_points);

// ---- evaluation tail (src/my_utilities.cpp:226-260,400-410,459-478)
vo::Iso3f augment_pose(const vo::Vec3f& pose);
float compute_scale(const std::vector<vo::Vec3f>& points_reconstructed, const std::vector<vo::Vec3f>& points_ground_truth);
float computeRotationError(const vo::Mat3f& R_err);
// Eigen::umeyama(P, Q, with_scaling = true) reduced to what exec/icp_test.cpp:164 uses: the scale factor.
float alignTrajectoriesScale(const std::vector<vo::Iso3f>& poses, const std::vector<vo::Iso3f>& gt_poses);
