// my_utilities.h — drop-in for the live part of the reference's src/my_utilities.h: the declarations its drivers use
// (same names, same parameter types), with match_points<> (src/my_utilities.h:70-120) forwarding to vo_match on the
// GPU.  The ~30 prototypes of the original that have no definition anywhere are not
// reproduced; create_plot (OpenCV highgui window, blocks on waitKey) only reports that plotting is skipped.
#pragma once
#include <Eigen/Cholesky>
#include <Eigen/Core>
#include <Eigen/Dense>
#include <algorithm>
#include <cmath>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <opencv2/core/eigen.hpp>
#include <opencv2/opencv.hpp>
#include <opencv2/viz.hpp>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

#include "data_point.h"
#include "defs.h"
#include "picp_solver.h"

using namespace pr;

struct Measurement {
  int seq;
  Eigen::Vector3f gt_pose;        // ground truth pose (x y theta)
  Eigen::Vector3f odometry_pose;  // odometry pose (x y theta)
  std::vector<Data_Point> data_points;
  Measurement() : seq(0), gt_pose(Eigen::Vector3f::Zero()), odometry_pose(Eigen::Vector3f::Zero()), data_points() {}
};

// src/my_utilities.h:44-47
inline const float DISTANCE_THRESHOLD = 0.2f;
inline const float FRAMES_DISTANCE_THRESHOLD = 0.1f;
inline const float RATIO_THRESHOLD = 0.8f;
inline const int PICP_RUNS = 10;

namespace vo_compat {
// vo_match on packed descriptor rows: (i, best_j) pairs in ascending i + the two counters of the printed line
void match_rows(const float* descA, long long n1, const float* descB, long long n2, int dim, const int* idA,
                const int* idB, std::vector<int>& pairs, long long stats[2]);
template <class Record>
inline void gather(const std::vector<Record>& v, std::vector<float>& desc, std::vector<int>& ids, int& dim) {
  dim = v.empty() ? 0 : v.front().descriptor.size();
  desc.resize(v.size() * (size_t)dim);
  ids.resize(v.size());
  for (size_t i = 0; i < v.size(); ++i) {
    for (int k = 0; k < dim; ++k) desc[i * dim + k] = v[i].descriptor[k];
    ids[i] = v[i].id_real;
  }
}
}  // namespace vo_compat

// Brute-force descriptor matching with ratio test (src/my_utilities.h:70-120): appends to `matches` and
// `correspondences` (never clears), ascending in the index of points1, prints the reference's summary line.
template <typename PointType1, typename PointType2>
void match_points(const std::vector<PointType1>& points1, const std::vector<PointType2>& points2,
                  std::vector<std::pair<PointType1, PointType2>>& matches, IntPairVector& correspondences) {
  long long stats[2] = {0, 0};
  if (!points1.empty() && !points2.empty()) {
    std::vector<float> dA, dB;
    std::vector<int> iA, iB, pairs;
    int dimA = 0, dimB = 0;
    vo_compat::gather(points1, dA, iA, dimA);
    vo_compat::gather(points2, dB, iB, dimB);
    vo_compat::match_rows(dA.data(), (long long)points1.size(), dB.data(), (long long)points2.size(), dimA, iA.data(),
                          iB.data(), pairs, stats);
    for (size_t k = 0; k + 1 < pairs.size(); k += 2) {
      matches.push_back(std::make_pair(points1[pairs[k]], points2[pairs[k + 1]]));
      correspondences.push_back(IntPair(pairs[k], pairs[k + 1]));
    }
  }
  std::cout << "Matches: Out of " << stats[0] << " possible matches, found " << matches.size() << ", of which "
            << stats[1] << " are correct" << std::endl;
}

void extract_coordinates_from_matches(std::vector<std::pair<Data_Point, Data_Point>> matches,
                                      std::vector<cv::Point2f>& matches1, std::vector<cv::Point2f>& matches2);
Vector2fVector extract_V2fV(const std::vector<Data_Point>& points);
Vector3fVector extract_V3fV(const std::vector<World_Point>& points);

std::vector<std::string> split(const std::string& str, const std::string& delimiter);
Measurement extract_measurement(const std::string& filename);
std::vector<Measurement> extract_measurements(const std::string& filename, int n_meas);
std::vector<Measurement> load_and_initialize_data(const std::string& path, int num_measurements);
std::vector<World_Point> load_world_points(const std::string& filename);

Eigen::Isometry3f oneRound(Eigen::Isometry3f last_pose_estimate, pr::Camera& pr_cam,
                           const pr::Vector3fVector& world_points, const pr::Vector2fVector& image_points,
                           const pr::IntPairVector& correspondences);
Eigen::Isometry3f augment_pose(const Eigen::Vector3f& pose);
float compute_scale(const std::vector<Eigen::Vector3f>& points_reconstructed,
                    const std::vector<Eigen::Vector3f>& points_ground_truth);
void create_plot(const std::vector<Eigen::Isometry3f>& gt_poses, const std::vector<Eigen::Isometry3f>& est_poses,
                 const std::string& title);
float computeRotationError(const Eigen::Matrix3f& R_err);
std::vector<std::pair<Data_Point, Data_Point>> add_new_world_points(
    std::vector<std::pair<Data_Point, World_Point>> img_world_matches,
    std::vector<std::pair<Data_Point, Data_Point>> img_matches);
int check_world_points_sanity(const std::vector<World_Point>& world_points);
Eigen::Affine3f alignTrajectories(const std::vector<Eigen::Isometry3f>& poses,
                                  const std::vector<Eigen::Isometry3f>& gt_poses);
