// vo_native.cpp — the reference's older driver (exec/vo.cpp:19-214) against the host mirror: Cam::initOneRound
// / Cam::oneRound (kernel threshold 1000, exactly five rounds, src/cam.cpp:178-224), 120 frames. It keeps the
// driver's own pose convention (getPose() = world-in-camera handed straight to triangulatePoints); the OpenCV
// plot at the end is replaced by a text dump of the poses.
//   vo_native [meas_path_prefix=./data/meas-] [out_file=output/vo_poses.txt] [n_meas=120]
#include <fstream>
#include <iostream>
#include <string>

#include "cam.h"
#include "my_utilities.h"

int main(int argc, char** argv) {
  const std::string meas_prefix = argc > 1 ? argv[1] : "./data/meas-";
  const std::string out_file = argc > 2 ? argv[2] : "output/vo_poses.txt";
  const int n_meas = argc > 3 ? std::atoi(argv[3]) : 120;
  std::vector<Measurement> measurements = load_and_initialize_data(meas_prefix, n_meas);

  Cam cam;
  std::vector<World_Point> world_points;
  std::vector<vo::Iso3f> estimated_poses, gt_poses;
  estimated_poses.push_back(vo::Iso3f::Identity());
  gt_poses.push_back(vo::Iso3f::Identity());

  for (int i = 0; i < n_meas - 1; ++i) {
    std::cout << "\nIteration: " << i + 1 << std::endl;
    const std::vector<Data_Point>& points1 = measurements[i].data_points;
    const std::vector<Data_Point>& points2 = measurements[i + 1].data_points;
    gt_poses.push_back(augment_pose(measurements[i].gt_pose));
    pr::IntPairVector img_correspondences, img_world_correspondences;
    if (i == 0) {
      std::vector<std::pair<Data_Point, Data_Point>> matches;
      match_points(points1, points2, matches, img_correspondences);
      const int inl = (int)matches.size();
      const int total = (int)std::max(points1.size(), points2.size());
      std::cout << "Number of inliers: " << inl << ", Number of outliers: " << total - inl << std::endl;
      std::vector<uint8_t> mask;
      cam.computeEssentialAndRecoverPose(matches, mask);
      cam.initOneRound(world_points, points2);
      const vo::Iso3f estimated_pose = cam.getPose();
      cam.triangulatePoints(vo::Iso3f::Identity(), estimated_pose, matches, world_points);
      std::cout << "Number of world points: " << world_points.size() << std::endl;
      estimated_poses.push_back(estimated_pose);
    } else {
      std::vector<std::pair<Data_Point, World_Point>> img_world_matches;
      match_points(points2, world_points, img_world_matches, img_world_correspondences);
      const vo::Iso3f previous_pose = estimated_poses.back();
      cam.setPose(previous_pose);
      cam.initOneRound(world_points, points2);
      cam.oneRound(img_world_correspondences);
      const vo::Iso3f estimated_pose = cam.getPose();
      estimated_poses.push_back(estimated_pose);
      std::vector<std::pair<Data_Point, Data_Point>> img_matches;
      match_points(points1, points2, img_matches, img_correspondences);
      std::vector<std::pair<Data_Point, Data_Point>> fresh = add_new_world_points(img_world_matches, img_matches);
      cam.triangulatePoints(previous_pose, estimated_pose, fresh, world_points);
      check_world_points_sanity(world_points);
      std::cout << "Number of world points: " << world_points.size() << std::endl;
    }
  }
  std::ofstream out(out_file);
  if (!out.is_open()) {
    std::cerr << "Error: Unable to open output file." << std::endl;
    return EXIT_FAILURE;
  }
  out.precision(9);
  for (const auto& T : estimated_poses) {
    for (int k = 0; k < 12; ++k) out << T.m[k] << (k == 11 ? "\n" : " ");
  }
  std::cout << "world points " << world_points.size() << std::endl;
  return 0;
}
