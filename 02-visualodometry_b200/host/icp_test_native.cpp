// icp_test_native.cpp — the reference's final pipeline (exec/icp_test.cpp:17-215) against the host
// mirror: same statements, same constants, same output files; no OpenCV window at the end.
//   icp_test_native [meas_path_prefix=./data/meas-] [output_dir=output] [n_meas=121]
#include <cmath>
#include <fstream>
#include <iostream>
#include <limits>
#include <string>

#include "cam.h"
#include "my_utilities.h"

static void print_pose(const char* title, const vo::Iso3f& T) {
  std::cout << title << "\n";
  for (int r = 0; r < 3; ++r) std::cout << T(r, 0) << " " << T(r, 1) << " " << T(r, 2) << " " << T(r, 3) << "\n";
  std::cout << "0 0 0 1" << std::endl;
}

int main(int argc, char** argv) {
  const std::string meas_prefix = argc > 1 ? argv[1] : "./data/meas-";
  const std::string out_dir = argc > 2 ? argv[2] : "output";
  const int n_meas = argc > 3 ? std::atoi(argv[3]) : 121;

  std::vector<Measurement> measurements = load_and_initialize_data(meas_prefix, n_meas);

  Cam cam;
  pr::Camera picp_cam(480, 640, cam.getEigenCamera(), vo::Iso3f::Identity());
  pr::PICPSolver picp_solver;

  WorldPointVector world_points;
  std::vector<vo::Iso3f> poses, gt_poses;
  poses.push_back(vo::Iso3f::Identity());

  std::cout << "\nIteration: 0" << std::endl;
  const std::vector<Data_Point>& first = measurements[0].data_points;
  const std::vector<Data_Point>& second = measurements[1].data_points;
  pr::IntPairVector img_correspondences;
  std::vector<std::pair<Data_Point, Data_Point>> initial_matches;
  match_points(first, second, initial_matches, img_correspondences);

  std::vector<uint8_t> mask;
  cam.computeEssentialAndRecoverPose(initial_matches, mask);
  const vo::Iso3f initial_estimated_pose = cam.getPose();
  print_pose("Pose: ", initial_estimated_pose);
  cam.triangulatePoints(vo::Iso3f::Identity(), initial_estimated_pose, initial_matches, world_points);

  for (int i = 0; i < n_meas - 1; ++i) {
    gt_poses.push_back(augment_pose(measurements[i].gt_pose));
    std::cout << "\nIteration: " << i << std::endl;
    const std::vector<Data_Point>& curr_points = measurements[i].data_points;
    const std::vector<Data_Point>& next_points = measurements[i + 1].data_points;

    // image points of the next frame against the map
    pr::IntPairVector img_world_correspondences;
    std::vector<std::pair<Data_Point, World_Point>> img_world_matches;
    match_points(next_points, world_points, img_world_matches, img_world_correspondences);

    const vo::Iso3f previous_pose = poses.back();
    picp_cam.setWorldInCameraPose(previous_pose.inverse());
    picp_solver.init(picp_cam, extract_V3fV(world_points), extract_V2fV(next_points));
    picp_solver.setKernelThreshold(3000.0f);

    const int maxIterations = 50;
    float prevError = std::numeric_limits<float>::max();
    const float convergenceThreshold = 0.00001f;
    bool converged = false;
    for (int j = 0; j < maxIterations; ++j) {
      if (!picp_solver.oneRound(img_world_correspondences, false)) {
        std::cerr << "Solver iteration " << j << " failed." << std::endl;
        break;
      }
      const float currentError = picp_solver.chiInliers();
      const float rel = (prevError > 1e-10) ? std::abs(prevError - currentError) / prevError : 0.0f;
      if (rel < convergenceThreshold) {
        converged = true;
        std::cout << "Convergence reached at iteration " << j << std::endl;
        break;
      }
      prevError = currentError;
    }
    if (!converged) std::cerr << "Convergence not reached." << std::endl;
    std::cout << "PICP inliers: " << picp_solver.numInliers() << "/" << img_world_correspondences.size() << std::endl;

    const vo::Iso3f estimated_pose = picp_solver.camera().worldInCameraPose().inverse();
    print_pose("Estimated pose", estimated_pose);
    poses.push_back(estimated_pose);

    // new landmarks: matches between the two frames that are not in the map yet
    pr::IntPairVector img_correspondences_local;
    std::vector<std::pair<Data_Point, Data_Point>> img_matches;
    match_points(curr_points, next_points, img_matches, img_correspondences_local);
    std::vector<std::pair<Data_Point, Data_Point>> fresh = add_new_world_points(img_world_matches, img_matches);
    cam.triangulatePoints(previous_pose, estimated_pose, fresh, world_points);
    std::cout << "Number of world points: " << world_points.size() << std::endl;
  }
  gt_poses.push_back(augment_pose(measurements[n_meas - 1].gt_pose));

  for (auto& p : poses) p = cam.cameraToImage() * p;
  const float scale = alignTrajectoriesScale(poses, gt_poses);

  std::ofstream f_traj(out_dir + "/estimated_trajectory.txt"), f_scaled(out_dir + "/estimated_trajectory_scaled.txt");
  std::ofstream f_err(out_dir + "/errors.txt"), f_world(out_dir + "/estimated_world_points.txt");
  if (!f_traj.is_open() || !f_scaled.is_open() || !f_err.is_open() || !f_world.is_open()) {
    std::cerr << "Error: Unable to open output file." << std::endl;
    return EXIT_FAILURE;
  }
  for (size_t j = 0; j < poses.size(); ++j) {
    const vo::Iso3f& gt = gt_poses[j];
    vo::Iso3f& pose = poses[j];
    const float angle_gt = std::atan2(gt(1, 0), gt(0, 0));
    float angle = std::atan2(pose(1, 0), pose(0, 0));
    const float pi = 3.1415926535897932384626433832795028841971693;
    angle += pi / 2.0;  // un-wrapped on purpose: the reference's errors.txt carries the 2*pi jumps
    f_traj << j << " " << pose.translation().x() << " " << pose.translation().y() << " " << angle << "\n";
    pose.setTranslation(pose.translation() * scale);
    f_scaled << j << " " << pose.translation().x() << " " << pose.translation().y() << " " << angle << "\n";
    const float e_t = (pose.translation() - gt.translation()).norm();
    f_err << j << " " << e_t << " " << std::abs(angle - angle_gt) << "\n";
  }
  for (int id = 0; id < 1000; ++id)
    for (const auto& wp : world_points)
      if (wp.id_real == id) {
        const vo::Vec3f p = (cam.cameraToImage() * vo::Vec3f(wp.coordinates.x, wp.coordinates.y, wp.coordinates.z)) * scale;
        f_world << id << " " << p.x() << " " << p.y() << " " << p.z() << "\n";
        break;
      }
  std::cout << "scale " << scale << ", world points " << world_points.size() << std::endl;
  return 0;
}
