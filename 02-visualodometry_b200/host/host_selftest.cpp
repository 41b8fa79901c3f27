// host_selftest.cpp — CPU-only checks of the host mirror's own logic (text loaders, gathers, evaluation tail,
// Camera::projectPoint, Iso3f algebra). No GPU call is made; prints "name value..." lines that
// tests/test_native_host.py compares against numpy / the oracle.
//   host_selftest <meas_path_prefix> <n_meas> [world.dat]
#include <cstdio>
#include <iostream>
#include <string>

#include "cam.h"
#include "my_utilities.h"

int main(int argc, char** argv) {
  if (argc < 3) {
    std::cerr << "usage: host_selftest <meas_prefix> <n_meas> [world.dat]" << std::endl;
    return 2;
  }
  const std::vector<Measurement> meas = load_and_initialize_data(argv[1], std::atoi(argv[2]));
  size_t total = 0;
  double sum_uv = 0, sum_desc = 0;
  long long sum_ids = 0;
  for (const auto& m : meas) {
    total += m.data_points.size();
    for (const auto& p : m.data_points) {
      sum_uv += p.coordinates.x + p.coordinates.y;
      for (float d : p.descriptor) sum_desc += d;
      sum_ids += p.id_real + 3 * p.id_meas;
    }
  }
  std::printf("frames %zu points %zu sum_uv %.6f sum_desc %.6f sum_ids %lld\n", meas.size(), total, sum_uv, sum_desc, sum_ids);
  std::printf("seq_last %d gt_last %.9g %.9g %.9g odom_last %.9g %.9g %.9g\n", meas.back().seq, meas.back().gt_pose[0],
              meas.back().gt_pose[1], meas.back().gt_pose[2], meas.back().odometry_pose[0], meas.back().odometry_pose[1],
              meas.back().odometry_pose[2]);
  // split()
  const auto tok = split("point  12 7   3.5 -1e-3", " ");
  std::printf("split %zu %s %s %s\n", tok.size(), tok[0].c_str(), tok[1].c_str(), tok.back().c_str());
  // gathers
  const pr::Vector2fVector v2 = extract_V2fV(meas[0].data_points);
  std::printf("v2 %zu %.9g %.9g\n", v2.size(), v2[0].x(), v2.back().y());
  // augment_pose + Iso3f algebra
  const vo::Iso3f G = augment_pose(meas.back().gt_pose);
  const vo::Iso3f Gi = G.inverse(), GG = G * Gi;
  std::printf("augment %.9g %.9g %.9g %.9g inv %.9g %.9g id %.9g %.9g %.9g\n", G(0, 0), G(0, 1), G(0, 3), G(1, 3), Gi(0, 3),
              Gi(1, 3), GG(0, 0), GG(0, 1), GG(0, 3));
  // evaluation tail: umeyama scale between a scaled odometry track and the ground-truth track, compute_scale,
  // rotation error
  std::vector<vo::Iso3f> P, Q;
  std::vector<vo::Vec3f> a, b;
  for (const auto& m : meas) {
    vo::Iso3f o = augment_pose(m.odometry_pose), g = augment_pose(m.gt_pose);
    o.setTranslation(o.translation() * 0.37f);
    P.push_back(o);
    Q.push_back(g);
    a.push_back(o.translation());
    b.push_back(g.translation());
  }
  std::printf("umeyama_scale %.7g compute_scale %.7g rot_err %.7g\n", alignTrajectoriesScale(P, Q), compute_scale(a, b),
              computeRotationError((Q.back().inverse() * P.back()).linear()));
  // Camera::projectPoint (host inline, reference float order)
  Cam cam_defaults;  // constructing a Cam makes no GPU call
  pr::Camera cam(480, 640, cam_defaults.getEigenCamera(), G);
  int inside = 0;
  double acc = 0;
  for (int i = 0; i < 1000; ++i) {
    vo::Vec2f uv;
    const vo::Vec3f p(0.013f * i - 6.f, 0.007f * i - 3.f, 0.02f * i - 4.f);
    if (cam.projectPoint(uv, p)) {
      inside++;
      acc += uv.x() + 2.0 * uv.y();
    }
  }
  std::printf("project inside %d acc %.9g\n", inside, acc);
  if (argc > 3) {
    const std::vector<World_Point> w = load_world_points(argv[3]);
    double s = 0;
    for (const auto& p : w) s += p.coordinates.x + p.coordinates.y + p.coordinates.z + p.descriptor[9] + p.id_real;
    std::printf("world %zu sum %.6f dup %d\n", w.size(), s, check_world_points_sanity(w));
  }
  return 0;
}
