// my_utilities.cpp — host mirror of the live functions of the reference's src/my_utilities.cpp
// (text loaders, gathers, frame-loop glue, evaluation tail). Host I/O and 121-element post-processing
// stay on the CPU; the anti-join and the PICP rounds go through the C-ABI.
#include "my_utilities.h"

#include <algorithm>
#include <cmath>
#include <fstream>
#include <iomanip>
#include <sstream>

std::vector<std::string> split(const std::string& str, const std::string& delimiter) {
  std::vector<std::string> out;
  size_t pos = 0;
  while (true) {
    const size_t hit = str.find(delimiter, pos);
    const std::string tok = str.substr(pos, hit == std::string::npos ? std::string::npos : hit - pos);
    if (!tok.empty()) out.push_back(tok);
    if (hit == std::string::npos) break;
    pos = hit + delimiter.size();
  }
  return out;
}

// meas-NNNNN.dat: "seq: n", "gt_pose: x y th", "odom_pose: x y th", "point id_meas id_real u v d0..d9"
Measurement extract_measurement(const std::string& filename) {
  std::ifstream in(filename);
  if (!in.is_open()) {
    std::cerr << "Error opening file " << filename << std::endl;
    exit(EXIT_FAILURE);
  }
  Measurement m;
  std::string line;
  while (std::getline(in, line)) {
    const std::vector<std::string> tok = split(line, " ");
    if (tok.empty()) continue;
    const std::string& key = tok[0];
    if (key == "seq:" && tok.size() >= 2) {
      m.seq = std::stoi(tok[1]);
    } else if ((key == "gt_pose:" || key == "odom_pose:") && tok.size() >= 4) {
      vo::Vec3f& dst = key == "gt_pose:" ? m.gt_pose : m.odometry_pose;
      for (int k = 0; k < 3; ++k) dst[k] = std::stof(tok[1 + k]);
    } else if (key == "point" && tok.size() >= 15) {
      vo::Descriptor d(10);
      for (int k = 0; k < 10; ++k) d[k] = std::stof(tok[5 + k]);
      m.data_points.emplace_back(std::stoi(tok[1]), std::stoi(tok[2]), vo::Point2f(std::stof(tok[3]), std::stof(tok[4])), d);
    } else {
      std::cerr << "Invalid line in file " << filename << ": " << line << std::endl;
    }
  }
  return m;
}

std::vector<Measurement> extract_measurements(const std::string& filename, int n_meas) {
  std::cout << "Extracting measurements from files..." << std::endl;
  std::vector<Measurement> all;
  all.reserve(n_meas);
  for (int i = 0; i < n_meas; ++i) {
    std::ostringstream name;
    name << filename << std::setfill('0') << std::setw(5) << i << ".dat";
    all.push_back(extract_measurement(name.str()));
  }
  std::cout << "Measurements extracted" << std::endl;
  return all;
}

std::vector<Measurement> load_and_initialize_data(const std::string& path, int num_measurements) {
  return extract_measurements(path, num_measurements);
}

// world.dat: "id x y z d0..d9"
std::vector<World_Point> load_world_points(const std::string& filename) {
  std::ifstream in(filename);
  if (!in.is_open()) {
    std::cerr << "Error opening world points file " << filename << std::endl;
    exit(EXIT_FAILURE);
  }
  std::vector<World_Point> pts;
  std::string line;
  while (std::getline(in, line)) {
    std::istringstream ss(line);
    int id;
    float x, y, z;
    if (!(ss >> id >> x >> y >> z)) continue;
    vo::Descriptor d(10);
    bool ok = true;
    for (int k = 0; k < 10 && ok; ++k) ok = bool(ss >> d[k]);
    if (ok) pts.emplace_back(vo::Point3f(x, y, z), d, id);
  }
  return pts;
}

void extract_coordinates_from_matches(const std::vector<std::pair<Data_Point, Data_Point>>& matches,
                                      std::vector<vo::Point2f>& matches1, std::vector<vo::Point2f>& matches2) {
  matches1.reserve(matches1.size() + matches.size());
  matches2.reserve(matches2.size() + matches.size());
  for (const auto& m : matches) {
    matches1.push_back(m.first.coordinates);
    matches2.push_back(m.second.coordinates);
  }
}

pr::Vector2fVector extract_V2fV(const std::vector<Data_Point>& points) {
  pr::Vector2fVector out;
  out.reserve(points.size());
  for (const auto& p : points) out.emplace_back(p.coordinates.x, p.coordinates.y);
  return out;
}

pr::Vector3fVector extract_V3fV(const std::vector<World_Point>& points) {
  pr::Vector3fVector out;
  out.reserve(points.size());
  for (const auto& p : points) out.emplace_back(p.coordinates.x, p.coordinates.y, p.coordinates.z);
  return out;
}

// Keep the image<->image matches whose second point is not already matched to a world point.
std::vector<std::pair<Data_Point, Data_Point>> add_new_world_points(
    const std::vector<std::pair<Data_Point, World_Point>>& img_world_matches,
    const std::vector<std::pair<Data_Point, Data_Point>>& img_matches) {
  std::vector<int32_t> matched(img_world_matches.size()), cand(img_matches.size());
  for (size_t i = 0; i < matched.size(); ++i) matched[i] = img_world_matches[i].first.id_meas;
  for (size_t j = 0; j < cand.size(); ++j) cand[j] = img_matches[j].second.id_meas;
  std::vector<uint8_t> keep(cand.size());
  int64_t n_keep = 0;
  vo::check(vo_anti_join(vo::default_ctx(), matched.data(), (int64_t)matched.size(), cand.data(), (int64_t)cand.size(),
                         keep.data(), &n_keep),
            "vo_anti_join");
  std::vector<std::pair<Data_Point, Data_Point>> out;
  out.reserve((size_t)n_keep);
  for (size_t j = 0; j < cand.size(); ++j)
    if (keep[j]) out.push_back(img_matches[j]);
  std::cout << "Points to be triangulated: " << out.size() << std::endl;
  return out;
}

// Third driver of the solver (src/my_utilities.cpp:263-315): threshold 100, outliers kept with the
// saturating kernel, at most 50 rounds, stop below 5 % relative improvement.
vo::Iso3f oneRound(vo::Iso3f last_pose_estimate, pr::Camera& pr_cam, const pr::Vector3fVector& world_points,
                   const pr::Vector2fVector& image_points, const pr::IntPairVector& correspondences) {
  if (correspondences.size() < 10) {
    std::cerr << "Warning: Not enough correspondences for pose estimation (" << correspondences.size()
              << " < 10), using previous pose" << std::endl;
    return last_pose_estimate;
  }
  pr::PICPSolver solver;
  pr_cam.setWorldInCameraPose(last_pose_estimate);
  solver.init(pr_cam, world_points, image_points);
  solver.setKernelThreshold(100.0f);
  std::cout << "Kernel threshold set to 100.0f" << std::endl;
  double prev = std::numeric_limits<double>::max();
  for (int i = 0; i < 50; ++i) {
    if (!solver.oneRound(correspondences, true)) {
      std::cerr << "Solver iteration " << i << " failed." << std::endl;
      break;
    }
    const double cur = solver.chiInliers();
    const double rel = (prev > 1e-10) ? std::abs(prev - cur) / prev : 0.0;
    std::cout << "Iteration " << i << ", current error: " << cur << ", relative improvement: " << rel
              << ", inliers: " << solver.numInliers() << std::endl;
    if (rel < 0.05) {
      std::cout << "Convergence reached at iteration " << i << std::endl;
      break;
    }
    prev = cur;
  }
  std::cout << "Final pose computed. Inliers: " << solver.numInliers() << " out of " << correspondences.size() << std::endl;
  return solver.camera().worldInCameraPose();
}

int check_world_points_sanity(const std::vector<World_Point>& world_points) {
  std::vector<int> seen(1000, 0);
  for (const auto& p : world_points)
    if (p.id_real >= 0 && p.id_real < 1000) seen[p.id_real]++;
  const int dup = (int)std::count_if(seen.begin(), seen.end(), [](int c) { return c > 1; });
  std::cout << "Number of duplicate world points: " << dup << std::endl;
  return dup;
}

vo::Iso3f augment_pose(const vo::Vec3f& pose) {
  vo::Iso3f T;
  const float th = pose[2];
  vo::Mat3f R;
  R(0, 0) = std::cos(th); R(0, 1) = -std::sin(th);
  R(1, 0) = std::sin(th); R(1, 1) = std::cos(th);
  R(2, 2) = 1.f;
  T.setLinear(R);
  T.setTranslation(vo::Vec3f(pose[0], pose[1], 0.f));
  return T;
}

float compute_scale(const std::vector<vo::Vec3f>& rec, const std::vector<vo::Vec3f>& gt) {
  float total = 0.f;
  int n = 0;
  for (size_t i = 0; i < rec.size() && i < gt.size(); ++i) {
    const float a = rec[i].norm(), b = gt[i].norm();
    if (a > 0 && b > 0) {
      total += b / a;
      n++;
    }
  }
  return n > 0 ? total / n : 1.0f;
}

float computeRotationError(const vo::Mat3f& R) {
  float c = (R(0, 0) + R(1, 1) + R(2, 2) - 1.0f) / 2.0f;
  c = std::max(-1.0f, std::min(1.0f, c));
  return std::acos(c);
}

namespace {
// singular values of a 3x3 matrix and sign(det U * det V) by one-sided Jacobi (double)
void svd3_values(const double S[9], double w[3], double& det_sign) {
  double A[3][3], V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) A[i][j] = S[3 * i + j];
  for (int sweep = 0; sweep < 60; ++sweep) {
    bool changed = false;
    for (int i = 0; i < 2; ++i)
      for (int j = i + 1; j < 3; ++j) {
        double a = 0, b = 0, p = 0;
        for (int k = 0; k < 3; ++k) {
          a += A[k][i] * A[k][i];
          b += A[k][j] * A[k][j];
          p += A[k][i] * A[k][j];
        }
        if (std::fabs(p) <= 1e-15 * std::sqrt(a * b) || p == 0) continue;
        changed = true;
        const double z = (b - a) / (2 * p), t = (z >= 0 ? 1.0 : -1.0) / (std::fabs(z) + std::sqrt(1 + z * z));
        const double c = 1 / std::sqrt(1 + t * t), s = c * t;
        for (int k = 0; k < 3; ++k) {
          const double x = A[k][i], y = A[k][j];
          A[k][i] = c * x - s * y;
          A[k][j] = s * x + c * y;
          const double vx = V[k][i], vy = V[k][j];
          V[k][i] = c * vx - s * vy;
          V[k][j] = s * vx + c * vy;
        }
      }
    if (!changed) break;
  }
  for (int j = 0; j < 3; ++j) w[j] = std::sqrt(A[0][j] * A[0][j] + A[1][j] * A[1][j] + A[2][j] * A[2][j]);
  // det(S) = det(U) det(V) prod(w): its sign is the sign of det(U) det(V) for a full-rank S
  const double det = S[0] * (S[4] * S[8] - S[5] * S[7]) - S[1] * (S[3] * S[8] - S[5] * S[6]) + S[2] * (S[3] * S[7] - S[4] * S[6]);
  det_sign = det < 0 ? -1.0 : 1.0;
}
}  // namespace

// umeyama with scaling (src/my_utilities.cpp:459-478); only ||linear.col(0)|| = the scale c is consumed by
// the caller: c = sum(d_i * S_i) / var(src), S = diag(1,1,sign)
float alignTrajectoriesScale(const std::vector<vo::Iso3f>& poses, const std::vector<vo::Iso3f>& gt_poses) {
  const size_t n = poses.size();
  double mp[3] = {0, 0, 0}, mq[3] = {0, 0, 0};
  for (size_t i = 0; i < n; ++i)
    for (int k = 0; k < 3; ++k) {
      mp[k] += poses[i].translation()[k];
      mq[k] += gt_poses[i].translation()[k];
    }
  for (int k = 0; k < 3; ++k) {
    mp[k] /= (double)n;
    mq[k] /= (double)n;
  }
  double var = 0, sigma[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (size_t i = 0; i < n; ++i) {
    double p[3], q[3];
    for (int k = 0; k < 3; ++k) {
      p[k] = poses[i].translation()[k] - mp[k];
      q[k] = gt_poses[i].translation()[k] - mq[k];
      var += p[k] * p[k];
    }
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) sigma[3 * r + c] += q[r] * p[c];
  }
  var /= (double)n;
  for (double& s : sigma) s /= (double)n;
  double w[3], sign;
  svd3_values(sigma, w, sign);
  std::sort(w, w + 3, [](double a, double b) { return a > b; });
  // a rank-2 covariance (planar ground truth) has w[2] = 0: the reflection sign then multiplies zero
  return (float)((w[0] + w[1] + sign * w[2]) / var);
}
