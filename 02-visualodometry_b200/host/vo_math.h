// vo_math.h — dependency-free stand-ins for the handful of Eigen / OpenCV value types that cross the
// reference's hot-path interface (src/defs.h:21-42,205-211; src/data_point.h). Layout-compatible with
// the reference's containers: Vec3f is 12 bytes, Vec2f 8 bytes, IntPair is std::pair<int,int>, so a
// std::vector of them goes through the C-ABI as a plain float / int32 array without conversion.
#pragma once
#include <array>
#include <cmath>
#include <cstdint>
#include <utility>
#include <vector>

#include "../../include/vo_b200.h"

namespace vo {

struct Vec2f {
  float v[2];
  Vec2f() : v{0.f, 0.f} {}
  Vec2f(float a, float b) : v{a, b} {}
  float& x() { return v[0]; }
  float& y() { return v[1]; }
  float x() const { return v[0]; }
  float y() const { return v[1]; }
  float& operator[](int i) { return v[i]; }
  float operator[](int i) const { return v[i]; }
};

struct Vec3f {
  float v[3];
  Vec3f() : v{0.f, 0.f, 0.f} {}
  Vec3f(float a, float b, float c) : v{a, b, c} {}
  float& x() { return v[0]; }
  float& y() { return v[1]; }
  float& z() { return v[2]; }
  float x() const { return v[0]; }
  float y() const { return v[1]; }
  float z() const { return v[2]; }
  float& operator[](int i) { return v[i]; }
  float operator[](int i) const { return v[i]; }
  Vec3f operator-(const Vec3f& o) const { return {v[0] - o.v[0], v[1] - o.v[1], v[2] - o.v[2]}; }
  Vec3f operator*(float s) const { return {v[0] * s, v[1] * s, v[2] * s}; }
  float norm() const { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
};

static_assert(sizeof(Vec2f) == 8 && sizeof(Vec3f) == 12, "point vectors must be tightly packed");
static_assert(sizeof(std::pair<int, int>) == 8, "IntPair must be two packed int32");

// row-major 3x3
struct Mat3f {
  float m[9];
  Mat3f() : m{0, 0, 0, 0, 0, 0, 0, 0, 0} {}
  static Mat3f Identity() {
    Mat3f r;
    r.m[0] = r.m[4] = r.m[8] = 1.f;
    return r;
  }
  float& operator()(int r, int c) { return m[3 * r + c]; }
  float operator()(int r, int c) const { return m[3 * r + c]; }
  const float* data() const { return m; }
};

// Rigid transform [R|t], row-major 3x4, with the semantics of Eigen::Isometry3f as the reference uses it
// (products and inverse are evaluated by the library in Eigen's float32 order, see vo_pose_mul/inverse).
struct Iso3f {
  float m[12];
  Iso3f() : m{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0} {}
  static Iso3f Identity() { return Iso3f(); }
  float& operator()(int r, int c) { return m[4 * r + c]; }
  float operator()(int r, int c) const { return m[4 * r + c]; }
  Mat3f linear() const {
    Mat3f R;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) R(r, c) = m[4 * r + c];
    return R;
  }
  Mat3f rotation() const { return linear(); }
  void setLinear(const Mat3f& R) {
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) m[4 * r + c] = R(r, c);
  }
  Vec3f translation() const { return {m[3], m[7], m[11]}; }
  void setTranslation(const Vec3f& t) {
    m[3] = t[0];
    m[7] = t[1];
    m[11] = t[2];
  }
  Iso3f inverse() const {
    Iso3f r;
    vo_pose_inverse(m, r.m);
    return r;
  }
  Iso3f operator*(const Iso3f& o) const {
    Iso3f r;
    vo_pose_mul(m, o.m, r.m);
    return r;
  }
  Vec3f operator*(const Vec3f& p) const {  // t + R p, x0 + (x1 + x2) per row
    Vec3f r;
    for (int i = 0; i < 3; ++i) {
      const float x0 = m[4 * i] * p[0], x1 = m[4 * i + 1] * p[1], x2 = m[4 * i + 2] * p[2];
      r[i] = m[4 * i + 3] + (x0 + (x1 + x2));
    }
    return r;
  }
  bool isApprox(const Iso3f& o, float prec = 1e-5f) const {
    double d = 0, n1 = 0, n2 = 0;
    for (int i = 0; i < 12; ++i) {
      d += double(m[i] - o.m[i]) * (m[i] - o.m[i]);
      n1 += double(m[i]) * m[i];
      n2 += double(o.m[i]) * o.m[i];
    }
    return d <= double(prec) * prec * std::min(n1 + 1, n2 + 1);
  }
  const float* data() const { return m; }
};

// cv::Point2f / cv::Point3f stand-ins (src/data_point.h)
struct Point2f {
  float x, y;
  Point2f() : x(0), y(0) {}
  Point2f(float a, float b) : x(a), y(b) {}
};
struct Point3f {
  float x, y, z;
  Point3f() : x(0), y(0), z(0) {}
  Point3f(float a, float b, float c) : x(a), y(b), z(c) {}
};

using Descriptor = std::vector<float>;  // Eigen::VectorXf in the reference

// process-wide context of the drop-in classes (the reference API has no context argument).
// Device index: $VO_B200_DEVICE (default 0). Throws std::runtime_error when no GPU is usable.
vo_ctx* default_ctx();
void check(int status, const char* what);

}  // namespace vo

namespace pr {
using Vector3fVector = std::vector<vo::Vec3f>;
using Vector2fVector = std::vector<vo::Vec2f>;
using IntPair = std::pair<int, int>;
using IntPairVector = std::vector<IntPair>;
}  // namespace pr
