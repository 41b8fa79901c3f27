// defs.h — drop-in for the reference's src/defs.h: the typedefs its interface is written in (src/defs.h:21-42,
// 205-211) on the Eigen type shim, plus v2tEuler (src/defs.h:100-136).  The image / OpenCV typedefs of the original
// (src/defs.h:148-203) are unused by any caller and not reproduced.
#pragma once
#include <Eigen/Core>
#include <Eigen/Geometry>
#include <Eigen/StdVector>
#include <cmath>
#include <utility>
#include <vector>

namespace pr {

typedef std::vector<Eigen::Vector3f, Eigen::aligned_allocator<Eigen::Vector3f> > Vector3fVector;
typedef std::vector<Eigen::Vector2f, Eigen::aligned_allocator<Eigen::Vector2f> > Vector2fVector;
typedef Eigen::Matrix<float, 2, 3> Matrix2_3f;
typedef Eigen::Matrix<float, 2, 6> Matrix2_6f;
typedef Eigen::Matrix<float, 3, 6> Matrix3_6f;
typedef Eigen::Matrix<float, 6, 6> Matrix6f;
typedef Eigen::Matrix<float, 6, 1> Vector6f;
typedef std::pair<int, int> IntPair;
typedef std::vector<IntPair> IntPairVector;

static_assert(sizeof(Eigen::Vector3f) == 12 && sizeof(Eigen::Vector2f) == 8 && sizeof(IntPair) == 8,
              "the point / pair vectors go through the C-ABI as packed float / int32 arrays");

inline Eigen::Matrix3f Rx(float a) {
  Eigen::Matrix3f R;
  R << 1, 0, 0, 0, std::cos(a), -std::sin(a), 0, std::sin(a), std::cos(a);
  return R;
}
inline Eigen::Matrix3f Ry(float a) {
  Eigen::Matrix3f R;
  R << std::cos(a), 0, std::sin(a), 0, 1, 0, -std::sin(a), 0, std::cos(a);
  return R;
}
inline Eigen::Matrix3f Rz(float a) {
  Eigen::Matrix3f R;
  R << std::cos(a), -std::sin(a), 0, std::sin(a), std::cos(a), 0, 0, 0, 1;
  return R;
}
inline Eigen::Isometry3f v2tEuler(const Vector6f& v) {
  Eigen::Isometry3f T;
  T.linear() = Rx(v[3]) * Ry(v[4]) * Rz(v[5]);
  T.translation() = v.head<3>();
  return T;
}
inline Eigen::Matrix3f skew(const Eigen::Vector3f& v) {
  Eigen::Matrix3f S;
  S << 0, -v[2], v[1], v[2], 0, -v[0], -v[1], v[0], 0;
  return S;
}

}  // namespace pr
