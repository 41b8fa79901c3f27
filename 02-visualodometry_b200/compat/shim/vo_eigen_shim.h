// vo_eigen_shim.h — the compile-time surface of Eigen that the reference's drivers (exec/icp_test.cpp, exec/vo.cpp)
// and the replacement headers in compat/src/ touch, and nothing more (last row).  Eigen itself is an
// un-vendored dependency of the reference and is not in this image; this is a TYPE shim so that the unchanged mains
// compile and link against libvo_b200.so - every number that matters is computed by the library.  Where a driver
// does arithmetic of its own on these types (Isometry3f products / inverse, translation scaling, norms) the
// operations round like Eigen's float32 code does (row sums as x0 + (x1 + x2), see vo_pose_mul / vo_pose_inverse).
#pragma once
#include <cmath>
#include <cstddef>
#include <initializer_list>
#include <memory>
#include <ostream>
#include <vector>

#include "../../../include/vo_b200.h"

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#define EIGEN_WORLD_VERSION 3

namespace Eigen {

const int Dynamic = -1;

template <typename T, int R, int C>
class Matrix {
 public:
  typedef T Scalar;
  T m[R * C];  // row-major; a vector (C == 1) is R contiguous scalars like Eigen's

  Matrix() {
    for (int i = 0; i < R * C; ++i) m[i] = T(0);
  }
  Matrix(T x, T y) {
    static_assert(R * C == 2, "two-coefficient constructor");
    m[0] = x; m[1] = y;
  }
  Matrix(T x, T y, T z) {
    static_assert(R * C == 3, "three-coefficient constructor");
    m[0] = x; m[1] = y; m[2] = z;
  }
  Matrix(T x, T y, T z, T w) {
    static_assert(R * C == 4, "four-coefficient constructor");
    m[0] = x; m[1] = y; m[2] = z; m[3] = w;
  }
  static Matrix Zero() { return Matrix(); }
  static Matrix Identity() {
    Matrix r;
    for (int i = 0; i < (R < C ? R : C); ++i) r.m[i * C + i] = T(1);
    return r;
  }
  void setZero() { *this = Zero(); }
  void setIdentity() { *this = Identity(); }
  int rows() const { return R; }
  int cols() const { return C; }
  int size() const { return R * C; }
  T* data() { return m; }
  const T* data() const { return m; }
  Matrix& matrix() { return *this; }
  const Matrix& matrix() const { return *this; }
  T& operator()(int r, int c) { return m[r * C + c]; }
  const T& operator()(int r, int c) const { return m[r * C + c]; }
  T& operator()(int i) { return m[i]; }
  const T& operator()(int i) const { return m[i]; }
  T& operator[](int i) { return m[i]; }
  const T& operator[](int i) const { return m[i]; }
  T& x() { return m[0]; }
  T& y() { return m[1]; }
  T& z() { return m[2]; }
  T& w() { return m[3]; }
  const T& x() const { return m[0]; }
  const T& y() const { return m[1]; }
  const T& z() const { return m[2]; }
  const T& w() const { return m[3]; }

  Matrix operator+(const Matrix& o) const { Matrix r; for (int i = 0; i < R * C; ++i) r.m[i] = m[i] + o.m[i]; return r; }
  Matrix operator-(const Matrix& o) const { Matrix r; for (int i = 0; i < R * C; ++i) r.m[i] = m[i] - o.m[i]; return r; }
  Matrix operator-() const { Matrix r; for (int i = 0; i < R * C; ++i) r.m[i] = -m[i]; return r; }
  Matrix operator*(T s) const { Matrix r; for (int i = 0; i < R * C; ++i) r.m[i] = m[i] * s; return r; }
  Matrix operator/(T s) const { Matrix r; for (int i = 0; i < R * C; ++i) r.m[i] = m[i] / s; return r; }
  Matrix& operator+=(const Matrix& o) { for (int i = 0; i < R * C; ++i) m[i] += o.m[i]; return *this; }
  Matrix& operator-=(const Matrix& o) { for (int i = 0; i < R * C; ++i) m[i] -= o.m[i]; return *this; }
  Matrix& operator*=(T s) { for (int i = 0; i < R * C; ++i) m[i] *= s; return *this; }
  template <int C2>
  Matrix<T, R, C2> operator*(const Matrix<T, C, C2>& o) const {  // coefficient-wise sum_k a_ik b_kj, x0 + (x1 + (x2 ...))
    Matrix<T, R, C2> r;
    for (int i = 0; i < R; ++i)
      for (int j = 0; j < C2; ++j) {
        T s = m[i * C + C - 1] * o.m[(C - 1) * C2 + j];
        for (int k = C - 2; k >= 0; --k) s = m[i * C + k] * o.m[k * C2 + j] + s;
        r.m[i * C2 + j] = s;
      }
    return r;
  }
  Matrix<T, C, R> transpose() const {
    Matrix<T, C, R> r;
    for (int i = 0; i < R; ++i)
      for (int j = 0; j < C; ++j) r.m[j * R + i] = m[i * C + j];
    return r;
  }
  T dot(const Matrix& o) const { T s = T(0); for (int i = 0; i < R * C; ++i) s += m[i] * o.m[i]; return s; }
  T squaredNorm() const { return dot(*this); }
  T norm() const { return std::sqrt(squaredNorm()); }
  T trace() const { T s = T(0); for (int i = 0; i < (R < C ? R : C); ++i) s += m[i * C + i]; return s; }
  Matrix normalized() const { return *this / norm(); }
  Matrix<T, R, 1> col(int j) const { Matrix<T, R, 1> r; for (int i = 0; i < R; ++i) r.m[i] = m[i * C + j]; return r; }
  Matrix<T, 1, C> row(int i) const { Matrix<T, 1, C> r; for (int j = 0; j < C; ++j) r.m[j] = m[i * C + j]; return r; }
  template <int N>
  Matrix<T, N, 1> head() const { Matrix<T, N, 1> r; for (int i = 0; i < N; ++i) r.m[i] = m[i]; return r; }
  template <int BR, int BC>
  Matrix<T, BR, BC> block(int r0, int c0) const {
    Matrix<T, BR, BC> r;
    for (int i = 0; i < BR; ++i)
      for (int j = 0; j < BC; ++j) r.m[i * BC + j] = m[(r0 + i) * C + c0 + j];
    return r;
  }
  bool isApprox(const Matrix& o, T prec = T(1e-5)) const {
    const T d = (*this - o).squaredNorm(), a = squaredNorm(), b = o.squaredNorm();
    return d <= prec * prec * (a < b ? a : b);
  }
  // Eigen::MatrixBase::eulerAngles (Geometry/EulerAngles.h) for a 3x3 rotation
  Matrix<T, 3, 1> eulerAngles(int a0, int a1, int a2) const {
    static_assert(R == 3 && C == 3, "eulerAngles needs a 3x3 matrix");
    Matrix<T, 3, 1> res;
    const int odd = ((a0 + 1) % 3 == a1) ? 0 : 1;
    const int i = a0, j = (a0 + 1 + odd) % 3, k = (a0 + 2 - odd) % 3;
    auto c = [&](int r, int cc) { return m[r * 3 + cc]; };
    const T pi = T(3.14159265358979323846);
    if (a0 == a2) {
      res[0] = std::atan2(c(j, i), c(k, i));
      if ((odd && res[0] < T(0)) || ((!odd) && res[0] > T(0))) {
        res[0] = (res[0] > T(0)) ? res[0] - pi : res[0] + pi;
        const T s2 = std::sqrt(c(j, i) * c(j, i) + c(k, i) * c(k, i));
        res[1] = -std::atan2(s2, c(i, i));
      } else {
        const T s2 = std::sqrt(c(j, i) * c(j, i) + c(k, i) * c(k, i));
        res[1] = std::atan2(s2, c(i, i));
      }
      const T s1 = std::sin(res[0]), c1 = std::cos(res[0]);
      res[2] = std::atan2(c1 * c(j, k) - s1 * c(k, k), c1 * c(j, j) - s1 * c(k, j));
    } else {
      res[0] = std::atan2(c(j, k), c(k, k));
      const T c2 = std::sqrt(c(i, i) * c(i, i) + c(i, j) * c(i, j));
      if ((odd && res[0] < T(0)) || ((!odd) && res[0] > T(0))) {
        res[0] = (res[0] > T(0)) ? res[0] - pi : res[0] + pi;
        res[1] = std::atan2(-c(i, k), -c2);
      } else {
        res[1] = std::atan2(-c(i, k), c2);
      }
      const T s1 = std::sin(res[0]), c1 = std::cos(res[0]);
      res[2] = std::atan2(s1 * c(k, i) - c1 * c(j, i), c1 * c(j, j) - s1 * c(k, j));
    }
    if (!odd) res = -res;
    return res;
  }

  // comma initialiser:  M << a, b, c, ...;  (row-major order, like Eigen)
  struct CommaInit {
    Matrix* mat;
    int n;
    CommaInit& operator,(T v) {
      if (n < R * C) mat->m[n++] = v;
      return *this;
    }
  };
  CommaInit operator<<(T v) {
    m[0] = v;
    return CommaInit{this, 1};
  }
};

template <typename T, int R, int C>
Matrix<T, R, C> operator*(T s, const Matrix<T, R, C>& a) { return a * s; }

template <typename T, int R, int C>
std::ostream& operator<<(std::ostream& os, const Matrix<T, R, C>& a) {
  for (int i = 0; i < R; ++i) {
    for (int j = 0; j < C; ++j) os << (j ? " " : "") << a(i, j);
    if (i + 1 < R) os << "\n";
  }
  return os;
}

typedef Matrix<float, 2, 1> Vector2f;
typedef Matrix<float, 3, 1> Vector3f;
typedef Matrix<float, 4, 1> Vector4f;
typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<int, 2, 1> Vector2i;
typedef Matrix<float, 2, 2> Matrix2f;
typedef Matrix<float, 3, 3> Matrix3f;
typedef Matrix<float, 4, 4> Matrix4f;
typedef Matrix<double, 3, 3> Matrix3d;

// dynamic float vector: the descriptor of a Data_Point / World_Point (src/data_point.h)
class VectorXf {
 public:
  VectorXf() {}
  explicit VectorXf(int n) : v_(n > 0 ? n : 0, 0.f) {}
  static VectorXf Zero(int n) { return VectorXf(n); }
  int size() const { return (int)v_.size(); }
  int rows() const { return size(); }
  void resize(int n) { v_.resize(n); }
  void setZero() { for (auto& x : v_) x = 0.f; }
  float& operator()(int i) { return v_[i]; }
  float operator()(int i) const { return v_[i]; }
  float& operator[](int i) { return v_[i]; }
  float operator[](int i) const { return v_[i]; }
  float* data() { return v_.data(); }
  const float* data() const { return v_.data(); }
  VectorXf operator-(const VectorXf& o) const {
    VectorXf r(size());
    for (int i = 0; i < size(); ++i) r.v_[i] = v_[i] - o.v_[i];
    return r;
  }
  float squaredNorm() const { float s = 0.f; for (float x : v_) s += x * x; return s; }
  float norm() const { return std::sqrt(squaredNorm()); }

 private:
  std::vector<float> v_;
};
inline std::ostream& operator<<(std::ostream& os, const VectorXf& a) {
  for (int i = 0; i < a.size(); ++i) os << (i ? "\n" : "") << a[i];
  return os;
}

enum TransformTraits { Isometry = 0x1, Affine = 0x2, AffineCompact = 0x10 | Affine, Projective = 0x20 };

// Transform<float, 3, Isometry / Affine>: linear part + translation, with the member functions the drivers use
template <typename T, int Dim, int Mode>
class Transform {
  static_assert(Dim == 3, "only 3-D transforms are used by the reference");

 public:
  typedef Matrix<T, 3, 3> LinearMatrixType;
  typedef Matrix<T, 3, 1> VectorType;
  Transform() { lin_.setIdentity(); }
  static Transform Identity() { return Transform(); }
  void setIdentity() { *this = Transform(); }
  LinearMatrixType& linear() { return lin_; }
  const LinearMatrixType& linear() const { return lin_; }
  LinearMatrixType rotation() const { return lin_; }  // (an isometry's linear part IS its rotation)
  VectorType& translation() { return tr_; }
  const VectorType& translation() const { return tr_; }
  Matrix<T, 4, 4> matrix() const {
    Matrix<T, 4, 4> M = Matrix<T, 4, 4>::Identity();
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) M(i, j) = lin_(i, j);
      M(i, 3) = tr_[i];
    }
    return M;
  }
  // Isometry3f::inverse() is (R^T, -R^T t); products are R1 R2, R1 t2 + t1 - evaluated by the library's host helpers
  // so that they round exactly like the reference's Eigen float32 code
  Transform inverse() const {
    float a[12], b[12];
    to12(a);
    vo_pose_inverse(a, b);
    return from12(b);
  }
  Transform operator*(const Transform& o) const {
    float a[12], b[12], c[12];
    to12(a);
    o.to12(b);
    vo_pose_mul(a, b, c);
    return from12(c);
  }
  VectorType operator*(const VectorType& p) const {
    VectorType r;
    for (int i = 0; i < 3; ++i) {
      const T x0 = lin_(i, 0) * p[0], x1 = lin_(i, 1) * p[1], x2 = lin_(i, 2) * p[2];
      r[i] = tr_[i] + (x0 + (x1 + x2));
    }
    return r;
  }
  bool isApprox(const Transform& o, T prec = T(1e-5)) const { return matrix().isApprox(o.matrix(), prec); }
  void to12(float* a) const {
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) a[4 * i + j] = (float)lin_(i, j);
      a[4 * i + 3] = (float)tr_[i];
    }
  }
  static Transform from12(const float* a) {
    Transform t;
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) t.lin_(i, j) = (T)a[4 * i + j];
      t.tr_[i] = (T)a[4 * i + 3];
    }
    return t;
  }

 private:
  LinearMatrixType lin_;
  VectorType tr_;
};
typedef Transform<float, 3, Isometry> Isometry3f;
typedef Transform<float, 3, Affine> Affine3f;

template <class T>
using aligned_allocator = std::allocator<T>;

}  // namespace Eigen
