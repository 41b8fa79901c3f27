// vo_cv_shim.h — the compile-time surface of OpenCV that the reference's drivers and the replacement headers touch:
// cv::Point2f / cv::Point3f (fields of Data_Point / World_Point, src/data_point.h) and a byte-matrix cv::Mat (the
// `mask` argument of Cam::computeEssentialAndRecoverPose, src/cam.h).  A TYPE shim: OpenCV's algorithms
// (findEssentialMat, recoverPose, triangulatePoints) run inside libvo_b200.so, restated for the GPU.
#pragma once
#include <cstdint>
#include <ostream>
#include <vector>

#define CV_8U 0
#define CV_32F 5
#define CV_64F 6

typedef unsigned char uchar;  // OpenCV's global typedef (exec/pose_recovery_test.cpp:47)

namespace cv {

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

// rows x cols matrix of bytes (CV_8U) or doubles (CV_64F), just enough for masks and the stored R / t
class Mat {
 public:
  int rows = 0, cols = 0;
  Mat() {}
  Mat(int r, int c, int type) { create(r, c, type); }
  void create(int r, int c, int type) {
    rows = r;
    cols = c;
    type_ = type;
    bytes_.assign((size_t)r * c * (type == CV_64F ? 8 : type == CV_32F ? 4 : 1), 0);
  }
  bool empty() const { return rows == 0 || cols == 0; }
  int type() const { return type_; }
  size_t total() const { return (size_t)rows * cols; }
  template <typename T>
  T& at(int i, int j = 0) { return reinterpret_cast<T*>(bytes_.data())[(size_t)i * cols + j]; }
  template <typename T>
  const T& at(int i, int j = 0) const { return reinterpret_cast<const T*>(bytes_.data())[(size_t)i * cols + j]; }
  unsigned char* data() { return bytes_.data(); }

 private:
  int type_ = CV_8U;
  std::vector<unsigned char> bytes_;
};

inline void setRNGSeed(int) {}  // (no effect on findEssentialMat in OpenCV either: it seeds its own RNG)

}  // namespace cv
