// five_point.cpp — TEST INFRASTRUCTURE (part of libvo_oracle.so, never linked into the product).
//
// CPU restatement of cv::findEssentialMat(points1, points2, K, cv::RANSAC) with OpenCV's defaults
// (prob 0.999, threshold 1.0 px, maxIters 1000), which is what the reference calls at src/cam.cpp:49.
// OpenCV is an un-vendored, un-pinned dependency of the reference (CMakeLists.txt:12) and its C++ side is not in
// this image; the published algorithm of OpenCV 4.x is restated here from
//   modules/calib3d/src/five-point.cpp   EMEstimatorCallback::runKernel / computeError, findEssentialMat
//   modules/calib3d/src/ptsetreg.cpp     RANSACPointSetRegistrator::run / getSubset / findInliers, RANSACUpdateNumIters
//   modules/core/src/lapack.cpp          JacobiSVDImpl_ (the FULL_UV completion decides the null-space basis)
//   modules/core/src/mathfuncs.cpp       solvePoly (Durand-Kerner; its iteration decides the ORDER of the roots,
//                                        and the first best hypothesis wins ties)
//   modules/core/include/opencv2/core.hpp  cv::RNG (multiply-with-carry), seeds (uint64)-1 and 0x12345678
// and PINNED against the black box itself: cv2 4.13.0's findEssentialMat on 32 committed fixtures
// (tests/golden/cv2_fixtures.npz, cv2_recoverpose.npz: the bundled dataset's frame pairs, synthetic motions with
// noise and 10 % gross outliers, up to 122 RANSAC iterations) - E agrees to 1e-10 (1e-5 on three samples whose
// polynomial has clustered roots), inlier masks identical: tests/test_oracle_golden.py::test_ransac_*.
// exp/five_point_proto.py is the numpy prototype this file follows.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstring>
#include <vector>

#include "vo_oracle.h"

namespace {

struct CvRng {  // cv::RNG
  uint64_t state;
  explicit CvRng(uint64_t s) : state(s ? s : 0xffffffffull) {}
  unsigned next() {
    state = (uint64_t)(unsigned)state * 4164903690u + (unsigned)(state >> 32);
    return (unsigned)state;
  }
  int uniform(int a, int b) { return a == b ? a : (int)(next() % (unsigned)(b - a) + a); }
};

// cv::JacobiSVDImpl_<double> (modules/core/src/lapack.cpp), restated.  At: n rows of length m (the transposed input),
// orthogonalised in place by one-sided Jacobi rotations of row pairs (the same rotations accumulate in Vt, n x n,
// nullable), rows sorted by norm (descending), then the first n1 rows are normalised; a row whose norm is <= DBL_MIN
// (and every row i >= n) is replaced by a +-1/m sign vector from RNG(0x12345678) Gram-Schmidt'ed twice against all
// previous rows.  W = singular values.
void cv_jacobi_svd(double* At, int astep, double* W, double* Vt, int vstep, int m, int n, int n1) {
  const double eps = DBL_EPSILON * 10;
  for (int i = 0; i < n; ++i) {
    double sd = 0;
    for (int k = 0; k < m; ++k) sd += At[i * astep + k] * At[i * astep + k];
    W[i] = sd;
    if (Vt) {
      for (int k = 0; k < n; ++k) Vt[i * vstep + k] = 0;
      Vt[i * vstep + i] = 1;
    }
  }
  const int max_iter = std::max(m, 30);
  for (int iter = 0; iter < max_iter; ++iter) {
    bool changed = false;
    for (int i = 0; i < n - 1; ++i)
      for (int j = i + 1; j < n; ++j) {
        double* Ai = At + i * astep;
        double* Aj = At + j * astep;
        double a = W[i], p = 0, b = W[j];
        for (int k = 0; k < m; ++k) p += Ai[k] * Aj[k];
        if (std::fabs(p) <= eps * std::sqrt(a * b)) continue;
        p *= 2;
        const double beta = a - b, gamma = hypot(p, beta);
        double c, s;
        if (beta < 0) {
          const double delta = (gamma - beta) * 0.5;
          s = std::sqrt(delta / gamma);
          c = p / (gamma * s * 2);
        } else {
          c = std::sqrt((gamma + beta) / (gamma * 2));
          s = p / (gamma * c * 2);
        }
        a = b = 0;
        for (int k = 0; k < m; ++k) {
          const double t0 = c * Ai[k] + s * Aj[k], t1 = -s * Ai[k] + c * Aj[k];
          Ai[k] = t0;
          Aj[k] = t1;
          a += t0 * t0;
          b += t1 * t1;
        }
        W[i] = a;
        W[j] = b;
        changed = true;
        if (Vt) {
          double* Vi = Vt + i * vstep;
          double* Vj = Vt + j * vstep;
          for (int k = 0; k < n; ++k) {
            const double t0 = c * Vi[k] + s * Vj[k], t1 = -s * Vi[k] + c * Vj[k];
            Vi[k] = t0;
            Vj[k] = t1;
          }
        }
      }
    if (!changed) break;
  }
  for (int i = 0; i < n; ++i) {
    double sd = 0;
    for (int k = 0; k < m; ++k) sd += At[i * astep + k] * At[i * astep + k];
    W[i] = std::sqrt(sd);
  }
  for (int i = 0; i < n - 1; ++i) {
    int j = i;
    for (int k = i + 1; k < n; ++k)
      if (W[j] < W[k]) j = k;
    if (i != j) {
      std::swap(W[i], W[j]);
      if (Vt) {
        for (int k = 0; k < m; ++k) std::swap(At[i * astep + k], At[j * astep + k]);
        for (int k = 0; k < n; ++k) std::swap(Vt[i * vstep + k], Vt[j * vstep + k]);
      }
    }
  }
  if (!Vt) return;
  CvRng rng(0x12345678);
  for (int i = 0; i < n1; ++i) {
    double sd = i < n ? W[i] : 0;
    double* Ai = At + i * astep;
    for (int ii = 0; ii < 100 && sd <= DBL_MIN; ++ii) {
      const double val0 = 1. / m;
      for (int k = 0; k < m; ++k) Ai[k] = (rng.next() & 256) != 0 ? val0 : -val0;
      for (int iter = 0; iter < 2; ++iter)
        for (int j = 0; j < i; ++j) {
          const double* Aj = At + j * astep;
          sd = 0;
          for (int k = 0; k < m; ++k) sd += Ai[k] * Aj[k];
          double asum = 0;
          for (int k = 0; k < m; ++k) {
            const double t = Ai[k] - sd * Aj[k];
            Ai[k] = t;
            asum += std::fabs(t);
          }
          asum = asum > eps * 100 ? 1 / asum : 0;
          for (int k = 0; k < m; ++k) Ai[k] *= asum;
        }
      sd = 0;
      for (int k = 0; k < m; ++k) sd += Ai[k] * Ai[k];
      sd = std::sqrt(sd);
    }
    const double s = sd > DBL_MIN ? 1 / sd : 0.;
    for (int k = 0; k < m; ++k) Ai[k] *= s;
  }
}

// rows 5..8 of Vt of cv::SVD::compute(Q 5x9, MODIFY_A | FULL_UV): m < n, so OpenCV factors the transpose - the five
// rows of Q are the "columns", completed to nine orthonormal rows
void cv_null_basis(const double Q[5][9], double EE[4][9]) {
  double At[9][9] = {}, W[9], Vt[5][5];
  for (int i = 0; i < 5; ++i)
    for (int k = 0; k < 9; ++k) At[i][k] = Q[i][k];
  cv_jacobi_svd(&At[0][0], 9, W, &Vt[0][0], 5, 9, 5, 9);
  for (int i = 0; i < 4; ++i)
    for (int k = 0; k < 9; ++k) EE[i][k] = At[5 + i][k];
}

// cv::solvePoly: Durand-Kerner from the start values (1+i)^k, in-place updates, 300 sweeps. c[k] = coefficient of
// z^k. Returns the number of roots written (OpenCV's order).
int cv_solve_poly(const double* c, int deg, std::complex<double>* roots, int max_iters = 300) {
  typedef std::complex<double> C;
  int n = deg;
  for (; n > 1; --n)
    if (std::fabs(c[n]) > DBL_EPSILON) break;
  C p(1, 0), r(1, 1);
  for (int i = 0; i < n; ++i) {
    roots[i] = p;
    p = p * r;
  }
  for (int iter = 0; iter < max_iters; ++iter) {
    double max_diff = 0;
    for (int i = 0; i < n; ++i) {
      p = roots[i];
      C num(c[n], 0), den(c[n], 0);
      for (int j = 0; j < n; ++j) {
        num = num * p + c[n - j - 1];
        if (j != i && (p.real() != roots[j].real() || p.imag() != roots[j].imag())) den = den * (p - roots[j]);
      }
      num /= den;
      roots[i] = p - num;
      max_diff = std::max(max_diff, std::abs(num));
    }
    if (max_diff <= 0) break;
  }
  return n;
}

// polynomials in (x, y, z) of total degree <= 3, coefficient of x^i y^j z^k at [i][j][k]
struct Poly3 {
  double c[4][4][4];
  Poly3() { std::memset(c, 0, sizeof(c)); }
};
Poly3 pmul(const Poly3& a, const Poly3& b) {
  Poly3 r;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; i + j < 4; ++j)
      for (int k = 0; i + j + k < 4; ++k) {
        const double av = a.c[i][j][k];
        if (av == 0) continue;
        for (int u = 0; i + j + k + u < 4; ++u)
          for (int v = 0; i + j + k + u + v < 4; ++v)
            for (int w = 0; i + j + k + u + v + w < 4; ++w) r.c[i + u][j + v][k + w] += av * b.c[u][v][w];
      }
  return r;
}
Poly3 padd(const Poly3& a, const Poly3& b, double s = 1.0) {
  Poly3 r = a;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j)
      for (int k = 0; k < 4; ++k) r.c[i][j][k] += s * b.c[i][j][k];
  return r;
}

// Nister's elimination order: the first ten monomials are eliminated, the last ten are {x, y, 1} x powers of z
const int kMono[20][3] = {{3, 0, 0}, {0, 3, 0}, {2, 1, 0}, {1, 2, 0}, {2, 0, 1}, {2, 0, 0}, {0, 2, 1},
                          {0, 2, 0}, {1, 1, 1}, {1, 1, 0}, {1, 0, 2}, {1, 0, 1}, {1, 0, 0}, {0, 1, 2},
                          {0, 1, 1}, {0, 1, 0}, {0, 0, 3}, {0, 0, 2}, {0, 0, 1}, {0, 0, 0}};

// X = A1^-1 A2 for the 10x20 system [A1 | A2] (Gauss-Jordan, partial pivoting); false if singular
bool reduce10(double A[10][20]) {
  for (int col = 0; col < 10; ++col) {
    int piv = col;
    for (int r = col + 1; r < 10; ++r)
      if (std::fabs(A[r][col]) > std::fabs(A[piv][col])) piv = r;
    if (std::fabs(A[piv][col]) < DBL_MIN) return false;
    if (piv != col)
      for (int k = 0; k < 20; ++k) std::swap(A[piv][k], A[col][k]);
    const double inv = 1.0 / A[col][col];
    for (int k = 0; k < 20; ++k) A[col][k] *= inv;
    for (int r = 0; r < 10; ++r) {
      if (r == col) continue;
      const double f = A[r][col];
      if (f == 0) continue;
      for (int k = 0; k < 20; ++k) A[r][k] -= f * A[col][k];
    }
  }
  return true;
}

// univariate polynomials, c[k] = coefficient of z^k
struct Poly1 {
  double c[11];
  int deg;
  Poly1() : deg(0) { std::memset(c, 0, sizeof(c)); }
  double at(double z) const {
    double v = 0;
    for (int k = deg; k >= 0; --k) v = v * z + c[k];
    return v;
  }
};
Poly1 p1mul(const Poly1& a, const Poly1& b) {
  Poly1 r;
  r.deg = a.deg + b.deg;
  for (int i = 0; i <= a.deg; ++i)
    for (int j = 0; j <= b.deg; ++j) r.c[i + j] += a.c[i] * b.c[j];
  return r;
}
Poly1 p1sub(const Poly1& a, const Poly1& b) {
  Poly1 r;
  r.deg = std::max(a.deg, b.deg);
  for (int i = 0; i <= r.deg; ++i) r.c[i] = a.c[i] - b.c[i];
  return r;
}

// null vector of a 3x3 matrix: right singular vector of the smallest singular value (cv::SVD::solveZ)
void null3(const double B[3][3], double v[3]) {
  double A[9], V[9], w[3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) A[3 * i + j] = B[i][j];
  vo_ref_jacobi_svd(A, 3, 3, V, w);
  for (int k = 0; k < 3; ++k) v[k] = V[3 * k + 2];
}

// EMEstimatorCallback::runKernel: five normalised correspondences -> up to 10 essential matrices (unit Frobenius norm,
// row-major, x2^T E x1 = 0), in OpenCV's order
int five_point(const double q1[5][2], const double q2[5][2], double E_out[10][9]) {
  double Q[5][9];
  for (int i = 0; i < 5; ++i) {
    const double x1 = q1[i][0], y1 = q1[i][1], x2 = q2[i][0], y2 = q2[i][1];
    const double row[9] = {x2 * x1, x2 * y1, x2, y2 * x1, y2 * y1, y2, x1, y1, 1.0};
    std::memcpy(Q[i], row, sizeof(row));
  }
  double EE[4][9];
  cv_null_basis(Q, EE);
  // E(x, y, z) = x E0 + y E1 + z E2 + E3: nine degree-1 polynomials
  Poly3 Ep[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      Ep[i][j].c[1][0][0] = EE[0][3 * i + j];
      Ep[i][j].c[0][1][0] = EE[1][3 * i + j];
      Ep[i][j].c[0][0][1] = EE[2][3 * i + j];
      Ep[i][j].c[0][0][0] = EE[3][3 * i + j];
    }
  std::vector<Poly3> cons;
  {  // det E = 0
    Poly3 t = pmul(Ep[0][0], padd(pmul(Ep[1][1], Ep[2][2]), pmul(Ep[1][2], Ep[2][1]), -1));
    t = padd(t, pmul(Ep[0][1], padd(pmul(Ep[1][0], Ep[2][2]), pmul(Ep[1][2], Ep[2][0]), -1)), -1);
    t = padd(t, pmul(Ep[0][2], padd(pmul(Ep[1][0], Ep[2][1]), pmul(Ep[1][1], Ep[2][0]), -1)));
    cons.push_back(t);
  }
  Poly3 EEt[3][3], tr;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      for (int k = 0; k < 3; ++k) EEt[i][j] = padd(EEt[i][j], pmul(Ep[i][k], Ep[j][k]));
  tr = padd(padd(EEt[0][0], EEt[1][1]), EEt[2][2]);
  for (int i = 0; i < 3; ++i)  // 2 E E^T E - trace(E E^T) E = 0
    for (int j = 0; j < 3; ++j) {
      Poly3 s;
      for (int k = 0; k < 3; ++k) s = padd(s, pmul(EEt[i][k], Ep[k][j]), 2.0);
      cons.push_back(padd(s, pmul(tr, Ep[i][j]), -1.0));
    }
  double A[10][20];
  for (int r = 0; r < 10; ++r)
    for (int m = 0; m < 20; ++m) A[r][m] = cons[r].c[kMono[m][0]][kMono[m][1]][kMono[m][2]];
  if (!reduce10(A)) return 0;
  // rows 4..9 of the reduced system express x^2 z, x^2, y^2 z, y^2, xyz, xy: (row 2i+4) - z (row 2i+5) = 0 gives three
  // equations [cubic(z)] x + [cubic(z)] y + [quartic(z)] = 0
  Poly1 P[3][3];
  for (int i = 0; i < 3; ++i) {
    const double* a1 = &A[2 * i + 4][10];
    const double* a2 = &A[2 * i + 5][10];
    double b[13] = {};
    double r1[13] = {}, r2[13] = {};
    for (int k = 0; k < 3; ++k) { r1[1 + k] = a1[k]; r1[5 + k] = a1[3 + k]; r2[k] = a2[k]; r2[4 + k] = a2[3 + k]; }
    for (int k = 0; k < 4; ++k) { r1[9 + k] = a1[6 + k]; r2[8 + k] = a2[6 + k]; }
    for (int k = 0; k < 13; ++k) b[k] = r1[k] - r2[k];
    P[i][0].deg = 3; P[i][1].deg = 3; P[i][2].deg = 4;
    for (int k = 0; k < 4; ++k) { P[i][0].c[3 - k] = b[k]; P[i][1].c[3 - k] = b[4 + k]; }
    for (int k = 0; k < 5; ++k) P[i][2].c[4 - k] = b[8 + k];
  }
  Poly1 det = p1mul(P[0][0], p1sub(p1mul(P[1][1], P[2][2]), p1mul(P[1][2], P[2][1])));
  det = p1sub(det, p1mul(P[0][1], p1sub(p1mul(P[1][0], P[2][2]), p1mul(P[1][2], P[2][0]))));
  Poly1 last = p1mul(P[0][2], p1sub(p1mul(P[1][0], P[2][1]), p1mul(P[1][1], P[2][0])));
  for (int k = 0; k <= 10; ++k) det.c[k] += last.c[k];
  std::complex<double> roots[10];
  const int n_roots = cv_solve_poly(det.c, 10, roots);
  int count = 0;
  for (int i = 0; i < n_roots && count < 10; ++i) {
    if (std::fabs(roots[i].imag()) > 1e-10) continue;
    const double z = roots[i].real();
    double Bz[3][3];
    for (int j = 0; j < 3; ++j)
      for (int k = 0; k < 3; ++k) Bz[j][k] = P[j][k].at(z);
    double xy1[3];
    null3(Bz, xy1);
    if (std::fabs(xy1[2]) < 1e-10) continue;
    const double x = xy1[0] / xy1[2], y = xy1[1] / xy1[2];
    double nrm = 0;
    for (int k = 0; k < 9; ++k) {
      E_out[count][k] = EE[0][k] * x + EE[1][k] * y + EE[2][k] * z + EE[3][k];
      nrm += E_out[count][k] * E_out[count][k];
    }
    nrm = std::sqrt(nrm);
    for (int k = 0; k < 9; ++k) E_out[count][k] /= nrm;
    ++count;
  }
  return count;
}

// EMEstimatorCallback::computeError: Sampson distance, rounded to float like the cv::Mat it is stored in
inline float sampson(const double E[9], double x1, double y1, double x2, double y2) {
  const double Ex1[3] = {E[0] * x1 + E[1] * y1 + E[2], E[3] * x1 + E[4] * y1 + E[5], E[6] * x1 + E[7] * y1 + E[8]};
  const double Etx2[2] = {E[0] * x2 + E[3] * y2 + E[6], E[1] * x2 + E[4] * y2 + E[7]};
  const double x2tEx1 = x2 * Ex1[0] + y2 * Ex1[1] + Ex1[2];
  const double a = Ex1[0] * Ex1[0], b = Ex1[1] * Ex1[1], c = Etx2[0] * Etx2[0], d = Etx2[1] * Etx2[1];
  return (float)(x2tEx1 * x2tEx1 / (a + b + c + d));
}

int ransac_update_num_iters(double p, double ep, int model_points, int max_iters) {
  p = std::min(std::max(p, 0.), 1.);
  ep = std::min(std::max(ep, 0.), 1.);
  double num = std::max(1. - p, DBL_MIN);
  double denom = 1. - std::pow(1. - ep, model_points);
  if (denom < DBL_MIN) return 0;
  num = std::log(num);
  denom = std::log(denom);
  return denom >= 0 || -num >= max_iters * (-denom) ? max_iters : (int)std::nearbyint(num / denom);  // cvRound
}

}  // namespace

extern "C" {

// cv::SVD::compute(A 3x3, w, U, Vt): U and Vt row-major, singular values descending
void vo_ref_cv_svd3(const double A[9], double U[9], double w[3], double Vt[9]) {
  double At[9], V[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) At[3 * i + j] = A[3 * j + i];  // rows of At = columns of A
  cv_jacobi_svd(At, 3, w, V, 3, 3, 3, 3);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      U[3 * j + i] = At[3 * i + j];  // U = At^T: columns are the left singular vectors
      Vt[3 * i + j] = V[3 * i + j];
    }
}

int vo_ref_five_point(const double q1[10], const double q2[10], double E_out[90]) {
  double a[5][2], b[5][2], E[10][9];
  for (int i = 0; i < 5; ++i) {
    a[i][0] = q1[2 * i]; a[i][1] = q1[2 * i + 1];
    b[i][0] = q2[2 * i]; b[i][1] = q2[2 * i + 1];
  }
  const int n = five_point(a, b, E);
  std::memcpy(E_out, E, sizeof(double) * 9 * n);
  return n;
}

void vo_ref_ransac_subsets(int n, int n_subsets, int32_t* idx_out) {
  CvRng rng((uint64_t)-1);
  for (int s = 0; s < n_subsets; ++s) {
    int idx[5];
    for (int i = 0; i < 5; ++i) {
      int v = rng.uniform(0, n);
      while (std::find(idx, idx + i, v) != idx + i) v = rng.uniform(0, n);
      idx[i] = v;
    }
    for (int i = 0; i < 5; ++i) idx_out[5 * s + i] = idx[i];
  }
}

int vo_ref_find_essential_ransac(const float K[9], const float* x1, const float* x2, int64_t n, double prob,
                                 double threshold, int max_iters, double E[9], uint8_t* mask, int* iters_out) {
  const double fx = K[0], fy = K[4], cx = K[2], cy = K[5];
  if (iters_out) *iters_out = 0;
  if (n < 5) return 0;
  // (points - c) / f as one scaled conversion: x * (1/f) + (-c * (1/f))   (cv::MatExpr of the normalisation)
  const double ifx = 1. / fx, ify = 1. / fy;
  std::vector<double> q1(2 * n), q2(2 * n);
  for (int64_t i = 0; i < n; ++i) {
    q1[2 * i] = (double)x1[2 * i] * ifx + (-cx * ifx);
    q1[2 * i + 1] = (double)x1[2 * i + 1] * ify + (-cy * ify);
    q2[2 * i] = (double)x2[2 * i] * ifx + (-cx * ifx);
    q2[2 * i + 1] = (double)x2[2 * i + 1] * ify + (-cy * ify);
  }
  const double thr = threshold / ((fx + fy) / 2);
  const float t = (float)(thr * thr);
  std::vector<uint8_t> cur(n), best_mask(n, 0);
  int max_good = 0;
  double best[9] = {};
  int niters = std::max(max_iters, 1), iter = 0;
  CvRng rng((uint64_t)-1);
  if (n == 5) {  // count == modelPoints: the only sample, first model, all points inliers
    double a[5][2], b[5][2], Es[10][9];
    for (int i = 0; i < 5; ++i) { a[i][0] = q1[2 * i]; a[i][1] = q1[2 * i + 1]; b[i][0] = q2[2 * i]; b[i][1] = q2[2 * i + 1]; }
    const int nm = five_point(a, b, Es);
    if (nm <= 0) return 0;
    std::memcpy(E, Es[0], sizeof(best));
    if (mask) std::memset(mask, 1, n);
    return 5;
  }
  for (; iter < niters; ++iter) {
    int idx[5];
    for (int i = 0; i < 5; ++i) {
      int v = rng.uniform(0, (int)n);
      while (std::find(idx, idx + i, v) != idx + i) v = rng.uniform(0, (int)n);
      idx[i] = v;
    }
    double a[5][2], b[5][2], Es[10][9];
    for (int i = 0; i < 5; ++i) {
      a[i][0] = q1[2 * idx[i]]; a[i][1] = q1[2 * idx[i] + 1];
      b[i][0] = q2[2 * idx[i]]; b[i][1] = q2[2 * idx[i] + 1];
    }
    const int nm = five_point(a, b, Es);
    for (int m = 0; m < nm; ++m) {
      int good = 0;
      for (int64_t i = 0; i < n; ++i) {
        const bool in = sampson(Es[m], q1[2 * i], q1[2 * i + 1], q2[2 * i], q2[2 * i + 1]) <= t;
        cur[i] = in;
        good += in;
      }
      if (good > std::max(max_good, 4)) {
        best_mask.swap(cur);
        std::memcpy(best, Es[m], sizeof(best));
        max_good = good;
        niters = ransac_update_num_iters(prob, (double)(n - good) / (double)n, 5, niters);
      }
    }
  }
  if (iters_out) *iters_out = iter;
  if (max_good <= 0) return 0;
  std::memcpy(E, best, sizeof(best));
  if (mask) std::memcpy(mask, best_mask.data(), n);
  return max_good;
}

// src/cam.cpp:37-91 as the reference runs it: findEssentialMat(RANSAC, defaults) then recoverPose
int vo_ref_essential_recover_ransac(const float K[9], const float* x1, const float* x2, int64_t n, double E[9],
                                    double R[9], double t[3], uint8_t* mask) {
  if (vo_ref_find_essential_ransac(K, x1, x2, n, 0.999, 1.0, 1000, E, nullptr, nullptr) <= 0) return -1;
  return vo_ref_recover_pose(E, K, x1, x2, n, R, t, mask);
}

}  // extern "C"
