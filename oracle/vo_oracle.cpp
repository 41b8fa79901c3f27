// vo_oracle.cpp — CPU ORACLE (test infrastructure, NOT the product). See vo_oracle.h.
//
// Restates the reference algorithms with the float32 evaluation order Eigen uses
// (SURVEY.md Appendix A).  No Eigen, no OpenCV, no CUDA.  Build:
//   g++ -O2 -std=c++17 -ffp-contract=off -fPIC -shared -pthread (oracle/Makefile)
#include "vo_oracle.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

namespace {

// ---- Eigen fixed-size coefficient products (Appendix A.3): length-3 inner
// products reduce as x0 + (x1 + x2); length-2 as x0 + x1; never fused.
inline float dot3(float a0, float b0, float a1, float b1, float a2, float b2) {
  float x0 = a0 * b0, x1 = a1 * b1, x2 = a2 * b2;
  return x0 + (x1 + x2);
}

// Isometry3f * Vector3f (Appendix A.2): res = t; res += R*p
inline void iso_apply(const float T[12], const float p[3], float c[3]) {
  for (int i = 0; i < 3; ++i)
    c[i] = T[4 * i + 3] + dot3(T[4 * i], p[0], T[4 * i + 1], p[1], T[4 * i + 2], p[2]);
}

inline void mat3_apply(const float K[9], const float c[3], float q[3]) {
  for (int i = 0; i < 3; ++i) q[i] = dot3(K[3 * i], c[0], K[3 * i + 1], c[1], K[3 * i + 2], c[2]);
}

// reference: src/camera.h:24-36 (Camera::projectPoint)
inline bool project_full(const float K[9], int rows, int cols, const float T[12], const float p[3],
                         float c[3], float q[3], float uv[2]) {
  iso_apply(T, p, c);
  if (c[2] <= 0) return false;
  mat3_apply(K, c, q);
  float iz = (float)(1. / (double)q[2]);  // camera.h:30: double reciprocal promoted back to float
  uv[0] = q[0] * iz;
  uv[1] = q[1] * iz;
  if (uv[0] < 0 || uv[0] > (float)(cols - 1)) return false;
  if (uv[1] < 0 || uv[1] > (float)(rows - 1)) return false;
  return true;
}

// reference: src/picp_solver.cpp:26-54 (PICPSolver::errorAndJacobian)
inline bool error_jacobian(const float K[9], int rows, int cols, const float T[12], const float p[3],
                           const float z[2], float e[2], float J[12]) {
  float c[3], q[3], uv[2];
  if (!project_full(K, rows, cols, T, p, c, q, uv)) return false;
  e[0] = uv[0] - z[0];
  e[1] = uv[1] - z[1];
  // Jr = [I | skew(-c)]   (defs.h:139-145 with v = -c)
  float v0 = -c[0], v1 = -c[1], v2 = -c[2];
  float Jr[3][6] = {{1, 0, 0, 0, -v2, v1}, {0, 1, 0, v2, 0, -v0}, {0, 0, 1, -v1, v0, 0}};
  float iz = (float)(1. / (double)q[2]);
  float iz2 = iz * iz;
  float Jp[2][3] = {{iz, 0, -q[0] * iz2}, {0, iz, -q[1] * iz2}};
  float A[2][3];  // Jp*K materialised first (left-associated product)
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 3; ++j) A[i][j] = dot3(Jp[i][0], K[j], Jp[i][1], K[3 + j], Jp[i][2], K[6 + j]);
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 6; ++j) J[6 * i + j] = dot3(A[i][0], Jr[0][j], A[i][1], Jr[1][j], A[i][2], Jr[2][j]);
  return true;
}

struct Lin32 {
  float H[36], b[6], chi_in, chi_out;
  int64_t n_in;
};

// reference: src/picp_solver.cpp:56-91 (PICPSolver::linearize), float32 sequential
void linearize32(const float K[9], int rows, int cols, const float T[12], const float* W,
                 const float* Z, const int32_t* pairs, int64_t lo, int64_t hi, float thr,
                 bool keep, Lin32& out, uint8_t* status) {
  Lin32 o;  // thread-local accumulators (the per-thread slots of `parts` share cache lines)
  std::memset(&o, 0, sizeof(o));
  for (int64_t n = lo; n < hi; ++n) {
    int ref_idx = pairs[2 * n], curr_idx = pairs[2 * n + 1];
    float e[2], J[12];
    bool inside = error_jacobian(K, rows, cols, T, W + 3 * (int64_t)curr_idx, Z + 2 * (int64_t)ref_idx, e, J);
    if (!inside) {
      if (status) status[n] = VO_REF_SKIPPED;
      continue;
    }
    float chi = e[0] * e[0] + e[1] * e[1];
    float lambda = 1;
    bool inl = true;
    if (chi > thr) {
      lambda = std::sqrt(thr / chi);
      inl = false;
      o.chi_out += chi;
    } else {
      o.chi_in += chi;
      o.n_in++;
    }
    if (status) status[n] = inl ? VO_REF_INLIER : VO_REF_OUTLIER;
    if (inl || keep) {
      for (int i = 0; i < 6; ++i) {
        for (int j = 0; j < 6; ++j) o.H[6 * i + j] += (J[i] * J[j] + J[6 + i] * J[6 + j]) * lambda;
        o.b[i] += (J[i] * e[0] + J[6 + i] * e[1]) * lambda;
      }
    }
  }
  out = o;
}

// same float32 per-correspondence terms, float64 accumulators (tolerance anchor)
void linearize64(const float K[9], int rows, int cols, const float T[12], const float* W,
                 const float* Z, const int32_t* pairs, int64_t n_pairs, float thr, bool keep,
                 double H[36], double b[6], double& chi_in, double& chi_out, int64_t& n_in,
                 uint8_t* status) {
  std::fill(H, H + 36, 0.0);
  std::fill(b, b + 6, 0.0);
  chi_in = chi_out = 0;
  n_in = 0;
  for (int64_t n = 0; n < n_pairs; ++n) {
    float e[2], J[12];
    bool inside = error_jacobian(K, rows, cols, T, W + 3 * (int64_t)pairs[2 * n + 1],
                                 Z + 2 * (int64_t)pairs[2 * n], e, J);
    if (!inside) {
      if (status) status[n] = VO_REF_SKIPPED;
      continue;
    }
    float chi = e[0] * e[0] + e[1] * e[1];
    double lambda = 1;
    bool inl = true;
    if (chi > thr) {
      lambda = std::sqrt((double)thr / (double)chi);
      inl = false;
      chi_out += chi;
    } else {
      chi_in += chi;
      n_in++;
    }
    if (status) status[n] = inl ? VO_REF_INLIER : VO_REF_OUTLIER;
    if (inl || keep) {
      for (int i = 0; i < 6; ++i) {
        for (int j = 0; j < 6; ++j)
          H[6 * i + j] += ((double)J[i] * J[j] + (double)J[6 + i] * J[6 + j]) * lambda;
        b[i] += ((double)J[i] * e[0] + (double)J[6 + i] * e[1]) * lambda;
      }
    }
  }
}

// Eigen::LDLT<Matrix6f> (diagonal pivoting, lower storage), float32.
// Follows the published unblocked algorithm (Eigen/src/Cholesky/LDLT.h).
void ldlt_solve6(const float Ain[36], const float rhs[6], float x[6]) {
  const int n = 6;
  float m[6][6];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) m[i][j] = Ain[6 * i + j];
  int tr[6];
  for (int k = 0; k < n; ++k) {
    int big = k;
    float bigv = std::fabs(m[k][k]);
    for (int i = k + 1; i < n; ++i)
      if (std::fabs(m[i][i]) > bigv) {
        bigv = std::fabs(m[i][i]);
        big = i;
      }
    tr[k] = big;
    if (big != k) {  // symmetric row/col swap touching only the lower triangle
      for (int j = 0; j < k; ++j) std::swap(m[k][j], m[big][j]);
      for (int i = big + 1; i < n; ++i) std::swap(m[i][k], m[i][big]);
      std::swap(m[k][k], m[big][big]);
      for (int i = k + 1; i < big; ++i) std::swap(m[i][k], m[big][i]);
    }
    if (k > 0) {
      float temp[6];
      for (int j = 0; j < k; ++j) temp[j] = m[j][j] * m[k][j];
      float s = 0;
      for (int j = 0; j < k; ++j) s += m[k][j] * temp[j];
      m[k][k] -= s;
      for (int i = k + 1; i < n; ++i) {
        float a = 0;
        for (int j = 0; j < k; ++j) a += m[i][j] * temp[j];
        m[i][k] -= a;
      }
    }
    float akk = m[k][k];
    bool valid = std::fabs(akk) > 0;
    if (k == 0 && !valid) {  // zero matrix
      for (int j = 0; j < n; ++j) tr[j] = j;
      break;
    }
    if (valid)
      for (int i = k + 1; i < n; ++i) m[i][k] /= akk;
  }
  float d[6];
  for (int i = 0; i < n; ++i) d[i] = rhs[i];
  for (int k = 0; k < n; ++k) std::swap(d[k], d[tr[k]]);  // P
  for (int i = 0; i < n; ++i)                            // L^-1
    for (int j = 0; j < i; ++j) d[i] -= m[i][j] * d[j];
  const float tol = FLT_MIN;  // Eigen 3.4: numeric_limits<float>::min()
  for (int i = 0; i < n; ++i) d[i] = (std::fabs(m[i][i]) > tol) ? d[i] / m[i][i] : 0.f;
  for (int i = n - 1; i >= 0; --i)  // L^-T
    for (int j = i + 1; j < n; ++j) d[i] -= m[j][i] * d[j];
  for (int k = n - 1; k >= 0; --k) std::swap(d[k], d[tr[k]]);  // P^T
  for (int i = 0; i < n; ++i) x[i] = d[i];
}

inline void mat3_mul(const float A[9], const float B[9], float C[9]) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      C[3 * i + j] = dot3(A[3 * i], B[j], A[3 * i + 1], B[3 + j], A[3 * i + 2], B[6 + j]);
}

// reference: src/defs.h:100-136 (Rx,Ry,Rz,v2tEuler) + Isometry3f product (Appendix A.2)
void pose_update(const float dx[6], float T[12]) {
  float cx = std::cos(dx[3]), sx = std::sin(dx[3]);
  float cy = std::cos(dx[4]), sy = std::sin(dx[4]);
  float cz = std::cos(dx[5]), sz = std::sin(dx[5]);
  float Rx[9] = {1, 0, 0, 0, cx, -sx, 0, sx, cx};
  float Ry[9] = {cy, 0, sy, 0, 1, 0, -sy, 0, cy};
  float Rz[9] = {cz, -sz, 0, sz, cz, 0, 0, 0, 1};
  float Rxy[9], Rd[9];
  mat3_mul(Rx, Ry, Rxy);
  mat3_mul(Rxy, Rz, Rd);
  float D[12] = {Rd[0], Rd[1], Rd[2], dx[0], Rd[3], Rd[4], Rd[5], dx[1], Rd[6], Rd[7], Rd[8], dx[2]};
  float out[12];
  vo_ref_pose_mul(D, T, out);
  std::memcpy(T, out, sizeof(out));
}

void solve_and_update(Lin32& L, float damping, float T[12]) {
  for (int i = 0; i < 6; ++i) L.H[7 * i] += 1.f * damping;  // H += I*damping (picp_solver.cpp:96)
  float nb[6], dx[6];
  for (int i = 0; i < 6; ++i) nb[i] = -L.b[i];
  ldlt_solve6(L.H, nb, dx);
  pose_update(dx, T);
}

// ---------------------------------------------------------------- double LA
// One-sided (Hestenes) Jacobi SVD: A is m x n row-major, columns get
// orthogonalised in place (A <- U*Sigma); V (n x n row-major) accumulates the
// right singular vectors in its columns; w = singular values. Sorted descending.
void jacobi_svd(double* A, int m, int n, double* V, double* w) {
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) V[i * n + j] = (i == j);
  const double eps = 4 * DBL_EPSILON;
  for (int sweep = 0; sweep < 80; ++sweep) {
    bool changed = false;
    for (int i = 0; i < n - 1; ++i)
      for (int j = i + 1; j < n; ++j) {
        double a = 0, b = 0, p = 0;
        for (int k = 0; k < m; ++k) {
          double x = A[k * n + i], y = A[k * n + j];
          a += x * x;
          b += y * y;
          p += x * y;
        }
        if (std::fabs(p) <= eps * std::sqrt(a * b) || p == 0) continue;
        changed = true;
        double zeta = (b - a) / (2 * p);
        double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1 + zeta * zeta));
        double c = 1 / std::sqrt(1 + t * t), s = c * t;
        for (int k = 0; k < m; ++k) {
          double x = A[k * n + i], y = A[k * n + j];
          A[k * n + i] = c * x - s * y;
          A[k * n + j] = s * x + c * y;
        }
        for (int k = 0; k < n; ++k) {
          double x = V[k * n + i], y = V[k * n + j];
          V[k * n + i] = c * x - s * y;
          V[k * n + j] = s * x + c * y;
        }
      }
    if (!changed) break;
  }
  for (int i = 0; i < n; ++i) {
    double s = 0;
    for (int k = 0; k < m; ++k) s += A[k * n + i] * A[k * n + i];
    w[i] = std::sqrt(s);
  }
  for (int i = 0; i < n - 1; ++i) {  // selection sort, descending
    int j = i;
    for (int k = i + 1; k < n; ++k)
      if (w[k] > w[j]) j = k;
    if (j != i) {
      std::swap(w[i], w[j]);
      for (int k = 0; k < m; ++k) std::swap(A[k * n + i], A[k * n + j]);
      for (int k = 0; k < n; ++k) std::swap(V[k * n + i], V[k * n + j]);
    }
  }
}

// full 3x3 SVD E = U diag(w) V^T with U,V orthogonal (third column of U completed)
void svd3(const double E[9], double U[9], double w[3], double V[9]) {
  double A[9];
  std::memcpy(A, E, sizeof(A));
  jacobi_svd(A, 3, 3, V, w);
  double u[3][3];
  for (int j = 0; j < 2; ++j)
    for (int k = 0; k < 3; ++k) u[j][k] = (w[j] > 0) ? A[k * 3 + j] / w[j] : 0;
  u[2][0] = u[0][1] * u[1][2] - u[0][2] * u[1][1];
  u[2][1] = u[0][2] * u[1][0] - u[0][0] * u[1][2];
  u[2][2] = u[0][0] * u[1][1] - u[0][1] * u[1][0];
  for (int j = 0; j < 3; ++j)
    for (int k = 0; k < 3; ++k) U[k * 3 + j] = u[j][k];
}

inline double det3(const double M[9]) {
  return M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) +
         M[2] * (M[3] * M[7] - M[4] * M[6]);
}

inline void mm3(const double A[9], const double B[9], double C[9]) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += A[3 * i + k] * B[3 * k + j];
      C[3 * i + j] = s;
    }
}

// OpenCV triangulatePoints (DLT): rows x*P[2]-P[0], y*P[2]-P[1] per view; X =
// right singular vector of the smallest singular value of the 4x4 system.
void dlt_point(const double P1[12], const double P2[12], double x1, double y1, double x2, double y2,
               double X[4]) {
  double A[16], V[16], w[4];
  for (int k = 0; k < 4; ++k) {
    A[0 * 4 + k] = x1 * P1[8 + k] - P1[k];
    A[1 * 4 + k] = y1 * P1[8 + k] - P1[4 + k];
    A[2 * 4 + k] = x2 * P2[8 + k] - P2[k];
    A[3 * 4 + k] = y2 * P2[8 + k] - P2[4 + k];
  }
  jacobi_svd(A, 4, 4, V, w);
  for (int k = 0; k < 4; ++k) X[k] = V[k * 4 + 3];
}

// OpenCV decomposeEssentialMat + recoverPose (calib3d/five-point.cpp), restated.
int recover_pose(const double E[9], const float K[9], const float* x1, const float* x2, int64_t n,
                 double R[9], double t[3], uint8_t* mask_out) {
  const double fx = K[0], fy = K[4], cx = K[2], cy = K[5];
  const double dist_thr = 50.0;
  double U[9], w[3], Vt[9];
  vo_ref_cv_svd3(E, U, w, Vt);  // decomposeEssentialMat: SVD::compute(E, D, U, Vt) with OpenCV's own Jacobi SVD
  if (det3(U) < 0)
    for (double& v : U) v = -v;
  if (det3(Vt) < 0)
    for (double& v : Vt) v = -v;
  const double Wm[9] = {0, 1, 0, -1, 0, 0, 0, 0, 1};
  const double Wt[9] = {0, -1, 0, 1, 0, 0, 0, 0, 1};
  double UW[9], R1[9], R2[9], tt[3] = {U[2], U[5], U[8]};
  mm3(U, Wm, UW);
  mm3(UW, Vt, R1);
  mm3(U, Wt, UW);
  mm3(UW, Vt, R2);
  const double* Rs[4] = {R1, R2, R1, R2};
  const double sg[4] = {1, 1, -1, -1};
  std::vector<uint8_t> masks[4];
  int64_t good[4] = {0, 0, 0, 0};
  const double P0[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
  for (int cnd = 0; cnd < 4; ++cnd) {
    double P[12];
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) P[4 * i + j] = Rs[cnd][3 * i + j];
      P[4 * i + 3] = sg[cnd] * tt[i];
    }
    masks[cnd].assign(n, 0);
    for (int64_t i = 0; i < n; ++i) {
      double a1 = ((double)x1[2 * i] - cx) / fx, b1 = ((double)x1[2 * i + 1] - cy) / fy;
      double a2 = ((double)x2[2 * i] - cx) / fx, b2 = ((double)x2[2 * i + 1] - cy) / fy;
      double Q[4];
      dlt_point(P0, P, a1, b1, a2, b2, Q);
      bool ok = Q[2] * Q[3] > 0;
      double q0 = Q[0] / Q[3], q1 = Q[1] / Q[3], q2 = Q[2] / Q[3];
      ok = ok && (q2 < dist_thr);
      double z2 = P[8] * q0 + P[9] * q1 + P[10] * q2 + P[11];
      ok = ok && (z2 > 0) && (z2 < dist_thr);
      masks[cnd][i] = ok ? 255 : 0;
      good[cnd] += ok;
    }
  }
  int pick;
  if (good[0] >= good[1] && good[0] >= good[2] && good[0] >= good[3]) pick = 0;
  else if (good[1] >= good[0] && good[1] >= good[2] && good[1] >= good[3]) pick = 1;
  else if (good[2] >= good[0] && good[2] >= good[1] && good[2] >= good[3]) pick = 2;
  else pick = 3;
  std::memcpy(R, Rs[pick], 9 * sizeof(double));
  for (int i = 0; i < 3; ++i) t[i] = sg[pick] * tt[i];
  if (mask_out) std::memcpy(mask_out, masks[pick].data(), n);
  return (int)good[pick];
}

// descriptor distance (my_utilities.h:92) in Eigen's SSE redux order (Appendix A.7)
inline float sqdist_eigen(const float* a, const float* b, int dim) {
  if (dim < 4) {
    float r = 0;
    for (int k = 0; k < dim; ++k) {
      float d = a[k] - b[k];
      float x = d * d;
      r = (k == 0) ? x : r + x;
    }
    return r;
  }
  const int n4 = dim / 4 * 4, n8 = dim / 8 * 8;
  float p0[4], p1[4];
  auto sq = [&](int k) {
    float d = a[k] - b[k];
    return d * d;
  };
  for (int l = 0; l < 4; ++l) p0[l] = sq(l);
  if (n4 > 4) {
    for (int l = 0; l < 4; ++l) p1[l] = sq(4 + l);
    for (int i = 8; i < n8; i += 8)
      for (int l = 0; l < 4; ++l) {
        p0[l] = p0[l] + sq(i + l);
        p1[l] = p1[l] + sq(i + 4 + l);
      }
    for (int l = 0; l < 4; ++l) p0[l] = p0[l] + p1[l];
    if (n4 > n8)
      for (int l = 0; l < 4; ++l) p0[l] = p0[l] + sq(n8 + l);
  }
  float r = (p0[0] + p0[2]) + (p0[1] + p0[3]);  // SSE2 predux
  for (int k = n4; k < dim; ++k) r = r + sq(k);
  return r;
}

inline float sqdist_plain(const float* a, const float* b, int dim) {
  float r = 0;
  for (int k = 0; k < dim; ++k) {
    float d = a[k] - b[k];
    float x = d * d;
    r = (k == 0) ? x : r + x;
  }
  return r;
}

template <class F>
void parallel_ranges(int64_t lo, int64_t hi, int n_threads, F f) {
  if (n_threads <= 1 || hi - lo < 2) {
    f(0, lo, hi);
    return;
  }
  std::vector<std::thread> th;
  int64_t per = (hi - lo + n_threads - 1) / n_threads;
  for (int t = 0; t < n_threads; ++t) {
    int64_t a = lo + t * per, b = std::min(hi, a + per);
    if (a >= b) break;
    th.emplace_back([=] { f(t, a, b); });
  }
  for (auto& x : th) x.join();
}

}  // namespace

extern "C" {

int vo_ref_num_threads(void) {
  unsigned n = std::thread::hardware_concurrency();
  return n ? (int)n : 1;
}

int vo_ref_project_point(const float K[9], int rows, int cols, const float pose[12], const float p[3],
                         float uv[2]) {
  float c[3], q[3];
  return project_full(K, rows, cols, pose, p, c, q, uv) ? 1 : 0;
}

// reference: src/camera.cpp:14-35
int vo_ref_project_points(const float K[9], int rows, int cols, const float pose[12],
                          const float* world_xyz, int n, int keep_indices, float* out_uv, int* n_out) {
  int num = 0, inside_n = 0;
  for (int i = 0; i < n; ++i) {
    float uv[2];
    bool inside = vo_ref_project_point(K, rows, cols, pose, world_xyz + 3 * (int64_t)i, uv);
    if (inside) inside_n++;
    else uv[0] = uv[1] = -1.f;
    if (keep_indices || inside) {
      out_uv[2 * (int64_t)num] = uv[0];
      out_uv[2 * (int64_t)num + 1] = uv[1];
      num++;
    }
  }
  if (n_out) *n_out = num;
  return inside_n;
}

int vo_ref_error_jacobian(const float K[9], int rows, int cols, const float pose[12], const float p[3],
                          const float z[2], float e[2], float J[12]) {
  return error_jacobian(K, rows, cols, pose, p, z, e, J) ? 1 : 0;
}

void vo_ref_linearize(const float K[9], int rows, int cols, const float pose[12], const float* world_xyz,
                      const float* image_xy, const int32_t* pairs, int64_t n_pairs, float thr,
                      int keep_outliers, int accum_mode, double H[36], double b[6], double* chi_in,
                      double* chi_out, int64_t* n_inliers, uint8_t* status) {
  if (accum_mode == 0) {
    Lin32 L;
    linearize32(K, rows, cols, pose, world_xyz, image_xy, pairs, 0, n_pairs, thr, keep_outliers != 0, L,
                status);
    for (int i = 0; i < 36; ++i) H[i] = L.H[i];
    for (int i = 0; i < 6; ++i) b[i] = L.b[i];
    *chi_in = L.chi_in;
    *chi_out = L.chi_out;
    *n_inliers = L.n_in;
  } else {
    linearize64(K, rows, cols, pose, world_xyz, image_xy, pairs, n_pairs, thr, keep_outliers != 0, H, b,
                *chi_in, *chi_out, *n_inliers, status);
  }
}

void vo_ref_ldlt_solve6(const float A[36], const float rhs[6], float x[6]) { ldlt_solve6(A, rhs, x); }

void vo_ref_pose_update(const float dx[6], float pose[12]) { pose_update(dx, pose); }

void vo_ref_one_round(const float K[9], int rows, int cols, float pose[12], const float* world_xyz,
                      const float* image_xy, const int32_t* pairs, int64_t n_pairs, float thr,
                      float damping, int keep_outliers, float* chi_in, float* chi_out, int* n_inliers) {
  Lin32 L;
  linearize32(K, rows, cols, pose, world_xyz, image_xy, pairs, 0, n_pairs, thr, keep_outliers != 0, L,
              nullptr);
  if (chi_in) *chi_in = L.chi_in;
  if (chi_out) *chi_out = L.chi_out;
  if (n_inliers) *n_inliers = (int)L.n_in;
  solve_and_update(L, damping, pose);
}

void vo_ref_one_round_mt(const float K[9], int rows, int cols, float pose[12], const float* world_xyz,
                         const float* image_xy, const int32_t* pairs, int64_t n_pairs, float thr,
                         float damping, int keep_outliers, int n_threads, float* chi_in, float* chi_out,
                         int* n_inliers) {
  if (n_threads < 1) n_threads = 1;
  std::vector<Lin32> parts(n_threads);
  for (auto& p : parts) std::memset(&p, 0, sizeof(p));
  parallel_ranges(0, n_pairs, n_threads, [&](int t, int64_t lo, int64_t hi) {
    linearize32(K, rows, cols, pose, world_xyz, image_xy, pairs, lo, hi, thr, keep_outliers != 0,
                parts[t], nullptr);
  });
  Lin32 L = parts[0];
  for (int t = 1; t < n_threads; ++t) {
    for (int i = 0; i < 36; ++i) L.H[i] += parts[t].H[i];
    for (int i = 0; i < 6; ++i) L.b[i] += parts[t].b[i];
    L.chi_in += parts[t].chi_in;
    L.chi_out += parts[t].chi_out;
    L.n_in += parts[t].n_in;
  }
  if (chi_in) *chi_in = L.chi_in;
  if (chi_out) *chi_out = L.chi_out;
  if (n_inliers) *n_inliers = (int)L.n_in;
  solve_and_update(L, damping, pose);
}

// reference: src/my_utilities.h:70-120 (match_points)
int64_t vo_ref_match(const float* descA, int64_t n1, const float* descB, int64_t n2, int dim,
                     float dist_thr, float ratio_thr, const int32_t* idA, const int32_t* idB,
                     int64_t row_begin, int64_t row_end, int order_mode, int n_threads,
                     int32_t* pairs_out, int64_t stats[2], float* best_o, float* second_o,
                     int32_t* idx_o) {
  (void)n1;
  const int64_t rows = row_end - row_begin;
  std::vector<int32_t> bidx(rows);
  std::vector<uint8_t> acc(rows);
  if (n_threads < 1) n_threads = 1;
  std::vector<int64_t> possible(n_threads, 0);
  parallel_ranges(row_begin, row_end, n_threads, [&](int t, int64_t lo, int64_t hi) {
    int64_t poss = 0;
    for (int64_t i = lo; i < hi; ++i) {
      const float* a = descA + i * dim;
      float best = FLT_MAX, second = FLT_MAX;
      int bi = -1;
      for (int64_t j = 0; j < n2; ++j) {
        if (idA && idB && idA[i] == idB[j]) poss++;
        float d = order_mode == 0 ? sqdist_eigen(a, descB + j * dim, dim)
                                  : sqdist_plain(a, descB + j * dim, dim);
        if (d < best) {
          second = best;
          best = d;
          bi = (int)j;
        } else if (d < second) {
          second = d;
        }
      }
      bool ok = bi != -1 && best < dist_thr && best / second < ratio_thr;
      bidx[i - row_begin] = bi;
      acc[i - row_begin] = ok;
      if (best_o) best_o[i - row_begin] = best;
      if (second_o) second_o[i - row_begin] = second;
      if (idx_o) idx_o[i - row_begin] = bi;
    }
    possible[t] = poss;
  });
  int64_t cnt = 0, correct = 0, poss = 0;
  for (int t = 0; t < n_threads; ++t) poss += possible[t];
  for (int64_t r = 0; r < rows; ++r)
    if (acc[r]) {
      if (pairs_out) {
        pairs_out[2 * cnt] = (int32_t)(row_begin + r);
        pairs_out[2 * cnt + 1] = bidx[r];
      }
      if (idA && idB && idA[row_begin + r] == idB[bidx[r]]) correct++;
      cnt++;
    }
  if (stats) {
    stats[0] = poss;
    stats[1] = correct;
  }
  return cnt;
}

// Eigen Isometry3f::inverse() (Appendix A.2): R^T, -(R^T) t
void vo_ref_pose_inverse(const float T[12], float out[12]) {
  float Rt[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) Rt[3 * i + j] = T[4 * j + i];
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) out[4 * i + j] = Rt[3 * i + j];
    out[4 * i + 3] = dot3(-Rt[3 * i], T[3], -Rt[3 * i + 1], T[7], -Rt[3 * i + 2], T[11]);
  }
}

// Isometry3f * Isometry3f (Appendix A.2): R = Ra Rb ; t = Ra tb + ta
void vo_ref_pose_mul(const float A[12], const float B[12], float out[12]) {
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j)
      out[4 * i + j] = dot3(A[4 * i], B[j], A[4 * i + 1], B[4 + j], A[4 * i + 2], B[8 + j]);
    out[4 * i + 3] = dot3(A[4 * i], B[3], A[4 * i + 1], B[7], A[4 * i + 2], B[11]) + A[4 * i + 3];
  }
}

// reference: src/cam.cpp:94-140. P = K * T^-1[0:3,:] as a float32 cv::Mat product
// (cv::gemm accumulates float products in double, then rounds to float).
void vo_ref_triangulate(const float K[9], const float T1[12], const float T2[12], const float* x1,
                        const float* x2, int64_t n, float* xyz_out) {
  float Ti[2][12];
  vo_ref_pose_inverse(T1, Ti[0]);
  vo_ref_pose_inverse(T2, Ti[1]);
  double P[2][12];
  for (int v = 0; v < 2; ++v)
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 4; ++j) {
        double s = 0;
        for (int k = 0; k < 3; ++k) s += (double)K[3 * i + k] * (double)Ti[v][4 * k + j];
        P[v][4 * i + j] = (double)(float)s;
      }
  for (int64_t i = 0; i < n; ++i) {
    double X[4];
    dlt_point(P[0], P[1], x1[2 * i], x1[2 * i + 1], x2[2 * i], x2[2 * i + 1], X);
    float Xf[4] = {(float)X[0], (float)X[1], (float)X[2], (float)X[3]};
    float scale = Xf[3] != 0.f ? 1.f / Xf[3] : 1.f;  // convertPointsFromHomogeneous
    xyz_out[3 * i] = Xf[0] * scale;
    xyz_out[3 * i + 1] = Xf[1] * scale;
    xyz_out[3 * i + 2] = Xf[2] * scale;
  }
}

void vo_ref_jacobi_svd(double* A, int m, int n, double* V, double* w) { jacobi_svd(A, m, n, V, w); }

int vo_ref_recover_pose(const double E[9], const float K[9], const float* x1, const float* x2, int64_t n,
                        double R[9], double t[3], uint8_t* mask) {
  return recover_pose(E, K, x1, x2, n, R, t, mask);
}

// Essential matrix by the normalised 8-point algorithm on all matches (RMS-isotropic
// Hartley normalisation, SVD null vector, projection onto the essential manifold),
// then recoverPose.  The reference calls cv::findEssentialMat(RANSAC) here
// (cam.cpp:49) — a minimal-sample 5-point hypothesis; see DESIGN.md for why the
// all-inlier linear estimator is the B200 design and what tolerance that implies.
int vo_ref_essential_recover(const float K[9], const float* x1, const float* x2, int64_t n, double E[9],
                             double R[9], double t[3], uint8_t* mask) {
  if (n < 8) return -1;
  const double fx = K[0], fy = K[4], cx = K[2], cy = K[5];
  std::vector<double> a(2 * n), b(2 * n);
  double m1[2] = {0, 0}, m2[2] = {0, 0};
  for (int64_t i = 0; i < n; ++i) {
    a[2 * i] = ((double)x1[2 * i] - cx) / fx;
    a[2 * i + 1] = ((double)x1[2 * i + 1] - cy) / fy;
    b[2 * i] = ((double)x2[2 * i] - cx) / fx;
    b[2 * i + 1] = ((double)x2[2 * i + 1] - cy) / fy;
    m1[0] += a[2 * i];
    m1[1] += a[2 * i + 1];
    m2[0] += b[2 * i];
    m2[1] += b[2 * i + 1];
  }
  for (int k = 0; k < 2; ++k) {
    m1[k] /= (double)n;
    m2[k] /= (double)n;
  }
  double v1 = 0, v2 = 0;
  for (int64_t i = 0; i < n; ++i) {
    v1 += (a[2 * i] - m1[0]) * (a[2 * i] - m1[0]) + (a[2 * i + 1] - m1[1]) * (a[2 * i + 1] - m1[1]);
    v2 += (b[2 * i] - m2[0]) * (b[2 * i] - m2[0]) + (b[2 * i + 1] - m2[1]) * (b[2 * i + 1] - m2[1]);
  }
  double s1 = std::sqrt(2.0 / (v1 / (double)n)), s2 = std::sqrt(2.0 / (v2 / (double)n));
  std::vector<double> A(9 * n);
  for (int64_t i = 0; i < n; ++i) {
    double px = s1 * (a[2 * i] - m1[0]), py = s1 * (a[2 * i + 1] - m1[1]);
    double qx = s2 * (b[2 * i] - m2[0]), qy = s2 * (b[2 * i + 1] - m2[1]);
    double* r = &A[9 * i];
    r[0] = qx * px; r[1] = qx * py; r[2] = qx;
    r[3] = qy * px; r[4] = qy * py; r[5] = qy;
    r[6] = px;      r[7] = py;      r[8] = 1;
  }
  double V[81], w[9];
  jacobi_svd(A.data(), (int)n, 9, V, w);
  double Fh[9];
  for (int k = 0; k < 9; ++k) Fh[k] = V[k * 9 + 8];
  // E = T2^T Fh T1, Ti = [s 0 -s mx; 0 s -s my; 0 0 1]
  double T1m[9] = {s1, 0, -s1 * m1[0], 0, s1, -s1 * m1[1], 0, 0, 1};
  double T2t[9] = {s2, 0, 0, 0, s2, 0, -s2 * m2[0], -s2 * m2[1], 1};
  double tmp[9], E0[9];
  mm3(T2t, Fh, tmp);
  mm3(tmp, T1m, E0);
  double U[9], sw[3], Vv[9];
  svd3(E0, U, sw, Vv);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) E[3 * i + j] = U[3 * i] * Vv[3 * j] + U[3 * i + 1] * Vv[3 * j + 1];
  int big = 0;
  for (int k = 1; k < 9; ++k)
    if (std::fabs(E[k]) > std::fabs(E[big])) big = k;
  if (E[big] < 0)
    for (int k = 0; k < 9; ++k) E[k] = -E[k];
  return recover_pose(E, K, x1, x2, n, R, t, mask);
}

// reference: src/my_utilities.cpp:413-434
int64_t vo_ref_anti_join(const int32_t* matched_id_meas, int64_t n_matched,
                         const int32_t* cand_second_id_meas, int64_t n_cand, uint8_t* keep) {
  int64_t cnt = 0;
  for (int64_t j = 0; j < n_cand; ++j) {
    bool found = false;
    for (int64_t i = 0; i < n_matched; ++i)
      if (cand_second_id_meas[j] == matched_id_meas[i]) {
        found = true;
        break;
      }
    keep[j] = !found;
    cnt += !found;
  }
  return cnt;
}

}  // extern "C"
