// five_point.cuh — cv::findEssentialMat(RANSAC)'s hypothesis generator on the device (reference src/cam.cpp:49).
//
// The reference delegates to OpenCV (un-vendored, CMakeLists.txt:12): five-point.cpp EMEstimatorCallback::runKernel
// (Nister's five-point solver) inside ptsetreg.cpp's RANSAC loop.  Which hypothesis wins depends on details outside
// the textbook algorithm - the sampling order of cv::RNG, the null-space basis cv::SVD's FULL_UV completion happens
// to produce (it fixes the parametrisation E = x E0 + y E1 + z E2 + E3 and so the polynomial in z), and the order
// in which cv::solvePoly's Durand-Kerner iteration delivers the roots (the FIRST best hypothesis wins ties) - so
// those are restated as well.  One thread solves one minimal sample in double precision; the RANSAC loop around it
// (csrc/essential5.cu) evaluates batches of samples in parallel and replays OpenCV's sequential bookkeeping over
// the per-hypothesis inlier counts.
#pragma once
#include "vo_device.cuh"

namespace {

// RANSACPointSetRegistrator::getSubset (ptsetreg.cpp): five distinct indices, a duplicate is simply redrawn
__host__ __device__ inline void ransac_next_subset(CvRng& rng, int n, int idx[5]) {
  for (int i = 0; i < 5; ++i) {
    int v;
    bool dup;
    do {
      v = rng.uniform(0, n);
      dup = false;
      for (int j = 0; j < i; ++j) dup = dup || (idx[j] == v);
    } while (dup);
    idx[i] = v;
  }
}

// RANSACUpdateNumIters (ptsetreg.cpp)
__host__ __device__ inline int ransac_update_num_iters(double p, double ep, int model_points, int max_iters) {
  p = fmin(fmax(p, 0.), 1.);
  ep = fmin(fmax(ep, 0.), 1.);
  double num = fmax(1. - p, DBL_MIN);
  double denom = 1. - pow(1. - ep, (double)model_points);
  if (denom < DBL_MIN) return 0;
  num = log(num);
  denom = log(denom);
  return denom >= 0 || -num >= max_iters * (-denom) ? max_iters : (int)rint(num / denom);  // cvRound
}

// EMEstimatorCallback::computeError: Sampson distance in double, stored as float
__device__ __forceinline__ float sampson_dev(const double* E, double x1, double y1, double x2, double y2) {
  const double e0 = E[0] * x1 + E[1] * y1 + E[2], e1 = E[3] * x1 + E[4] * y1 + E[5], e2 = E[6] * x1 + E[7] * y1 + E[8];
  const double t0 = E[0] * x2 + E[3] * y2 + E[6], t1 = E[1] * x2 + E[4] * y2 + E[7];
  const double x2tEx1 = x2 * e0 + y2 * e1 + e2;
  return (float)(x2tEx1 * x2tEx1 / (e0 * e0 + e1 * e1 + t0 * t0 + t1 * t1));
}

// rows 5..8 of Vt of cv::SVD::compute(Q 5x9, MODIFY_A | FULL_UV): m < n, so OpenCV factors the transpose - the five
// rows of Q are orthogonalised and completed to nine orthonormal rows (cv_jacobi_svd_dev, vo_device.cuh)
__device__ void cv_null_basis_dev(double (&At)[9][9]) {  // in: rows 0..4 = Q; out: rows 5..8 = the basis
  double W[9], Vt[5][5];
  cv_jacobi_svd_dev(&At[0][0], 9, W, &Vt[0][0], 5, 9, 5, 9);
}

// cv::solvePoly (mathfuncs.cpp): Durand-Kerner from the start values (1+i)^k with in-place updates.  OpenCV always
// runs 300 sweeps (its exit test is an exact zero step).  The iteration converges quadratically: it needs 15-30 sweeps
// to bring the largest step below 1e-9 of the root magnitude and ONE more to reach the rounding floor, where it then
// jitters for the remaining ~270 sweeps (measured on the fixtures, exp/five_point_proto.py).  Here it stops two sweeps
// after the first sweep below 1e-9: the roots equal the 300-sweep ones to rounding, and their ORDER - which is what
// the RANSAC tie-break depends on - was settled long before.
// c[k] = coefficient of z^k.  Returns the number of roots.
__device__ int cv_solve_poly_dev(const double* c, int deg, double* re, double* im) {
  int n = deg;
  for (; n > 1; --n)
    if (fabs(c[n]) > DBL_EPSILON) break;
  {
    double pr = 1, pi = 0;
    for (int i = 0; i < n; ++i) {
      re[i] = pr;
      im[i] = pi;
      const double nr = pr - pi, ni = pr + pi;  // * (1 + i)
      pr = nr;
      pi = ni;
    }
  }
  int settle = 1 << 30;
  for (int iter = 0; iter < 300 && iter <= settle; ++iter) {
    double max_diff = 0, max_abs = 0;
    for (int i = 0; i < n; ++i) {
      const double pr = re[i], pi = im[i];
      double nr = c[n], ni = 0, dr = c[n], di = 0;
      for (int j = 0; j < n; ++j) {
        const double tr = nr * pr - ni * pi + c[n - j - 1], ti = nr * pi + ni * pr;
        nr = tr;
        ni = ti;
        if (j != i) {
          const double qr = pr - re[j], qi = pi - im[j];
          if (qr != 0 || qi != 0) {
            const double ur = dr * qr - di * qi, ui = dr * qi + di * qr;
            dr = ur;
            di = ui;
          }
        }
      }
      // num /= denom (std::complex division as cv::Complex does it: straightforward formula)
      const double den = dr * dr + di * di;
      const double qr = (nr * dr + ni * di) / den, qi = (ni * dr - nr * di) / den;
      re[i] = pr - qr;
      im[i] = pi - qi;
      max_diff = fmax(max_diff, hypot(qr, qi));
      max_abs = fmax(max_abs, fabs(re[i]) + fabs(im[i]));
    }
    if (max_diff <= 0) break;
    if (settle == (1 << 30) && max_diff <= 1e-9 * fmax(max_abs, 1e-300)) settle = iter + 2;
  }
  return n;
}

// column of the monomial x^i y^j z^k in Nister's elimination order (the first ten are eliminated, the last ten are
// {x, y, 1} times powers of z): x3 y3 x2y xy2 x2z x2 y2z y2 xyz xy | xz2 xz x yz2 yz y z3 z2 z 1
__device__ __forceinline__ int mono_col(int i, int j, int k) {
  // packed lookup indexed by 16 i + 4 j + k
  switch (16 * i + 4 * j + k) {
    case 48: return 0;  case 12: return 1;  case 36: return 2;  case 24: return 3;  case 33: return 4;
    case 32: return 5;  case 9: return 6;   case 8: return 7;   case 21: return 8;  case 20: return 9;
    case 18: return 10; case 17: return 11; case 16: return 12; case 6: return 13;  case 5: return 14;
    case 4: return 15;  case 3: return 16;  case 2: return 17;  case 1: return 18;  default: return 19;
  }
}

// EMEstimatorCallback::runKernel (five-point.cpp): five normalised correspondences -> up to 10 essential matrices
// (row-major, unit Frobenius norm, x2^T E x1 = 0) in OpenCV's order.  q1, q2: x,y interleaved.
__device__ int five_point_dev(const double* q1, const double* q2, double* E_out /* [10][9] */) {
  double At[9][9];
  for (int i = 0; i < 9; ++i)
    for (int k = 0; k < 9; ++k) At[i][k] = 0;
  for (int i = 0; i < 5; ++i) {
    const double x1 = q1[2 * i], y1 = q1[2 * i + 1], x2 = q2[2 * i], y2 = q2[2 * i + 1];
    At[i][0] = x2 * x1; At[i][1] = x2 * y1; At[i][2] = x2;
    At[i][3] = y2 * x1; At[i][4] = y2 * y1; At[i][5] = y2;
    At[i][6] = x1;      At[i][7] = y1;      At[i][8] = 1.0;
  }
  cv_null_basis_dev(At);
  const double(*EE)[9] = &At[5];  // E(x, y, z) = x EE[0] + y EE[1] + z EE[2] + EE[3]
  // the ten cubic constraints det E = 0 and 2 E E^T E - tr(E E^T) E = 0, expanded over the 64 ordered triples
  // (a, b, c) of basis matrices: triple (a,b,c) contributes to the monomial v_a v_b v_c, v = (x, y, z, 1)
  double A[10][20];
  for (int r = 0; r < 10; ++r)
    for (int m = 0; m < 20; ++m) A[r][m] = 0;
  for (int a = 0; a < 4; ++a)
    for (int b = 0; b < 4; ++b) {
      double G[9];  // E_a E_b^T
      double tr = 0;
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
          double s = 0;
          for (int k = 0; k < 3; ++k) s += EE[a][3 * i + k] * EE[b][3 * j + k];
          G[3 * i + j] = s;
        }
      tr = G[0] + G[4] + G[8];
      for (int c = 0; c < 4; ++c) {
        int ex[4] = {0, 0, 0, 0};
        ex[a]++; ex[b]++; ex[c]++;
        const int col = mono_col(ex[0], ex[1], ex[2]);
        const double* Ec = EE[c];
        // det: E_a row 0 . (E_b row 1 x E_c row 2)
        const double* r0 = EE[a];
        const double* r1 = EE[b] + 3;
        const double* r2 = Ec + 6;
        A[0][col] += r0[0] * (r1[1] * r2[2] - r1[2] * r2[1]) - r0[1] * (r1[0] * r2[2] - r1[2] * r2[0]) +
                     r0[2] * (r1[0] * r2[1] - r1[1] * r2[0]);
        for (int i = 0; i < 3; ++i)
          for (int j = 0; j < 3; ++j) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += G[3 * i + k] * Ec[3 * k + j];
            A[1 + 3 * i + j][col] += 2 * s - tr * Ec[3 * i + j];
          }
      }
    }
  // A <- A1^-1 A2 (Gauss-Jordan with partial pivoting on the first ten columns)
  for (int col = 0; col < 10; ++col) {
    int piv = col;
    for (int r = col + 1; r < 10; ++r)
      if (fabs(A[r][col]) > fabs(A[piv][col])) piv = r;
    if (fabs(A[piv][col]) < DBL_MIN) return 0;
    if (piv != col)
      for (int k = 0; k < 20; ++k) { const double t = A[piv][k]; A[piv][k] = A[col][k]; A[col][k] = t; }
    const double inv = 1.0 / A[col][col];
    for (int k = 0; k < 20; ++k) A[col][k] *= inv;
    for (int r = 0; r < 10; ++r) {
      if (r == col) continue;
      const double f = A[r][col];
      if (f == 0) continue;
      for (int k = 0; k < 20; ++k) A[r][k] -= f * A[col][k];
    }
  }
  // (row 2i+4) - z (row 2i+5): [cubic] x + [cubic] y + [quartic] = 0; P[i][0..1][k], P[i][2][k] = coefficient of z^k
  double P[3][3][5];
  for (int i = 0; i < 3; ++i) {
    const double* a1 = &A[2 * i + 4][10];
    const double* a2 = &A[2 * i + 5][10];
    double b[13];
    for (int k = 0; k < 13; ++k) b[k] = 0;
    for (int k = 0; k < 3; ++k) { b[1 + k] += a1[k]; b[5 + k] += a1[3 + k]; b[k] -= a2[k]; b[4 + k] -= a2[3 + k]; }
    for (int k = 0; k < 4; ++k) { b[9 + k] += a1[6 + k]; b[8 + k] -= a2[6 + k]; }
    for (int k = 0; k < 4; ++k) { P[i][0][3 - k] = b[k]; P[i][1][3 - k] = b[4 + k]; }
    P[i][0][4] = 0; P[i][1][4] = 0;
    for (int k = 0; k < 5; ++k) P[i][2][4 - k] = b[8 + k];
  }
  // det of the 3x3 polynomial matrix: degree 10
  double det[11];
  for (int k = 0; k <= 10; ++k) det[k] = 0;
  auto minor_acc = [&](int c0, int c1, int c2, double sign) {  // sign * P[0][c0] * (P[1][c1] P[2][c2] - P[1][c2] P[2][c1])
    double mn[9];
    for (int k = 0; k < 9; ++k) mn[k] = 0;
    for (int u = 0; u < 5; ++u)
      for (int v = 0; v < 5; ++v) mn[u + v] += P[1][c1][u] * P[2][c2][v] - P[1][c2][u] * P[2][c1][v];
    for (int u = 0; u < 5; ++u)
      for (int v = 0; v < 9; ++v)
        if (u + v <= 10) det[u + v] += sign * P[0][c0][u] * mn[v];
  };
  minor_acc(0, 1, 2, 1.0);
  minor_acc(1, 0, 2, -1.0);
  minor_acc(2, 0, 1, 1.0);
  double rr[10], ri[10];
  const int n_roots = cv_solve_poly_dev(det, 10, rr, ri);
  int count = 0;
  for (int i = 0; i < n_roots && count < 10; ++i) {
    if (fabs(ri[i]) > 1e-10) continue;
    const double z = rr[i];
    double Bz[3][3], V[3][3], w[3];
    for (int j = 0; j < 3; ++j)
      for (int k = 0; k < 3; ++k) {
        double v = 0;
        for (int d = 4; d >= 0; --d) v = v * z + P[j][k][d];
        Bz[j][k] = v;
      }
    jacobi_svd_dev<3, 3>(Bz, V, w);  // cv::SVD::solveZ: right singular vector of the smallest singular value
    if (fabs(V[2][2]) < 1e-10) continue;
    const double x = V[0][2] / V[2][2], y = V[1][2] / V[2][2];
    double nrm = 0;
    double* Eo = E_out + 9 * count;
    for (int k = 0; k < 9; ++k) {
      Eo[k] = EE[0][k] * x + EE[1][k] * y + EE[2][k] * z + EE[3][k];
      nrm += Eo[k] * Eo[k];
    }
    nrm = sqrt(nrm);
    for (int k = 0; k < 9; ++k) Eo[k] /= nrm;
    ++count;
  }
  return count;
}

}  // namespace
