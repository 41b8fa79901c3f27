// picp.cu — projective-ICP Gauss-Newton rounds on sm_100a.
//
// Replaces pr::PICPSolver::{init, linearize, errorAndJacobian, oneRound} (reference
// src/picp_solver.cpp:17-105) and the inlined Camera::projectPoint (src/camera.h:24-36).
//
// Data layout in HBM (per solver handle)
//   world_xyz  float[3*Nw]  AoS exactly as Vector3fVector           (caller's layout)
//   image_xy   float[2*Ni]  AoS exactly as Vector2fVector
//   pairs      int32[2*C]   AoS exactly as IntPairVector (first: image, second: world)
//   packed     float[5][Cp] SoA stream gathered ONCE per correspondence set by
//              picp_pack_kernel: wx, wy, wz, zu, zv; Cp = C rounded up to 4.  Every
//              Gauss-Newton round then streams 20 B/correspondence with LDG.128, each
//              thread owning 4 consecutive correspondences per step (float4 per plane).
//   partials   float[grid][32]  per-block sums, fixed slot order
//   result     double[32]   21 upper-triangular H terms, 6 b terms, chi_in, chi_out,
//              n_inliers, n_outliers (+1 pad): the unit all-reduced across GPUs
//   dev        PicpDev      pose, round counter, stats ring: the pose never leaves HBM
//                           between rounds
//
// Reduction is deterministic: per-thread float accumulators over a fixed grid-stride
// assignment -> warp shuffle tree -> shared-memory sum over warps in warp order ->
// per-block partial; the block that takes the last ticket sums the partials in block order
// in float64 (pass 2), then (single GPU) solves the damped 6x6 system and updates the pose
// in the same launch.  With a communicator, pass 2 stops at `result`, NCCL all-reduces the
// 32 doubles over NVLink, and a one-warp kernel solves identically on every rank.
//
// Rounding contract: everything that decides the inlier mask (camera point, projection,
// reciprocal, error, chi) uses explicit round-to-nearest intrinsics in the reference's
// evaluation order and is never contracted to FMA; J, H and b are tolerance-level and use FMA.
#include "vo_common.cuh"

#include <float.h>
#include <math.h>

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kSlots = 32;       // partial row: 0..20 H, 21..26 b, 27 chi_in, 28 chi_out, 29 n_in, 30 n_out
constexpr int kCtasPerSm = 2;

struct PicpCam {
  float K[9];
  float umax, vmax;  // cols-1, rows-1 as float (camera.h:31-34 compares float against int)
};

struct PicpDev {
  float pose[12];
  unsigned int ticket;
  int round;      // rounds executed since the last reset
  int stop;       // set by the device-side convergence test
  int bad_index;  // set by the pack kernel when a correspondence is out of range
  float prev_chi;
  float rel_tol;  // < 0: no convergence test
  vo_picp_stats stats[VO_PICP_MAX_ROUNDS];
};

struct LinArgs {
  const float* pk;
  long long n;
  long long stride;
  PicpCam cam;
  float thr;
  float damping;
  PicpDev* dev;
  float* partials;
  double* result;
  unsigned char* status;
  int fuse_solve;
};

__device__ __forceinline__ float dot3_rn(float a0, float b0, float a1, float b1, float a2, float b2) {
  return __fadd_rn(__fmul_rn(a0, b0), __fadd_rn(__fmul_rn(a1, b1), __fmul_rn(a2, b2)));
}

__device__ __forceinline__ bool finite_f(float x) { return fabsf(x) <= FLT_MAX; }

// One correspondence. Returns VO_PICP_* and accumulates into acc/n_in/n_out.
template <bool KEEP, bool PINHOLE>
__device__ __forceinline__ int picp_point(const PicpCam& cam, const float* __restrict__ T, float thr,
                                          float px, float py, float pz, float zu, float zv,
                                          float (&acc)[29], int& n_in, int& n_out) {
  // ---- exact part: Camera::projectPoint, then e and chi (camera.h:24-36, picp_solver.cpp:36,74)
  const float c0 = __fadd_rn(T[3], dot3_rn(T[0], px, T[1], py, T[2], pz));
  const float c1 = __fadd_rn(T[7], dot3_rn(T[4], px, T[5], py, T[6], pz));
  const float c2 = __fadd_rn(T[11], dot3_rn(T[8], px, T[9], py, T[10], pz));
  if (c2 <= 0.f) return VO_PICP_SKIPPED;
  float q0, q1, q2;
  if (PINHOLE && finite_f(c0) && finite_f(c1)) {
    // K = [fx 0 cx; 0 fy cy; 0 0 1]: the zero products vanish exactly for finite c
    q0 = __fadd_rn(__fmul_rn(cam.K[0], c0), __fmul_rn(cam.K[2], c2));
    q1 = __fadd_rn(__fmul_rn(cam.K[4], c1), __fmul_rn(cam.K[5], c2));
    q2 = c2;
  } else {
    q0 = dot3_rn(cam.K[0], c0, cam.K[1], c1, cam.K[2], c2);
    q1 = dot3_rn(cam.K[3], c0, cam.K[4], c1, cam.K[5], c2);
    q2 = dot3_rn(cam.K[6], c0, cam.K[7], c1, cam.K[8], c2);
  }
  const float iz = __frcp_rn(q2);  // == (float)(1./(double)q2): correctly rounded reciprocal
  const float u = __fmul_rn(q0, iz);
  const float v = __fmul_rn(q1, iz);
  if (u < 0.f || u > cam.umax) return VO_PICP_SKIPPED;
  if (v < 0.f || v > cam.vmax) return VO_PICP_SKIPPED;
  const float e0 = __fsub_rn(u, zu);
  const float e1 = __fsub_rn(v, zv);
  const float chi = __fadd_rn(__fmul_rn(e0, e0), __fmul_rn(e1, e1));
  float lambda = 1.f;
  int st;
  if (chi > thr) {
    acc[28] += chi;
    n_out++;
    if (!KEEP) return VO_PICP_OUTLIER;
    lambda = __fsqrt_rn(__fdiv_rn(thr, chi));
    st = VO_PICP_OUTLIER;
  } else {
    acc[27] += chi;
    n_in++;
    st = VO_PICP_INLIER;
  }
  // ---- tolerance part: J = (Jp*K)*[I | skew(-c)], H += lambda J^T J, b += lambda J^T e
  const float iz2 = iz * iz;
  const float m0 = -q0 * iz2, m1 = -q1 * iz2;
  if (PINHOLE) {
    const float a = iz * cam.K[0], d = iz * cam.K[4];
    const float g = fmaf(iz, cam.K[2], m0), h = fmaf(iz, cam.K[5], m1);
    const float j3 = g * c1, j4 = fmaf(a, c2, -g * c0), j5 = -a * c1;
    const float k3 = fmaf(h, c1, -d * c2), k4 = -h * c0, k5 = d * c0;
    const float as = KEEP ? a * lambda : a, ds = KEEP ? d * lambda : d;
    const float gs = KEEP ? g * lambda : g, hs = KEEP ? h * lambda : h;
    const float j3s = KEEP ? j3 * lambda : j3, j4s = KEEP ? j4 * lambda : j4, j5s = KEEP ? j5 * lambda : j5;
    const float k3s = KEEP ? k3 * lambda : k3, k4s = KEEP ? k4 * lambda : k4, k5s = KEEP ? k5 * lambda : k5;
    acc[0] = fmaf(as, a, acc[0]);  // H00 ; H01 (acc[1]) is structurally zero
    acc[2] = fmaf(as, g, acc[2]);
    acc[3] = fmaf(as, j3, acc[3]);
    acc[4] = fmaf(as, j4, acc[4]);
    acc[5] = fmaf(as, j5, acc[5]);
    acc[6] = fmaf(ds, d, acc[6]);  // H11
    acc[7] = fmaf(ds, h, acc[7]);
    acc[8] = fmaf(ds, k3, acc[8]);
    acc[9] = fmaf(ds, k4, acc[9]);
    acc[10] = fmaf(ds, k5, acc[10]);
    acc[11] = fmaf(gs, g, fmaf(hs, h, acc[11]));  // H22
    acc[12] = fmaf(gs, j3, fmaf(hs, k3, acc[12]));
    acc[13] = fmaf(gs, j4, fmaf(hs, k4, acc[13]));
    acc[14] = fmaf(gs, j5, fmaf(hs, k5, acc[14]));
    acc[15] = fmaf(j3s, j3, fmaf(k3s, k3, acc[15]));  // H33
    acc[16] = fmaf(j3s, j4, fmaf(k3s, k4, acc[16]));
    acc[17] = fmaf(j3s, j5, fmaf(k3s, k5, acc[17]));
    acc[18] = fmaf(j4s, j4, fmaf(k4s, k4, acc[18]));  // H44
    acc[19] = fmaf(j4s, j5, fmaf(k4s, k5, acc[19]));
    acc[20] = fmaf(j5s, j5, fmaf(k5s, k5, acc[20]));  // H55
    acc[21] = fmaf(as, e0, acc[21]);
    acc[22] = fmaf(ds, e1, acc[22]);
    acc[23] = fmaf(gs, e0, fmaf(hs, e1, acc[23]));
    acc[24] = fmaf(j3s, e0, fmaf(k3s, e1, acc[24]));
    acc[25] = fmaf(j4s, e0, fmaf(k4s, e1, acc[25]));
    acc[26] = fmaf(j5s, e0, fmaf(k5s, e1, acc[26]));
  } else {
    float J0[6], J1[6];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      J0[j] = fmaf(iz, cam.K[j], m0 * cam.K[6 + j]);
      J1[j] = fmaf(iz, cam.K[3 + j], m1 * cam.K[6 + j]);
    }
    J0[3] = fmaf(J0[2], c1, -J0[1] * c2);
    J0[4] = fmaf(J0[0], c2, -J0[2] * c0);
    J0[5] = fmaf(J0[1], c0, -J0[0] * c1);
    J1[3] = fmaf(J1[2], c1, -J1[1] * c2);
    J1[4] = fmaf(J1[0], c2, -J1[2] * c0);
    J1[5] = fmaf(J1[1], c0, -J1[0] * c1);
    int k = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const float s0 = KEEP ? J0[i] * lambda : J0[i], s1 = KEEP ? J1[i] * lambda : J1[i];
#pragma unroll
      for (int j = i; j < 6; ++j, ++k) acc[k] = fmaf(s0, J0[j], fmaf(s1, J1[j], acc[k]));
      acc[21 + i] = fmaf(s0, e0, fmaf(s1, e1, acc[21 + i]));
    }
  }
  return st;
}

// ------------------------------------------------------------------ 6x6 solve + pose update
// Eigen::LDLT<Matrix6f> (diagonal pivoting) restated for one thread, float32
// (picp_solver.cpp:102), then v2tEuler(dx)*pose (defs.h:100-136, picp_solver.cpp:103).
__device__ void ldlt_solve6_dev(float (&m)[6][6], float (&d)[6]) {
  int tr[6];
  for (int k = 0; k < 6; ++k) {
    int big = k;
    float bigv = fabsf(m[k][k]);
    for (int i = k + 1; i < 6; ++i)
      if (fabsf(m[i][i]) > bigv) {
        bigv = fabsf(m[i][i]);
        big = i;
      }
    tr[k] = big;
    if (big != k) {
      for (int j = 0; j < k; ++j) { float t = m[k][j]; m[k][j] = m[big][j]; m[big][j] = t; }
      for (int i = big + 1; i < 6; ++i) { float t = m[i][k]; m[i][k] = m[i][big]; m[i][big] = t; }
      { float t = m[k][k]; m[k][k] = m[big][big]; m[big][big] = t; }
      for (int i = k + 1; i < big; ++i) { float t = m[i][k]; m[i][k] = m[big][i]; m[big][i] = t; }
    }
    if (k > 0) {
      float temp[6];
      float s = 0.f;
      for (int j = 0; j < k; ++j) {
        temp[j] = __fmul_rn(m[j][j], m[k][j]);
        s = __fadd_rn(s, __fmul_rn(m[k][j], temp[j]));
      }
      m[k][k] = __fsub_rn(m[k][k], s);
      for (int i = k + 1; i < 6; ++i) {
        float a = 0.f;
        for (int j = 0; j < k; ++j) a = __fadd_rn(a, __fmul_rn(m[i][j], temp[j]));
        m[i][k] = __fsub_rn(m[i][k], a);
      }
    }
    const float akk = m[k][k];
    const bool valid = fabsf(akk) > 0.f;
    if (k == 0 && !valid) {
      for (int j = 0; j < 6; ++j) tr[j] = j;
      break;
    }
    if (valid)
      for (int i = k + 1; i < 6; ++i) m[i][k] = __fdiv_rn(m[i][k], akk);
  }
  for (int k = 0; k < 6; ++k) { float t = d[k]; d[k] = d[tr[k]]; d[tr[k]] = t; }
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < i; ++j) d[i] = __fsub_rn(d[i], __fmul_rn(m[i][j], d[j]));
  for (int i = 0; i < 6; ++i) d[i] = (fabsf(m[i][i]) > FLT_MIN) ? __fdiv_rn(d[i], m[i][i]) : 0.f;
  for (int i = 5; i >= 0; --i)
    for (int j = i + 1; j < 6; ++j) d[i] = __fsub_rn(d[i], __fmul_rn(m[j][i], d[j]));
  for (int k = 5; k >= 0; --k) { float t = d[k]; d[k] = d[tr[k]]; d[tr[k]] = t; }
}

__device__ void mat3_mul_rn(const float* A, const float* B, float* C) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      C[3 * i + j] = dot3_rn(A[3 * i], B[j], A[3 * i + 1], B[3 + j], A[3 * i + 2], B[6 + j]);
}

// result[32] (double) -> damped solve -> pose update, stats ring, convergence flag. One thread.
__device__ void picp_solve_update(const double* __restrict__ res, float damping, PicpDev* dev) {
  float m[6][6], rhs[6];
  int k = 0;
  for (int i = 0; i < 6; ++i)
    for (int j = i; j < 6; ++j, ++k) {
      const float h = (float)res[k];
      m[i][j] = h;
      m[j][i] = h;
    }
  for (int i = 0; i < 6; ++i) {
    m[i][i] = __fadd_rn(m[i][i], damping);  // H += I*damping (picp_solver.cpp:96)
    rhs[i] = -(float)res[21 + i];
  }
  ldlt_solve6_dev(m, rhs);
  // libm cos/sin of a float argument: evaluate in double and round (matches cosf to the last bit
  // except in astronomically rare double-rounding cases)
  const float cx = (float)cos((double)rhs[3]), sx = (float)sin((double)rhs[3]);
  const float cy = (float)cos((double)rhs[4]), sy = (float)sin((double)rhs[4]);
  const float cz = (float)cos((double)rhs[5]), sz = (float)sin((double)rhs[5]);
  const float Rx[9] = {1, 0, 0, 0, cx, -sx, 0, sx, cx};
  const float Ry[9] = {cy, 0, sy, 0, 1, 0, -sy, 0, cy};
  const float Rz[9] = {cz, -sz, 0, sz, cz, 0, 0, 0, 1};
  float Rxy[9], Rd[9], T[12], out[12];
  mat3_mul_rn(Rx, Ry, Rxy);
  mat3_mul_rn(Rxy, Rz, Rd);
  for (int i = 0; i < 12; ++i) T[i] = dev->pose[i];
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 4; ++j)
      out[4 * i + j] = dot3_rn(Rd[3 * i], T[j], Rd[3 * i + 1], T[4 + j], Rd[3 * i + 2], T[8 + j]);
    out[4 * i + 3] = __fadd_rn(out[4 * i + 3], rhs[i]);
  }
  for (int i = 0; i < 12; ++i) dev->pose[i] = out[i];
  const int r = dev->round;
  vo_picp_stats st;
  st.chi_inliers = (float)res[27];
  st.chi_outliers = (float)res[28];
  st.num_inliers = (int)res[29];
  st.num_outliers = (int)res[30];
  if (r < VO_PICP_MAX_ROUNDS) dev->stats[r] = st;
  dev->round = r + 1;
  if (dev->rel_tol >= 0.f) {  // exec/icp_test.cpp:99-106
    const float prev = dev->prev_chi, cur = st.chi_inliers;
    const float rel = (prev > 1e-10f) ? __fdiv_rn(fabsf(__fsub_rn(prev, cur)), prev) : 0.f;
    if (rel < dev->rel_tol) dev->stop = 1;
    dev->prev_chi = cur;
  }
}

// ------------------------------------------------------------------ linearize + reduce
template <bool KEEP, bool STATUS, bool PINHOLE>
__global__ void __launch_bounds__(kThreads, kCtasPerSm) picp_linearize_kernel(const LinArgs a) {
  if (a.dev->stop) return;  // a converged device-side loop turns the remaining launches into no-ops
  __shared__ float s_pose[12];
  __shared__ float s_part[kWarps][kSlots];
  __shared__ double s_fin[kWarps][kSlots];
  __shared__ int s_last;
  if (threadIdx.x < 12) s_pose[threadIdx.x] = a.dev->pose[threadIdx.x];
  __syncthreads();
  float T[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) T[i] = s_pose[i];

  float acc[29];
#pragma unroll
  for (int i = 0; i < 29; ++i) acc[i] = 0.f;
  int n_in = 0, n_out = 0;

  const long long n_quads = (a.n + 3) >> 2;
  const long long step = (long long)gridDim.x * kThreads;
  const float4* __restrict__ p0 = reinterpret_cast<const float4*>(a.pk);
  const float4* __restrict__ p1 = reinterpret_cast<const float4*>(a.pk + a.stride);
  const float4* __restrict__ p2 = reinterpret_cast<const float4*>(a.pk + 2 * a.stride);
  const float4* __restrict__ p3 = reinterpret_cast<const float4*>(a.pk + 3 * a.stride);
  const float4* __restrict__ p4 = reinterpret_cast<const float4*>(a.pk + 4 * a.stride);

  long long q = (long long)blockIdx.x * kThreads + threadIdx.x;
  float4 wx, wy, wz, zu, zv;
  if (q < n_quads) {
    wx = ldg_stream4(p0 + q); wy = ldg_stream4(p1 + q); wz = ldg_stream4(p2 + q);
    zu = ldg_stream4(p3 + q); zv = ldg_stream4(p4 + q);
  }
  while (q < n_quads) {
    const long long qn = q + step;
    float4 nwx, nwy, nwz, nzu, nzv;
    if (qn < n_quads) {  // software prefetch of the next quad: 160 B in flight per thread
      nwx = ldg_stream4(p0 + qn); nwy = ldg_stream4(p1 + qn); nwz = ldg_stream4(p2 + qn);
      nzu = ldg_stream4(p3 + qn); nzv = ldg_stream4(p4 + qn);
    }
    const long long base = q << 2;
    int s0 = VO_PICP_SKIPPED, s1 = VO_PICP_SKIPPED, s2 = VO_PICP_SKIPPED, s3 = VO_PICP_SKIPPED;
    s0 = picp_point<KEEP, PINHOLE>(a.cam, T, a.thr, wx.x, wy.x, wz.x, zu.x, zv.x, acc, n_in, n_out);
    if (base + 1 < a.n) s1 = picp_point<KEEP, PINHOLE>(a.cam, T, a.thr, wx.y, wy.y, wz.y, zu.y, zv.y, acc, n_in, n_out);
    if (base + 2 < a.n) s2 = picp_point<KEEP, PINHOLE>(a.cam, T, a.thr, wx.z, wy.z, wz.z, zu.z, zv.z, acc, n_in, n_out);
    if (base + 3 < a.n) s3 = picp_point<KEEP, PINHOLE>(a.cam, T, a.thr, wx.w, wy.w, wz.w, zu.w, zv.w, acc, n_in, n_out);
    if (STATUS) {
      if (base + 3 < a.n) {
        *reinterpret_cast<uchar4*>(a.status + base) = make_uchar4(s0, s1, s2, s3);
      } else {
        a.status[base] = s0;
        if (base + 1 < a.n) a.status[base + 1] = s1;
        if (base + 2 < a.n) a.status[base + 2] = s2;
      }
    }
    wx = nwx; wy = nwy; wz = nwz; zu = nzu; zv = nzv;
    q = qn;
  }

  // ---- pass 1: warp shuffle tree, then warps summed in warp order
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 29; ++i) acc[i] = warp_sum(acc[i]);
  n_in = warp_sum_i(n_in);
  n_out = warp_sum_i(n_out);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 29; ++i) s_part[warp][i] = acc[i];
    s_part[warp][29] = __int_as_float(n_in);
    s_part[warp][30] = __int_as_float(n_out);
    s_part[warp][31] = 0.f;
  }
  __syncthreads();
  if (threadIdx.x < kSlots) {
    const int c = threadIdx.x;
    float v;
    if (c == 29 || c == 30) {
      int s = 0;
      for (int w = 0; w < kWarps; ++w) s += __float_as_int(s_part[w][c]);
      v = __int_as_float(s);
    } else {
      v = 0.f;
      for (int w = 0; w < kWarps; ++w) v += s_part[w][c];
    }
    a.partials[(size_t)blockIdx.x * kSlots + c] = v;
  }
  // ---- pass 2: the block that takes the last ticket reduces all partials in block order
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(&a.dev->ticket, 1u);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  {
    const int c = threadIdx.x & 31, chunk = threadIdx.x >> 5;
    double v = 0.0;
    long long iv = 0;
    for (unsigned b = chunk; b < gridDim.x; b += kWarps) {
      const float x = __ldcg(a.partials + (size_t)b * kSlots + c);
      if (c == 29 || c == 30) iv += __float_as_int(x);
      else v += (double)x;
    }
    s_fin[chunk][c] = (c == 29 || c == 30) ? (double)iv : v;
  }
  __syncthreads();
  if (threadIdx.x < kSlots) {
    double v = 0.0;
    for (int w = 0; w < kWarps; ++w) v += s_fin[w][threadIdx.x];
    a.result[threadIdx.x] = v;
    s_fin[0][threadIdx.x] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    a.dev->ticket = 0;
    if (a.fuse_solve) picp_solve_update(s_fin[0], a.damping, a.dev);
  }
}

// multi-GPU tail: after the all-reduce every rank runs the identical solve
__global__ void picp_solve_kernel(const double* result, float damping, PicpDev* dev) {
  if (threadIdx.x == 0 && !dev->stop) picp_solve_update(result, damping, dev);
}

// ------------------------------------------------------------------ gather once per correspondence set
__global__ void __launch_bounds__(256) picp_pack_kernel(const int2* __restrict__ pairs, long long n,
                                                        const float* __restrict__ world, long long n_world,
                                                        const float* __restrict__ image, long long n_image,
                                                        float* __restrict__ pk, long long stride, PicpDev* dev) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int2 pr = __ldg(pairs + i);  // (first: image index, second: world index)
  if (pr.x < 0 || pr.x >= n_image || pr.y < 0 || pr.y >= n_world) {
    dev->bad_index = 1;
    pk[i] = 0.f; pk[stride + i] = 0.f; pk[2 * stride + i] = -1.f;  // behind the camera at identity
    pk[3 * stride + i] = 0.f; pk[4 * stride + i] = 0.f;
    return;
  }
  const float* w = world + 3ll * pr.y;
  const float2 z = __ldg(reinterpret_cast<const float2*>(image) + pr.x);
  pk[i] = __ldg(w);
  pk[stride + i] = __ldg(w + 1);
  pk[2 * stride + i] = __ldg(w + 2);
  pk[3 * stride + i] = z.x;
  pk[4 * stride + i] = z.y;
}

__global__ void picp_reset_kernel(PicpDev* dev, float rel_tol) {
  dev->round = 0;
  dev->stop = 0;
  dev->prev_chi = FLT_MAX;
  dev->rel_tol = rel_tol;
}

bool is_pinhole(const float K[9]) {
  return K[1] == 0.f && K[3] == 0.f && K[6] == 0.f && K[7] == 0.f && K[8] == 1.f;
}

}  // namespace

struct vo_picp {
  vo_ctx* ctx = nullptr;
  PicpCam cam{};
  bool have_cam = false;
  bool pinhole = false;
  long long n_world = 0, n_image = 0, n_pairs = -1, n_pad = 0;
  float* d_world = nullptr;
  float* d_image = nullptr;
  bool own_points = false;
  size_t world_cap = 0, image_cap = 0;
  int32_t* d_pairs = nullptr;  // owned staging for host correspondences
  size_t pairs_cap = 0;
  float* d_pk = nullptr;
  size_t pk_cap = 0;
  unsigned char* d_status = nullptr;
  size_t status_cap = 0;
  float* d_partials = nullptr;
  double* d_result = nullptr;
  PicpDev* d_dev = nullptr;
  int max_grid = 0;
  int last_rounds = 0;
};

namespace {

int grow(vo_ctx* ctx, void** p, size_t* cap, size_t bytes) {
  if (bytes <= *cap) return VO_OK;
  VO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (*p) cudaFree(*p);
  *p = nullptr;
  *cap = 0;
  size_t want = vo_align_up(bytes, 256);
  cudaError_t e = cudaMalloc(p, want);
  if (e != cudaSuccess) return vo_set_error(ctx, VO_ERR_NOMEM, "cudaMalloc", cudaGetErrorString(e));
  *cap = want;
  return VO_OK;
}

int grid_for(const vo_picp* s) {
  const long long quads = (s->n_pairs + 3) / 4;
  long long g = (quads + kThreads - 1) / kThreads;
  if (g < 1) g = 1;
  if (g > s->max_grid) g = s->max_grid;
  return (int)g;
}

template <bool KEEP, bool STATUS>
void launch_lin2(bool pinhole, int grid, cudaStream_t st, const LinArgs& a) {
  if (pinhole) picp_linearize_kernel<KEEP, STATUS, true><<<grid, kThreads, 0, st>>>(a);
  else picp_linearize_kernel<KEEP, STATUS, false><<<grid, kThreads, 0, st>>>(a);
}

int launch_linearize(vo_picp* s, float thr, float damping, bool keep, bool status, bool fuse_solve) {
  vo_ctx* ctx = s->ctx;
  LinArgs a;
  a.pk = s->d_pk;
  a.n = s->n_pairs;
  a.stride = s->n_pad;
  a.cam = s->cam;
  a.thr = thr;
  a.damping = damping;
  a.dev = s->d_dev;
  a.partials = s->d_partials;
  a.result = s->d_result;
  a.status = status ? s->d_status : nullptr;
  a.fuse_solve = fuse_solve ? 1 : 0;
  const int grid = grid_for(s);
  if (keep) {
    if (status) launch_lin2<true, true>(s->pinhole, grid, ctx->stream, a);
    else launch_lin2<true, false>(s->pinhole, grid, ctx->stream, a);
  } else {
    if (status) launch_lin2<false, true>(s->pinhole, grid, ctx->stream, a);
    else launch_lin2<false, false>(s->pinhole, grid, ctx->stream, a);
  }
  VO_CHECK_LAUNCH(ctx, "picp_linearize_kernel");
  return VO_OK;
}

// one Gauss-Newton round on the stream (single GPU: 1 launch; with a communicator: 2 + all-reduce)
int enqueue_round(vo_picp* s, float thr, float damping, bool keep) {
  vo_ctx* ctx = s->ctx;
  const bool multi = ctx->nccl_comm != nullptr;
  int st = launch_linearize(s, thr, damping, keep, false, !multi);
  if (st) return st;
  if (multi) {
    st = vo_comm_allreduce_f64(ctx, s->d_result, kSlots);
    if (st) return st;
    picp_solve_kernel<<<1, 32, 0, ctx->stream>>>(s->d_result, damping, s->d_dev);
    VO_CHECK_LAUNCH(ctx, "picp_solve_kernel");
  }
  return VO_OK;
}

int ready(vo_picp* s) {
  if (!s) return VO_ERR_INVALID;
  if (!s->have_cam) return vo_set_error(s->ctx, VO_ERR_STATE, "picp", "set_camera not called");
  if (!s->d_world || !s->d_image) return vo_set_error(s->ctx, VO_ERR_STATE, "picp", "set_points not called");
  if (s->n_pairs < 0) return vo_set_error(s->ctx, VO_ERR_STATE, "picp", "set_correspondences not called");
  return vo_ctx_activate(s->ctx);
}

int reset_rounds(vo_picp* s, float rel_tol) {
  picp_reset_kernel<<<1, 1, 0, s->ctx->stream>>>(s->d_dev, rel_tol);
  VO_CHECK_LAUNCH(s->ctx, "picp_reset_kernel");
  return VO_OK;
}

}  // namespace

extern "C" {

int vo_picp_create(vo_ctx* ctx, vo_picp** out) {
  if (!ctx || !out) return VO_ERR_INVALID;
  *out = nullptr;
  int st = vo_ctx_activate(ctx);
  if (st) return st;
  vo_picp* s = new (std::nothrow) vo_picp();
  if (!s) return VO_ERR_NOMEM;
  s->ctx = ctx;
  s->max_grid = ctx->sm_count * kCtasPerSm;
  cudaError_t e = cudaMalloc((void**)&s->d_partials, (size_t)s->max_grid * kSlots * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc((void**)&s->d_result, kSlots * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc((void**)&s->d_dev, sizeof(PicpDev));
  if (e == cudaSuccess) e = cudaMemsetAsync(s->d_dev, 0, sizeof(PicpDev), ctx->stream);
  if (e != cudaSuccess) {
    vo_set_error(ctx, VO_ERR_CUDA, "vo_picp_create", cudaGetErrorString(e));
    vo_picp_destroy(s);
    return VO_ERR_CUDA;
  }
  st = reset_rounds(s, -1.f);
  if (st) {
    vo_picp_destroy(s);
    return st;
  }
  *out = s;
  return VO_OK;
}

int vo_picp_destroy(vo_picp* s) {
  if (!s) return VO_OK;
  cudaSetDevice(s->ctx->device);
  cudaStreamSynchronize(s->ctx->stream);
  if (s->own_points) {
    if (s->d_world) cudaFree(s->d_world);
    if (s->d_image) cudaFree(s->d_image);
  }
  if (s->d_pairs) cudaFree(s->d_pairs);
  if (s->d_pk) cudaFree(s->d_pk);
  if (s->d_status) cudaFree(s->d_status);
  if (s->d_partials) cudaFree(s->d_partials);
  if (s->d_result) cudaFree(s->d_result);
  if (s->d_dev) cudaFree(s->d_dev);
  delete s;
  return VO_OK;
}

int vo_picp_set_pose(vo_picp* s, const float pose[12]) {
  if (!s || !pose) return VO_ERR_INVALID;
  int st = vo_ctx_activate(s->ctx);
  if (st) return st;
  void* h;
  st = vo_pinned(s->ctx, 64, &h);
  if (st) return st;
  VO_CUDA(s->ctx, cudaStreamSynchronize(s->ctx->stream));  // the staging buffer may still be in flight
  memcpy(h, pose, 12 * sizeof(float));
  VO_CUDA(s->ctx, cudaMemcpyAsync(s->d_dev, h, 12 * sizeof(float), cudaMemcpyHostToDevice, s->ctx->stream));
  VO_CUDA(s->ctx, cudaStreamSynchronize(s->ctx->stream));
  return VO_OK;
}

int vo_picp_get_pose(vo_picp* s, float pose[12]) {
  if (!s || !pose) return VO_ERR_INVALID;
  int st = vo_ctx_activate(s->ctx);
  if (st) return st;
  void* h;
  st = vo_pinned(s->ctx, 64, &h);
  if (st) return st;
  VO_CUDA(s->ctx, cudaMemcpyAsync(h, s->d_dev, 12 * sizeof(float), cudaMemcpyDeviceToHost, s->ctx->stream));
  VO_CUDA(s->ctx, cudaStreamSynchronize(s->ctx->stream));
  memcpy(pose, h, 12 * sizeof(float));
  return VO_OK;
}

int vo_picp_set_camera(vo_picp* s, const float K[9], int rows, int cols, const float pose[12]) {
  if (!s || !K || !pose || rows <= 0 || cols <= 0) return VO_ERR_INVALID;
  for (int i = 0; i < 9; ++i) s->cam.K[i] = K[i];
  s->cam.umax = (float)(cols - 1);
  s->cam.vmax = (float)(rows - 1);
  s->pinhole = is_pinhole(K);
  s->have_cam = true;
  return vo_picp_set_pose(s, pose);
}

int vo_picp_set_points(vo_picp* s, const float* world_xyz, int64_t n_world, const float* image_xy,
                       int64_t n_image) {
  if (!s || n_world < 0 || n_image < 0 || (n_world && !world_xyz) || (n_image && !image_xy)) return VO_ERR_INVALID;
  vo_ctx* ctx = s->ctx;
  int st = vo_ctx_activate(ctx);
  if (st) return st;
  if (!s->own_points) {
    s->d_world = s->d_image = nullptr;
    s->world_cap = s->image_cap = 0;
    s->own_points = true;
  }
  st = grow(ctx, (void**)&s->d_world, &s->world_cap, (size_t)(n_world > 0 ? n_world : 1) * 12);
  if (st) return st;
  st = grow(ctx, (void**)&s->d_image, &s->image_cap, (size_t)(n_image > 0 ? n_image : 1) * 8);
  if (st) return st;
  if (n_world) VO_CUDA(ctx, cudaMemcpyAsync(s->d_world, world_xyz, (size_t)n_world * 12, cudaMemcpyHostToDevice, ctx->stream));
  if (n_image) VO_CUDA(ctx, cudaMemcpyAsync(s->d_image, image_xy, (size_t)n_image * 8, cudaMemcpyHostToDevice, ctx->stream));
  VO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // caller may free its vectors right after init()
  s->n_world = n_world;
  s->n_image = n_image;
  s->n_pairs = -1;
  return VO_OK;
}

int vo_picp_set_points_dev(vo_picp* s, const float* d_world_xyz, int64_t n_world, const float* d_image_xy,
                           int64_t n_image) {
  if (!s || n_world < 0 || n_image < 0 || !d_world_xyz || !d_image_xy) return VO_ERR_INVALID;
  if (s->own_points) {
    cudaStreamSynchronize(s->ctx->stream);
    if (s->d_world) cudaFree(s->d_world);
    if (s->d_image) cudaFree(s->d_image);
    s->world_cap = s->image_cap = 0;
    s->own_points = false;
  }
  s->d_world = const_cast<float*>(d_world_xyz);
  s->d_image = const_cast<float*>(d_image_xy);
  s->n_world = n_world;
  s->n_image = n_image;
  s->n_pairs = -1;
  return VO_OK;
}

int vo_picp_set_correspondences_dev(vo_picp* s, const int32_t* d_pairs, int64_t n_pairs) {
  if (!s || n_pairs < 0 || (n_pairs && !d_pairs)) return VO_ERR_INVALID;
  vo_ctx* ctx = s->ctx;
  if (!s->d_world || !s->d_image) return vo_set_error(ctx, VO_ERR_STATE, "picp", "set_points not called");
  int st = vo_ctx_activate(ctx);
  if (st) return st;
  const long long n_pad = (n_pairs + 3) / 4 * 4;
  st = grow(ctx, (void**)&s->d_pk, &s->pk_cap, (size_t)(n_pad > 0 ? n_pad : 4) * 5 * sizeof(float));
  if (st) return st;
  s->n_pad = n_pad > 0 ? n_pad : 4;
  s->n_pairs = n_pairs;
  if (n_pairs) {
    const long long blocks = (n_pairs + 255) / 256;
    picp_pack_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(reinterpret_cast<const int2*>(d_pairs), n_pairs,
                                                               s->d_world, s->n_world, s->d_image, s->n_image,
                                                               s->d_pk, s->n_pad, s->d_dev);
    VO_CHECK_LAUNCH(ctx, "picp_pack_kernel");
  }
  return VO_OK;
}

int vo_picp_set_correspondences(vo_picp* s, const int32_t* pairs, int64_t n_pairs) {
  if (!s || n_pairs < 0 || (n_pairs && !pairs)) return VO_ERR_INVALID;
  vo_ctx* ctx = s->ctx;
  int st = vo_ctx_activate(ctx);
  if (st) return st;
  st = grow(ctx, (void**)&s->d_pairs, &s->pairs_cap, (size_t)(n_pairs > 0 ? n_pairs : 1) * 8);
  if (st) return st;
  if (n_pairs) VO_CUDA(ctx, cudaMemcpyAsync(s->d_pairs, pairs, (size_t)n_pairs * 8, cudaMemcpyHostToDevice, ctx->stream));
  st = vo_picp_set_correspondences_dev(s, s->d_pairs, n_pairs);
  if (st) return st;
  // range check result + the caller may release `pairs` after return
  void* h;
  st = vo_pinned(ctx, 64, &h);
  if (st) return st;
  VO_CUDA(ctx, cudaMemcpyAsync(h, &s->d_dev->bad_index, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  VO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (*(int*)h) {
    VO_CUDA(ctx, cudaMemsetAsync(&s->d_dev->bad_index, 0, sizeof(int), ctx->stream));
    s->n_pairs = -1;
    return vo_set_error(ctx, VO_ERR_INVALID, "vo_picp_set_correspondences", "index out of range");
  }
  return VO_OK;
}

int vo_picp_linearize(vo_picp* s, float thr, int keep_outliers, float H[36], float b[6], vo_picp_stats* stats,
                      uint8_t* status) {
  int st = ready(s);
  if (st) return st;
  vo_ctx* ctx = s->ctx;
  if (status) {
    st = grow(ctx, (void**)&s->d_status, &s->status_cap, (size_t)s->n_pad);
    if (st) return st;
  }
  VO_CUDA(ctx, cudaMemsetAsync(&s->d_dev->stop, 0, sizeof(int), ctx->stream));
  st = launch_linearize(s, thr, 0.f, keep_outliers != 0, status != nullptr, false);
  if (st) return st;
  st = vo_comm_allreduce_f64(ctx, s->d_result, kSlots);
  if (st) return st;
  void* h;
  st = vo_pinned(ctx, kSlots * sizeof(double), &h);
  if (st) return st;
  VO_CUDA(ctx, cudaMemcpyAsync(h, s->d_result, kSlots * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (status && s->n_pairs)
    VO_CUDA(ctx, cudaMemcpyAsync(status, s->d_status, (size_t)s->n_pairs, cudaMemcpyDeviceToHost, ctx->stream));
  VO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  const double* r = (const double*)h;
  if (H) {
    int k = 0;
    for (int i = 0; i < 6; ++i)
      for (int j = i; j < 6; ++j, ++k) H[6 * i + j] = H[6 * j + i] = (float)r[k];
  }
  if (b)
    for (int i = 0; i < 6; ++i) b[i] = (float)r[21 + i];
  if (stats) {
    stats->chi_inliers = (float)r[27];
    stats->chi_outliers = (float)r[28];
    stats->num_inliers = (int32_t)r[29];
    stats->num_outliers = (int32_t)r[30];
  }
  return VO_OK;
}

int vo_picp_enqueue_rounds(vo_picp* s, float thr, float damping, int keep_outliers, int n_rounds) {
  int st = ready(s);
  if (st) return st;
  if (n_rounds < 0 || n_rounds > VO_PICP_MAX_ROUNDS) return vo_set_error(s->ctx, VO_ERR_INVALID, "vo_picp_enqueue_rounds", "n_rounds");
  st = reset_rounds(s, -1.f);
  if (st) return st;
  for (int r = 0; r < n_rounds; ++r) {
    st = enqueue_round(s, thr, damping, keep_outliers != 0);
    if (st) return st;
  }
  s->last_rounds = n_rounds;
  return VO_OK;
}

int vo_picp_fetch_stats(vo_picp* s, vo_picp_stats* stats_out, int n_rounds) {
  if (!s) return VO_ERR_INVALID;
  vo_ctx* ctx = s->ctx;
  int st = vo_ctx_activate(ctx);
  if (st) return st;
  if (n_rounds < 0 || n_rounds > VO_PICP_MAX_ROUNDS) return VO_ERR_INVALID;
  if (stats_out && n_rounds) {
    void* h;
    st = vo_pinned(ctx, sizeof(vo_picp_stats) * VO_PICP_MAX_ROUNDS, &h);
    if (st) return st;
    VO_CUDA(ctx, cudaMemcpyAsync(h, s->d_dev->stats, sizeof(vo_picp_stats) * n_rounds, cudaMemcpyDeviceToHost, ctx->stream));
    VO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(stats_out, h, sizeof(vo_picp_stats) * n_rounds);
  } else {
    VO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return VO_OK;
}

int vo_picp_one_round(vo_picp* s, float thr, float damping, int keep_outliers, vo_picp_stats* stats) {
  int st = vo_picp_enqueue_rounds(s, thr, damping, keep_outliers, 1);
  if (st) return st;
  vo_picp_stats tmp;
  return vo_picp_fetch_stats(s, stats ? stats : &tmp, 1);
}

int vo_picp_solve(vo_picp* s, float thr, float damping, int keep_outliers, int max_rounds, float rel_tol,
                  int* rounds_done, vo_picp_stats* last) {
  int st = ready(s);
  if (st) return st;
  vo_ctx* ctx = s->ctx;
  if (max_rounds < 1 || max_rounds > VO_PICP_MAX_ROUNDS || rel_tol < 0.f) return vo_set_error(ctx, VO_ERR_INVALID, "vo_picp_solve", "max_rounds / rel_tol");
  st = reset_rounds(s, rel_tol);
  if (st) return st;
  struct Head { int round, stop; };
  void* h;
  st = vo_pinned(ctx, sizeof(vo_picp_stats) * VO_PICP_MAX_ROUNDS, &h);
  if (st) return st;
  const int chunk = 8;  // rounds enqueued between host polls of the device-side stop flag
  int done = 0;
  for (int r = 0; r < max_rounds;) {
    const int m = (max_rounds - r < chunk) ? max_rounds - r : chunk;
    for (int i = 0; i < m; ++i) {
      st = enqueue_round(s, thr, damping, keep_outliers != 0);
      if (st) return st;
    }
    r += m;
    VO_CUDA(ctx, cudaMemcpyAsync(h, &s->d_dev->round, sizeof(Head), cudaMemcpyDeviceToHost, ctx->stream));
    VO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    done = ((Head*)h)->round;
    if (((Head*)h)->stop) break;
  }
  if (rounds_done) *rounds_done = done;
  if (last && done > 0) {
    VO_CUDA(ctx, cudaMemcpyAsync(h, &s->d_dev->stats[done - 1], sizeof(vo_picp_stats), cudaMemcpyDeviceToHost, ctx->stream));
    VO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(last, h, sizeof(vo_picp_stats));
  }
  s->last_rounds = done;
  return VO_OK;
}

}  // extern "C"
