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
//              Gauss-Newton round then streams 20 B/correspondence: tiles of 1408
//              correspondences (5 x 5.5 KB) are staged into a 4-deep shared-memory ring by one
//              producer lane with cp.async.bulk (TMA) + mbarrier, and 11 consumer warps read
//              them back as float4, each thread owning 4 consecutive correspondences per tile
//              (two pairs; each pair runs through packed f32x2 arithmetic).
//   partials   float[grid][32]  per-block sums, fixed slot order
//   result     double[32]   21 upper-triangular H terms, 6 b terms, chi_in, chi_out,
//              n_inliers, n_outliers (+1 pad): the unit all-reduced across GPUs
//   dev        PicpDev      pose, round counter, stats ring: the pose never leaves HBM
//                           between rounds
//
// Reduction is deterministic: per-thread float accumulators over a fixed tile -> CTA
// assignment -> warp shuffle tree -> shared-memory sum over warps in warp order ->
// per-block partial; the block that takes the last ticket sums the partials in block order
// in float64 (pass 2), then (single GPU) solves the damped 6x6 system and updates the pose
// in the same launch.  With peers attached (vo_ctx_peer_attach) the same block first exchanges the 32
// sums with the other GPUs through IPC-mapped mailboxes (picp_peer_allreduce: self-validating 8-byte
// words over NVLink, summed in rank order) and every rank solves the identical system; with only an
// NCCL communicator, pass 2 stops at `result`, ncclAllReduce runs on the stream and a one-warp kernel solves.
// Rounds of a frame are chained with programmatic dependent launch (prologue + first tiles of round r+1 under
// the tail of round r).
//
// Rounding contract: everything that decides the inlier mask (camera point, projection,
// reciprocal, error, chi) uses explicit round-to-nearest intrinsics in the reference's
// evaluation order and is never contracted to FMA; J, H and b are tolerance-level and use FMA.
#include "vo_device.cuh"

#include <float.h>
#include <math.h>

namespace {

#ifndef VO_LIN_THREADS
#define VO_LIN_THREADS 352
#endif
#ifndef VO_LIN_CTAS
#define VO_LIN_CTAS 1
#endif
#ifndef VO_LIN_STAGES
#define VO_LIN_STAGES 4
#endif
// 352 consumer threads + 1 producer warp = 12 warps, one CTA per SM: the packed even/odd accumulators want
// ~150 registers per thread. Measured on B200 (exp/ab.sh, us per 10M-correspondence round): 256x2 CTAs spills,
// 256: 58.2, 320: 57.3, 352: 53.7, 384: 55.7, 448: 59.5, 480: 56.3.
constexpr int kThreads = VO_LIN_THREADS;  // consumer threads per CTA of the linearize kernel
constexpr int kWarps = kThreads / 32;
constexpr int kSlots = 32;       // partial row: 0..20 H, 21..26 b, 27 chi_in, 28 chi_out, 29 n_in, 30 n_out
constexpr int kCtasPerSm = VO_LIN_CTAS;

struct PicpDev {
  float pose[12];
  unsigned int ticket;
  int round;      // rounds executed since the last reset
  int stop;       // set by the device-side convergence test
  int bad_index;  // set by the pack / gather / check kernels when a correspondence index is out of range
  int timeout;    // set by the resident kernel when a wait for another CTA / GPU expired
  float prev_chi;
  float rel_tol;  // < 0: no convergence test
  vo_picp_stats stats[VO_PICP_MAX_ROUNDS];
};

#ifdef VO_PROFILE_STAMPS
__device__ unsigned long long g_stamps[8];
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define VO_STAMP(i) do { g_stamps[i] = gtime(); } while (0)
#else
#define VO_STAMP(i) do { } while (0)
#endif

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
  // fused multi-GPU exchange (peer_n > 1): every rank's mailbox, mine at index peer_rank
  int peer_n, peer_rank;
  VoMailbox* peers[VO_MAX_PEERS];
};

// ---- packed f32x2 arithmetic (sm_100a): one instruction, two IEEE round-to-nearest float ops.
// The FP32 pipe retires the same lanes per clock either way; packing halves the issue slots, which
// is what bounds this kernel (profiles/r01_picp_linearize_v1.md).
// (f2, pack2, unpack2, add2, sub2, fma2: vo_device.cuh)
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
  f2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f2 neg2(f2 a) { return a ^ 0x8000000080000000ull; }
// a*b rounded once and never contracted into a following add (see the note in picp_pair)
__device__ __forceinline__ f2 prod2(f2 a, f2 b) { return fma2(a, b, 0ull); }

// 1/z of two lanes: rcp.approx + one Newton step on FMA.  For 1e-30 <= z <= 1e30 this IS the correctly rounded
// reciprocal (the sequence __frcp_rn runs after its range check); vo_selftest_reciprocal checks all 2^32 inputs.
__device__ __forceinline__ f2 rcp2_newton(f2 z, float z0, float z1) {
  float r0, r1;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(z0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(z1));
  const f2 r = pack2(r0, r1);
  return fma2(r, fma2(z, pack2(-r0, -r1), pack2(1.f, 1.f)), r);
}

// Four correspondences (one quad = two packed pairs) per call: the exact part per point, then J, H += lambda J^T J
// and b += lambda J^T e in packed f32x2 (lane 0 = first point of a pair, lane 1 = second; acc2[k] holds the two
// lanes' partial sums of slot k).  Points that do not contribute get all-zero Jacobian rows instead of a branch.
//
// Instruction budget (the kernels are bound by FP32 / ALU issue, not by HBM: profiles/r02_picp_*): everything stays
// packed from the shared-memory load to the accumulators, and the select / compare / integer work is kept off the
// common path:
//  * a dropped point is removed by zeroing its 1/z ALONE - every Jacobian entry carries a factor 1/z, so J = 0
//    exactly provided the point's c, q and e are finite.  `odd` collects the pairs for which that is not certain
//    (|c0 + c1 + c2| or chi not finite - conservative: a sum may overflow where no term does); such a quad takes the
//    accumulation that zeroes every input of a dropped point by select (the semantics of the first version).
//  * no negated copies of c: the pinhole Jacobian is built from -g, -h and -fx/z, which turns columns 2 and 3 of BOTH
//    rows into their exact negatives; round-to-nearest is sign-symmetric, so the affected sums (H[i][j] with exactly
//    one index in {2,3}; b[2], b[3]) are the exact negatives of the true ones and are flipped back once per round
//    where the packed accumulators are unpacked (picp_slot_flipped, picp_slot_total).
//  * the inlier / outlier counters are predicated adds, chi sums are plain packed adds of the selected chi.
struct PairState {
  f2 c0, c1, c2, q0, q1, iz, e0, e1, chi;
  bool fix0, fix1;  // this lane needs the reference's arithmetic verbatim (general K, IEEE reciprocal)
  bool odd;         // some term of the pair may be non-finite: dropped points need every input zeroed
  bool use0, use1, out0, out1;
};

// sign of accumulator slot k (0..20 upper-triangular H row by row, 21..26 b) in the column-flipped pinhole form
__host__ __device__ constexpr bool picp_slot_flipped(int k) {
  if (k >= 27) return false;
  if (k >= 21) return k - 21 == 2 || k - 21 == 3;
  int i = 0, first = 0;
  while (k >= first + (6 - i)) { first += 6 - i; ++i; }
  const int j = i + (k - first);
  return ((i == 2 || i == 3) != (j == 2 || j == 3));
}

__device__ __forceinline__ void count_if(int& n, bool p) {
  asm("{\n.reg .pred q;\nsetp.ne.u32 q, %1, 0;\n@q add.s32 %0, %0, 1;\n}" : "+r"(n) : "r"((unsigned)p));
}

// stage 1: c = R p + t and the pinhole shortcut values of q and 1/z for both lanes
template <bool PINHOLE>
__device__ __forceinline__ void pair_front(const PicpCam& cam, const float* __restrict__ T, float px0, float py0, float pz0,
                                           float px1, float py1, float pz1, PairState& s) {
  const f2 px = pack2(px0, px1), py = pack2(py0, py1), pz = pack2(pz0, pz1);
  auto bc = [](float x) { return pack2(x, x); };
  s.c0 = add2(bc(T[3]), add2(prod2(bc(T[0]), px), add2(prod2(bc(T[1]), py), prod2(bc(T[2]), pz))));
  s.c1 = add2(bc(T[7]), add2(prod2(bc(T[4]), px), add2(prod2(bc(T[5]), py), prod2(bc(T[6]), pz))));
  s.c2 = add2(bc(T[11]), add2(prod2(bc(T[8]), px), add2(prod2(bc(T[9]), py), prod2(bc(T[10]), pz))));
  float z0, z1, f0, f1;
  unpack2(s.c2, z0, z1);
  unpack2(add2(add2(s.c0, s.c1), s.c2), f0, f1);  // finite <=> c0, c1, c2 all finite (and their sum does not overflow)
  const bool fin0 = finite_f(f0), fin1 = finite_f(f1);
  s.q0 = 0ull; s.q1 = 0ull; s.iz = 0ull;
  if (PINHOLE) {  // shortcut values for both lanes (see picp_project<> for why they are exact)
    s.q0 = add2(prod2(bc(cam.K[0]), s.c0), prod2(bc(cam.K[2]), s.c2));
    s.q1 = add2(prod2(bc(cam.K[4]), s.c1), prod2(bc(cam.K[5]), s.c2));
    s.iz = rcp2_newton(s.c2, z0, z1);
  }
  const bool sc0 = PINHOLE && fin0 && (z0 >= 1e-30f) && (z0 <= 1e30f);
  const bool sc1 = PINHOLE && fin1 && (z1 >= 1e-30f) && (z1 <= 1e30f);
  s.fix0 = !sc0 && !(z0 <= 0.f);
  s.fix1 = !sc1 && !(z1 <= 0.f);
  s.odd = !fin0 || !fin1;
}

// stage 2 (rare): the reference's arithmetic verbatim for the lanes that need it
__device__ __forceinline__ void pair_fix(const PicpCam& cam, PairState& s) {
  float c00, c01, c10, c11, c20, c21, q00, q01, q10, q11, iz0, iz1;
  unpack2(s.c0, c00, c01);
  unpack2(s.c1, c10, c11);
  unpack2(s.c2, c20, c21);
  unpack2(s.q0, q00, q01);
  unpack2(s.q1, q10, q11);
  unpack2(s.iz, iz0, iz1);
  if (s.fix0) {
    q00 = dot3_rn(cam.K[0], c00, cam.K[1], c10, cam.K[2], c20);
    q10 = dot3_rn(cam.K[3], c00, cam.K[4], c10, cam.K[5], c20);
    iz0 = __frcp_rn(dot3_rn(cam.K[6], c00, cam.K[7], c10, cam.K[8], c20));
  }
  if (s.fix1) {
    q01 = dot3_rn(cam.K[0], c01, cam.K[1], c11, cam.K[2], c21);
    q11 = dot3_rn(cam.K[3], c01, cam.K[4], c11, cam.K[5], c21);
    iz1 = __frcp_rn(dot3_rn(cam.K[6], c01, cam.K[7], c11, cam.K[8], c21));
  }
  s.q0 = pack2(q00, q01);
  s.q1 = pack2(q10, q11);
  s.iz = pack2(iz0, iz1);
}

// stage 3: u, inside test, error, chi, status, counters and chi sums
template <bool KEEP>
__device__ __forceinline__ void pair_mid(const PicpCam& cam, float thr, PairState& s, float zu0, float zv0, float zu1,
                                         float zv1, bool v0, bool v1, f2 (&acc2)[29], int& n_in, int& n_out, int& st0,
                                         int& st1) {
  const f2 u = prod2(s.q0, s.iz), v = prod2(s.q1, s.iz);
  float u0, u1, w0, w1, z0, z1, chi0, chi1;
  unpack2(u, u0, u1);
  unpack2(v, w0, w1);
  unpack2(s.c2, z0, z1);
  // camera.h:27,31-34 with their NaN behaviour: a NaN compares false and stays "inside"
  const bool ins0 = v0 && !(z0 <= 0.f) && !(u0 < 0.f) && !(u0 > cam.umax) && !(w0 < 0.f) && !(w0 > cam.vmax);
  const bool ins1 = v1 && !(z1 <= 0.f) && !(u1 < 0.f) && !(u1 > cam.umax) && !(w1 < 0.f) && !(w1 > cam.vmax);
  s.e0 = sub2(u, pack2(zu0, zu1));
  s.e1 = sub2(v, pack2(zv0, zv1));
  s.chi = add2(prod2(s.e0, s.e0), prod2(s.e1, s.e1));
  unpack2(s.chi, chi0, chi1);
  const bool gt0 = chi0 > thr, gt1 = chi1 > thr;
  const bool in0 = ins0 && !gt0, in1 = ins1 && !gt1;
  s.out0 = ins0 && gt0;
  s.out1 = ins1 && gt1;
  st0 = ins0 ? (gt0 ? VO_PICP_OUTLIER : VO_PICP_INLIER) : VO_PICP_SKIPPED;
  st1 = ins1 ? (gt1 ? VO_PICP_OUTLIER : VO_PICP_INLIER) : VO_PICP_SKIPPED;
  count_if(n_in, in0);
  count_if(n_in, in1);
  count_if(n_out, s.out0);
  count_if(n_out, s.out1);
  acc2[27] = add2(acc2[27], pack2(in0 ? chi0 : 0.f, in1 ? chi1 : 0.f));
  acc2[28] = add2(acc2[28], pack2(s.out0 ? chi0 : 0.f, s.out1 ? chi1 : 0.f));
  s.use0 = in0 || (KEEP && s.out0);
  s.use1 = in1 || (KEEP && s.out1);
  s.odd = s.odd || !(chi0 <= FLT_MAX) || !(chi1 <= FLT_MAX);  // chi is >= 0, +inf or NaN
}

// the accumulation's multiply-add: packed.  (Two scalar FFMA on the halves - the same two IEEE operations, twice the
// issue slots, no three-wide-operand register reads - measured 61.2 instead of 47.0 us per round at 10,485,760.)
__device__ __forceinline__ f2 acc_fma2(f2 a, f2 b, f2 c) { return fma2(a, b, c); }

// stage 4: J, H += w J^T J and b += w J^T e (tolerance part).  ODD: every input of a dropped point is zeroed.
template <bool KEEP, bool PINHOLE, bool ODD>
__device__ __forceinline__ void pair_acc(const PicpCam& cam, float thr, const PairState& s, f2 (&acc2)[29]) {
  const bool use0 = s.use0, use1 = s.use1;
  auto keep = [&](f2 x) {
    if (!ODD) return x;
    float a, b;
    unpack2(x, a, b);
    return pack2(use0 ? a : 0.f, use1 ? b : 0.f);
  };
  float iz0, iz1;
  unpack2(s.iz, iz0, iz1);
  const f2 iz = pack2(use0 ? iz0 : 0.f, use1 ? iz1 : 0.f);
  const f2 q0 = keep(s.q0), q1 = keep(s.q1), c0 = keep(s.c0), c1 = keep(s.c1), c2 = keep(s.c2);
  const f2 e0 = keep(s.e0), e1 = keep(s.e1);
  // contribution weight: 1 (inlier) or lambda = sqrt(thr/chi) (kept outlier, picp_solver.cpp:78); a dropped point's
  // weight multiplies zeros
  f2 w = 0ull;
  if (KEEP) {
    float chi0, chi1;
    unpack2(s.chi, chi0, chi1);
    float w0 = (ODD && !use0) ? 0.f : 1.f, w1 = (ODD && !use1) ? 0.f : 1.f;
    if (s.out0) w0 = __fsqrt_rn(__fdiv_rn(thr, chi0));
    if (s.out1) w1 = __fsqrt_rn(__fdiv_rn(thr, chi1));
    w = pack2(w0, w1);
  }
  // ---- J = (Jp*K)*[I | skew(-c)]
  const f2 iz2 = mul2(iz, iz);
  const f2 p0 = mul2(q0, iz2), p1 = mul2(q1, iz2);  // q*iz^2
  f2 J0[6], J1[6];
  if (PINHOLE) {
    auto bc = [](float x) { return pack2(x, x); };
    const f2 a = mul2(iz, bc(cam.K[0])), na = mul2(iz, bc(-cam.K[0])), d = mul2(iz, bc(cam.K[4]));
    const f2 ng = fma2(iz, bc(-cam.K[2]), p0), nh = fma2(iz, bc(-cam.K[5]), p1);  // -(cx/z - q0/z^2), -(cy/z - q1/z^2)
    // columns 2 and 3 carry the opposite sign in both rows (picp_slot_flipped)
    J0[0] = a; J0[1] = 0ull; J0[2] = ng;
    J0[3] = mul2(ng, c1); J0[4] = fma2(ng, c0, mul2(a, c2)); J0[5] = mul2(na, c1);
    J1[0] = 0ull; J1[1] = d; J1[2] = nh;
    J1[3] = fma2(d, c2, mul2(nh, c1)); J1[4] = mul2(nh, c0); J1[5] = mul2(d, c0);
  } else {
    const f2 m0 = neg2(p0), m1 = neg2(p1);
    const f2 nc0 = neg2(c0), nc1 = neg2(c1), nc2 = neg2(c2);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      J0[j] = fma2(iz, pack2(cam.K[j], cam.K[j]), mul2(m0, pack2(cam.K[6 + j], cam.K[6 + j])));
      J1[j] = fma2(iz, pack2(cam.K[3 + j], cam.K[3 + j]), mul2(m1, pack2(cam.K[6 + j], cam.K[6 + j])));
    }
    J0[3] = fma2(J0[1], nc2, mul2(J0[2], c1));
    J0[4] = fma2(J0[2], nc0, mul2(J0[0], c2));
    J0[5] = fma2(J0[0], nc1, mul2(J0[1], c0));
    J1[3] = fma2(J1[1], nc2, mul2(J1[2], c1));
    J1[4] = fma2(J1[2], nc0, mul2(J1[0], c2));
    J1[5] = fma2(J1[0], nc1, mul2(J1[1], c0));
  }
  // ---- H += w J^T J (21 upper-triangular slots), b += w J^T e; structural zeros of the pinhole
  // Jacobian (J0[1] = J1[0] = 0) are skipped at compile time
  int k = 0;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const bool z0 = PINHOLE && i == 1, z1 = PINHOLE && i == 0;  // row entries known to be zero
    const f2 s0 = KEEP ? mul2(J0[i], w) : J0[i], s1 = KEEP ? mul2(J1[i], w) : J1[i];
#pragma unroll
    for (int j = i; j < 6; ++j, ++k) {
      const bool y0 = z0 || (PINHOLE && j == 1), y1 = z1 || (PINHOLE && j == 0);
      if (!y0) acc2[k] = acc_fma2(s0, J0[j], acc2[k]);
      if (!y1) acc2[k] = acc_fma2(s1, J1[j], acc2[k]);
    }
    if (!z0) acc2[21 + i] = acc_fma2(s0, e0, acc2[21 + i]);
    if (!z1) acc2[21 + i] = acc_fma2(s1, e1, acc2[21 + i]);
  }
}

// one quad through all stages; v1..v3: validity of points 1..3 (false only in the last quad of a set)
template <bool KEEP, bool PINHOLE>
__device__ __forceinline__ void picp_quad(const PicpCam& cam, const float* __restrict__ T, float thr, const float4& wx,
                                          const float4& wy, const float4& wz, const float4& zu, const float4& zv, bool v1,
                                          bool v2, bool v3, f2 (&acc2)[29], int& n_in, int& n_out, int& s0, int& s1,
                                          int& s2, int& s3) {
  // both pairs of the quad go through the stages together: one (rare) branch each for the verbatim arithmetic and
  // for the zero-everything accumulation
  PairState pa, pb;
  pair_front<PINHOLE>(cam, T, wx.x, wy.x, wz.x, wx.y, wy.y, wz.y, pa);
  pair_front<PINHOLE>(cam, T, wx.z, wy.z, wz.z, wx.w, wy.w, wz.w, pb);
  if (pa.fix0 || pa.fix1 || pb.fix0 || pb.fix1) {
    pair_fix(cam, pa);
    pair_fix(cam, pb);
  }
  pair_mid<KEEP>(cam, thr, pa, zu.x, zv.x, zu.y, zv.y, true, v1, acc2, n_in, n_out, s0, s1);
  pair_mid<KEEP>(cam, thr, pb, zu.z, zv.z, zu.w, zv.w, v2, v3, acc2, n_in, n_out, s2, s3);
  if (pa.odd || pb.odd) {
    pair_acc<KEEP, PINHOLE, true>(cam, thr, pa, acc2);
    pair_acc<KEEP, PINHOLE, true>(cam, thr, pb, acc2);
  } else {
    pair_acc<KEEP, PINHOLE, false>(cam, thr, pa, acc2);
    pair_acc<KEEP, PINHOLE, false>(cam, thr, pb, acc2);
  }
}

// picp_quad in two halves for a quad with four valid points, so that the resident kernel can put the shared-memory
// loads of its NEXT quad between them: the loaded values of this quad are dead after stage 3, and the ~110 packed
// FMAs of stage 4 cover the latency of the loads (1,048,576 correspondences: 6.86 -> 6.82 us per round, 1,310,720:
// 7.86 -> 7.64).  The streaming kernel does NOT do this: there the early wait on the next tile's barrier costs more
// than the loads it hides (46.98 -> 50.82 us per round at 10,485,760).
template <bool KEEP, bool PINHOLE>
__device__ __forceinline__ void picp_quad_head(const PicpCam& cam, const float* __restrict__ T, float thr, const float4& wx,
                                               const float4& wy, const float4& wz, const float4& zu, const float4& zv,
                                               PairState& pa, PairState& pb, f2 (&acc2)[29], int& n_in, int& n_out) {
  int s0, s1, s2, s3;
  pair_front<PINHOLE>(cam, T, wx.x, wy.x, wz.x, wx.y, wy.y, wz.y, pa);
  pair_front<PINHOLE>(cam, T, wx.z, wy.z, wz.z, wx.w, wy.w, wz.w, pb);
  if (pa.fix0 || pa.fix1 || pb.fix0 || pb.fix1) {
    pair_fix(cam, pa);
    pair_fix(cam, pb);
  }
  pair_mid<KEEP>(cam, thr, pa, zu.x, zv.x, zu.y, zv.y, true, true, acc2, n_in, n_out, s0, s1);
  pair_mid<KEEP>(cam, thr, pb, zu.z, zv.z, zu.w, zv.w, true, true, acc2, n_in, n_out, s2, s3);
}
template <bool KEEP, bool PINHOLE>
__device__ __forceinline__ void picp_quad_tail(const PicpCam& cam, float thr, const PairState& pa, const PairState& pb,
                                               f2 (&acc2)[29]) {
  if (pa.odd || pb.odd) {
    pair_acc<KEEP, PINHOLE, true>(cam, thr, pa, acc2);
    pair_acc<KEEP, PINHOLE, true>(cam, thr, pb, acc2);
  } else {
    pair_acc<KEEP, PINHOLE, false>(cam, thr, pa, acc2);
    pair_acc<KEEP, PINHOLE, false>(cam, thr, pb, acc2);
  }
}

// the thread's total of slot i from its packed accumulator (undoing the column flips of the pinhole form)
template <bool PINHOLE>
__device__ __forceinline__ float picp_slot_total(const f2 (&acc2)[29], int i) {
  float lo, hi;
  unpack2(acc2[i], lo, hi);
  const float v = lo + hi;
  return (PINHOLE && picp_slot_flipped(i)) ? -v : v;
}

// result[32] (double) -> damped solve -> pose update, stats ring, convergence flag. One thread.
__device__ void picp_solve_update(const double* __restrict__ res, float damping, PicpDev* dev) {
  float Hu[21], bb[6], pose[12];
#pragma unroll
  for (int k = 0; k < 21; ++k) Hu[k] = (float)res[k];
#pragma unroll
  for (int k = 0; k < 6; ++k) bb[k] = (float)res[21 + k];
#pragma unroll
  for (int i = 0; i < 12; ++i) pose[i] = dev->pose[i];
  picp_gn_step(Hu, bb, damping, pose);
#pragma unroll
  for (int i = 0; i < 12; ++i) dev->pose[i] = pose[i];
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

// lane `lane` reads term `lane` of every rank out of THIS GPU's mailbox (parity buffer `par`) until all words carry
// sequence number `seq`, and adds the ranks' doubles in rank order.  false: gave up after ~5 s (a peer never came).
__device__ __forceinline__ bool peer_collect(VoMailbox* me, int par, unsigned seq, int peer_n, int lane, double& total) {
  unsigned long long r0[VO_MAX_PEERS], r1[VO_MAX_PEERS];
  bool ok = false;
  for (long long spins = 0; spins < (1ll << 23); ++spins) {
    bool all = true;
#pragma unroll
    for (int q = 0; q < VO_MAX_PEERS; ++q) {
      if (q < peer_n) {
        const volatile unsigned long long* src = &me->ll[par][q][lane][0];
        r0[q] = src[0];
        r1[q] = src[1];
      }
    }
#pragma unroll
    for (int q = 0; q < VO_MAX_PEERS; ++q)
      if (q < peer_n) all = all && ((unsigned)(r0[q] >> 32) == seq) && ((unsigned)(r1[q] >> 32) == seq);
    if (all) {
      ok = true;
      break;
    }
  }
  total = 0.0;
#pragma unroll
  for (int q = 0; q < VO_MAX_PEERS; ++q)
    if (q < peer_n) total += __longlong_as_double((long long)((r0[q] & 0xffffffffull) | (r1[q] << 32)));
  return ok;
}

// ------------------------------------------------------------------ fused all-reduce over NVLink peer memory
// One warp. Lane c owns term c. Push my 32 terms into every rank's mailbox (mine included) as self-validating
// 8-byte words {half of the double | sequence number}, poll my own mailbox until every rank's words carry this
// round's number, and sum them in RANK order: all ranks add the same numbers in the same order, so they solve
// bit-identical systems without a broadcast.  (A first version wrote plain doubles, a system-scope fence, a flag
// per rank and a second fence before reading: ~6.5 us per round; see profiles/r01_final_summary.md.)
// Mailboxes are double-buffered by round parity; a rank can reach round s+2 only after every rank has
// published s+1, i.e. finished reading round s, so a parity buffer is never overwritten while in use.
__device__ __forceinline__ double picp_peer_allreduce(const LinArgs& a, double mine, int lane) {
  VoMailbox* me = a.peers[a.peer_rank];
  const unsigned seq = *(volatile unsigned*)&me->seq + 1u;  // this round's sequence number (same on all ranks)
  const int par = (int)(seq & 1u);
  const unsigned long long bits = (unsigned long long)__double_as_longlong(mine), tag = (unsigned long long)seq << 32;
  const unsigned long long w0 = (bits & 0xffffffffull) | tag, w1 = (bits >> 32) | tag;
  for (int p = 0; p < a.peer_n; ++p) {  // 512 B per peer, every 8-byte word self-validating
    volatile unsigned long long* dst = &a.peers[p]->ll[par][a.peer_rank][lane][0];
    dst[0] = w0;
    dst[1] = w1;
  }
  // lane c collects term c of every rank: all 2 x n loads in flight at once (one L2 round trip per poll, not one per
  // rank), repeated until every word carries this round's number; the sum then runs in RANK order
  double total = 0.0;
  const bool ok = peer_collect(me, par, seq, a.peer_n, lane, total);
  const bool all_ok = __all_sync(0xffffffffu, ok);
  if (lane == 0) {
    *(volatile unsigned*)&me->seq = seq;
    if (!all_ok) me->timeout = 1u;
  }
  return total;
}

// ------------------------------------------------------------------ linearize + reduce
// Tile ring: kStages x (5 planes x kTile floats) in shared memory, filled by one producer lane
// with cp.async.bulk and consumed by 8 warps; tile t belongs to CTA (t mod gridDim.x) and quad
// `threadIdx.x` of every tile to the same thread, so the summation order is fixed.
constexpr int kTile = 4 * kThreads;  // correspondences per tile
constexpr int kStages = VO_LIN_STAGES;
constexpr int kLinThreads = kThreads + 32;  // 8 consumer warps + 1 producer warp
constexpr size_t kLinSmemBytes = (size_t)kStages * 5 * kTile * sizeof(float);

template <bool KEEP, bool STATUS, bool PINHOLE>
#ifdef VO_LIN_MAXNREG
__global__ void __maxnreg__(VO_LIN_MAXNREG) picp_linearize_kernel(const LinArgs a) {
#else
__global__ void __launch_bounds__(kLinThreads, kCtasPerSm) picp_linearize_kernel(const LinArgs a) {
#endif
  // Programmatic dependent launch: the next round's grid may start launching right away; its CTAs run this
  // prologue (barrier init, first TMA tiles: the packed planes do not change between rounds) on the SMs this
  // grid has already left, i.e. under the solve tail, and block in griddepcontrol.wait until this grid has
  // completed and its pose / partials / ticket are visible.  Both instructions are no-ops in a normal launch.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (blockIdx.x == 0 && threadIdx.x == 0) VO_STAMP(0);
  extern __shared__ __align__(128) float s_tiles[];
  __shared__ __align__(8) unsigned long long s_full[kStages], s_empty[kStages];
  __shared__ float s_pose[12];
  __shared__ float s_part[kWarps][kSlots];
  __shared__ double s_fin[kWarps][kSlots];
  __shared__ int s_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&s_full[s], 1);        // the producer's arrive.expect_tx
      mbar_init(&s_empty[s], kWarps);  // one arrival per consumer warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const long long n_tiles = (a.n + kTile - 1) / kTile, n_full_tiles = a.n / kTile;
  f2 acc2[29];
#pragma unroll
  for (int i = 0; i < 29; ++i) acc2[i] = 0ull;
  int n_in = 0, n_out = 0;

  if (warp == kWarps) {
    // ---------------- producer: one elected lane keeps the ring full
    if (lane == 0) {
      auto issue = [&](long long t, int it) {
        const int stage = it % kStages;
        const long long first = t * kTile;
        const long long left = ((a.n + 3) & ~3ll) - first;
        const unsigned bytes = (unsigned)((left < kTile ? left : kTile) * sizeof(float));
        float* dst = s_tiles + (size_t)stage * 5 * kTile;
        mbar_arrive_expect_tx(&s_full[stage], 5u * bytes);
#pragma unroll
        for (int p = 0; p < 5; ++p) bulk_g2s(dst + p * kTile, a.pk + p * a.stride + first, bytes, &s_full[stage]);
      };
      // the first kStages tiles do not depend on the previous round: fetch them before the dependency resolves
      int it = 0;
      long long t = blockIdx.x;
      for (; t < n_tiles && it < kStages; t += gridDim.x, ++it) issue(t, it);
      asm volatile("griddepcontrol.wait;" ::: "memory");
      if (*(volatile int*)&a.dev->stop) {
        // converged device-side loop: this launch is a no-op, but the copies in flight must land before exit
        for (int k = 0; k < it; ++k) mbar_wait(&s_full[k], 0u);
        return;
      }
      for (; t < n_tiles; t += gridDim.x, ++it) {
        const int stage = it % kStages;
        const unsigned phase = (unsigned)(it / kStages) & 1u;
        mbar_wait_relaxed(&s_empty[stage], phase ^ 1u);
        issue(t, it);
      }
    } else {
      asm volatile("griddepcontrol.wait;" ::: "memory");
      if (*(volatile int*)&a.dev->stop) return;
    }
  } else {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (*(volatile int*)&a.dev->stop) return;  // a converged device-side loop turns the remaining launches into no-ops
    if (threadIdx.x < 12) s_pose[threadIdx.x] = a.dev->pose[threadIdx.x];
    asm volatile("bar.sync 1, %0;" ::"n"(kThreads) : "memory");  // consumers only
    // ---------------- consumers
    float T[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) T[i] = s_pose[i];
    int it = 0;
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
      const int stage = it % kStages;
      const unsigned phase = (unsigned)(it / kStages) & 1u;
      mbar_wait(&s_full[stage], phase);
      const float4* tile = reinterpret_cast<const float4*>(s_tiles + (size_t)stage * 5 * kTile);
      const long long base = t * kTile + 4ll * threadIdx.x;
      float4 wx, wy, wz, zu, zv;
      int s0, s1, s2, s3;
      if (t < n_full_tiles) {  // every tile but the last of the stream: no per-point validity
        wx = tile[threadIdx.x];
        wy = tile[kThreads + threadIdx.x];
        wz = tile[2 * kThreads + threadIdx.x];
        zu = tile[3 * kThreads + threadIdx.x];
        zv = tile[4 * kThreads + threadIdx.x];
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[stage]);  // values are in registers: hand the stage back
        picp_quad<KEEP, PINHOLE>(a.cam, T, a.thr, wx, wy, wz, zu, zv, true, true, true, acc2, n_in, n_out, s0, s1, s2, s3);
      } else {
        const bool any = base < a.n;
        if (any) {
          wx = tile[threadIdx.x];
          wy = tile[kThreads + threadIdx.x];
          wz = tile[2 * kThreads + threadIdx.x];
          zu = tile[3 * kThreads + threadIdx.x];
          zv = tile[4 * kThreads + threadIdx.x];
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[stage]);
        if (!any) continue;
        const long long left = a.n - base;  // >= 1; < 4 only in the last quad of the stream
        picp_quad<KEEP, PINHOLE>(a.cam, T, a.thr, wx, wy, wz, zu, zv, left > 1, left > 2, left > 3, acc2, n_in, n_out, s0,
                                 s1, s2, s3);
      }
      if (STATUS) {
        if (base + 3 < a.n) {
          *reinterpret_cast<uchar4*>(a.status + base) = make_uchar4(s0, s1, s2, s3);
        } else {
          a.status[base] = s0;
          if (base + 1 < a.n) a.status[base + 1] = s1;
          if (base + 2 < a.n) a.status[base + 2] = s2;
        }
      }
    }
    float acc[29];
#pragma unroll
    for (int i = 0; i < 29; ++i) acc[i] = picp_slot_total<PINHOLE>(acc2, i);
    // ---- pass 1: warp shuffle tree, then warps summed in warp order
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
  }
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x == 0) VO_STAMP(1);
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
  if (threadIdx.x == 0) VO_STAMP(2);
  __threadfence();
  if (threadIdx.x < kThreads) {
    const int c = threadIdx.x & 31, chunk = threadIdx.x >> 5;
    double v = 0.0;
    long long iv = 0;
    constexpr int kBatch = 8;  // independent L2 loads in flight per thread (the sum order stays fixed)
    for (unsigned b0 = chunk; b0 < gridDim.x; b0 += kWarps * kBatch) {
      float x[kBatch];
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const unsigned b = b0 + u * kWarps;
        x[u] = (b < gridDim.x) ? __ldcg(a.partials + (size_t)b * kSlots + c) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        if (c == 29 || c == 30) iv += __float_as_int(x[u]);
        else v += (double)x[u];
      }
    }
    s_fin[chunk][c] = (c == 29 || c == 30) ? (double)iv : v;
  }
  __syncthreads();
  if (threadIdx.x < kSlots) {
    double v = 0.0;
    for (int w = 0; w < kWarps; ++w) v += s_fin[w][threadIdx.x];
    if (a.peer_n > 1) v = picp_peer_allreduce(a, v, threadIdx.x);  // warp 0: all ranks' terms, rank order
    a.result[threadIdx.x] = v;
    s_fin[0][threadIdx.x] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    VO_STAMP(3);
    a.dev->ticket = 0;
    if (a.fuse_solve) picp_solve_update(s_fin[0], a.damping, a.dev);
    VO_STAMP(4);
  }
}

// ------------------------------------------------------------------ resident multi-round kernel
// A correspondence set that fits the shared memory of the machine (148 SMs x 2816 quads = 1.67 M correspondences:
// BASELINE config 2's 1 M frame, config 3's 10 M frame sharded over 8 GPUs, every frame of the bundled dataset) is
// gathered ONCE into shared memory (28 B per correspondence read from HBM, nothing written back) and stays there for
// all Gauss-Newton rounds of the frame: one cooperative launch, rounds 2..n touch neither HBM nor the launch path.
// Per round and CTA: linearize the resident quads (same device functions, same rounding as the streaming kernel)
// -> warp recursive-halving reduction -> CTA partial -> ONE exchange between the CTAs: every CTA publishes its 32
// sums as self-validating 8-byte words {float bits | round sequence number} (payload and flag in one store: no
// fence, no separate barrier) and every CTA polls all partials out of L2, adds them in CTA order in float64 and
// solves the 6x6 system itself.  All CTAs add the same numbers in the same order, so they all hold the
// bit-identical new pose without a broadcast and without a second grid-wide synchronisation.  With peers attached
// only CTA 0 collects the GPU's partials; it pushes the GPU's 32 sums into every rank's mailbox over NVLink
// (picp_peer_allreduce's protocol) and the warp 0 of EVERY CTA on every GPU polls its GPU's mailbox and adds the
// ranks' sums in rank order.  Word buffers are double-buffered by round parity (a CTA can reach round r+2 only
// after every CTA has published r+1, i.e. finished reading round r).  The convergence test of
// exec/icp_test.cpp:99-106 runs inside the kernel: every CTA takes the same decision from the same sums.
#ifndef VO_RES_THREADS
#define VO_RES_THREADS 384
#endif
constexpr int kResThreads = VO_RES_THREADS;
constexpr int kResWarps = kResThreads / 32;
constexpr int kResMaxGrid = 160;   // CTAs whose partials one CTA can collect (B200: 148 SMs)
constexpr int kResMaxQuads = 2816;  // quads per CTA: 5 planes x 2816 x 16 B = 225,280 B of the 232,448 B a CTA can own
constexpr long long kResSpinLimit = 1ll << 21;  // ~2 s of polling: flag a timeout instead of hanging the GPU

struct ResArgs {
  const int2* pairs;
  long long n;
  const float* world;
  long long n_world;
  const float* image;
  long long n_image;
  PicpCam cam;
  float thr, damping, rel_tol;
  int n_rounds;
  int quads_per_cta;
  PicpDev* dev;
  unsigned long long* ll;  // [2 parities][ll_stride CTAs][32 terms]
  unsigned ll_seq0;        // round r of this launch carries sequence number ll_seq0 + 1 + r
  int ll_stride;
  int peer_n, peer_rank;
  VoMailbox* peers[VO_MAX_PEERS];
  int use_pose0;     // start from pose0 (a vo_picp_set_pose not yet on the device) instead of dev->pose
  float pose0[12];
};

__device__ __forceinline__ unsigned long long ld_poll(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_poll(unsigned long long* p, unsigned long long v) {
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// named-barrier helpers: the persistent kernels synchronise their COMPUTE threads only (the streaming variant has a
// producer warp that must not take part)
__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ bool named_bar_or(int id, int n, bool pred) {
  int r;
  asm volatile(
      "{\n.reg .pred p, q;\nsetp.ne.s32 p, %3, 0;\nbar.red.or.pred q, %1, %2, p;\nselp.s32 %0, 1, 0, q;\n}\n"
      : "=r"(r)
      : "r"(id), "r"(n), "r"((int)pred)
      : "memory");
  return r != 0;
}

struct RoundCtx {
  PicpDev* dev;
  unsigned long long* ll;
  unsigned ll_seq0;
  int ll_stride;
  float damping, rel_tol;
  int peer_n, peer_rank;
  VoMailbox* const* peers;
};

// The end of one Gauss-Newton round of a persistent kernel, executed by the N_THREADS compute threads of every CTA
// (named barrier BAR): s_part[warp][term] holds the warps' sums.  Publishes the CTA's partial as self-validating
// words, collects all CTAs' partials (every CTA on one GPU; CTA 0 only with peers, which then pushes the GPU's sums
// into every rank's mailbox while every CTA polls its own GPU's mailbox), solves the damped 6x6 system redundantly in
// every CTA, applies the increment to s_pose and sets *s_stop (0 go on, 1 converged, 2 a wait timed out).
// WARP_SOLVE picks the form of the 6x6 solve (same bits either way, measured with exp/picp_ab.py on a B200): in the
// resident kernel nothing else runs on the SM while warp 0 solves, and one lane's instruction-level parallelism beats
// the shuffle latencies of the warp form (6.95 vs 7.43 us per round at 1,048,576); in the streaming kernel the warp
// form is the faster one (49.7 vs 50.5 us at 10,485,760).
template <int BAR, int N_THREADS, bool WARP_SOLVE>
__device__ __forceinline__ void round_exchange_and_solve(const RoundCtx& a, int r, unsigned mb_seq0, VoMailbox* me,
                                                         float (*s_part)[kSlots], double (*s_fin)[kSlots], double* s_tot,
                                                         float* s_pose, float* s_dx, int* s_stop, float& prev_chi) {
  constexpr int WARPS = N_THREADS / 32;
  constexpr int BATCH = (kResMaxGrid + WARPS - 1) / WARPS;  // L2 loads in flight per polling thread
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool multi_cta = gridDim.x > 1;
  const bool peers = a.peer_n > 1;
  const bool collect = multi_cta && (!peers || blockIdx.x == 0);  // this CTA adds up the GPU's partials
  named_bar_sync(BAR, N_THREADS);
  const unsigned seq = a.ll_seq0 + 1u + (unsigned)r;
  bool timed_out = false;
  if (multi_cta) {
    unsigned long long* buf = a.ll + (size_t)(seq & 1u) * a.ll_stride * kSlots;
    if (warp == 0) {  // publish this CTA's partial: term `lane`, warps added in warp order
      float p = 0.f;
#pragma unroll
      for (int w = 0; w < WARPS; ++w) p += s_part[w][lane];
      st_poll(buf + (size_t)blockIdx.x * kSlots + lane, (unsigned long long)__float_as_uint(p) | ((unsigned long long)seq << 32));
    }
    if (collect) {  // thread (warp, lane): term `lane` of CTAs warp, warp + WARPS, ...: all loads in flight at once
      unsigned long long w[BATCH];
      long long spins = 0;
      for (;;) {
        bool ok = true;
#pragma unroll
        for (int u = 0; u < BATCH; ++u) {
          const unsigned b = warp + u * WARPS;
          w[u] = (b < gridDim.x) ? ld_poll(buf + (size_t)b * kSlots + lane) : ((unsigned long long)seq << 32);
        }
#pragma unroll
        for (int u = 0; u < BATCH; ++u) ok = ok && ((unsigned)(w[u] >> 32) == seq);
        if (ok) break;
        if (++spins > kResSpinLimit) {
          timed_out = true;
          break;
        }
      }
      double v = 0.0;
#pragma unroll
      for (int u = 0; u < BATCH; ++u) {
        const unsigned b = warp + u * WARPS;
        if (b < gridDim.x) v += (double)__uint_as_float((unsigned)w[u]);
      }
      s_fin[warp][lane] = v;
    }
  }
  if (named_bar_or(BAR, N_THREADS, timed_out)) {
    if (tid == 0) {
      a.dev->timeout = 1;
      *s_stop = 2;
    }
    named_bar_sync(BAR, N_THREADS);
    return;
  }
  if (warp == 0) {
    double tot = 0.0;
    if (collect) {
#pragma unroll
      for (int w = 0; w < WARPS; ++w) tot += s_fin[w][lane];
    } else if (!multi_cta) {
#pragma unroll
      for (int w = 0; w < WARPS; ++w) tot += (double)s_part[w][lane];
    }
    bool ok = true;
    if (peers) {
      const unsigned mseq = mb_seq0 + 1u + (unsigned)r;
      const int par = (int)(mseq & 1u);
      if (blockIdx.x == 0) {  // the GPU's sums -> every rank's mailbox (mine included), 8-byte self-validating words
        const unsigned long long bits = (unsigned long long)__double_as_longlong(tot), tag = (unsigned long long)mseq << 32;
        const unsigned long long w0 = (bits & 0xffffffffull) | tag, w1 = (bits >> 32) | tag;
        for (int p = 0; p < a.peer_n; ++p) {
          volatile unsigned long long* dst = &a.peers[p]->ll[par][a.peer_rank][lane][0];
          dst[0] = w0;
          dst[1] = w1;
        }
      }
      // every CTA: term `lane` of every rank out of this GPU's mailbox, added in RANK order
      ok = peer_collect(me, par, mseq, a.peer_n, lane, tot);
      ok = __all_sync(0xffffffffu, ok);
    }
    s_tot[lane] = tot;
    float* hb = s_part[0];  // this lane has read its column of s_part for the last time this round: 21 H, 6 b as float
    hb[lane] = (float)tot;
    __syncwarp();
    float dx[6];
    if (WARP_SOLVE) {  // all 32 lanes: six columns of shuffles instead of one lane's chain
      picp_gn_solve_warp(hb, a.damping, lane, s_dx, dx);
    } else {  // lane 0 alone (the same bits): shorter when nothing else competes for the scheduler, see the callers
      if (lane == 0) {
        float Hu[21], bb[6], x[6];
#pragma unroll
        for (int k = 0; k < 21; ++k) Hu[k] = hb[k];
#pragma unroll
        for (int k = 0; k < 6; ++k) bb[k] = hb[21 + k];
        picp_gn_solve(Hu, bb, a.damping, x);
#pragma unroll
        for (int k = 0; k < 6; ++k) s_dx[k] = x[k];
      }
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 6; ++k) dx[k] = s_dx[k];
    }
    if (lane == 0) {
      vo_picp_stats st;
      st.chi_inliers = (float)s_tot[27];
      st.chi_outliers = (float)s_tot[28];
      st.num_inliers = (int)s_tot[29];
      st.num_outliers = (int)s_tot[30];
      if (blockIdx.x == 0 && r < VO_PICP_MAX_ROUNDS) a.dev->stats[r] = st;
      int stop = 0;
      if (a.rel_tol >= 0.f) {  // exec/icp_test.cpp:99-106
        const float cur = st.chi_inliers;
        const float rel = (prev_chi > 1e-10f) ? __fdiv_rn(fabsf(__fsub_rn(prev_chi, cur)), prev_chi) : 0.f;
        if (rel < a.rel_tol) stop = 1;
        prev_chi = cur;
      }
      if (!ok) {
        stop = 2;
        me->timeout = 1u;
      }
      *s_stop = stop;
    }
    picp_apply_dx_warp(dx, s_pose, lane);
  }
  named_bar_sync(BAR, N_THREADS);
}

template <bool KEEP, bool PINHOLE>
__global__ void __launch_bounds__(kResThreads, 1) picp_resident_kernel(const ResArgs a) {
  extern __shared__ __align__(16) float4 s_pl[];  // [5 planes][quads_per_cta]: wx wy wz zu zv of 4 correspondences
  __shared__ float s_part[kResWarps][kSlots];
  __shared__ double s_fin[kResWarps][kSlots];
  __shared__ double s_tot[kSlots];
  __shared__ float s_pose[12];
  __shared__ float s_dx[6];
  __shared__ int s_stop;  // 0 go on, 1 converged, 2 a wait timed out
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int qpc = a.quads_per_cta;
  const long long n_quads = (a.n + 3) >> 2;
  const long long qb = (long long)blockIdx.x * qpc;
  const long long qe = (qb + qpc < n_quads) ? qb + qpc : n_quads;
  const int nq = qe > qb ? (int)(qe - qb) : 0;
  // quads with four valid points: all but the last quad of the set when n is not a multiple of 4
  const int nq_full = (nq > 0 && qe == n_quads && (a.n & 3)) ? nq - 1 : nq;
  const bool peers = a.peer_n > 1;

  // ---- gather once: (first: image index, second: world index) -> the CTA's planes in shared memory
  {
    bool bad = false;
    for (int l = tid; l < nq; l += kResThreads) {
      const long long i0 = (qb + l) * 4;
      float wx[4], wy[4], wz[4], zu[4], zv[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        wx[k] = wy[k] = wz[k] = zu[k] = zv[k] = 0.f;  // padding lanes of the last quad: finite, masked by `left`
        if (i0 + k < a.n) {
          const int2 pr = __ldg(a.pairs + i0 + k);
          if (pr.x < 0 || pr.x >= a.n_image || pr.y < 0 || pr.y >= a.n_world) {
            bad = true;
            wz[k] = -1.f;
          } else {
            const float* w = a.world + 3ll * pr.y;
            const float2 z = __ldg(reinterpret_cast<const float2*>(a.image) + pr.x);
            wx[k] = __ldg(w);
            wy[k] = __ldg(w + 1);
            wz[k] = __ldg(w + 2);
            zu[k] = z.x;
            zv[k] = z.y;
          }
        }
      }
      s_pl[l] = make_float4(wx[0], wx[1], wx[2], wx[3]);
      s_pl[qpc + l] = make_float4(wy[0], wy[1], wy[2], wy[3]);
      s_pl[2 * qpc + l] = make_float4(wz[0], wz[1], wz[2], wz[3]);
      s_pl[3 * qpc + l] = make_float4(zu[0], zu[1], zu[2], zu[3]);
      s_pl[4 * qpc + l] = make_float4(zv[0], zv[1], zv[2], zv[3]);
    }
    if (bad) a.dev->bad_index = 1;
  }
  if (tid < 12) s_pose[tid] = a.use_pose0 ? a.pose0[tid] : a.dev->pose[tid];
  if (tid == 0) s_stop = 0;
  VoMailbox* me = peers ? a.peers[a.peer_rank] : nullptr;
  const unsigned mb_seq0 = peers ? *(volatile unsigned*)&me->seq : 0u;
  float prev_chi = FLT_MAX;  // (lane 0 of warp 0)
  __syncthreads();

  int r = 0;
  for (; r < a.n_rounds; ++r) {
    float T[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) T[i] = s_pose[i];
    f2 acc2[29];
#pragma unroll
    for (int i = 0; i < 29; ++i) acc2[i] = 0ull;
    int n_in = 0, n_out = 0;
    {  // each quad is loaded one step ahead, between stages 3 and 4 of the quad before it (picp_quad_head / _tail)
      int l = tid;
      float4 wx, wy, wz, zu, zv;
      if (l < nq_full) {
        wx = s_pl[l]; wy = s_pl[qpc + l]; wz = s_pl[2 * qpc + l]; zu = s_pl[3 * qpc + l]; zv = s_pl[4 * qpc + l];
      }
      while (l < nq_full) {
        PairState pa, pb;
        picp_quad_head<KEEP, PINHOLE>(a.cam, T, a.thr, wx, wy, wz, zu, zv, pa, pb, acc2, n_in, n_out);
        l += kResThreads;
        if (l < nq_full) {
          wx = s_pl[l]; wy = s_pl[qpc + l]; wz = s_pl[2 * qpc + l]; zu = s_pl[3 * qpc + l]; zv = s_pl[4 * qpc + l];
        }
        picp_quad_tail<KEEP, PINHOLE>(a.cam, a.thr, pa, pb, acc2);
      }
    }
    if (nq_full < nq && nq_full % kResThreads == tid) {  // the partial last quad of the set: the last quad of its thread
      const int l = nq_full;
      const float4 wx = s_pl[l], wy = s_pl[qpc + l], wz = s_pl[2 * qpc + l], zu = s_pl[3 * qpc + l], zv = s_pl[4 * qpc + l];
      const long long left = a.n - (qb + l) * 4;  // 1..3
      int s0, s1, s2, s3;
      picp_quad<KEEP, PINHOLE>(a.cam, T, a.thr, wx, wy, wz, zu, zv, left > 1, left > 2, left > 3, acc2, n_in, n_out, s0, s1,
                               s2, s3);
    }
    // ---- warp reduction: lane L ends up with the warp's total of term L (counts are exact in float: < 2^24)
    {
      float v[32];
#pragma unroll
      for (int i = 0; i < 29; ++i) v[i] = picp_slot_total<PINHOLE>(acc2, i);
      v[29] = (float)n_in;
      v[30] = (float)n_out;
      v[31] = 0.f;
      s_part[warp][lane] = warp_sum32_scatter(v, lane);
    }
    RoundCtx rc;
    rc.dev = a.dev; rc.ll = a.ll; rc.ll_seq0 = a.ll_seq0; rc.ll_stride = a.ll_stride; rc.damping = a.damping;
    rc.rel_tol = a.rel_tol; rc.peer_n = a.peer_n; rc.peer_rank = a.peer_rank; rc.peers = a.peers;
    round_exchange_and_solve<0, kResThreads, false>(rc, r, mb_seq0, me, s_part, s_fin, s_tot, s_pose, s_dx, &s_stop, prev_chi);
    if (s_stop) {
      ++r;
      break;
    }
  }
  if (blockIdx.x == 0) {
    if (tid < 12) a.dev->pose[tid] = s_pose[tid];
    if (tid == 0) {
      a.dev->round = r;
      a.dev->stop = (s_stop == 1);
      a.dev->prev_chi = prev_chi;
      a.dev->rel_tol = a.rel_tol;
      if (peers) *(volatile unsigned*)&me->seq = mb_seq0 + (unsigned)r;
    }
  }
}

// ------------------------------------------------------------------ persistent streaming kernel (all rounds, one launch)
// For a correspondence set too large for shared memory (the 10 M frame on 1, 2 or 4 GPUs): the tile ring, the
// producer lane and the consumer arithmetic of picp_linearize_kernel, wrapped in the round loop of the resident
// kernel.  The packed planes do not depend on the pose, so the producer simply keeps streaming - round after round,
// the first tiles of round r+1 already in the ring while the consumers reduce, exchange and solve round r - and the
// serial tail of the one-launch-per-round design (ticket, last-CTA pass 2, solve, launch hand-over: ~7.6 us) is
// replaced by the one all-to-all exchange of round_exchange_and_solve (~2 us).  With the device-side convergence
// test (rel_tol >= 0) the producer waits for a round's verdict before it fetches the next round's tiles.
struct StreamArgs {
  const float* pk;
  long long n;
  long long stride;
  PicpCam cam;
  float thr, damping, rel_tol;
  int n_rounds;
  PicpDev* dev;
  unsigned long long* ll;
  unsigned ll_seq0;
  int ll_stride;
  int peer_n, peer_rank;
  VoMailbox* peers[VO_MAX_PEERS];
  int use_pose0;  // as in ResArgs
  float pose0[12];
};

__device__ __forceinline__ bool mbar_try_wait_hint(unsigned long long* bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(20000u)
      : "memory");
  return ok != 0;
}

template <bool KEEP, bool PINHOLE>
__global__ void __launch_bounds__(kLinThreads, 1) picp_stream_rounds_kernel(const StreamArgs a) {
  extern __shared__ __align__(128) float s_tiles[];
  __shared__ __align__(8) unsigned long long s_full[kStages], s_empty[kStages];
  __shared__ float s_pose[12];
  __shared__ float s_part[kWarps][kSlots];
  __shared__ double s_fin[kWarps][kSlots];
  __shared__ double s_tot[kSlots];
  __shared__ float s_dx[6];
  __shared__ int s_stop;        // 0 go on, 1 converged, 2 a wait timed out
  __shared__ int s_round_done;  // rounds whose verdict is in s_stop (read by the producer)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
#pragma unroll
    for (int st = 0; st < kStages; ++st) {
      mbar_init(&s_full[st], 1);
      mbar_init(&s_empty[st], kWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    s_stop = 0;
    s_round_done = 0;
  }
  if (tid < 12) s_pose[tid] = a.use_pose0 ? a.pose0[tid] : a.dev->pose[tid];
  __syncthreads();
  // This CTA's share: a contiguous range of quads, the same for every CTA to within one quad (tiles dealt out round
  // robin leave 47 of 148 CTAs with 51 tiles and the rest with 50 at 10,485,760 correspondences - and 26 / 25, 13 / 12
  // on a 2- / 4-GPU shard - and every round ends when the longest CTA does).  The range is cut into tiles of kTile
  // correspondences; only its last tile can be short.
  const long long n_quads = (a.n + 3) >> 2;
  const long long c_begin = 4 * ((long long)blockIdx.x * n_quads / gridDim.x);
  const long long c_end = 4 * ((long long)(blockIdx.x + 1) * n_quads / gridDim.x);  // (padded to whole quads)
  const int my_tiles = (int)((c_end - c_begin + kTile - 1) / kTile);
  const long long c_lim = c_end < a.n ? c_end : a.n;  // the valid correspondences of the range end here

  if (warp == kWarps) {
    // ---------------- producer: one elected lane streams every round's tiles through the ring
    if (lane != 0) return;
    long long it = 0;
    bool aborted = false;
    auto issue = [&](int j, long long i) {
      const int stage = (int)(i % kStages);
      const long long first = c_begin + (long long)j * kTile;
      const long long left = c_end - first;
      const unsigned bytes = (unsigned)((left < kTile ? left : kTile) * sizeof(float));
      float* dst = s_tiles + (size_t)stage * 5 * kTile;
      mbar_arrive_expect_tx(&s_full[stage], 5u * bytes);
#pragma unroll
      for (int p = 0; p < 5; ++p) bulk_g2s(dst + p * kTile, a.pk + p * a.stride + first, bytes, &s_full[stage]);
    };
    for (int r = 0; r < a.n_rounds && !aborted; ++r) {
      if (a.rel_tol >= 0.f && r > 0) {  // the verdict of round r-1 first: never fetch a round that will not run
        while (*(volatile int*)&s_round_done < r && *(volatile int*)&s_stop != 2) {}
        if (*(volatile int*)&s_stop) break;
      }
      for (int j = 0; j < my_tiles; ++j, ++it) {
        if (it >= kStages) {
          const int stage = (int)(it % kStages);
          const unsigned phase = (unsigned)(it / kStages) & 1u;
          while (!mbar_try_wait_hint(&s_empty[stage], phase ^ 1u)) {
            if (*(volatile int*)&s_stop == 2) {  // the consumers gave up (a wait for another CTA / GPU expired)
              aborted = true;
              break;
            }
          }
          if (aborted) break;
        }
        issue(j, it);
      }
    }
    if (aborted) {  // copies in flight must land before the CTA may exit
      const long long first = it > kStages ? it - kStages : 0;
      for (long long j = first; j < it; ++j) mbar_wait(&s_full[j % kStages], (unsigned)(j / kStages) & 1u);
    }
    return;
  }

  // ---------------- consumers
  VoMailbox* me = a.peer_n > 1 ? a.peers[a.peer_rank] : nullptr;
  const unsigned mb_seq0 = me ? *(volatile unsigned*)&me->seq : 0u;
  float prev_chi = FLT_MAX;  // (lane 0 of warp 0)
  RoundCtx rc;
  rc.dev = a.dev; rc.ll = a.ll; rc.ll_seq0 = a.ll_seq0; rc.ll_stride = a.ll_stride; rc.damping = a.damping;
  rc.rel_tol = a.rel_tol; rc.peer_n = a.peer_n; rc.peer_rank = a.peer_rank; rc.peers = a.peers;
  // ring position of the consumers (carried across the rounds) and this thread's quad slot, as shared-window addresses
  unsigned stage = 0, phase = 0;
  const unsigned full_s = smem_u32(s_full), empty_s = smem_u32(s_empty);
  const unsigned slot_s = smem_u32(s_tiles) + 16u * (unsigned)tid;
  constexpr unsigned kPlaneBytes = kTile * sizeof(float), kStageBytes = 5 * kPlaneBytes;
  const int n_full_i = (int)((c_lim - c_begin) / kTile);  // tiles of this CTA with kTile valid correspondences
  int r = 0;
  for (; r < a.n_rounds; ++r) {
    float T[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) T[i] = s_pose[i];
    f2 acc2[29];
#pragma unroll
    for (int i = 0; i < 29; ++i) acc2[i] = 0ull;
    int n_in = 0, n_out = 0;
    for (int t = 0; t < my_tiles; ++t) {
      mbar_wait_s(full_s + 8u * stage, phase);
      const unsigned q = slot_s + stage * kStageBytes;
      const unsigned empty = empty_s + 8u * stage;
      if (++stage == kStages) {
        stage = 0;
        phase ^= 1u;
      }
      float4 wx, wy, wz, zu, zv;
      int s0, s1, s2, s3;
      if (t < n_full_i) {  // every tile but the last of the range: no per-point validity
        wx = lds128(q);
        wy = lds128(q + kPlaneBytes);
        wz = lds128(q + 2 * kPlaneBytes);
        zu = lds128(q + 3 * kPlaneBytes);
        zv = lds128(q + 4 * kPlaneBytes);
        __syncwarp();
        if (lane == 0) mbar_arrive_s(empty);  // values are in registers: hand the stage back
        picp_quad<KEEP, PINHOLE>(a.cam, T, a.thr, wx, wy, wz, zu, zv, true, true, true, acc2, n_in, n_out, s0, s1, s2, s3);
      } else {
        const long long base = c_begin + (long long)t * kTile + 4ll * tid;
        const bool any = base < c_lim;
        if (any) {
          wx = lds128(q);
          wy = lds128(q + kPlaneBytes);
          wz = lds128(q + 2 * kPlaneBytes);
          zu = lds128(q + 3 * kPlaneBytes);
          zv = lds128(q + 4 * kPlaneBytes);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive_s(empty);
        if (!any) continue;
        const long long left = c_lim - base;  // < 4 only in the last quad of the whole set
        picp_quad<KEEP, PINHOLE>(a.cam, T, a.thr, wx, wy, wz, zu, zv, left > 1, left > 2, left > 3, acc2, n_in, n_out, s0,
                                 s1, s2, s3);
      }
    }
    {  // warp reduction: lane L ends up with the warp's total of term L (per-warp counts are exact in float)
      float v[32];
#pragma unroll
      for (int i = 0; i < 29; ++i) v[i] = picp_slot_total<PINHOLE>(acc2, i);
      v[29] = (float)n_in;
      v[30] = (float)n_out;
      v[31] = 0.f;
      s_part[warp][lane] = warp_sum32_scatter(v, lane);
    }
    round_exchange_and_solve<1, kThreads, true>(rc, r, mb_seq0, me, s_part, s_fin, s_tot, s_pose, s_dx, &s_stop, prev_chi);
    if (tid == 0) {
      __threadfence_block();
      *(volatile int*)&s_round_done = r + 1;
    }
    if (s_stop) {
      ++r;
      break;
    }
  }
  if (blockIdx.x == 0) {
    if (tid < 12) a.dev->pose[tid] = s_pose[tid];
    if (tid == 0) {
      a.dev->round = r;
      a.dev->stop = (s_stop == 1);
      a.dev->prev_chi = prev_chi;
      a.dev->rel_tol = a.rel_tol;
      if (me) *(volatile unsigned*)&me->seq = mb_seq0 + (unsigned)r;
    }
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
  if (i >= n) {
    if (i < ((n + 3) & ~3ll)) {  // padding lanes of the last quad: finite values, masked out by the consumer
      pk[i] = 0.f; pk[stride + i] = 0.f; pk[2 * stride + i] = 0.f; pk[3 * stride + i] = 0.f; pk[4 * stride + i] = 0.f;
    }
    return;
  }
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

// range check of a correspondence set without packing it (host-buffer entry point; the planes are built lazily)
__global__ void __launch_bounds__(256) picp_check_kernel(const int2* __restrict__ pairs, long long n, long long n_world,
                                                         long long n_image, PicpDev* dev) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int2 pr = __ldg(pairs + i);
  if (pr.x < 0 || pr.x >= n_image || pr.y < 0 || pr.y >= n_world) dev->bad_index = 1;
}

// exhaustive check of the reciprocal shortcut: every float bit pattern inside the gate of pair_front / picp_project
// against __frcp_rn and IEEE 1.f / z.  out: [0] inputs inside the gate, [1] mismatches of the packed form,
// [2] mismatches of the scalar form (vo_device.cuh picp_project), [3] first mismatching input + 1
__global__ void __launch_bounds__(256) picp_rcp_selftest_kernel(unsigned long long* out) {
  unsigned long long in_gate = 0, bad_packed = 0, bad_scalar = 0;
  const unsigned long long total = 1ull << 32, step = (unsigned long long)gridDim.x * blockDim.x * 2;
  for (unsigned long long i = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) * 2; i < total; i += step) {
    const float z0 = __uint_as_float((unsigned)i), z1 = __uint_as_float((unsigned)(i + 1));
    float p0, p1;
    unpack2(rcp2_newton(pack2(z0, z1), z0, z1), p0, p1);
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const float z = k ? z1 : z0, p = k ? p1 : p0;
      if (!((z >= 1e-30f) && (z <= 1e30f))) continue;
      ++in_gate;
      const float ref = __frcp_rn(z), ref2 = __fdiv_rn(1.f, z);
      float r;
      asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(z));
      const float sc = fmaf(r, fmaf(-z, r, 1.f), r);
      const bool bp = __float_as_uint(p) != __float_as_uint(ref) || __float_as_uint(ref) != __float_as_uint(ref2);
      const bool bs = __float_as_uint(sc) != __float_as_uint(ref);
      bad_packed += bp;
      bad_scalar += bs;
      if (bp || bs) atomicMin(out + 3, (unsigned long long)__float_as_uint(z) + 1ull);
    }
  }
  atomicAdd(out + 0, in_gate);
  if (bad_packed) atomicAdd(out + 1, bad_packed);
  if (bad_scalar) atomicAdd(out + 2, bad_scalar);
}

struct PoseArg { float p[12]; };
// the pose travels as a kernel argument: no staging buffer, no host synchronisation
__global__ void picp_set_pose_kernel(PicpDev* dev, PoseArg pose) {
  if (threadIdx.x < 12) dev->pose[threadIdx.x] = pose.p[threadIdx.x];
}

__global__ void picp_reset_kernel(PicpDev* dev, float rel_tol) {
  dev->round = 0;
  dev->stop = 0;
  dev->prev_chi = FLT_MAX;
  dev->rel_tol = rel_tol;
}

template <bool KEEP, bool STATUS, bool PINHOLE>
cudaError_t lin_opt_in() {
  return cudaFuncSetAttribute(picp_linearize_kernel<KEEP, STATUS, PINHOLE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)kLinSmemBytes);
}

cudaError_t lin_opt_in_all() {
  cudaError_t e = cudaSuccess;
  if (e == cudaSuccess) e = lin_opt_in<false, false, false>();
  if (e == cudaSuccess) e = lin_opt_in<false, false, true>();
  if (e == cudaSuccess) e = lin_opt_in<false, true, false>();
  if (e == cudaSuccess) e = lin_opt_in<false, true, true>();
  if (e == cudaSuccess) e = lin_opt_in<true, false, false>();
  if (e == cudaSuccess) e = lin_opt_in<true, false, true>();
  if (e == cudaSuccess) e = lin_opt_in<true, true, false>();
  if (e == cudaSuccess) e = lin_opt_in<true, true, true>();
  const int res_smem = (int)(kResMaxQuads * 5 * sizeof(float4));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(picp_resident_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, res_smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(picp_resident_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, res_smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(picp_resident_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, res_smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(picp_resident_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, res_smem);
  const int lin_smem = (int)kLinSmemBytes;
  if (e == cudaSuccess) e = cudaFuncSetAttribute(picp_stream_rounds_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, lin_smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(picp_stream_rounds_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lin_smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(picp_stream_rounds_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, lin_smem);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(picp_stream_rounds_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lin_smem);
  return e;
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
  const int32_t* d_pairs_cur = nullptr;  // the current correspondence set (borrowed, or d_pairs)
  bool packed = false;                   // d_pk holds the planes of the current set
  unsigned long long* d_ll = nullptr;    // resident kernel: [2][kResMaxGrid][32] self-validating words
  unsigned ll_seq = 0;                   // sequence number of the last round exchanged through d_ll
  int mode = VO_PICP_MODE_AUTO;
  // vo_picp_set_pose only records the pose: the persistent kernels take it as a launch argument (one tiny kernel and
  // one launch gap less per frame), every other reader of dev->pose flushes it first (flush_pose)
  bool pose_pending = false;
  float h_pose[12] = {0};
};

namespace {

int flush_pose(vo_picp* s) {
  if (!s->pose_pending) return VO_OK;
  PoseArg a;
  memcpy(a.p, s->h_pose, sizeof(a.p));
  picp_set_pose_kernel<<<1, 32, 0, s->ctx->stream>>>(s->d_dev, a);
  VO_CHECK_LAUNCH(s->ctx, "picp_set_pose_kernel");
  s->pose_pending = false;
  return VO_OK;
}

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
  long long g = (s->n_pairs + kTile - 1) / kTile;  // one tile = kTile correspondences
  if (g < 1) g = 1;
  if (g > s->max_grid) g = s->max_grid;
  return (int)g;
}

template <bool KEEP, bool STATUS, bool PINHOLE>
void launch_lin3(int grid, cudaStream_t st, const LinArgs& a, bool pdl) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kLinThreads);
  cfg.dynamicSmemBytes = kLinSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, picp_linearize_kernel<KEEP, STATUS, PINHOLE>, a);
}

template <bool KEEP, bool STATUS>
void launch_lin2(bool pinhole, int grid, cudaStream_t st, const LinArgs& a, bool pdl) {
  if (pinhole) launch_lin3<KEEP, STATUS, true>(grid, st, a, pdl);
  else launch_lin3<KEEP, STATUS, false>(grid, st, a, pdl);
}

// the streaming kernel reads the packed planes: build them on first use (picp_pack_kernel, once per set)
int ensure_packed(vo_picp* s) {
  if (s->packed) return VO_OK;
  vo_ctx* ctx = s->ctx;
  int st = grow(ctx, (void**)&s->d_pk, &s->pk_cap, (size_t)s->n_pad * 5 * sizeof(float));
  if (st) return st;
  if (s->n_pairs) {
    const long long blocks = (s->n_pad + 255) / 256;
    picp_pack_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(reinterpret_cast<const int2*>(s->d_pairs_cur), s->n_pairs,
                                                               s->d_world, s->n_world, s->d_image, s->n_image,
                                                               s->d_pk, s->n_pad, s->d_dev);
    VO_CHECK_LAUNCH(ctx, "picp_pack_kernel");
  }
  s->packed = true;
  return VO_OK;
}

int launch_linearize(vo_picp* s, float thr, float damping, bool keep, bool status, bool fuse_solve) {
  vo_ctx* ctx = s->ctx;
  int st = ensure_packed(s);
  if (st) return st;
  st = flush_pose(s);  // this kernel reads dev->pose
  if (st) return st;
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
  a.peer_n = ctx->peer_n > 1 ? ctx->peer_n : 0;
  a.peer_rank = ctx->peer_rank;
  for (int p = 0; p < VO_MAX_PEERS; ++p) a.peers[p] = (VoMailbox*)ctx->peer_mailbox[p];
  const int grid = grid_for(s);
  // programmatic dependent launch only between the fused single-launch rounds (no NCCL call in between)
  const bool pdl = fuse_solve && !status;
  if (keep) {
    if (status) launch_lin2<true, true>(s->pinhole, grid, ctx->stream, a, pdl);
    else launch_lin2<true, false>(s->pinhole, grid, ctx->stream, a, pdl);
  } else {
    if (status) launch_lin2<false, true>(s->pinhole, grid, ctx->stream, a, pdl);
    else launch_lin2<false, false>(s->pinhole, grid, ctx->stream, a, pdl);
  }
  VO_CHECK_LAUNCH(ctx, "picp_linearize_kernel");
  return VO_OK;
}

// ---- resident path: geometry of the launch, or grid = 0 when the set does not fit / the mode forbids it
struct ResPlan { int grid, quads_per_cta; size_t smem; };

ResPlan resident_plan(const vo_picp* s, int n_rounds) {
  ResPlan p = {0, 0, 0};
  const vo_ctx* ctx = s->ctx;
  if (s->mode == VO_PICP_MODE_STREAM || s->mode == VO_PICP_MODE_STREAM_PERSISTENT) return p;
  if (ctx->nccl_comm != nullptr && ctx->peer_n <= 1) return p;  // NCCL-only exchange happens between launches
  if (s->mode == VO_PICP_MODE_AUTO && n_rounds < 2 && s->packed) return p;  // one round of a packed set: stream 20 B
  const long long n_quads = (s->n_pairs + 3) / 4;
  int max_grid = s->max_grid < kResMaxGrid ? s->max_grid : kResMaxGrid;
  long long g = (n_quads + kResThreads - 1) / kResThreads;  // at least one quad per thread and round
  if (g < 1) g = 1;
  if (g > max_grid) g = max_grid;
  long long qpc = (n_quads + g - 1) / g;
  if (qpc < 1) qpc = 1;
  if (qpc > kResMaxQuads) return p;
  p.grid = (int)g;
  p.quads_per_cta = (int)qpc;
  p.smem = (size_t)qpc * 5 * sizeof(float4);
  return p;
}

template <bool KEEP, bool PINHOLE>
cudaError_t launch_res2(const ResPlan& pl, cudaStream_t st, const ResArgs& a) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)pl.grid);
  cfg.blockDim = dim3(kResThreads);
  cfg.dynamicSmemBytes = pl.smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;  // all CTAs co-resident: they wait for each other's partials
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pl.grid > 1 ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, picp_resident_kernel<KEEP, PINHOLE>, a);
}

// n_rounds Gauss-Newton rounds (fewer if rel_tol >= 0 and the driver's test converges) in ONE launch
int launch_resident(vo_picp* s, const ResPlan& pl, float thr, float damping, bool keep, int n_rounds, float rel_tol) {
  vo_ctx* ctx = s->ctx;
  if (s->ll_seq > 0xF0000000u) {  // sequence numbers about to wrap: start over from clean words
    VO_CUDA(ctx, cudaMemsetAsync(s->d_ll, 0, sizeof(unsigned long long) * 2 * kResMaxGrid * kSlots, ctx->stream));
    s->ll_seq = 0;
  }
  ResArgs a;
  a.pairs = reinterpret_cast<const int2*>(s->d_pairs_cur);
  a.n = s->n_pairs;
  a.world = s->d_world;
  a.n_world = s->n_world;
  a.image = s->d_image;
  a.n_image = s->n_image;
  a.cam = s->cam;
  a.thr = thr;
  a.damping = damping;
  a.rel_tol = rel_tol;
  a.n_rounds = n_rounds;
  a.quads_per_cta = pl.quads_per_cta;
  a.dev = s->d_dev;
  a.ll = s->d_ll;
  a.ll_seq0 = s->ll_seq;
  a.ll_stride = kResMaxGrid;
  a.peer_n = ctx->peer_n > 1 ? ctx->peer_n : 0;
  a.peer_rank = ctx->peer_rank;
  for (int p = 0; p < VO_MAX_PEERS; ++p) a.peers[p] = (VoMailbox*)ctx->peer_mailbox[p];
  a.use_pose0 = s->pose_pending ? 1 : 0;  // (the kernel leaves the final pose in dev->pose either way)
  memcpy(a.pose0, s->h_pose, sizeof(a.pose0));
  s->pose_pending = false;
  s->ll_seq += (unsigned)n_rounds;
  cudaError_t e;
  if (keep) e = s->pinhole ? launch_res2<true, true>(pl, ctx->stream, a) : launch_res2<true, false>(pl, ctx->stream, a);
  else e = s->pinhole ? launch_res2<false, true>(pl, ctx->stream, a) : launch_res2<false, false>(pl, ctx->stream, a);
  ctx->launches++;
  if (e != cudaSuccess) return vo_set_error(ctx, VO_ERR_CUDA, "picp_resident_kernel", cudaGetErrorString(e));
  return VO_OK;
}

template <bool KEEP, bool PINHOLE>
cudaError_t launch_sr2(int grid, cudaStream_t st, const StreamArgs& a) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kLinThreads);
  cfg.dynamicSmemBytes = kLinSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;  // all CTAs co-resident: they wait for each other's partials
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = grid > 1 ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, picp_stream_rounds_kernel<KEEP, PINHOLE>, a);
}

// can this solve run as ONE persistent streaming launch?  (needs the in-kernel exchange: no NCCL-only communicator)
bool stream_rounds_ok(const vo_picp* s, int n_rounds) {
  const vo_ctx* ctx = s->ctx;
  if (ctx->nccl_comm != nullptr && ctx->peer_n <= 1) return false;
  if (s->mode == VO_PICP_MODE_STREAM_PERSISTENT) return true;
  return s->mode == VO_PICP_MODE_AUTO && n_rounds >= 2;
}

// n_rounds Gauss-Newton rounds over the packed planes in ONE persistent launch (picp_stream_rounds_kernel)
int launch_stream_rounds(vo_picp* s, float thr, float damping, bool keep, int n_rounds, float rel_tol) {
  vo_ctx* ctx = s->ctx;
  int st = ensure_packed(s);
  if (st) return st;
  if (s->ll_seq > 0xF0000000u) {
    VO_CUDA(ctx, cudaMemsetAsync(s->d_ll, 0, sizeof(unsigned long long) * 2 * kResMaxGrid * kSlots, ctx->stream));
    s->ll_seq = 0;
  }
  StreamArgs a;
  a.pk = s->d_pk;
  a.n = s->n_pairs;
  a.stride = s->n_pad;
  a.cam = s->cam;
  a.thr = thr;
  a.damping = damping;
  a.rel_tol = rel_tol;
  a.n_rounds = n_rounds;
  a.dev = s->d_dev;
  a.ll = s->d_ll;
  a.ll_seq0 = s->ll_seq;
  a.ll_stride = kResMaxGrid;
  a.peer_n = ctx->peer_n > 1 ? ctx->peer_n : 0;
  a.peer_rank = ctx->peer_rank;
  for (int p = 0; p < VO_MAX_PEERS; ++p) a.peers[p] = (VoMailbox*)ctx->peer_mailbox[p];
  a.use_pose0 = s->pose_pending ? 1 : 0;  // (the kernel leaves the final pose in dev->pose either way)
  memcpy(a.pose0, s->h_pose, sizeof(a.pose0));
  s->pose_pending = false;
  s->ll_seq += (unsigned)n_rounds;
  int grid = grid_for(s);
  if (grid > kResMaxGrid) grid = kResMaxGrid;
  cudaError_t e;
  if (keep) e = s->pinhole ? launch_sr2<true, true>(grid, ctx->stream, a) : launch_sr2<true, false>(grid, ctx->stream, a);
  else e = s->pinhole ? launch_sr2<false, true>(grid, ctx->stream, a) : launch_sr2<false, false>(grid, ctx->stream, a);
  ctx->launches++;
  if (e != cudaSuccess) return vo_set_error(ctx, VO_ERR_CUDA, "picp_stream_rounds_kernel", cudaGetErrorString(e));
  return VO_OK;
}

// one Gauss-Newton round on the stream (single GPU: 1 launch; with a communicator: 2 + all-reduce)
int enqueue_round(vo_picp* s, float thr, float damping, bool keep) {
  vo_ctx* ctx = s->ctx;
  const bool multi = ctx->nccl_comm != nullptr && ctx->peer_n <= 1;  // the fused peer exchange replaces NCCL
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
  cudaError_t e = lin_opt_in_all();
  if (e == cudaSuccess) e = cudaMalloc((void**)&s->d_partials, (size_t)s->max_grid * kSlots * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc((void**)&s->d_result, kSlots * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc((void**)&s->d_dev, sizeof(PicpDev));
  if (e == cudaSuccess) e = cudaMemsetAsync(s->d_dev, 0, sizeof(PicpDev), ctx->stream);
  const size_t ll_bytes = sizeof(unsigned long long) * 2 * kResMaxGrid * kSlots;
  if (e == cudaSuccess) e = cudaMalloc((void**)&s->d_ll, ll_bytes);
  if (e == cudaSuccess) e = cudaMemsetAsync(s->d_ll, 0, ll_bytes, ctx->stream);
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
  if (s->d_ll) cudaFree(s->d_ll);
  delete s;
  return VO_OK;
}

int vo_picp_set_pose(vo_picp* s, const float pose[12]) {
  if (!s || !pose) return VO_ERR_INVALID;
  memcpy(s->h_pose, pose, sizeof(s->h_pose));  // reaches the device with the next kernel that reads the pose
  s->pose_pending = true;
  return VO_OK;
}

int vo_picp_get_pose(vo_picp* s, float pose[12]) {
  if (!s || !pose) return VO_ERR_INVALID;
  int st = vo_ctx_activate(s->ctx);
  if (st) return st;
  st = flush_pose(s);
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
  s->packed = false;
  return VO_OK;
}

int vo_picp_set_points_dev(vo_picp* s, const float* d_world_xyz, int64_t n_world, const float* d_image_xy,
                           int64_t n_image) {
  if (!s || n_world < 0 || n_image < 0 || !d_world_xyz || !d_image_xy) return VO_ERR_INVALID;
  if ((reinterpret_cast<uintptr_t>(d_image_xy) & 7u) != 0)  // read as float2 (include/vo_b200.h, conventions)
    return vo_set_error(s->ctx, VO_ERR_INVALID, "vo_picp_set_points_dev", "d_image_xy must be 8-byte aligned");
  int st = vo_ctx_activate(s->ctx);  // the frees below must run on this context's device
  if (st) return st;
  if (s->own_points) {
    VO_CUDA(s->ctx, cudaStreamSynchronize(s->ctx->stream));
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
  s->packed = false;
  return VO_OK;
}

int vo_picp_set_correspondences_dev(vo_picp* s, const int32_t* d_pairs, int64_t n_pairs) {
  if (!s || n_pairs < 0 || (n_pairs && !d_pairs)) return VO_ERR_INVALID;
  vo_ctx* ctx = s->ctx;
  if (!s->d_world || !s->d_image) return vo_set_error(ctx, VO_ERR_STATE, "picp", "set_points not called");
  if ((reinterpret_cast<uintptr_t>(d_pairs) & 7u) != 0)  // read as int2
    return vo_set_error(ctx, VO_ERR_INVALID, "vo_picp_set_correspondences_dev", "d_pairs must be 8-byte aligned");
  int st = vo_ctx_activate(ctx);
  if (st) return st;
  // Nothing is gathered here: the resident kernel gathers straight into shared memory, the streaming kernel
  // packs its planes on first use (ensure_packed).  d_pairs stays borrowed until the next set_correspondences*.
  const long long n_pad = (n_pairs + 3) / 4 * 4;
  s->n_pad = n_pad > 0 ? n_pad : 4;
  s->n_pairs = n_pairs;
  s->d_pairs_cur = d_pairs;
  s->packed = false;
  // a fresh set starts with a clean range-check flag (the gather / pack / check kernels raise it)
  VO_CUDA(ctx, cudaMemsetAsync(&s->d_dev->bad_index, 0, sizeof(int), ctx->stream));
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
  // synchronous contract of the host entry point: indices are range-checked before it returns, and the caller
  // may release `pairs` afterwards
  if (n_pairs) {
    picp_check_kernel<<<(unsigned)((n_pairs + 255) / 256), 256, 0, ctx->stream>>>(
        reinterpret_cast<const int2*>(s->d_pairs), n_pairs, s->n_world, s->n_image, s->d_dev);
    VO_CHECK_LAUNCH(ctx, "picp_check_kernel");
  }
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

int vo_selftest_reciprocal(vo_ctx* ctx, uint64_t out[4]) {
  if (!ctx || !out) return VO_ERR_INVALID;
  int st = vo_ctx_activate(ctx);
  if (st) return st;
  void* d;
  st = vo_scratch(ctx, 64, &d);
  if (st) return st;
  const unsigned long long init[4] = {0, 0, 0, ~0ull};
  VO_CUDA(ctx, cudaMemcpyAsync(d, init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
  picp_rcp_selftest_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>((unsigned long long*)d);
  VO_CHECK_LAUNCH(ctx, "picp_rcp_selftest_kernel");
  unsigned long long h[4];
  VO_CUDA(ctx, cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  VO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  out[0] = h[0];
  out[1] = h[1];
  out[2] = h[2];
  out[3] = (h[3] == ~0ull) ? 0 : h[3];
  return VO_OK;
}

int vo_picp_pack(vo_picp* s) {
  int st = ready(s);
  if (st) return st;
  return ensure_packed(s);
}

int vo_picp_set_mode(vo_picp* s, int mode) {
  if (!s || mode < VO_PICP_MODE_AUTO || mode > VO_PICP_MODE_STREAM_PERSISTENT) return VO_ERR_INVALID;
  s->mode = mode;
  return VO_OK;
}

int vo_picp_resident_capacity(const vo_picp* s, int64_t* n_pairs_max) {
  if (!s || !n_pairs_max) return VO_ERR_INVALID;
  const int g = s->max_grid < kResMaxGrid ? s->max_grid : kResMaxGrid;
  *n_pairs_max = (int64_t)g * kResMaxQuads * 4;
  return VO_OK;
}

// the sticky per-set flags of the device state: an out-of-range index on the _dev path, an expired wait
static int check_dev_flags(vo_picp* s, const char* who) {
  vo_ctx* ctx = s->ctx;
  struct Flags { int bad_index, timeout; };
  void* h;
  int st = vo_pinned(ctx, 64, &h);
  if (st) return st;
  VO_CUDA(ctx, cudaMemcpyAsync(h, &s->d_dev->bad_index, sizeof(Flags), cudaMemcpyDeviceToHost, ctx->stream));
  VO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  const Flags f = *(const Flags*)h;
  if (f.timeout) {
    VO_CUDA(ctx, cudaMemsetAsync(&s->d_dev->timeout, 0, sizeof(int), ctx->stream));
    return vo_set_error(ctx, VO_ERR_STATE, who, "a wait inside the resident kernel timed out (CTAs / ranks out of step)");
  }
  if (f.bad_index) return vo_set_error(ctx, VO_ERR_INVALID, who, "correspondence index out of range");
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
  if (ctx->peer_n <= 1) {  // (with peers attached the kernel has already exchanged the terms)
    st = vo_comm_allreduce_f64(ctx, s->d_result, kSlots);
    if (st) return st;
  }
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
  s->last_rounds = n_rounds;
  const ResPlan pl = resident_plan(s, n_rounds);
  if (pl.grid > 0) return n_rounds ? launch_resident(s, pl, thr, damping, keep_outliers != 0, n_rounds, -1.f) : VO_OK;
  if (s->mode == VO_PICP_MODE_RESIDENT)
    return vo_set_error(s->ctx, VO_ERR_CAPACITY, "vo_picp_enqueue_rounds", "VO_PICP_MODE_RESIDENT: the set does not fit shared memory");
  if (stream_rounds_ok(s, n_rounds)) return n_rounds ? launch_stream_rounds(s, thr, damping, keep_outliers != 0, n_rounds, -1.f) : VO_OK;
  st = reset_rounds(s, -1.f);
  if (st) return st;
  for (int r = 0; r < n_rounds; ++r) {
    st = enqueue_round(s, thr, damping, keep_outliers != 0);
    if (st) return st;
  }
  return VO_OK;
}

int vo_picp_fetch_stats(vo_picp* s, vo_picp_stats* stats_out, int n_rounds) {
  if (!s) return VO_ERR_INVALID;
  vo_ctx* ctx = s->ctx;
  int st = vo_ctx_activate(ctx);
  if (st) return st;
  if (n_rounds < 0 || n_rounds > VO_PICP_MAX_ROUNDS) return VO_ERR_INVALID;
  if (ctx->peer_n > 1) {  // a fused exchange that gave up waiting for a peer invalidates the rounds
    void* hp;
    st = vo_pinned(ctx, 64, &hp);
    if (st) return st;
    VO_CUDA(ctx, cudaMemcpyAsync(hp, &((VoMailbox*)ctx->mailbox)->timeout, sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
    VO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (*(unsigned*)hp) return vo_set_error(ctx, VO_ERR_STATE, "vo_picp_fetch_stats", "peer exchange timed out (ranks out of step)");
  }
  st = check_dev_flags(s, "vo_picp_fetch_stats");
  if (st) return st;
  if (stats_out && n_rounds) {
    void* h;
    st = vo_pinned(ctx, sizeof(vo_picp_stats) * VO_PICP_MAX_ROUNDS, &h);
    if (st) return st;
    VO_CUDA(ctx, cudaMemcpyAsync(h, s->d_dev->stats, sizeof(vo_picp_stats) * n_rounds, cudaMemcpyDeviceToHost, ctx->stream));
    VO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(stats_out, h, sizeof(vo_picp_stats) * n_rounds);
  }
  return VO_OK;
}

#ifdef VO_PROFILE_STAMPS
int vo_debug_stamps(unsigned long long out[8]) {
  return cudaMemcpyFromSymbol(out, g_stamps, sizeof(unsigned long long) * 8) == cudaSuccess ? 0 : 2;
}
#endif

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
  struct Head { int round, stop; };
  void* h;
  st = vo_pinned(ctx, sizeof(vo_picp_stats) * VO_PICP_MAX_ROUNDS, &h);
  if (st) return st;
  int done = 0;
  const ResPlan pl = resident_plan(s, max_rounds);
  if (pl.grid > 0) {  // the whole driver loop in one launch: the convergence test runs inside the kernel
    st = launch_resident(s, pl, thr, damping, keep_outliers != 0, max_rounds, rel_tol);
    if (st) return st;
    st = check_dev_flags(s, "vo_picp_solve");
    if (st) return st;
    VO_CUDA(ctx, cudaMemcpyAsync(h, &s->d_dev->round, sizeof(Head), cudaMemcpyDeviceToHost, ctx->stream));
    VO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    done = ((Head*)h)->round;
  } else if (s->mode != VO_PICP_MODE_RESIDENT && stream_rounds_ok(s, max_rounds)) {
    st = launch_stream_rounds(s, thr, damping, keep_outliers != 0, max_rounds, rel_tol);
    if (st) return st;
    st = check_dev_flags(s, "vo_picp_solve");
    if (st) return st;
    VO_CUDA(ctx, cudaMemcpyAsync(h, &s->d_dev->round, sizeof(Head), cudaMemcpyDeviceToHost, ctx->stream));
    VO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    done = ((Head*)h)->round;
  } else {
    if (s->mode == VO_PICP_MODE_RESIDENT)
      return vo_set_error(ctx, VO_ERR_CAPACITY, "vo_picp_solve", "VO_PICP_MODE_RESIDENT: the set does not fit shared memory");
    st = reset_rounds(s, rel_tol);
    if (st) return st;
    const int chunk = 8;  // rounds enqueued between host polls of the device-side stop flag
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
    st = check_dev_flags(s, "vo_picp_solve");
    if (st) return st;
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
