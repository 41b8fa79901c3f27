// geometry.cu — batched per-correspondence geometry kernels on sm_100a.
//
//   vo_triangulate         Cam::triangulatePoints (reference src/cam.cpp:94-140): OpenCV's DLT
//                          (4x4 system in double, right singular vector of sigma_min) per pair,
//                          float32 dehomogenisation (convertPointsFromHomogeneous).
//   vo_essential_recover   Cam::computeEssentialAndRecoverPose (src/cam.cpp:37-91): essential
//                          matrix from all matches (normalised 8-point, moments reduced on the
//                          GPU) + OpenCV recoverPose (decomposeEssentialMat, 4 candidates,
//                          per-correspondence cheirality vote).
//   vo_project_points      Camera::projectPoints (src/camera.cpp:14-35).
//   vo_anti_join           add_new_world_points (src/my_utilities.cpp:413-434).
//
// These are latency / FP64-issue bound, not HBM bound: one thread per correspondence, no shared
// staging; small dense solves (9x9 eigen, 3x3 SVD) run on one thread between the data-parallel
// passes so nothing returns to the host mid-pipeline.
#include "five_point.cuh"

#include <float.h>
#include <algorithm>
#include <math.h>

namespace {

struct ProjPair { double P1[12], P2[12]; };

__global__ void __launch_bounds__(128) triangulate_kernel(ProjPair pp, const float2* __restrict__ x1,
                                                          const float2* __restrict__ x2, long long n,
                                                          float* __restrict__ xyz) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float2 a = __ldg(x1 + i), b = __ldg(x2 + i);
  triangulate_point_dev(pp.P1, pp.P2, a.x, a.y, b.x, b.y, xyz + 3 * i);
}

// ---------------------------------------------------------------- essential: moments pass
constexpr int kEssThreads = 128;

struct EssCam { double fx, fy, cx, cy; };

__global__ void __launch_bounds__(kEssThreads) essential_moments_kernel(EssCam cam, const float2* __restrict__ x1,
                                                                       const float2* __restrict__ x2, long long n,
                                                                       double* __restrict__ partials) {
  double m[kMom];
#pragma unroll
  for (int k = 0; k < kMom; ++k) m[k] = 0.0;
  for (long long i = (long long)blockIdx.x * kEssThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kEssThreads) {
    const float2 p = __ldg(x1 + i), q = __ldg(x2 + i);
    const double a0 = ((double)p.x - cam.cx) / cam.fx, a1 = ((double)p.y - cam.cy) / cam.fy;
    const double b0 = ((double)q.x - cam.cx) / cam.fx, b1 = ((double)q.y - cam.cy) / cam.fy;
    const double r[9] = {b0 * a0, b0 * a1, b0, b1 * a0, b1 * a1, b1, a0, a1, 1.0};
    int k = 0;
#pragma unroll
    for (int u = 0; u < 9; ++u)
#pragma unroll
      for (int v = u; v < 9; ++v, ++k) m[k] += r[u] * r[v];
  }
  __shared__ double s[kEssThreads / 32][kMom];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < kMom; ++k) {
    double v = m[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) s[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < kMom) {
    double v = 0;
    for (int w = 0; w < kEssThreads / 32; ++w) v += s[w][threadIdx.x];
    partials[(size_t)blockIdx.x * kMom + threadIdx.x] = v;
  }
}

__global__ void essential_solve_kernel(const double* __restrict__ partials, int n_blocks, EssState* st) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double m[kMom];
  for (int k = 0; k < kMom; ++k) {
    double v = 0;
    for (int b = 0; b < n_blocks; ++b) v += partials[(size_t)b * kMom + k];
    m[k] = v;
  }
  essential_from_moments(m, st);
}

// recoverPose's cheirality vote: each correspondence triangulated against the 4 candidates
__global__ void __launch_bounds__(128) essential_cheirality_kernel(EssCam cam, const float2* __restrict__ x1,
                                                                  const float2* __restrict__ x2, long long n,
                                                                  EssState* st, unsigned char* __restrict__ masks) {
  __shared__ double sR[2][9];
  __shared__ double sT[3];
  if (threadIdx.x < 9) {
    sR[0][threadIdx.x] = st->R1[threadIdx.x];
    sR[1][threadIdx.x] = st->R2[threadIdx.x];
  }
  if (threadIdx.x < 3) sT[threadIdx.x] = st->t[threadIdx.x];
  __syncthreads();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int ok[4] = {0, 0, 0, 0};
  if (i < n) {
    const float2 p = __ldg(x1 + i), q = __ldg(x2 + i);
    const double a0 = ((double)p.x - cam.cx) / cam.fx, a1 = ((double)p.y - cam.cy) / cam.fy;
    const double b0 = ((double)q.x - cam.cx) / cam.fx, b1 = ((double)q.y - cam.cy) / cam.fy;
    for (int c = 0; c < 4; ++c) {
      ok[c] = cheirality_ok(sR[0], sR[1], sT, c, a0, a1, b0, b1);
      masks[(size_t)c * n + i] = ok[c] ? 255 : 0;
    }
  }
  for (int c = 0; c < 4; ++c) {
    const int cnt = __syncthreads_count(ok[c]);
    if (threadIdx.x == 0 && cnt) atomicAdd(&st->good[c], cnt);
  }
}

__global__ void essential_pick_kernel(EssState* st) {
  if (threadIdx.x != 0) return;
  const int* g = st->good;
  const int pick = recover_pose_pick(g);
  const double* R = (pick & 1) ? st->R2 : st->R1;
  const double sg = (pick < 2) ? 1.0 : -1.0;
  for (int k = 0; k < 9; ++k) st->R[k] = R[k];
  for (int k = 0; k < 3; ++k) st->tt[k] = sg * st->t[k];
  st->pick = pick;
  st->n_good = g[pick];
}

// ---------------------------------------------------------------- essential: cv::findEssentialMat(RANSAC) restated
// OpenCV's loop is sequential (sample, solve, count inliers, maybe shrink the iteration budget); here a batch of
// samples is solved and scored in parallel (one thread per minimal sample, one CTA per hypothesis for the Sampson
// counts) and ONE thread then replays OpenCV's bookkeeping over the batch in sample order, so the winner, the
// iteration count and the tie-breaks are those of the sequential loop (ptsetreg.cpp RANSACPointSetRegistrator::run).
constexpr int kMaxModels = 10;

struct Ransac5State {
  int iter;      // samples consumed (OpenCV's `iter`)
  int niters;    // current iteration budget
  int max_good;
  int done;
  int best_sample, best_model;
  double E[9];
};

__global__ void ess5_init_kernel(Ransac5State* st, int max_iters) {
  if (threadIdx.x != 0) return;
  st->iter = 0;
  st->niters = max_iters > 1 ? max_iters : 1;
  st->max_good = 0;
  st->done = 0;
  st->best_sample = st->best_model = -1;
  for (int k = 0; k < 9; ++k) st->E[k] = 0;
}

// (x - c) / f exactly as findEssentialMat's matrix expression evaluates it: x * (1/f) + (-c * (1/f))
__device__ __forceinline__ double ess5_norm(float x, double inv_f, double c) { return (double)x * inv_f + (-c * inv_f); }

__global__ void __launch_bounds__(32) ess5_solve_kernel(EssCam cam, const float2* __restrict__ x1,
                                                       const float2* __restrict__ x2, const int* __restrict__ subsets,
                                                       int s0, int count, double* __restrict__ models,
                                                       int* __restrict__ n_models, const Ransac5State* st) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= count || st->done) return;
  const double ifx = 1.0 / cam.fx, ify = 1.0 / cam.fy;
  double q1[10], q2[10];
  for (int i = 0; i < 5; ++i) {
    const int id = subsets[5 * (s0 + b) + i];
    const float2 p = __ldg(x1 + id), q = __ldg(x2 + id);
    q1[2 * i] = ess5_norm(p.x, ifx, cam.cx);
    q1[2 * i + 1] = ess5_norm(p.y, ify, cam.cy);
    q2[2 * i] = ess5_norm(q.x, ifx, cam.cx);
    q2[2 * i + 1] = ess5_norm(q.y, ify, cam.cy);
  }
  n_models[b] = five_point_dev(q1, q2, models + (size_t)b * kMaxModels * 9);
}

// blockIdx.x = sample * kMaxModels + model: number of correspondences with Sampson error <= t (float compare)
__global__ void __launch_bounds__(128) ess5_score_kernel(EssCam cam, const float2* __restrict__ x1,
                                                        const float2* __restrict__ x2, long long n,
                                                        const double* __restrict__ models, const int* __restrict__ n_models,
                                                        float t, int* __restrict__ counts, const Ransac5State* st) {
  const int b = blockIdx.x / kMaxModels, m = blockIdx.x % kMaxModels;
  if (st->done || m >= n_models[b]) return;
  __shared__ double sE[9];
  __shared__ int s_cnt;
  if (threadIdx.x < 9) sE[threadIdx.x] = models[((size_t)b * kMaxModels + m) * 9 + threadIdx.x];
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  const double ifx = 1.0 / cam.fx, ify = 1.0 / cam.fy;
  int good = 0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const float2 p = __ldg(x1 + i), q = __ldg(x2 + i);
    good += sampson_dev(sE, ess5_norm(p.x, ifx, cam.cx), ess5_norm(p.y, ify, cam.cy), ess5_norm(q.x, ifx, cam.cx),
                        ess5_norm(q.y, ify, cam.cy)) <= t;
  }
  good = warp_sum_i(good);
  if ((threadIdx.x & 31) == 0 && good) atomicAdd(&s_cnt, good);  // integer sum: order independent
  __syncthreads();
  if (threadIdx.x == 0) counts[blockIdx.x] = s_cnt;
}

// OpenCV's sequential bookkeeping over one batch (ptsetreg.cpp:run): a hypothesis replaces the best one only when it
// has strictly more inliers (the first best wins ties), every improvement shrinks the iteration budget, and sample
// `iter` is looked at only while iter < niters
__global__ void ess5_select_kernel(Ransac5State* st, const double* __restrict__ models, const int* __restrict__ n_models,
                                   const int* __restrict__ counts, int s0, int count, long long n, double prob) {
  if (threadIdx.x != 0 || st->done) return;
  for (int b = 0; b < count; ++b) {
    if (st->iter >= st->niters) break;
    for (int m = 0; m < n_models[b]; ++m) {
      const int good = counts[b * kMaxModels + m];
      if (good > max(st->max_good, 4)) {
        st->max_good = good;
        st->best_sample = s0 + b;
        st->best_model = m;
        for (int k = 0; k < 9; ++k) st->E[k] = models[((size_t)b * kMaxModels + m) * 9 + k];
        st->niters = ransac_update_num_iters(prob, (double)(n - good) / (double)n, 5, st->niters);
      }
    }
    st->iter++;
  }
  if (st->iter >= st->niters) st->done = 1;
}

// inlier mask of the winning hypothesis (findEssentialMat's optional mask output: 0 / 1)
__global__ void __launch_bounds__(128) ess5_mask_kernel(EssCam cam, const float2* __restrict__ x1,
                                                       const float2* __restrict__ x2, long long n,
                                                       const Ransac5State* st, float t, unsigned char* __restrict__ mask) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double ifx = 1.0 / cam.fx, ify = 1.0 / cam.fy;
  const float2 p = __ldg(x1 + i), q = __ldg(x2 + i);
  mask[i] = sampson_dev(st->E, ess5_norm(p.x, ifx, cam.cx), ess5_norm(p.y, ify, cam.cy), ess5_norm(q.x, ifx, cam.cx),
                        ess5_norm(q.y, ify, cam.cy)) <= t;
}

__global__ void ess5_decompose_kernel(const Ransac5State* rs, EssState* st) {
  if (threadIdx.x == 0 && blockIdx.x == 0) essential_decompose_dev(rs->E, st);
}

// ---------------------------------------------------------------- Camera::projectPoints
struct ProjCam { float K[9]; float T[12]; float umax, vmax; };

// camera.h:24-36 in the reference's float32 evaluation order (no FMA)
__device__ __forceinline__ bool project_point_dev(const ProjCam& c, float px, float py, float pz, float& u, float& v) {
  const float c0 = __fadd_rn(c.T[3], dot3_rn(c.T[0], px, c.T[1], py, c.T[2], pz));
  const float c1 = __fadd_rn(c.T[7], dot3_rn(c.T[4], px, c.T[5], py, c.T[6], pz));
  const float c2 = __fadd_rn(c.T[11], dot3_rn(c.T[8], px, c.T[9], py, c.T[10], pz));
  if (c2 <= 0.f) return false;
  const float q0 = dot3_rn(c.K[0], c0, c.K[1], c1, c.K[2], c2);
  const float q1 = dot3_rn(c.K[3], c0, c.K[4], c1, c.K[5], c2);
  const float q2 = dot3_rn(c.K[6], c0, c.K[7], c1, c.K[8], c2);
  const float iz = __frcp_rn(q2);
  u = __fmul_rn(q0, iz);
  v = __fmul_rn(q1, iz);
  if (u < 0.f || u > c.umax) return false;
  if (v < 0.f || v > c.vmax) return false;
  return true;
}

__global__ void __launch_bounds__(256) project_points_kernel(ProjCam cam, const float* __restrict__ world, long long n,
                                                             float2* __restrict__ uv, unsigned char* __restrict__ flags,
                                                             int* __restrict__ block_counts) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  int inside = 0;
  if (i < n) {
    float u = -1.f, v = -1.f;
    inside = project_point_dev(cam, __ldg(world + 3 * i), __ldg(world + 3 * i + 1), __ldg(world + 3 * i + 2), u, v);
    uv[i] = inside ? make_float2(u, v) : make_float2(-1.f, -1.f);
    flags[i] = (unsigned char)inside;
  }
  const int cnt = __syncthreads_count(inside);
  if (threadIdx.x == 0) block_counts[blockIdx.x] = cnt;
}

__global__ void __launch_bounds__(256) compact_uv_kernel(const float2* __restrict__ uv, const unsigned char* __restrict__ flags,
                                                         const int* __restrict__ block_offsets, long long n,
                                                         float2* __restrict__ out) {
  __shared__ int s_warp[8];
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  const int f = (i < n) ? flags[i] : 0;
  const unsigned bal = __ballot_sync(0xffffffffu, f);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) s_warp[warp] = __popc(bal);
  __syncthreads();
  int off = block_offsets[blockIdx.x];
  for (int w = 0; w < warp; ++w) off += s_warp[w];
  off += __popc(bal & ((1u << lane) - 1u));
  if (f) out[off] = uv[i];
}

// ---------------------------------------------------------------- add_new_world_points anti-join
__global__ void __launch_bounds__(256) anti_join_kernel(const int* __restrict__ matched, long long n_matched,
                                                        const int* __restrict__ cand, long long n_cand,
                                                        unsigned char* __restrict__ keep, unsigned long long* n_keep) {
  const long long j = (long long)blockIdx.x * 256 + threadIdx.x;
  int k = 0;
  if (j < n_cand) {
    const int id = cand[j];
    bool found = false;
    for (long long i = 0; i < n_matched && !found; ++i) found = (__ldg(matched + i) == id);
    k = !found;
    keep[j] = (unsigned char)k;
  }
  const int c = __syncthreads_count(k);
  if (threadIdx.x == 0 && c) atomicAdd(n_keep, (unsigned long long)c);
}

// P = K * T^-1[0:3,:]: float32 cv::Mat product (cv::gemm accumulates float products in double and
// rounds to float), T^-1 from the isometry inverse (cam.cpp:108-112)
void projection_matrix(const float K[9], const float T[12], double P[12]) {
  float Ti[12];
  vo_pose_inverse(T, Ti);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 4; ++j) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += (double)K[3 * i + k] * (double)Ti[4 * k + j];
      P[4 * i + j] = (double)(float)s;
    }
}

}  // namespace

extern "C" {

int vo_triangulate_dev(vo_ctx* ctx, const float K[9], const float T1[12], const float T2[12], const float* d_x1,
                       const float* d_x2, int64_t n, float* d_xyz_out) {
  if (!ctx) return VO_ERR_INVALID;
  int st = vo_ctx_activate(ctx);
  if (st) return st;
  VO_REQUIRE(ctx, K && T1 && T2 && n >= 0, "vo_triangulate: arguments");
  if (n == 0) return VO_OK;  // cam.cpp:103-106: nothing to do
  VO_REQUIRE(ctx, d_x1 && d_x2 && d_xyz_out, "vo_triangulate: null buffers");
  VO_REQUIRE(ctx, ((reinterpret_cast<uintptr_t>(d_x1) | reinterpret_cast<uintptr_t>(d_x2)) & 7u) == 0,
             "vo_triangulate_dev: d_x1 / d_x2 must be 8-byte aligned");  // read as float2
  ProjPair pp;
  projection_matrix(K, T1, pp.P1);
  projection_matrix(K, T2, pp.P2);
  triangulate_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(
      pp, reinterpret_cast<const float2*>(d_x1), reinterpret_cast<const float2*>(d_x2), n, d_xyz_out);
  VO_CHECK_LAUNCH(ctx, "triangulate_kernel");
  return VO_OK;
}

int vo_triangulate(vo_ctx* ctx, const float K[9], const float T1[12], const float T2[12], const float* x1,
                   const float* x2, int64_t n, float* xyz_out) {
  if (!ctx) return VO_ERR_INVALID;
  int st = vo_ctx_activate(ctx);
  if (st) return st;
  VO_REQUIRE(ctx, n >= 0, "vo_triangulate: n");
  if (n == 0) return VO_OK;
  VO_REQUIRE(ctx, x1 && x2 && xyz_out, "vo_triangulate: null buffers");
  char* base;
  const size_t bx = vo_align_up((size_t)n * 8, 256);
  st = vo_scratch(ctx, 2 * bx + (size_t)n * 12, (void**)&base);
  if (st) return st;
  float* dx1 = (float*)base;
  float* dx2 = (float*)(base + bx);
  float* dout = (float*)(base + 2 * bx);
  VO_CUDA(ctx, cudaMemcpyAsync(dx1, x1, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
  VO_CUDA(ctx, cudaMemcpyAsync(dx2, x2, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
  st = vo_triangulate_dev(ctx, K, T1, T2, dx1, dx2, n, dout);
  if (st) return st;
  VO_CUDA(ctx, cudaMemcpyAsync(xyz_out, dout, (size_t)n * 12, cudaMemcpyDeviceToHost, ctx->stream));
  VO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return VO_OK;
}

int vo_essential_recover_ex(vo_ctx* ctx, const float K[9], const float* x1, const float* x2, int64_t n, int method,
                            double prob, double threshold, int max_iters, double E[9], double R[9], double t[3],
                            uint8_t* mask, int* n_good, uint8_t* ransac_mask, int* ransac_inliers, int* ransac_iters) {
  if (!ctx) return VO_ERR_INVALID;
  int st = vo_ctx_activate(ctx);
  if (st) return st;
  VO_REQUIRE(ctx, K && x1 && x2, "vo_essential_recover: null buffers");
  VO_REQUIRE(ctx, method == VO_ESSENTIAL_RANSAC5 || method == VO_ESSENTIAL_LINEAR8, "vo_essential_recover: method");
  const bool ransac = method == VO_ESSENTIAL_RANSAC5;
  if (ransac) {
    VO_REQUIRE(ctx, n >= 5, "vo_essential_recover: at least 5 correspondences");
    VO_REQUIRE(ctx, n < 0x7fffffff, "vo_essential_recover: too many correspondences");
    VO_REQUIRE(ctx, prob > 0 && prob < 1 && threshold > 0 && max_iters >= 1 && max_iters <= 100000, "vo_essential_recover: RANSAC parameters");
  } else {
    VO_REQUIRE(ctx, n >= 8, "vo_essential_recover: at least 8 correspondences");
  }
  if (ransac_inliers) *ransac_inliers = 0;
  if (ransac_iters) *ransac_iters = 0;
  const EssCam cam = {(double)K[0], (double)K[4], (double)K[2], (double)K[5]};
  long long blocks = (n + kEssThreads - 1) / kEssThreads;
  if (blocks > ctx->sm_count) blocks = ctx->sm_count;
  constexpr int kBatch0 = 32, kBatch = 128;  // samples solved per launch: most clean problems stop within the first batch
  size_t off = 0;
  auto carve = [&](size_t bytes) { size_t o = off; off = vo_align_up(off + bytes, 256); return o; };
  const size_t o_x1 = carve((size_t)n * 8), o_x2 = carve((size_t)n * 8);
  const size_t o_part = carve((size_t)blocks * kMom * 8), o_state = carve(sizeof(EssState));
  const size_t o_masks = carve((size_t)n * 4);
  const size_t o_rs = carve(sizeof(Ransac5State)), o_sub = carve(ransac ? (size_t)max_iters * 5 * 4 : 0);
  const size_t o_models = carve(ransac ? (size_t)kBatch * kMaxModels * 9 * 8 : 0);
  const size_t o_nm = carve(ransac ? (size_t)kBatch * 4 : 0), o_cnt = carve(ransac ? (size_t)kBatch * kMaxModels * 4 : 0);
  const size_t o_rmask = carve(ransac ? (size_t)n : 0);
  char* base;
  st = vo_scratch(ctx, off, (void**)&base);
  if (st) return st;
  float2* dx1 = (float2*)(base + o_x1);
  float2* dx2 = (float2*)(base + o_x2);
  double* part = (double*)(base + o_part);
  EssState* dstate = (EssState*)(base + o_state);
  unsigned char* masks = (unsigned char*)(base + o_masks);
  VO_CUDA(ctx, cudaMemcpyAsync(dx1, x1, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
  VO_CUDA(ctx, cudaMemcpyAsync(dx2, x2, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
  if (ransac) {
    Ransac5State* rs = (Ransac5State*)(base + o_rs);
    int* dsub = (int*)(base + o_sub);
    double* models = (double*)(base + o_models);
    int* n_models = (int*)(base + o_nm);
    int* counts = (int*)(base + o_cnt);
    unsigned char* rmask = (unsigned char*)(base + o_rmask);
    // the sample sequence depends only on n: cv::RNG((uint64)-1) + getSubset, generated on the host
    int total = max_iters;
    if (n == 5) total = 1;  // count == modelPoints: OpenCV solves the one sample and keeps its first model
    void* hp;
    st = vo_pinned(ctx, (size_t)total * 5 * 4 + 256, &hp);
    if (st) return st;
    int* hsub = (int*)hp;
    if (n == 5) {
      for (int i = 0; i < 5; ++i) hsub[i] = i;
    } else {
      CvRng rng(~0ull);
      for (int s_ = 0; s_ < total; ++s_) ransac_next_subset(rng, (int)n, hsub + 5 * s_);
    }
    VO_CUDA(ctx, cudaMemcpyAsync(dsub, hsub, (size_t)total * 5 * 4, cudaMemcpyHostToDevice, ctx->stream));
    VO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the pinned buffer is reused for read-backs below
    const double thr = threshold / ((cam.fx + cam.fy) / 2);
    const float tf = (n == 5) ? 3.0e38f : (float)(thr * thr);  // (n == 5: every point counts as an inlier)
    ess5_init_kernel<<<1, 32, 0, ctx->stream>>>(rs, total);
    VO_CHECK_LAUNCH(ctx, "ess5_init_kernel");
    Ransac5State hs_r;
    for (int s0 = 0; s0 < total;) {
      const int count = std::min(s0 == 0 ? kBatch0 : kBatch, total - s0);
      ess5_solve_kernel<<<(count + 31) / 32, 32, 0, ctx->stream>>>(cam, dx1, dx2, dsub, s0, count, models, n_models, rs);
      VO_CHECK_LAUNCH(ctx, "ess5_solve_kernel");
      ess5_score_kernel<<<count * kMaxModels, 128, 0, ctx->stream>>>(cam, dx1, dx2, n, models, n_models, tf, counts, rs);
      VO_CHECK_LAUNCH(ctx, "ess5_score_kernel");
      ess5_select_kernel<<<1, 32, 0, ctx->stream>>>(rs, models, n_models, counts, s0, count, n, prob);
      VO_CHECK_LAUNCH(ctx, "ess5_select_kernel");
      s0 += count;
      VO_CUDA(ctx, cudaMemcpyAsync(hp, rs, sizeof(Ransac5State), cudaMemcpyDeviceToHost, ctx->stream));
      VO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      memcpy(&hs_r, hp, sizeof(hs_r));
      if (hs_r.done) break;
    }
    if (ransac_iters) *ransac_iters = hs_r.iter;
    if (ransac_inliers) *ransac_inliers = hs_r.max_good;
    if (hs_r.max_good <= 0)  // cv::findEssentialMat returns an empty matrix; the reference exits (src/cam.cpp:56-59)
      return vo_set_error(ctx, VO_ERR_STATE, "vo_essential_recover", "RANSAC found no essential matrix");
    if (ransac_mask) {
      ess5_mask_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(cam, dx1, dx2, n, rs, tf, rmask);
      VO_CHECK_LAUNCH(ctx, "ess5_mask_kernel");
      VO_CUDA(ctx, cudaMemcpyAsync(ransac_mask, rmask, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    }
    ess5_decompose_kernel<<<1, 32, 0, ctx->stream>>>(rs, dstate);
    VO_CHECK_LAUNCH(ctx, "ess5_decompose_kernel");
  } else {
    essential_moments_kernel<<<(unsigned)blocks, kEssThreads, 0, ctx->stream>>>(cam, dx1, dx2, n, part);
    VO_CHECK_LAUNCH(ctx, "essential_moments_kernel");
    essential_solve_kernel<<<1, 32, 0, ctx->stream>>>(part, (int)blocks, dstate);
    VO_CHECK_LAUNCH(ctx, "essential_solve_kernel");
  }
  essential_cheirality_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(cam, dx1, dx2, n, dstate, masks);
  VO_CHECK_LAUNCH(ctx, "essential_cheirality_kernel");
  essential_pick_kernel<<<1, 32, 0, ctx->stream>>>(dstate);
  VO_CHECK_LAUNCH(ctx, "essential_pick_kernel");
  void* h;
  st = vo_pinned(ctx, sizeof(EssState), &h);
  if (st) return st;
  VO_CUDA(ctx, cudaMemcpyAsync(h, dstate, sizeof(EssState), cudaMemcpyDeviceToHost, ctx->stream));
  VO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  const EssState* hs = (const EssState*)h;
  if (E) memcpy(E, hs->E, sizeof(double) * 9);
  if (R) memcpy(R, hs->R, sizeof(double) * 9);
  if (t) memcpy(t, hs->tt, sizeof(double) * 3);
  if (n_good) *n_good = hs->n_good;
  const int pick = hs->pick;
  if (mask) {
    VO_CUDA(ctx, cudaMemcpyAsync(mask, masks + (size_t)pick * n, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    VO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return VO_OK;
}

int vo_essential_recover(vo_ctx* ctx, const float K[9], const float* x1, const float* x2, int64_t n, double E[9],
                         double R[9], double t[3], uint8_t* mask, int* n_good) {
  // the reference's call: cv::findEssentialMat(p1, p2, K, cv::RANSAC) with OpenCV's defaults (src/cam.cpp:49)
  return vo_essential_recover_ex(ctx, K, x1, x2, n, VO_ESSENTIAL_RANSAC5, 0.999, 1.0, 1000, E, R, t, mask, n_good,
                                 nullptr, nullptr, nullptr);
}

int vo_project_points(vo_ctx* ctx, const float K[9], int rows, int cols, const float pose[12], const float* world_xyz,
                      int64_t n, int keep_indices, float* out_uv, int64_t* n_out, int64_t* n_inside) {
  if (!ctx) return VO_ERR_INVALID;
  int st = vo_ctx_activate(ctx);
  if (st) return st;
  VO_REQUIRE(ctx, K && pose && n >= 0 && rows > 0 && cols > 0, "vo_project_points: arguments");
  if (n_out) *n_out = 0;
  if (n_inside) *n_inside = 0;
  if (n == 0) return VO_OK;
  VO_REQUIRE(ctx, world_xyz && out_uv, "vo_project_points: null buffers");
  ProjCam cam;
  for (int i = 0; i < 9; ++i) cam.K[i] = K[i];
  for (int i = 0; i < 12; ++i) cam.T[i] = pose[i];
  cam.umax = (float)(cols - 1);
  cam.vmax = (float)(rows - 1);
  const long long blocks = (n + 255) / 256;
  size_t off = 0;
  auto carve = [&](size_t bytes) { size_t o = off; off = vo_align_up(off + bytes, 256); return o; };
  const size_t o_w = carve((size_t)n * 12), o_uv = carve((size_t)n * 8), o_out = carve((size_t)n * 8);
  const size_t o_flags = carve((size_t)n), o_counts = carve((size_t)blocks * 4), o_total = carve(8);
  char* base;
  st = vo_scratch(ctx, off, (void**)&base);
  if (st) return st;
  float* dw = (float*)(base + o_w);
  float2* duv = (float2*)(base + o_uv);
  float2* dout = (float2*)(base + o_out);
  unsigned char* flags = (unsigned char*)(base + o_flags);
  int* counts = (int*)(base + o_counts);
  long long* dtotal = (long long*)(base + o_total);
  VO_CUDA(ctx, cudaMemcpyAsync(dw, world_xyz, (size_t)n * 12, cudaMemcpyHostToDevice, ctx->stream));
  project_points_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(cam, dw, n, duv, flags, counts);
  VO_CHECK_LAUNCH(ctx, "project_points_kernel");
  st = vo_scan_block_counts(ctx, counts, blocks, dtotal);
  if (st) return st;
  void* h;
  st = vo_pinned(ctx, 64, &h);
  if (st) return st;
  if (!keep_indices) {
    compact_uv_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(duv, flags, counts, n, dout);
    VO_CHECK_LAUNCH(ctx, "compact_uv_kernel");
  }
  VO_CUDA(ctx, cudaMemcpyAsync(h, dtotal, 8, cudaMemcpyDeviceToHost, ctx->stream));
  VO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  const long long inside = *(long long*)h;
  const long long rows_out = keep_indices ? n : inside;
  if (rows_out)
    VO_CUDA(ctx, cudaMemcpyAsync(out_uv, keep_indices ? (void*)duv : (void*)dout, (size_t)rows_out * 8,
                                 cudaMemcpyDeviceToHost, ctx->stream));
  VO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (n_out) *n_out = rows_out;
  if (n_inside) *n_inside = inside;
  return VO_OK;
}

int vo_anti_join(vo_ctx* ctx, const int32_t* matched_id, int64_t n_matched, const int32_t* cand_id, int64_t n_cand,
                 uint8_t* keep, int64_t* n_keep) {
  if (!ctx) return VO_ERR_INVALID;
  int st = vo_ctx_activate(ctx);
  if (st) return st;
  VO_REQUIRE(ctx, n_matched >= 0 && n_cand >= 0, "vo_anti_join: sizes");
  if (n_keep) *n_keep = 0;
  if (n_cand == 0) return VO_OK;
  VO_REQUIRE(ctx, cand_id && keep && (n_matched == 0 || matched_id), "vo_anti_join: null buffers");
  size_t off = 0;
  auto carve = [&](size_t bytes) { size_t o = off; off = vo_align_up(off + bytes, 256); return o; };
  const size_t o_m = carve((size_t)(n_matched ? n_matched : 1) * 4), o_c = carve((size_t)n_cand * 4);
  const size_t o_k = carve((size_t)n_cand), o_n = carve(8);
  char* base;
  st = vo_scratch(ctx, off, (void**)&base);
  if (st) return st;
  int* dm = (int*)(base + o_m);
  int* dc = (int*)(base + o_c);
  unsigned char* dk = (unsigned char*)(base + o_k);
  unsigned long long* dn = (unsigned long long*)(base + o_n);
  if (n_matched) VO_CUDA(ctx, cudaMemcpyAsync(dm, matched_id, (size_t)n_matched * 4, cudaMemcpyHostToDevice, ctx->stream));
  VO_CUDA(ctx, cudaMemcpyAsync(dc, cand_id, (size_t)n_cand * 4, cudaMemcpyHostToDevice, ctx->stream));
  VO_CUDA(ctx, cudaMemsetAsync(dn, 0, 8, ctx->stream));
  anti_join_kernel<<<(unsigned)((n_cand + 255) / 256), 256, 0, ctx->stream>>>(dm, n_matched, dc, n_cand, dk, dn);
  VO_CHECK_LAUNCH(ctx, "anti_join_kernel");
  void* h;
  st = vo_pinned(ctx, 64, &h);
  if (st) return st;
  VO_CUDA(ctx, cudaMemcpyAsync(keep, dk, (size_t)n_cand, cudaMemcpyDeviceToHost, ctx->stream));
  VO_CUDA(ctx, cudaMemcpyAsync(h, dn, 8, cudaMemcpyDeviceToHost, ctx->stream));
  VO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (n_keep) *n_keep = (int64_t) * (unsigned long long*)h;
  return VO_OK;
}

}  // extern "C"
