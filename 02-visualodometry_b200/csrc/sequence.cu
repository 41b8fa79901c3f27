// sequence.cu — batched independent sequences (BASELINE config 5).
//
// One CTA owns one sequence and runs the reference's final pipeline (exec/icp_test.cpp:40-136) for all of
// its frames without returning to the host: descriptor matching against the map, PICP rounds with the
// driver's convergence test, matching against the previous frame, the anti-join, triangulation and the
// append to the map.  The map (points + descriptors + ids) of every sequence stays resident in HBM; frame
// k+1 depends on frame k, so there is nothing to exchange between CTAs (replicas only).
//
// Every arithmetic step is the device function the stand-alone kernels use (vo_device.cuh): matching is
// bit-exact, the PICP inlier decisions are bit-exact given the pose, H/b are reduced over the CTA instead
// of sequentially (tolerance-level, like the large-N kernel).
#include "vo_device.cuh"

namespace {

constexpr int kSeqThreads = 128;  // = maximum points per frame (one thread per image point)
constexpr int kSeqWarps = kSeqThreads / 32;
constexpr int kDim = 10;
constexpr int kDimPad = 12;

struct SeqArgs {
  vo_seq_params p;
  int n_frames, max_pts, world_cap;
  const int* cnt;        // [S][F]
  const float* uv;       // [S][F][P][2]
  const float* desc;     // [S][F][P][10]
  const int* id_real;    // [S][F][P]
  float* poses;          // [S][F][12]  camera-in-world, frame 0 = identity
  float* w_xyz;          // [S][W][3]
  float* w_desc;         // [S][W][10]
  int* w_id;             // [S][W]
  int* w_cnt;            // [S]
  int* rounds;           // [S][F] nullable
  int* inliers;          // [S][F][2] nullable: (inliers of the last round, correspondences)
  int* status;           // [S]
};

struct FrameBuf {
  float desc[kSeqThreads][kDimPad];
  float2 uv[kSeqThreads];
  int id[kSeqThreads];
  int n;
};

// exclusive position of every set flag in thread order + total (stable compaction inside the CTA)
__device__ __forceinline__ int block_compact(bool flag, int* s_warp, int& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned bal = __ballot_sync(0xffffffffu, flag);
  __syncthreads();
  if (lane == 0) s_warp[warp] = __popc(bal);
  __syncthreads();
  int off = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < kSeqWarps; ++w) {
    if (w < warp) off += s_warp[w];
    tot += s_warp[w];
  }
  total = tot;
  return off + __popc(bal & ((1u << lane) - 1u));
}

__device__ __forceinline__ void load_frame(const SeqArgs& a, long long seq, int f, FrameBuf& fb) {
  const long long base = (seq * a.n_frames + f) * (long long)a.max_pts;
  const int n = min(a.cnt[seq * a.n_frames + f], kSeqThreads);
  __syncthreads();
  if (threadIdx.x == 0) fb.n = n;
  for (int t = threadIdx.x; t < n * kDim; t += kSeqThreads) {
    const int i = t / kDim, k = t - i * kDim;
    fb.desc[i][k] = __ldg(a.desc + (base + i) * kDim + k);
  }
  if ((int)threadIdx.x < n) {
    fb.uv[threadIdx.x] = __ldg(reinterpret_cast<const float2*>(a.uv) + base + threadIdx.x);
    fb.id[threadIdx.x] = __ldg(a.id_real + base + threadIdx.x);
  }
  __syncthreads();
}

// match_points of my_utilities.h:70-120 for one query row held by this thread against `n` candidates in
// shared memory (row stride kDimPad). Updates (best, second, idx); indices are offset by `first`.
__device__ __forceinline__ void scan_candidates(const float (&q)[kDim], const float* cand, int n, int first, float& best,
                                                float& second, int& idx) {
  for (int j = 0; j < n; ++j) update_best(sqdist_eigen<kDim>(q, cand + j * kDimPad), first + j, best, second, idx);
}

__device__ __forceinline__ bool accept_match(int idx, float best, float second, float dist_thr, float ratio_thr) {
  return (idx != -1) && (best < dist_thr) && (__fdiv_rn(best, second) < ratio_thr);
}

__constant__ float kIdentity12[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};

// -DVO_SEQ_STAGE_CLOCKS (exp/build_variant.sh seqclk): thread 0 of every CTA accumulates clock64() per stage of the
// frame loop: 0 frame loads, 1 match against the map, 2 PICP rounds, 3 pose + match against the previous frame +
// anti-join, 4 triangulate + append, 5 initialisation (frames 0/1).  Read with vo_debug_seq_stage_cycles.
#ifdef VO_SEQ_STAGE_CLOCKS
__device__ unsigned long long g_seq_stage[16];
#define VO_SEQ_CLK_DECL long long clk_t = clock64(); unsigned long long clk_acc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}
#define VO_SEQ_CLK(i) do { const long long n_ = clock64(); clk_acc[i] += (unsigned long long)(n_ - clk_t); clk_t = n_; } while (0)
#define VO_SEQ_CLK_FLUSH do { if (threadIdx.x == 0) for (int i_ = 0; i_ < 16; ++i_) atomicAdd(&g_seq_stage[i_], clk_acc[i_]); } while (0)
#else
#define VO_SEQ_CLK_DECL do { } while (0)
#define VO_SEQ_CLK(i) do { } while (0)
#define VO_SEQ_CLK_FLUSH do { } while (0)
#endif
// -DVO_SEQ_ROUND_CLOCKS (with VO_SEQ_STAGE_CLOCKS): thread 0 also splits a Gauss-Newton round: 8 linearize + warp
// reduction, 9 barrier 1, 10 cross-warp sums + gather to lane 0, 11 6x6 solve + verdict, 12 pose update, 13 barrier 2
#if defined(VO_SEQ_STAGE_CLOCKS) && defined(VO_SEQ_ROUND_CLOCKS)
#define VO_SEQ_RCLK(i) VO_SEQ_CLK(i)
#else
#define VO_SEQ_RCLK(i) do { } while (0)
#endif

#ifndef VO_SEQ_MINB
#define VO_SEQ_MINB 6
#endif
// MINB = resident CTAs per SM the register allocation is sized for: 6 (80 registers) for throughput when the batch fills
// the machine; 4 (128 registers, a third of the spills) when there are fewer sequences than resident slots anyway and
// only one sequence's latency counts (measured round 2: 512 sequences 12.8 -> 12.2 ms, 4096 sequences 64.2 -> 67.6 ms).
template <int MINB>
__global__ void __launch_bounds__(kSeqThreads, MINB) seq_pipeline_kernel(const SeqArgs a) {
  __shared__ FrameBuf s_curr, s_next;
  __shared__ __align__(16) float s_tile[kSeqThreads * kDimPad];  // map descriptors, 128 rows at a time
  __shared__ int2 s_iw[kSeqThreads];                              // (image idx in next, world idx)
  __shared__ int2 s_im[kSeqThreads];                              // (idx in curr, idx in next)
  __shared__ unsigned char s_matched[kSeqThreads];                // next-frame point already in the map
  __shared__ int s_warp[kSeqWarps];
  __shared__ float s_red[2][kSeqWarps][32];
  __shared__ int2 s_verdict[2];
  __shared__ float s_dx[6];
  __shared__ float s_hb[32];
  __shared__ double s_mom[kSeqWarps][kMom];
  __shared__ float s_pose[12];   // world-in-camera during PICP
  __shared__ float s_prev[12];   // camera-in-world of the previous frame
  __shared__ float s_est[12];    // camera-in-world estimate of the next frame
  __shared__ double s_P1[12], s_P2[12];
  __shared__ EssState s_ess;
  __shared__ int s_flag, s_wcnt;
  __shared__ float s_prev_chi;

  const long long seq = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  VO_SEQ_CLK_DECL;
  const PicpCam cam = {{a.p.K[0], a.p.K[1], a.p.K[2], a.p.K[3], a.p.K[4], a.p.K[5], a.p.K[6], a.p.K[7], a.p.K[8]},
                       (float)(a.p.cols - 1), (float)(a.p.rows - 1)};
  const bool pinhole = a.p.K[1] == 0.f && a.p.K[3] == 0.f && a.p.K[6] == 0.f && a.p.K[7] == 0.f && a.p.K[8] == 1.f;
  float* const w_xyz = a.w_xyz + seq * (long long)a.world_cap * 3;
  float* const w_desc = a.w_desc + seq * (long long)a.world_cap * kDim;
  int* const w_id = a.w_id + seq * (long long)a.world_cap;
  float* const poses = a.poses + seq * (long long)a.n_frames * 12;
  // (constant memory, not a per-thread array: a dynamically indexed local copy was found clobbered by the
  // cheirality stage's stack temporaries in one build - see profiles/r01_sequences.md)
#ifdef VO_SEQ_LOCAL_IDENTITY  // the round-1 form, kept for exp/seq_clobber_probe.py (profiles/r02_sequences.md)
  const float I12_local[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
  const float* const I12 = I12_local;
#else
  const float* const I12 = kIdentity12;
#endif

  if (tid < 12)
    for (int f = 0; f < a.n_frames; ++f) poses[f * 12 + tid] = I12[tid];
  if (tid == 0) {
    s_wcnt = 0;
    a.status[seq] = 0;
  }
  __syncthreads();
  if (a.n_frames < 2) return;

  // ------------------------------------------------------------ frames 0/1: match, essential, triangulate
  load_frame(a, seq, 0, s_curr);
  load_frame(a, seq, 1, s_next);
  {
    float best = FLT_MAX, second = FLT_MAX;
    int idx = -1;
    if (tid < s_curr.n) {
      float q[kDim];
#pragma unroll
      for (int k = 0; k < kDim; ++k) q[k] = s_curr.desc[tid][k];
      scan_candidates(q, &s_next.desc[0][0], s_next.n, 0, best, second, idx);
    }
    const bool acc = tid < s_curr.n && accept_match(idx, best, second, a.p.dist_thr, a.p.ratio_thr);
    int n01;
    const int pos = block_compact(acc, s_warp, n01);
    if (acc) s_im[pos] = make_int2(tid, idx);
    __syncthreads();
    if (n01 < 8) {  // the reference would exit here (empty essential matrix, cam.cpp:56-59)
      if (tid == 0) {
        a.status[seq] = 2;
        a.w_cnt[seq] = 0;
      }
      return;
    }
    // 8-point moments over the matches (double), reduced over the CTA
    const double fx = a.p.K[0], fy = a.p.K[4], cx = a.p.K[2], cy = a.p.K[5];
    double a0 = 0, a1 = 0, b0 = 0, b1 = 0;
    double m[kMom];
#pragma unroll
    for (int k = 0; k < kMom; ++k) m[k] = 0.0;
    if (tid < n01) {
      const float2 p = s_curr.uv[s_im[tid].x], q = s_next.uv[s_im[tid].y];
      a0 = ((double)p.x - cx) / fx; a1 = ((double)p.y - cy) / fy;
      b0 = ((double)q.x - cx) / fx; b1 = ((double)q.y - cy) / fy;
      const double r[9] = {b0 * a0, b0 * a1, b0, b1 * a0, b1 * a1, b1, a0, a1, 1.0};
      int k = 0;
#pragma unroll
      for (int u = 0; u < 9; ++u)
#pragma unroll
        for (int v = u; v < 9; ++v, ++k) m[k] = r[u] * r[v];
    }
#pragma unroll
    for (int k = 0; k < kMom; ++k) {
      double v = m[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) s_mom[warp][k] = v;
    }
    __syncthreads();
    if (tid == 0) {
      double tot[kMom];
      for (int k = 0; k < kMom; ++k) {
        double v = 0;
        for (int w = 0; w < kSeqWarps; ++w) v += s_mom[w][k];
        tot[k] = v;
      }
      essential_from_moments(tot, &s_ess);
    }
    __syncthreads();
    int good[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const bool ok = tid < n01 && cheirality_ok(s_ess.R1, s_ess.R2, s_ess.t, c, a0, a1, b0, b1);
      good[c] = __syncthreads_count(ok);
    }
    if (tid == 0) {
      const int pick = recover_pose_pick(good);
      const double* R = (pick & 1) ? s_ess.R2 : s_ess.R1;
      const double sg = (pick < 2) ? 1.0 : -1.0;
      float T[12];  // [R|t], cv2eigen: CV_64F -> float
      for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) T[4 * r + c] = (float)R[3 * r + c];
        T[4 * r + 3] = (float)(sg * s_ess.t[r]);
      }
      pose_inverse_dev(T, s_est);  // camera-2-in-world (cam.cpp:78-81)
      projection_matrix_dev(a.p.K, I12, s_P1);
      projection_matrix_dev(a.p.K, s_est, s_P2);
      s_wcnt = min(n01, a.world_cap);
    }
    __syncthreads();
    if (tid < n01 && tid < a.world_cap) {
      const int i0 = s_im[tid].x, i1 = s_im[tid].y;
      triangulate_point_dev(s_P1, s_P2, s_curr.uv[i0].x, s_curr.uv[i0].y, s_next.uv[i1].x, s_next.uv[i1].y, w_xyz + 3 * tid);
#pragma unroll
      for (int k = 0; k < kDim; ++k) w_desc[tid * kDim + k] = s_curr.desc[i0][k];
      w_id[tid] = s_curr.id[i0];
    }
    __syncthreads();
  }

  VO_SEQ_CLK(5);
  // ------------------------------------------------------------ frame loop (icp_test.cpp:61-136)
  for (int f = 0; f + 1 < a.n_frames; ++f) {
    load_frame(a, seq, f, s_curr);
    load_frame(a, seq, f + 1, s_next);
    VO_SEQ_CLK(0);
    const int W = s_wcnt;
    // (1) next frame against the map
    float best = FLT_MAX, second = FLT_MAX;
    int idx = -1;
    float q[kDim];
    const bool have_q = tid < s_next.n;
#pragma unroll
    for (int k = 0; k < kDim; ++k) q[k] = have_q ? s_next.desc[tid][k] : 0.f;
    for (int w0 = 0; w0 < W; w0 += kSeqThreads) {
      const int cnt = min(kSeqThreads, W - w0);
      __syncthreads();
      for (int t = tid; t < cnt * kDim; t += kSeqThreads) {
        const int j = t / kDim, k = t - j * kDim;
        s_tile[j * kDimPad + k] = w_desc[(w0 + j) * kDim + k];
      }
      __syncthreads();
      if (have_q) scan_candidates(q, s_tile, cnt, w0, best, second, idx);
    }
    const bool acc_w = have_q && accept_match(idx, best, second, a.p.dist_thr, a.p.ratio_thr);
    int C;
    const int pos_w = block_compact(acc_w, s_warp, C);
    if (acc_w) s_iw[pos_w] = make_int2(tid, idx);
    s_matched[tid] = acc_w ? 1 : 0;  // id_meas == index inside the frame
    if (tid < 12) s_prev[tid] = poses[f * 12 + tid];
    __syncthreads();
    VO_SEQ_CLK(1);
    // (2) PICP from the previous pose (icp_test.cpp:78-107)
    if (tid == 0) {
      pose_inverse_dev(s_prev, s_pose);
      s_prev_chi = FLT_MAX;
      s_flag = 0;
    }
    __syncthreads();
    float wx = 0, wy = 0, wz = 0, zu = 0, zv = 0;
    if (tid < C) {
      const int wi = s_iw[tid].y, ii = s_iw[tid].x;
      wx = w_xyz[3 * wi]; wy = w_xyz[3 * wi + 1]; wz = w_xyz[3 * wi + 2];
      zu = s_next.uv[ii].x; zv = s_next.uv[ii].y;
    }
    int rounds_done = 0, last_inl = 0;
    for (int r = 0; r < a.p.max_rounds; ++r) {
      float T[12];
#pragma unroll
      for (int i = 0; i < 12; ++i) T[i] = s_pose[i];
      float acc[32];  // 21 H + 6 b + chi_in + chi_out + n_in + n_out (small counts are exact in float) + pad
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[i] = 0.f;
      int n_in = 0, n_out = 0;
      if (tid < C) {
        const PointTerms t = pinhole ? picp_project<true>(cam, T, a.p.kernel_threshold, wx, wy, wz, zu, zv, true)
                                     : picp_project<false>(cam, T, a.p.kernel_threshold, wx, wy, wz, zu, zv, true);
        float lambda = 0.f;
        if (t.st == VO_PICP_INLIER) {
          acc[27] = t.chi; n_in = 1; lambda = 1.f;
        } else if (t.st == VO_PICP_OUTLIER) {
          acc[28] = t.chi; n_out = 1;
          if (a.p.keep_outliers) lambda = __fsqrt_rn(__fdiv_rn(a.p.kernel_threshold, t.chi));
        }
        if (lambda != 0.f) {  // J = (Jp*K)*[I | skew(-c)] for general K (tolerance-level arithmetic)
          const float iz2 = t.iz * t.iz, m0 = -t.q0 * iz2, m1 = -t.q1 * iz2;
          float J0[6], J1[6];
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            J0[j] = fmaf(t.iz, cam.K[j], m0 * cam.K[6 + j]);
            J1[j] = fmaf(t.iz, cam.K[3 + j], m1 * cam.K[6 + j]);
          }
          J0[3] = fmaf(J0[2], t.c1, -J0[1] * t.c2); J0[4] = fmaf(J0[0], t.c2, -J0[2] * t.c0); J0[5] = fmaf(J0[1], t.c0, -J0[0] * t.c1);
          J1[3] = fmaf(J1[2], t.c1, -J1[1] * t.c2); J1[4] = fmaf(J1[0], t.c2, -J1[2] * t.c0); J1[5] = fmaf(J1[1], t.c0, -J1[0] * t.c1);
          int k = 0;
#pragma unroll
          for (int i = 0; i < 6; ++i) {
            const float u0 = J0[i] * lambda, u1 = J1[i] * lambda;
#pragma unroll
            for (int j = i; j < 6; ++j, ++k) acc[k] = fmaf(u0, J0[j], u1 * J1[j]);
            acc[21 + i] = fmaf(u0, t.e0, u1 * t.e1);
          }
        }
      }
      acc[29] = (float)n_in;
      acc[30] = (float)n_out;
      const float mine = warp_sum32_scatter(acc, lane);  // lane k: this warp's total of term k
      // Two barriers per round.  The per-warp partials and the round's verdict are double-buffered by round parity,
      // so a warp that runs ahead into the next round never overwrites what a slower warp still has to read.
      const int par = r & 1;
      s_red[par][warp][lane] = mine;
      VO_SEQ_RCLK(8);
      __syncthreads();
      VO_SEQ_RCLK(9);
      if (warp == 0) {
        // lane k adds the warps' partials of term k in warp order, in double; lane 0 collects them by shuffle,
        // solves the 6x6 system, and the warp applies the increment (12 pose entries in parallel)
        double v = 0;
        int inl = 0;
#pragma unroll
        for (int w = 0; w < kSeqWarps; ++w) {
          v += (double)s_red[par][w][lane];
          inl += (int)s_red[par][w][29];
        }
        const float vf = (float)v;
        s_hb[lane] = vf;  // lane k holds term k: 21 H, 6 b, chi_in, chi_out, ...: through shared memory to the solver
        __syncwarp();
        const float cur = s_hb[27];
        VO_SEQ_RCLK(10);
        float dx[6];
        if (MINB <= 4) {  // below one wave nothing competes for the scheduler: one lane's ILP beats the shuffle chain (12.2 -> 12.0 ms)
          if (tid == 0) {
            float Hu[21], bb[6], x[6];
#pragma unroll
            for (int k = 0; k < 21; ++k) Hu[k] = s_hb[k];
#pragma unroll
            for (int k = 0; k < 6; ++k) bb[k] = s_hb[21 + k];
            picp_gn_solve(Hu, bb, a.p.damping, x);
#pragma unroll
            for (int k = 0; k < 6; ++k) s_dx[k] = x[k];
          }
          __syncwarp();
#pragma unroll
          for (int k = 0; k < 6; ++k) dx[k] = s_dx[k];
        } else {
          picp_gn_solve_warp(s_hb, a.p.damping, lane, s_dx, dx);
        }
        if (tid == 0) {
          const float prev = s_prev_chi;
          const float rel = (prev > 1e-10f) ? __fdiv_rn(fabsf(__fsub_rn(prev, cur)), prev) : 0.f;
          s_prev_chi = cur;
          s_verdict[par] = make_int2((rel < a.p.rel_tol) ? 1 : 0, inl);  // icp_test.cpp:99-106
        }
        VO_SEQ_RCLK(11);
        picp_apply_dx_warp(dx, s_pose, lane);
        VO_SEQ_RCLK(12);
      }
      __syncthreads();
      VO_SEQ_RCLK(13);
      rounds_done = r + 1;
      const int2 verdict = s_verdict[par];
      last_inl = verdict.y;
      const bool stop = verdict.x != 0;
      if (stop) break;
    }
    VO_SEQ_CLK(2);
    // (3) estimated camera-in-world pose of the next frame
    if (tid == 0) {
      pose_inverse_dev(s_pose, s_est);
      bool fin = true;
      for (int i = 0; i < 12; ++i) fin = fin && finite_f(s_est[i]);
      s_flag = fin ? 0 : 1;
      if (a.rounds) a.rounds[seq * a.n_frames + f + 1] = rounds_done;
      if (a.inliers) {
        a.inliers[(seq * a.n_frames + f + 1) * 2] = last_inl;
        a.inliers[(seq * a.n_frames + f + 1) * 2 + 1] = C;
      }
    }
    __syncthreads();
    if (s_flag) {  // tracking lost (non-finite pose): the reference would carry NaNs on; flag it and stop here
      if (tid == 0) {
        a.status[seq] = 3;
        a.w_cnt[seq] = s_wcnt;
      }
      return;
    }
    if (tid < 12) poses[(f + 1) * 12 + tid] = s_est[tid];
    // (4) current frame against the next frame
    best = FLT_MAX; second = FLT_MAX; idx = -1;
    const bool have_c = tid < s_curr.n;
    if (have_c) {
#pragma unroll
      for (int k = 0; k < kDim; ++k) q[k] = s_curr.desc[tid][k];
      scan_candidates(q, &s_next.desc[0][0], s_next.n, 0, best, second, idx);
    }
    const bool acc_m = have_c && accept_match(idx, best, second, a.p.dist_thr, a.p.ratio_thr);
    // (5) anti-join (my_utilities.cpp:413-434): keep pairs whose next-frame point is not matched to the map
    const bool fresh = acc_m && !s_matched[idx];
    int n_new;
    const int pos_n = block_compact(fresh, s_warp, n_new);
    VO_SEQ_CLK(3);
    // (6) triangulate the new pairs between the two poses and append them to the map (cam.cpp:94-140)
    if (tid == 0) {
      projection_matrix_dev(a.p.K, s_prev, s_P1);
      projection_matrix_dev(a.p.K, s_est, s_P2);
    }
    __syncthreads();
    if (fresh) {
      const int slot = W + pos_n;
      if (slot < a.world_cap) {
        triangulate_point_dev(s_P1, s_P2, s_curr.uv[tid].x, s_curr.uv[tid].y, s_next.uv[idx].x, s_next.uv[idx].y, w_xyz + 3 * slot);
#pragma unroll
        for (int k = 0; k < kDim; ++k) w_desc[slot * kDim + k] = s_curr.desc[tid][k];
        w_id[slot] = s_curr.id[tid];
      }
    }
    __syncthreads();
    if (tid == 0) {
      if (W + n_new > a.world_cap) a.status[seq] = 1;  // map capacity reached: the overflow is dropped
      s_wcnt = min(W + n_new, a.world_cap);
    }
    __threadfence_block();
    __syncthreads();
    VO_SEQ_CLK(4);
  }
  if (tid == 0) a.w_cnt[seq] = s_wcnt;
  VO_SEQ_CLK_FLUSH;
}

}  // namespace

#ifdef VO_SEQ_STAGE_CLOCKS
extern "C" int vo_debug_seq_stage_cycles(unsigned long long out[16], int reset) {
  if (cudaMemcpyFromSymbol(out, g_seq_stage, 128) != cudaSuccess) return VO_ERR_CUDA;
  if (reset) {
    unsigned long long z[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (cudaMemcpyToSymbol(g_seq_stage, z, 128) != cudaSuccess) return VO_ERR_CUDA;
  }
  return VO_OK;
}
#endif

extern "C" {

int vo_seq_batch_run_dev(vo_ctx* ctx, const vo_seq_params* params, int n_seq, int n_frames, int max_pts, int world_cap,
                         const int32_t* d_cnt, const float* d_uv, const float* d_desc, const int32_t* d_id_real,
                         float* d_poses, float* d_world_xyz, int32_t* d_world_id, int32_t* d_world_cnt,
                         int32_t* d_rounds, int32_t* d_inliers, int32_t* d_status) {
  if (!ctx) return VO_ERR_INVALID;
  int st = vo_ctx_activate(ctx);
  if (st) return st;
  VO_REQUIRE(ctx, params && n_seq >= 0 && n_frames >= 1 && world_cap >= 1, "vo_seq_batch_run: arguments");
  VO_REQUIRE(ctx, max_pts >= 1 && max_pts <= kSeqThreads, "vo_seq_batch_run: at most 128 points per frame");
  VO_REQUIRE(ctx, params->max_rounds >= 1 && params->rel_tol >= 0.f, "vo_seq_batch_run: max_rounds / rel_tol");
  if (n_seq == 0) return VO_OK;
  VO_REQUIRE(ctx, d_cnt && d_uv && d_desc && d_id_real && d_poses && d_world_xyz && d_world_id && d_world_cnt && d_status,
             "vo_seq_batch_run: null buffers");
  char* base;
  st = vo_scratch(ctx, (size_t)n_seq * world_cap * kDim * sizeof(float), (void**)&base);  // map descriptors
  if (st) return st;
  SeqArgs a;
  a.p = *params;
  a.n_frames = n_frames;
  a.max_pts = max_pts;
  a.world_cap = world_cap;
  a.cnt = d_cnt; a.uv = d_uv; a.desc = d_desc; a.id_real = d_id_real;
  a.poses = d_poses; a.w_xyz = d_world_xyz; a.w_desc = (float*)base; a.w_id = d_world_id; a.w_cnt = d_world_cnt;
  a.rounds = d_rounds; a.inliers = d_inliers; a.status = d_status;
  if (n_seq <= (long long)ctx->sm_count * 4)
    seq_pipeline_kernel<4><<<(unsigned)n_seq, kSeqThreads, 0, ctx->stream>>>(a);
  else
    seq_pipeline_kernel<VO_SEQ_MINB><<<(unsigned)n_seq, kSeqThreads, 0, ctx->stream>>>(a);
  VO_CHECK_LAUNCH(ctx, "seq_pipeline_kernel");
  return VO_OK;
}

int vo_seq_batch_run(vo_ctx* ctx, const vo_seq_params* params, int n_seq, int n_frames, int max_pts, int world_cap,
                     const int32_t* cnt, const float* uv, const float* desc, const int32_t* id_real, float* poses,
                     float* world_xyz, int32_t* world_id, int32_t* world_cnt, int32_t* rounds, int32_t* inliers,
                     int32_t* status) {
  if (!ctx) return VO_ERR_INVALID;
  int st = vo_ctx_activate(ctx);
  if (st) return st;
  VO_REQUIRE(ctx, n_seq >= 0 && n_frames >= 1 && max_pts >= 1 && world_cap >= 1, "vo_seq_batch_run: sizes");
  if (n_seq == 0) return VO_OK;
  const size_t SF = (size_t)n_seq * n_frames, SFP = SF * max_pts, SW = (size_t)n_seq * world_cap;
  const size_t bytes[11] = {SF * 4, SFP * 8, SFP * kDim * 4, SFP * 4, SF * 48, SW * 12, SW * 4, (size_t)n_seq * 4, SF * 4, SF * 8, (size_t)n_seq * 4};
  // one staging arena per context (grown on demand, reused): no cudaMalloc / cudaFree per call.  Inputs first, then
  // the outputs as ONE block that is zeroed before the launch: a sequence that stops early (status 2 / 3) hands
  // zeros, not stale device memory, back to the caller.
  size_t off[12];
  off[0] = 0;
  for (int i = 0; i < 11; ++i) off[i + 1] = vo_align_up(off[i] + bytes[i], 256);
  char* base;
  st = vo_stage(ctx, off[11], (void**)&base);
  if (st) return st;
  void* d[11];
  for (int i = 0; i < 11; ++i) d[i] = base + off[i];
  VO_CUDA(ctx, cudaMemsetAsync(d[4], 0, off[11] - off[4], ctx->stream));
  const void* src[4] = {cnt, uv, desc, id_real};
  for (int i = 0; i < 4; ++i) VO_CUDA(ctx, cudaMemcpyAsync(d[i], src[i], bytes[i], cudaMemcpyHostToDevice, ctx->stream));
  st = vo_seq_batch_run_dev(ctx, params, n_seq, n_frames, max_pts, world_cap, (int32_t*)d[0], (float*)d[1], (float*)d[2],
                            (int32_t*)d[3], (float*)d[4], (float*)d[5], (int32_t*)d[6], (int32_t*)d[7], (int32_t*)d[8],
                            (int32_t*)d[9], (int32_t*)d[10]);
  if (st) return st;
  void* dst[7] = {poses, world_xyz, world_id, world_cnt, rounds, inliers, status};
  for (int i = 0; i < 7; ++i)
    if (dst[i]) VO_CUDA(ctx, cudaMemcpyAsync(dst[i], d[4 + i], bytes[4 + i], cudaMemcpyDeviceToHost, ctx->stream));
  VO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return VO_OK;
}

}  // extern "C"
