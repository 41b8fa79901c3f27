/*
 * vo_b200.h — C-ABI of libvo_b200.so: the B200 (sm_100a) hot path of llepa/02-VisualOdometry.
 *
 * The reference has no FFI layer: its "operator API" is the C++ surface of the four
 * translation units every target links (CMakeLists.txt:25-59: cam.cpp camera.cpp
 * picp_solver.cpp my_utilities.cpp) plus the inline/template bodies in camera.h and
 * my_utilities.h.  Each entry point below names the reference interface it replaces.
 * The C++ mirror of that surface (pr::Camera, pr::PICPSolver, match_points<>, Cam) lives in
 * 02-visualodometry_b200/host/ and forwards here; INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch/Eigen/OpenCV types.
 *   - every function returns a vo_status (0 = ok).  No exit(), no exceptions cross the ABI.
 *   - matrices are row-major; a pose is a 3x4 [R|t] row-major float[12].
 *   - `Vector3fVector`, `Vector2fVector`, `IntPairVector` of the reference (src/defs.h:22-23,
 *     209-211) are contiguous float[3]/float[2]/int32[2] arrays and are passed as-is.
 *   - functions without a suffix take HOST buffers and return finished host-visible results
 *     (the reference is synchronous); `_dev` variants take DEVICE pointers resident in HBM
 *     and only enqueue work on the context's stream unless stated otherwise.
 *   - alignment of DEVICE arrays: 2-vectors (image points, int32 pairs) are read as 8-byte words and must be 8-byte
 *     aligned - any cudaMalloc'd buffer, or an element-aligned view into one, is; misaligned pointers are rejected with
 *     VO_ERR_INVALID.  float[3] points and descriptor rows need only their natural 4-byte alignment.
 *   - a vo_ctx is bound to one GPU and one CUDA stream; handles are not thread-safe.
 *   - there is no CPU fallback: every call fails with VO_ERR_CUDA when no device is usable.
 */
#ifndef VO_B200_H
#define VO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VO_B200_VERSION 100

typedef enum vo_status {
  VO_OK = 0,
  VO_ERR_INVALID = 1,  /* bad argument (null pointer, negative size, index out of range) */
  VO_ERR_CUDA = 2,     /* CUDA runtime / driver error, see vo_last_error */
  VO_ERR_NCCL = 3,     /* NCCL error or libnccl not loadable */
  VO_ERR_NOMEM = 4,
  VO_ERR_STATE = 5,    /* call order violated (e.g. one_round before set_points) */
  VO_ERR_CAPACITY = 6  /* caller's output buffer too small */
} vo_status;

typedef struct vo_ctx vo_ctx;   /* device + stream + scratch arena (+ optional NCCL communicator) */
typedef struct vo_picp vo_picp; /* device-resident state of one pr::PICPSolver */

/* per-correspondence status written by vo_picp_linearize (src/picp_solver.cpp:71-83) */
#define VO_PICP_SKIPPED 0 /* errorAndJacobian returned false: behind camera / outside image */
#define VO_PICP_INLIER 1  /* chi <= kernel threshold */
#define VO_PICP_OUTLIER 2 /* chi >  kernel threshold */

/* PICPSolver::chiInliers/chiOutliers/numInliers (src/picp_solver.h:47-53) of one round */
typedef struct vo_picp_stats {
  float chi_inliers;
  float chi_outliers;
  int32_t num_inliers;
  int32_t num_outliers;
} vo_picp_stats;

/* ------------------------------------------------------------------ context */
const char* vo_status_str(int status);
int vo_version(void);
int vo_device_count(int* n);
/* cuda_stream: a cudaStream_t to run on (e.g. torch's current stream), or NULL to create one */
int vo_ctx_create(int device, void* cuda_stream, vo_ctx** out);
int vo_ctx_destroy(vo_ctx* ctx);
int vo_ctx_sync(vo_ctx* ctx);
const char* vo_last_error(const vo_ctx* ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
int64_t vo_ctx_kernel_launches(const vo_ctx* ctx);
void* vo_ctx_stream(const vo_ctx* ctx);

/* multi-GPU (one process per GPU). The 128-byte id is an ncclUniqueId made by rank 0 and
 * distributed by the caller's own plumbing (torch.distributed / MPI / a file). libnccl.so.2
 * is resolved at run time. After init every PICP round all-reduces its H/b/chi terms. */
int vo_comm_unique_id(uint8_t id[128]);
int vo_ctx_comm_init(vo_ctx* ctx, int n_ranks, int rank, const uint8_t id[128]);
int vo_ctx_comm_destroy(vo_ctx* ctx);
int vo_ctx_comm_size(const vo_ctx* ctx);

/* Fused exchange over NVLink peer memory (replaces the NCCL call + separate solve launch of a PICP round):
 * every rank exports a 64-byte CUDA IPC handle of its mailbox, the caller all-gathers the handles
 * (n_ranks * 64 bytes, rank order) and attaches them. The last CTA of the linearize kernel then
 * stores its 32 terms into every peer's mailbox, waits for all ranks' flags and sums in rank order, so
 * all ranks solve the identical system in the same launch. Ranks must issue their PICP rounds in the
 * same order. At most VO_MAX_PEERS ranks of one node. */
#define VO_MAX_PEERS 8
#define VO_IPC_HANDLE_BYTES 64
int vo_ctx_peer_export(vo_ctx* ctx, uint8_t handle[VO_IPC_HANDLE_BYTES]);
int vo_ctx_peer_attach(vo_ctx* ctx, int n_ranks, int rank, const uint8_t* handles);
int vo_ctx_peer_detach(vo_ctx* ctx);
/* 1 when PICP rounds on this context use the fused peer exchange */
int vo_ctx_peer_active(const vo_ctx* ctx);

/* --------------------------------------------------------- Isometry helpers
 * Eigen::Isometry3f inverse / product as the callers use them on the host
 * (exec/icp_test.cpp:79,114,142; src/cam.cpp:78-81). Pure host arithmetic. */
void vo_pose_inverse(const float T[12], float out[12]);
void vo_pose_mul(const float A[12], const float B[12], float out[12]);

/* ----------------------------------------------------------------- pr::Camera
 * Camera::projectPoints (src/camera.cpp:14-35) / projectPoint (src/camera.h:24-36).
 * keep_indices != 0: out_uv has n rows with (-1,-1) for invalid points; else compacted
 * in input order. *n_out = rows written, *n_inside = return value of the reference. */
int vo_project_points(vo_ctx* ctx, const float K[9], int rows, int cols, const float pose[12],
                      const float* world_xyz, int64_t n, int keep_indices,
                      float* out_uv, int64_t* n_out, int64_t* n_inside);

/* ------------------------------------------------------------- pr::PICPSolver */
int vo_picp_create(vo_ctx* ctx, vo_picp** out);                 /* PICPSolver() picp_solver.cpp:8-15 */
int vo_picp_destroy(vo_picp* s);
/* init(camera, ...) copies the Camera by value (picp_solver.cpp:17-23; camera.cpp:4-11) */
int vo_picp_set_camera(vo_picp* s, const float K[9], int rows, int cols, const float pose[12]);
/* Camera::setWorldInCameraPose. Host-side only: the pose is copied and reaches the device in stream order with the next
 * call that reads it (as a launch argument of the persistent solve kernels, otherwise through a one-warp kernel). */
int vo_picp_set_pose(vo_picp* s, const float pose[12]);
int vo_picp_get_pose(vo_picp* s, float pose[12]);               /* camera().worldInCameraPose() */
/* init(..., world_points, image_points): uploads (copies) the points. The reference keeps raw
 * pointers (picp_solver.cpp:21-22), which dangle in exec/icp_test.cpp:81-85; copying is the
 * safe reading of that contract. */
int vo_picp_set_points(vo_picp* s, const float* world_xyz, int64_t n_world,
                       const float* image_xy, int64_t n_image);
int vo_picp_set_points_dev(vo_picp* s, const float* d_world_xyz, int64_t n_world,
                           const float* d_image_xy, int64_t n_image); /* borrowed, not copied */
/* correspondences (first: image index, second: world index), src/picp_solver.cpp:62-70.
 * Indices are range-checked on the device; out of range -> VO_ERR_INVALID. */
int vo_picp_set_correspondences(vo_picp* s, const int32_t* pairs, int64_t n_pairs);
/* d_pairs is BORROWED until the next set_correspondences* call (nothing is gathered at this point: the resident
 * kernel gathers straight into shared memory, the streaming kernel packs its planes on first use). An index out
 * of range is reported by the next vo_picp_fetch_stats / vo_picp_solve / vo_picp_linearize (VO_ERR_INVALID). */
int vo_picp_set_correspondences_dev(vo_picp* s, const int32_t* d_pairs, int64_t n_pairs);
/* Which kernel runs the Gauss-Newton rounds (diagnostic; results agree to float rounding, masks bit for bit):
 * AUTO: a solve of >= 2 rounds per call runs as ONE persistent cooperative launch - with the correspondences
 *       RESIDENT in shared memory across all rounds when the set fits the machine's shared memory
 *       (vo_picp_resident_capacity, 1.67 M correspondences on a B200), else STREAM_PERSISTENT: the packed planes
 *       streamed through the TMA ring every round; the rounds exchange their sums inside the kernel.  A single round
 *       (vo_picp_one_round) and a context with an NCCL-only communicator use STREAM: one launch per round.
 * STREAM / RESIDENT / STREAM_PERSISTENT force one of the three (RESIDENT fails with VO_ERR_CAPACITY when the set
 * does not fit). */
#define VO_PICP_MODE_AUTO 0
#define VO_PICP_MODE_STREAM 1
#define VO_PICP_MODE_RESIDENT 2
#define VO_PICP_MODE_STREAM_PERSISTENT 3
int vo_picp_set_mode(vo_picp* s, int mode);
int vo_picp_resident_capacity(const vo_picp* s, int64_t* n_pairs_max);
/* Builds the streaming kernels' packed planes of the current set now (picp_pack_kernel, 48 B per correspondence of
 * traffic) instead of lazily inside the first streamed round; a no-op when they exist. */
int vo_picp_pack(vo_picp* s);
/* PICPSolver::linearize (picp_solver.cpp:56-91) at the current pose, no state change.
 * H is the full symmetric 6x6; status (nullable) gets one VO_PICP_* byte per correspondence. */
int vo_picp_linearize(vo_picp* s, float kernel_threshold, int keep_outliers,
                      float H[36], float b[6], vo_picp_stats* stats, uint8_t* status);
/* PICPSolver::oneRound (picp_solver.cpp:93-105): linearize, H += I*damping, LDLT solve,
 * pose <- v2tEuler(dx) * pose. Synchronous; stats are those of the linearization. */
int vo_picp_one_round(vo_picp* s, float kernel_threshold, float damping, int keep_outliers,
                      vo_picp_stats* stats);
/* n_rounds oneRound()s enqueued back to back with no host synchronisation in between
 * (pose, H, b stay in HBM). At most VO_PICP_MAX_ROUNDS per call. */
#define VO_PICP_MAX_ROUNDS 64
int vo_picp_enqueue_rounds(vo_picp* s, float kernel_threshold, float damping, int keep_outliers,
                           int n_rounds);
/* waits for the stream and returns the stats of the last enqueue_rounds call (nullable) */
int vo_picp_fetch_stats(vo_picp* s, vo_picp_stats* stats_out, int n_rounds);
/* the driver loop of exec/icp_test.cpp:88-107 run on the device: up to max_rounds rounds,
 * stopping after the first round whose relative chi_inliers change is < rel_tol.
 * Returns rounds executed in *rounds_done and the stats of the last executed round. */
int vo_picp_solve(vo_picp* s, float kernel_threshold, float damping, int keep_outliers,
                  int max_rounds, float rel_tol, int* rounds_done, vo_picp_stats* last);

/* Self-test of the one arithmetic shortcut on the bit-exact path (csrc/picp.cu pair_front, vo_device.cuh
 * picp_project): rcp.approx + one FMA Newton step in place of the IEEE reciprocal of src/camera.h:30 /
 * src/picp_solver.cpp:44 inside the gate 1e-30 <= z <= 1e30.  Runs ALL 2^32 float bit patterns on the device
 * against __frcp_rn and 1.f / z.  out[0] = inputs inside the gate, out[1] / out[2] = mismatches of the packed /
 * scalar form (must be 0), out[3] = first mismatching bit pattern + 1 (0: none). */
int vo_selftest_reciprocal(vo_ctx* ctx, uint64_t out[4]);

/* --------------------------------------------------------------- match_points
 * match_points<P1,P2> (src/my_utilities.h:70-120): for every row i of A the best and second
 * best squared descriptor distance over all rows of B (float32, Eigen's evaluation order),
 * lowest index on ties; accepted iff best < dist_thr && best/second < ratio_thr.
 * pairs_out receives (i, best_j) in ascending i for rows [row_begin,row_end) (row sharding).
 * idA/idB (nullable) are the id_real columns: stats[0] = #(i,j) with equal ids over the row
 * range, stats[1] = #accepted pairs with equal ids (the line printed at :116-119). */
int vo_match(vo_ctx* ctx, const float* descA, int64_t n1, const float* descB, int64_t n2, int dim,
             float dist_thr, float ratio_thr, const int32_t* idA, const int32_t* idB,
             int64_t row_begin, int64_t row_end,
             int32_t* pairs_out, int64_t capacity, int64_t* n_out, int64_t stats[2]);
/* device-resident variant: all pointers are device pointers except n_out/stats (host).
 * d_best/d_second/d_best_idx (nullable) receive the per-row results. Synchronises once to
 * read the match count. */
int vo_match_dev(vo_ctx* ctx, const float* d_descA, int64_t n1, const float* d_descB, int64_t n2, int dim,
                 float dist_thr, float ratio_thr, const int32_t* d_idA, const int32_t* d_idB,
                 int64_t row_begin, int64_t row_end,
                 int32_t* d_pairs_out, int64_t capacity, int64_t* n_out, int64_t stats[2],
                 float* d_best, float* d_second, int32_t* d_best_idx);

/* Sharded matching with its exchange step (BASELINE config 4 across the GPUs of a box): shard `shard` of `n_shards`
 * scans its share of ALL n1 rows against the replicated B and writes d_match_idx[n1] (device): the matched column
 * of a row it owns, -1 for rejected rows and for rows of other shards.  On the indexed path (dim 10, large sets) the
 * shards are contiguous segments of the rows' MORTON ORDER, not of their index range: every shard then sees the row
 * density of the unsharded problem and the index prunes as well as on one GPU (row blocks by index prune 2x worse
 * at 8 shards).  With a communicator attached (vo_ctx_comm_init, n_ranks == n_shards) the per-row results are then
 * merged by one element-wise MAX all-reduce of n1 int32 over NVLink, and the accepted pairs (i, best_j) are
 * compacted in ascending i: every rank ends with the identical, complete result of match_points.
 * vo_match_compact_dev is step 3 alone (for callers that exchange d_match_idx themselves). */
int vo_match_sharded_dev(vo_ctx* ctx, const float* d_descA, int64_t n1, const float* d_descB, int64_t n2, int dim,
                         float dist_thr, float ratio_thr, int shard, int n_shards, int32_t* d_match_idx,
                         int32_t* d_pairs_out, int64_t capacity, int64_t* n_out);
int vo_match_compact_dev(vo_ctx* ctx, const int32_t* d_match_idx, int64_t n1, int32_t* d_pairs_out, int64_t capacity,
                         int64_t* n_out);

/* Which of the matcher's paths runs (diagnostics, parity tests and bench.py; every path returns the identical
 * bit-exact result): AUTO picks by size; BRUTE = the plain tiled distance-matrix scan over ALL n1*n2 pairs
 * (any dim); ORDERED = Morton-ordered rows + packed exact scan with the early-exit bound (dim 10);
 * INDEXED_EXACT = Morton index walk, every visited tile evaluated in exact fp32 (dim 10, large sets);
 * INDEXED_FILTERED = the same walk behind the bf16 tensor-core lower-bound filter (what AUTO uses at scale). */
#define VO_MATCH_PATH_AUTO 0
#define VO_MATCH_PATH_BRUTE 1
#define VO_MATCH_PATH_ORDERED 2
#define VO_MATCH_PATH_INDEXED_EXACT 3
#define VO_MATCH_PATH_INDEXED_FILTERED 4
int vo_match_set_path(vo_ctx* ctx, int path);

/* ---------------------------------------------------------------------- Cam
 * Cam::triangulatePoints (src/cam.cpp:94-140): P = K*T^-1[0:3], OpenCV DLT in double per pair,
 * float32 dehomogenisation. T1,T2 are camera-in-world poses as the reference passes them. */
int vo_triangulate(vo_ctx* ctx, const float K[9], const float T1[12], const float T2[12],
                   const float* x1, const float* x2, int64_t n, float* xyz_out);
int vo_triangulate_dev(vo_ctx* ctx, const float K[9], const float T1[12], const float T2[12],
                       const float* d_x1, const float* d_x2, int64_t n, float* d_xyz_out);
/* Cam::computeEssentialAndRecoverPose (src/cam.cpp:37-91): cv::findEssentialMat(p1, p2, K, cv::RANSAC) with
 * OpenCV's defaults (five-point minimal solver inside RANSAC, confidence 0.999, threshold 1 px, <= 1000 iterations,
 * no refit - restated incl. OpenCV's sampling sequence, null-space basis and root order, so the SAME hypothesis
 * wins) followed by OpenCV's recoverPose (4 candidates, cheirality vote, |t| = 1, x2 = R x1 + t).
 * E,R,t are double (CV_64F in the reference); mask (nullable) = recoverPose's mask (0/255).
 * Fails with VO_ERR_STATE when no hypothesis reaches 5 inliers (the reference exits, cam.cpp:56-59). */
int vo_essential_recover(vo_ctx* ctx, const float K[9], const float* x1, const float* x2, int64_t n,
                         double E[9], double R[9], double t[3], uint8_t* mask, int* n_good);
/* The same with the estimator and its parameters explicit.  VO_ESSENTIAL_LINEAR8 is the batched alternative named
 * by north_star: the normalised 8-point estimator on ALL matches (no outlier rejection; what vo_seq_batch_run uses).
 * ransac_mask (nullable, 0/1), ransac_inliers, ransac_iters (nullable): findEssentialMat's mask output, the inlier
 * count of the winning hypothesis and the number of samples OpenCV's loop would have consumed. */
#define VO_ESSENTIAL_RANSAC5 0
#define VO_ESSENTIAL_LINEAR8 1
int vo_essential_recover_ex(vo_ctx* ctx, const float K[9], const float* x1, const float* x2, int64_t n, int method,
                            double prob, double threshold, int max_iters, double E[9], double R[9], double t[3],
                            uint8_t* mask, int* n_good, uint8_t* ransac_mask, int* ransac_inliers, int* ransac_iters);

/* add_new_world_points (src/my_utilities.cpp:413-434): keep[j] = 1 iff cand_id[j] is not in
 * matched_id[0..n_matched). */
int vo_anti_join(vo_ctx* ctx, const int32_t* matched_id, int64_t n_matched,
                 const int32_t* cand_id, int64_t n_cand, uint8_t* keep, int64_t* n_keep);

/* ------------------------------------------------------- batched independent sequences
 * BASELINE config 5: n_seq independent sequences, each run through the reference's final pipeline
 * (exec/icp_test.cpp:40-136: match(0,1) -> essential -> triangulate, then per frame match vs map, PICP
 * with the driver's convergence loop, match vs previous frame, anti-join, triangulate, append) by ONE CTA
 * without returning to the host; the map of every sequence stays in HBM. Layouts (S sequences, F frames,
 * P = max_pts <= 128 points per frame, W = world_cap):
 *   cnt[S][F]  uv[S][F][P][2]  desc[S][F][P][10]  id_real[S][F][P]      (id_meas = index inside the frame)
 *   poses[S][F][12] camera-in-world (frame 0 = identity)   world_xyz[S][W][3]  world_id[S][W]  world_cnt[S]
 *   rounds[S][F] (nullable)  inliers[S][F][2] = (inliers of the last round, correspondences) (nullable)
 *   status[S]: 0 ok, 1 map capacity reached (overflow dropped), 2 fewer than 8 initial matches (nothing done),
 *              3 tracking lost (a non-finite pose came out of PICP; frames from there on keep the identity) */
typedef struct vo_seq_params {
  float K[9];
  int32_t rows, cols;
  float dist_thr, ratio_thr;  /* my_utilities.h:44-46: 0.2, 0.8 */
  float kernel_threshold;     /* icp_test.cpp:86: 3000 */
  float damping;              /* picp_solver.cpp:11: 1 */
  int32_t keep_outliers;      /* icp_test.cpp:95: 0 */
  int32_t max_rounds;         /* icp_test.cpp:88: 50 */
  float rel_tol;              /* icp_test.cpp:91: 1e-5 */
} vo_seq_params;
int vo_seq_batch_run(vo_ctx* ctx, const vo_seq_params* params, int n_seq, int n_frames, int max_pts, int world_cap,
                     const int32_t* cnt, const float* uv, const float* desc, const int32_t* id_real,
                     float* poses, float* world_xyz, int32_t* world_id, int32_t* world_cnt,
                     int32_t* rounds, int32_t* inliers, int32_t* status);
int vo_seq_batch_run_dev(vo_ctx* ctx, const vo_seq_params* params, int n_seq, int n_frames, int max_pts, int world_cap,
                         const int32_t* d_cnt, const float* d_uv, const float* d_desc, const int32_t* d_id_real,
                         float* d_poses, float* d_world_xyz, int32_t* d_world_id, int32_t* d_world_cnt,
                         int32_t* d_rounds, int32_t* d_inliers, int32_t* d_status);

#ifdef __cplusplus
}
#endif
#endif /* VO_B200_H */
