/*
 * vo_oracle.h — CPU ORACLE (test infrastructure, NOT the product).
 *
 * A dependency-free C++17 restatement of the reference hot path of
 * llepa/02-VisualOdometry.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product
 * (02-visualodometry_b200/csrc + host/) never links or calls it.
 *
 * Parity status
 *   - PICP / camera / matching / glue: restated from the reference sources
 *     (file:line cited per function) following SURVEY.md Appendix A for Eigen's
 *     float32 evaluation order.  Eigen itself is absent from this image, so at
 *     function level the restatement DEFINES bit-exactness ("parity unpinned"
 *     at function level); it is pinned END-TO-END against the reference's only
 *     goldens, the output/ text files (tests/golden/dataset.npz, tests/test_oracle_replay.py).
 *   - triangulation / essential / recoverPose: the reference delegates to
 *     OpenCV (un-pinned, un-vendored; cam.cpp:49,61,115,118).  Restated from the
 *     published algorithms and pinned against cv2 4.13.0 fixtures generated in
 *     the build container (oracle/gen_golden.py -> tests/golden/cv2_fixtures.npz).
 *
 * All matrices are row-major. A pose is a 3x4 [R|t] row-major float[12].
 * Compile with -O2 -ffp-contract=off (no FMA contraction, no fast-math).
 */
#pragma once
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* status byte written per correspondence by the linearize oracle */
#define VO_REF_SKIPPED 0 /* projectPoint returned false: contributes to nothing */
#define VO_REF_INLIER 1
#define VO_REF_OUTLIER 2

/* src/camera.h:24-36 */
int vo_ref_project_point(const float K[9], int rows, int cols, const float pose[12],
                         const float p[3], float uv[2]);

/* src/camera.cpp:14-35; returns number of points inside, *n_out = output length */
int vo_ref_project_points(const float K[9], int rows, int cols, const float pose[12],
                          const float* world_xyz, int n, int keep_indices,
                          float* out_uv, int* n_out);

/* src/picp_solver.cpp:26-54; J is 2x6 row-major */
int vo_ref_error_jacobian(const float K[9], int rows, int cols, const float pose[12],
                          const float p[3], const float z[2], float e[2], float J[12]);

/* src/picp_solver.cpp:56-91. pairs = (first: image idx, second: world idx).
 * accum_mode 0: float32 sequential (the reference), 1: float64 accumulators
 * (same per-correspondence float32 terms) used as the tolerance anchor. */
void vo_ref_linearize(const float K[9], int rows, int cols, const float pose[12],
                      const float* world_xyz, const float* image_xy,
                      const int32_t* pairs, int64_t n_pairs,
                      float thr, int keep_outliers, int accum_mode,
                      double H[36], double b[6], double* chi_in, double* chi_out,
                      int64_t* n_inliers, uint8_t* status /* nullable */);

/* Eigen pivoted LDLT in float32: solves A x = rhs (picp_solver.cpp:102) */
void vo_ref_ldlt_solve6(const float A[36], const float rhs[6], float x[6]);

/* src/defs.h:100-136 + picp_solver.cpp:103: pose <- v2tEuler(dx) * pose */
void vo_ref_pose_update(const float dx[6], float pose[12]);

/* src/picp_solver.cpp:93-105 (one Gauss-Newton round, float32 sequential) */
void vo_ref_one_round(const float K[9], int rows, int cols, float pose[12],
                      const float* world_xyz, const float* image_xy,
                      const int32_t* pairs, int64_t n_pairs,
                      float thr, float damping, int keep_outliers,
                      float* chi_in, float* chi_out, int* n_inliers);

/* multi-threaded correspondence-parallel variant (CPU baseline only):
 * per-thread float32 partials in correspondence order, summed in thread order */
void vo_ref_one_round_mt(const float K[9], int rows, int cols, float pose[12],
                         const float* world_xyz, const float* image_xy,
                         const int32_t* pairs, int64_t n_pairs,
                         float thr, float damping, int keep_outliers, int n_threads,
                         float* chi_in, float* chi_out, int* n_inliers);

/* src/my_utilities.h:70-120.  order_mode 0: Eigen SSE squaredNorm order
 * (Appendix A.7), 1: plain left-to-right.  Row range [row_begin,row_end) of A.
 * Returns number of accepted matches; pairs_out = (i, best_j) ascending i.
 * stats[0] = #pairs with equal id (needs idA/idB), stats[1] = #correct matches.
 * best/second/best_idx are optional per-row outputs (length row_end-row_begin). */
int64_t vo_ref_match(const float* descA, int64_t n1, const float* descB, int64_t n2, int dim,
                     float dist_thr, float ratio_thr,
                     const int32_t* idA, const int32_t* idB,
                     int64_t row_begin, int64_t row_end, int order_mode, int n_threads,
                     int32_t* pairs_out, int64_t stats[2],
                     float* best, float* second, int32_t* best_idx);

/* src/cam.cpp:94-140 (cv::triangulatePoints DLT + convertPointsFromHomogeneous).
 * T1,T2 = camera-in-world poses (as passed by the reference). */
void vo_ref_triangulate(const float K[9], const float T1[12], const float T2[12],
                        const float* x1, const float* x2, int64_t n, float* xyz_out);

/* src/cam.cpp:37-91: essential matrix (normalised 8-point on all matches, see
 * DESIGN.md) + OpenCV recoverPose restated.  R,t double (CV_64F in the reference). */
int vo_ref_essential_recover(const float K[9], const float* x1, const float* x2, int64_t n,
                             double E[9], double R[9], double t[3], uint8_t* mask);

/* src/cam.cpp:49 AS THE REFERENCE RUNS IT: cv::findEssentialMat(p1, p2, K, RANSAC) restated (five_point.cpp:
 * OpenCV's RNG, subset selection, 5-point solver incl. its null-space basis and root order, Sampson inliers, adaptive
 * iteration count; no refit).  Returns the inlier count of the winning hypothesis (0: none). mask (nullable) 0/1. */
int vo_ref_find_essential_ransac(const float K[9], const float* x1, const float* x2, int64_t n, double prob,
                                 double threshold, int max_iters, double E[9], uint8_t* mask, int* iters_out);
/* ... followed by recoverPose (src/cam.cpp:61): the reference's computeEssentialAndRecoverPose */
int vo_ref_essential_recover_ransac(const float K[9], const float* x1, const float* x2, int64_t n, double E[9],
                                    double R[9], double t[3], uint8_t* mask);
/* EMEstimatorCallback::runKernel: 5 normalised correspondences (x,y interleaved) -> up to 10 E (row-major, unit norm) */
int vo_ref_five_point(const double q1[10], const double q2[10], double E_out[90]);
/* the first n_subsets 5-subsets RANSACPointSetRegistrator::getSubset draws for a set of n points */
void vo_ref_ransac_subsets(int n, int n_subsets, int32_t* idx_out);
/* cv::SVD::compute of a 3x3 matrix restated (OpenCV's JacobiSVDImpl_: its rotation order fixes the signs of the
 * singular vectors, which decide the candidate ORDER in recoverPose when cheirality votes tie) */
void vo_ref_cv_svd3(const double A[9], double U[9], double w[3], double Vt[9]);
/* one-sided Jacobi SVD (A m x n row-major -> U*Sigma in place, V n x n, w descending) */
void vo_ref_jacobi_svd(double* A, int m, int n, double* V, double* w);

/* OpenCV recoverPose alone, for pinning against cv2 with a given E */
int vo_ref_recover_pose(const double E[9], const float K[9], const float* x1, const float* x2,
                        int64_t n, double R[9], double t[3], uint8_t* mask);

/* src/my_utilities.cpp:413-434: keep[j]=1 iff img_match j's second id_meas is not
 * the id_meas of any image point already matched to a world point */
int64_t vo_ref_anti_join(const int32_t* matched_id_meas, int64_t n_matched,
                         const int32_t* cand_second_id_meas, int64_t n_cand, uint8_t* keep);

/* 3x4 isometry helpers (Eigen Isometry3f semantics, Appendix A.2) */
void vo_ref_pose_inverse(const float T[12], float out[12]);
void vo_ref_pose_mul(const float A[12], const float B[12], float out[12]);

int vo_ref_num_threads(void);

#ifdef __cplusplus
}
#endif
