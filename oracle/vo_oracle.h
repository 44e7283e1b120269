/*
 * oracle/vo_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the arithmetic behind the reference's hot-path
 * cv2 call sites (reference VisualOdometryPipeLine.py:229,256,281,287,308,343).
 * The arithmetic itself lives in OpenCV (third-party, `opencv-python`, pinned
 * ==4.6.0.66 in reference requirements.txt:6; the oracle is pinned against the
 * installed cv2 4.13.0 -- see tests/test_oracle_*.py and tests/golden/).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * load this library.  The product path (monocular_visual_odometry_va4mr_b200/)
 * never does.
 */
#ifndef VO_ORACLE_H
#define VO_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- pyramid / derivatives ---- */
void orc_pyr_down_u8(const uint8_t* src, int w, int h, size_t sstep, uint8_t* dst, size_t dstep);
void orc_scharr_s16(const uint8_t* src, int w, int h, size_t sstep, int16_t* dst /* h*w*2 interleaved */);
/* number of pyramid levels (incl. level 0) cv2 builds for (w,h,win,maxLevel) */
int orc_pyr_levels(int w, int h, int win_w, int win_h, int max_level);

/* ---- pyramidal LK (ref :281,:287) ---- */
int orc_calc_optical_flow_pyr_lk(const uint8_t* prev, const uint8_t* next, int rows, int cols,
                                 size_t prev_step, size_t next_step,
                                 const float* prev_pts, int n, int win_w, int win_h, int max_level,
                                 int crit_type, int crit_max_count, double crit_eps,
                                 int flags, double min_eig_thr,
                                 float* next_pts, uint8_t* status, float* err, int32_t* iters_out);

/* ---- Shi-Tomasi (ref :256) ---- */
void orc_min_eig_map(const uint8_t* img, int w, int h, size_t step, int block, float* eig);
int orc_good_features_to_track(const uint8_t* img, int rows, int cols, size_t step, int max_corners,
                               double quality, double min_dist, int block_size,
                               float* corners_xy, int* n_out);

/* ---- brute-force kNN(k=2) + ratio (ref :229, :218-224) ---- */
int orc_knn2_ratio(const float* q, int nq, const float* t, int nt, int dim, double ratio,
                   int32_t* idx2, float* dist2, uint8_t* accept);

/* ---- RANSAC framework ---- */
void orc_ransac_subsets(int count, int model_points, int n_iters, int32_t* subsets);
int orc_ransac_update_num_iters(double p, double ep, int model_points, int max_iters);
/* Replays cv2's sequential loop over per-model inlier counts.  counts[iter*max_models+m],
 * nmodels[iter].  Returns winning flat index (iter*max_models+m) or -1; *iters_run = loop trips. */
int orc_ransac_select(const int32_t* counts, const int32_t* nmodels, int max_models, int max_iters,
                      int n_points, int model_points, double conf, int* iters_run);

/* ---- PnP (ref :343) ---- */
/* P3P on 3 correspondences: object X[3][3], normalised image rays (x,y,1) -> up to 4 (R,t). */
int orc_p3p(const double X[9], const double xn[6], double R[4][9], double t[4][3]);
/* 4-point minimal solve as cv2's PnPRansacCallback: returns 1 and [rvec|tvec] or 0. */
int orc_pnp_minimal(const float* obj4, const float* img4, const double K[9], double rvec[3], double tvec[3]);
/* reprojection errors as cv2 computeError: FP64 project, round to f32, f32 squared error. */
void orc_pnp_errors(const float* obj, const float* img, int n, const double K[9],
                    const double rvec[3], const double tvec[3], float* err);
void orc_rodrigues_to_R(const double rvec[3], double R[9]);
void orc_R_to_rodrigues(const double R[9], double rvec[3]);
int orc_epnp(const double* obj, const double* img, int n, const double K[9], double R[9], double t[3]);
int orc_solve_pnp_ransac_p3p(const float* obj, const float* img, int n, const double K[9], int iters,
                             float reproj_err, double conf, double rvec[3], double tvec[3],
                             int32_t* inliers, int* n_inliers, int* success, int* iters_run);

/* ---- essential matrix (ref :308) ---- */
int orc_five_point(const double x1[10], const double x2[10], double E[10][9]);
void orc_sampson_errors(const double* x1n, const double* x2n, int n, const double E[9], float* err);
int orc_find_essential_mat_ransac(const float* p1, const float* p2, int n, const double K[9],
                                  double prob, double thr, int max_iters, double E[9],
                                  uint8_t* mask, int* found, int* iters_run);

/* OpenCV's small-matrix SVD (one-sided Jacobi, lapack.cpp JacobiSVDImpl_), m >= n <= 12, row-major */
void orc_jacobi_svd(const double* A, int m, int n, double* W, double* U, double* Vt);

/* ---- "next" rows, SURVEY 8f (the reference does these in numpy + cv2.triangulatePoints) ---- */
/* f2, ref :258: valid[i] = all_j ( sqrtf(fl(dx^2) + fl(dy^2)) > min_dist ) in float32 */
void orc_min_distance_mask(const float* pts, int n, const float* existing, int m, float min_dist, uint8_t* valid);
/* f1, ref :107-206 triangulate_landmarks: age gate, bearing-angle gate, cv2.triangulatePoints (4x4 DLT,
 * Jacobi SVD, float32 homogeneous output), depth window in both views.  poses are (R_CW row-major | t_CW)
 * as the reference stores them in self.transforms.  keep[i] = too_short_baseline[i]; accepted landmarks
 * and their keypoints are appended in candidate order. */
int orc_triangulate_landmarks(const double K[9], double min_dist, double max_dist, double min_angle_deg,
                              int min_frames, const float* first_keys, const float* keys,
                              const int32_t* first_pose, int n, const double* poses_cw, int n_poses,
                              const double cur_cw[12], uint8_t* keep, float* new_landmarks,
                              float* new_keypoints, int* n_new);

/* f3, ref :315 cv2.recoverPose(E, p1, p2, K) with the default distanceThresh = 50.  mask must have room
 * for 4*n bytes (the first n hold the result, 0 / 255).  Returns the index 0..3 of the winning (R, +-t). */
void orc_decompose_essential(const double E[9], double R1[9], double R2[9], double t[3]);
int orc_recover_pose(const double E[9], const float* p1, const float* p2, int n, const double K[9],
                     double dist_thresh, double R[9], double t[3], uint8_t* mask, int* n_good);

/* ---- f4, ref :35, :226-227: cv2.SIFT_create().detectAndCompute(img, None), default parameters (sift_oracle.c) ----
 * kps: max_kp x {x, y, size, angle, response, octave (int32 bits)}, desc: max_kp x 128 (may be NULL); returns the number
 * of keypoints found, in cv2's order (removeDuplicatedSorted); only the first max_kp are written. */
int orc_sift_detect_and_compute(const uint8_t* img, int rows, int cols, size_t step, int max_kp, float* kps, float* desc);
int orc_sift_gauss_image(const uint8_t* img, int rows, int cols, size_t step, int octave, int layer, float* out, int* orows, int* ocols);
int orc_gaussian_kernel_f32(double sigma, float* k, int max_n);
void orc_gaussian_blur_f32(const float* src, int rows, int cols, double sigma, float* dst);
float orc_fast_atan2_deg(float y, float x);

#ifdef __cplusplus
}
#endif
#endif
