/*
 * oracle/landmarks_oracle.c -- TEST INFRASTRUCTURE ONLY (see vo_oracle.h).
 *
 * CPU restatement of the two components next to the hot path:
 *   f2  the candidate min-distance filter, reference VisualOdometryPipeLine.py:258
 *         valid[i] = np.all(np.linalg.norm(pts[i,:] - self.potential_keys, axis=1) > min_dist)
 *       on float32 arrays: numpy squares, adds and takes the square root in float32.
 *   f1  triangulate_landmarks, reference :107-206: per candidate the age gate (:171-174), the
 *       bearing-angle gate check_baseline (:117-147), cv2.triangulatePoints (:188-193), the
 *       de-homogenisation in float32 (:194) and the depth window disambguate_landmark (:149-168).
 * cv2.triangulatePoints (OpenCV calib3d triangulate.cpp, 4.13): per point the 4x4 matrix
 *   A[2v]   = x_v * P_v[2,:] - P_v[0,:],   A[2v+1] = y_v * P_v[2,:] - P_v[1,:]     (v = 0, 1)
 * in double, cv::SVD::compute (small matrix -> OpenCV's own Jacobi, orc_jacobi_svd), the last row
 * of V^T cast to the type of the input points (float32 here).  Pinned by
 * tests/test_oracle_landmarks.py against live cv2 and against the recorded calls of the unmodified
 * reference class (tests/golden/reference_trace.npz).
 */
#include "vo_oracle.h"
#include <math.h>
#include <string.h>

void orc_min_distance_mask(const float* pts, int n, const float* existing, int m, float min_dist, uint8_t* valid)
{
    for (int i = 0; i < n; ++i) {
        int ok = 1;
        for (int j = 0; j < m && ok; ++j) {
            const float dx = pts[2 * i] - existing[2 * j], dy = pts[2 * i + 1] - existing[2 * j + 1];
            const float xx = dx * dx, yy = dy * dy;
            const float s = xx + yy;
            const float d = sqrtf(s);
            ok = d > min_dist;      /* NaN compares false, as in numpy */
        }
        valid[i] = (uint8_t)ok;
    }
}

static void invert_pose(const double* cw, double* R, double* t)   /* (R, t) -> (R^T, -R^T t), ref :60-76 */
{
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R[3 * i + j] = cw[3 * j + i];
    for (int i = 0; i < 3; ++i) t[i] = -(R[3 * i] * cw[9] + R[3 * i + 1] * cw[10] + R[3 * i + 2] * cw[11]);
}

static void proj_matrix(const double* K, const double* R, const double* t, double* P /* 3x4 */)
{
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) P[4 * i + j] = K[3 * i] * R[j] + K[3 * i + 1] * R[3 + j] + K[3 * i + 2] * R[6 + j];
        P[4 * i + 3] = K[3 * i] * t[0] + K[3 * i + 1] * t[1] + K[3 * i + 2] * t[2];
    }
}

int orc_triangulate_landmarks(const double K[9], double min_dist, double max_dist, double min_angle_deg,
                              int min_frames, const float* first_keys, const float* keys,
                              const int32_t* first_pose, int n, const double* poses_cw, int n_poses,
                              const double cur_cw[12], uint8_t* keep, float* new_landmarks,
                              float* new_keypoints, int* n_new)
{
    /* K^-1 of the upper-triangular pinhole matrix */
    const double Ki[9] = {1.0 / K[0], 0, -K[2] / K[0], 0, 1.0 / K[4], -K[5] / K[4], 0, 0, 1};
    double Rc[9], tc[3], Pc[12];
    invert_pose(cur_cw, Rc, tc);
    proj_matrix(K, Rc, tc, Pc);
    int cnt = 0;
    for (int i = 0; i < n; ++i) {
        keep[i] = 1;
        const int fp = first_pose[i];
        if (n_poses > 1 && n_poses - fp <= min_frames) continue;                     /* :171-174 */
        if (fp < 0 || fp >= n_poses) return -1;
        const double* past = poses_cw + 12 * fp;
        /* check_baseline :117-147 */
        const double u = keys[2 * i], v = keys[2 * i + 1], u0 = first_keys[2 * i], v0 = first_keys[2 * i + 1];
        const double a[3] = {Ki[0] * u + Ki[2], Ki[4] * v + Ki[5], 1.0};
        double rel[9];   /* (R_cur^T R_past)^T = R_past^T R_cur */
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) rel[3 * r + c] = past[r] * cur_cw[c] + past[3 + r] * cur_cw[3 + c] + past[6 + r] * cur_cw[6 + c];
        double M[9];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) M[3 * r + c] = rel[3 * r] * Ki[c] + rel[3 * r + 1] * Ki[3 + c] + rel[3 * r + 2] * Ki[6 + c];
        const double b[3] = {M[0] * u0 + M[1] * v0 + M[2], M[3] * u0 + M[4] * v0 + M[5], M[6] * u0 + M[7] * v0 + M[8]};
        double cs = (a[0] * b[0] + a[1] * b[1] + a[2] * b[2]) /
                    (sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]) * sqrt(b[0] * b[0] + b[1] * b[1] + b[2] * b[2]));
        cs = cs < -1.0 ? -1.0 : (cs > 1.0 ? 1.0 : cs);
        const double alpha = acos(cs) * (180.0 / 3.14159265358979323846);
        if (alpha < min_angle_deg) continue;
        /* cv2.triangulatePoints :188-193 */
        double Rp[9], tp[3], Pp[12];
        invert_pose(past, Rp, tp);
        proj_matrix(K, Rp, tp, Pp);
        double A[16], W[4], U[16], Vt[16];
        for (int k = 0; k < 4; ++k) {
            A[k] = u0 * Pp[8 + k] - Pp[k];
            A[4 + k] = v0 * Pp[8 + k] - Pp[4 + k];
            A[8 + k] = u * Pc[8 + k] - Pc[k];
            A[12 + k] = v * Pc[8 + k] - Pc[4 + k];
        }
        orc_jacobi_svd(A, 4, 4, W, U, Vt);
        const float X[4] = {(float)Vt[12], (float)Vt[13], (float)Vt[14], (float)Vt[15]};
        const float L[3] = {X[0] / X[3], X[1] / X[3], X[2] / X[3]};                    /* :194, float32 */
        /* disambguate_landmark :149-168 (depth in both cameras) */
        const double zc = Rc[6] * (double)L[0] + Rc[7] * (double)L[1] + Rc[8] * (double)L[2] + tc[2];
        const double zp = Rp[6] * (double)L[0] + Rp[7] * (double)L[1] + Rp[8] * (double)L[2] + tp[2];
        if (zc > min_dist && zp > min_dist && zc < max_dist && zp < max_dist) {
            keep[i] = 0;
            memcpy(new_landmarks + 3 * cnt, L, sizeof L);
            new_keypoints[2 * cnt] = keys[2 * i]; new_keypoints[2 * cnt + 1] = keys[2 * i + 1];
            ++cnt;
        }
    }
    *n_new = cnt;
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * f3: cv2.recoverPose(E, points1, points2, K) at reference :315 (OpenCV calib3d five-point.cpp,
 * 4.13): decomposeEssentialMat (SVD of E by OpenCV's small-matrix Jacobi, det sign fix,
 * R1 = U W Vt, R2 = U W^T Vt, t = U[:,2]); for the four (R, +-t) the cheirality test on points
 * triangulated with P0 = [I|0] (cv2.triangulatePoints on K-normalised float64 points):
 *   Q.z * Q.w > 0,  Q.z/Q.w < 50,  0 < (P Q/Q.w).z < 50;
 * the pose with the most passing points wins (ties in the order 1, 2, 3, 4); mask is 0 / 255.
 * ------------------------------------------------------------------------------------------ */
static void mat3_mul(const double* A, const double* B, double* C)
{
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
static double det3d(const double* M)
{
    return M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
}

void orc_decompose_essential(const double E[9], double R1[9], double R2[9], double t[3])
{
    double W[3], U[9], Vt[9], T[9];
    orc_jacobi_svd(E, 3, 3, W, U, Vt);
    if (det3d(U) < 0) for (int k = 0; k < 9; ++k) U[k] = -U[k];
    if (det3d(Vt) < 0) for (int k = 0; k < 9; ++k) Vt[k] = -Vt[k];
    const double Wm[9] = {0, 1, 0, -1, 0, 0, 0, 0, 1}, Wt[9] = {0, -1, 0, 1, 0, 0, 0, 0, 1};
    mat3_mul(U, Wm, T); mat3_mul(T, Vt, R1);
    mat3_mul(U, Wt, T); mat3_mul(T, Vt, R2);
    t[0] = U[2]; t[1] = U[5]; t[2] = U[8];
}

int orc_recover_pose(const double E[9], const float* p1, const float* p2, int n, const double K[9],
                     double dist_thresh, double R[9], double t[3], uint8_t* mask, int* n_good)
{
    double R1[9], R2[9], tt[3];
    orc_decompose_essential(E, R1, R2, tt);
    const double fx = K[0], fy = K[4], cx = K[2], cy = K[5];
    int good[4] = {0, 0, 0, 0};
    uint8_t* masks = mask;   /* caller provides 4*n bytes of scratch after... see wrapper: mask has room for 4n */
    for (int h = 0; h < 4; ++h) {
        const double* Rh = (h & 1) ? R2 : R1;
        const double sg = h < 2 ? 1.0 : -1.0;
        double P[12];
        for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) P[4 * i + j] = Rh[3 * i + j]; P[4 * i + 3] = sg * tt[i]; }
        for (int i = 0; i < n; ++i) {
            const double x1 = ((double)p1[2 * i] - cx) / fx, y1 = ((double)p1[2 * i + 1] - cy) / fy;
            const double x2 = ((double)p2[2 * i] - cx) / fx, y2 = ((double)p2[2 * i + 1] - cy) / fy;
            /* P0 = [I | 0] */
            double A[16] = {-1, 0, x1, 0, 0, -1, y1, 0, 0, 0, 0, 0, 0, 0, 0, 0}, Wd[4], U[16], Vt[16];
            for (int k = 0; k < 4; ++k) { A[8 + k] = x2 * P[8 + k] - P[k]; A[12 + k] = y2 * P[8 + k] - P[4 + k]; }
            orc_jacobi_svd(A, 4, 4, Wd, U, Vt);
            double Q[4] = {Vt[12], Vt[13], Vt[14], Vt[15]};
            int ok = Q[2] * Q[3] > 0;
            Q[0] /= Q[3]; Q[1] /= Q[3]; Q[2] /= Q[3]; Q[3] /= Q[3];
            ok = ok && (Q[2] < dist_thresh);
            const double z = P[8] * Q[0] + P[9] * Q[1] + P[10] * Q[2] + P[11] * Q[3];
            ok = ok && (z > 0) && (z < dist_thresh);
            masks[(size_t)h * n + i] = ok ? 255 : 0;
            good[h] += ok;
        }
    }
    int win;
    if (good[0] >= good[1] && good[0] >= good[2] && good[0] >= good[3]) win = 0;
    else if (good[1] >= good[0] && good[1] >= good[2] && good[1] >= good[3]) win = 1;
    else if (good[2] >= good[0] && good[2] >= good[1] && good[2] >= good[3]) win = 2;
    else win = 3;
    const double* Rw = (win & 1) ? R2 : R1;
    for (int k = 0; k < 9; ++k) R[k] = Rw[k];
    for (int k = 0; k < 3; ++k) t[k] = win < 2 ? tt[k] : -tt[k];
    if (win != 0) memmove(mask, masks + (size_t)win * n, (size_t)n);
    *n_good = good[win];
    return win;
}
