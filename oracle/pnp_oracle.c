/*
 * oracle/pnp_oracle.c -- TEST INFRASTRUCTURE ONLY (see vo_oracle.h).
 *
 * Restates what cv2.solvePnPRansac(obj, img, K, zeros(4), flags=SOLVEPNP_P3P, ...)
 * computes at the reference call site VisualOdometryPipeLine.py:343 (OpenCV
 * modules/calib3d/src/{solvepnp,ptsetreg,p3p,epnp,calibration}.cpp; third party, not
 * vendored).  Spec: SURVEY.md A.6-A.8 plus the facts probed while writing this file:
 *   - the 4-point minimal solve normalises the image points with K and ROUNDS THEM TO
 *     float32 (undistortPoints on CV_32F input) before P3P; with that rounding an exact
 *     conic-pencil P3P reproduces cv2 4.13's candidate poses to ~1e-14 (median);
 *   - the candidate kept is the one with the smallest summed squared pixel reprojection
 *     error over the 4 sample points (solveP3P sorts by it);
 *   - the model is stored as rvec|tvec, scoring re-expands it with Rodrigues;
 *   - scoring: FP64 projection, rounded to float32, float32 squared error (no FMA);
 *   - the returned pose is an EPnP refit on the inliers (float64 inputs).
 * The P3P below is an independent exact solver (pencil of the two depth quadrics, one
 * real root of the cubic, split of the degenerate conic into two planes, Gauss-Newton
 * polish), not OpenCV's code.  Pinned against cv2 in tests/test_oracle_pnp.py and
 * tests/golden/pnp_*.npz.
 */
#include "vo_oracle.h"
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ small linear algebra */
static double det3(const double* M)
{
    return M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
}
static int inv3(const double* M, double* I)
{
    double d = det3(M);
    if (d == 0 || !isfinite(d)) return 0;
    double id = 1.0 / d;
    I[0] = (M[4] * M[8] - M[5] * M[7]) * id; I[1] = (M[2] * M[7] - M[1] * M[8]) * id; I[2] = (M[1] * M[5] - M[2] * M[4]) * id;
    I[3] = (M[5] * M[6] - M[3] * M[8]) * id; I[4] = (M[0] * M[8] - M[2] * M[6]) * id; I[5] = (M[2] * M[3] - M[0] * M[5]) * id;
    I[6] = (M[3] * M[7] - M[4] * M[6]) * id; I[7] = (M[1] * M[6] - M[0] * M[7]) * id; I[8] = (M[0] * M[4] - M[1] * M[3]) * id;
    return 1;
}
static void mat3mul(const double* A, const double* B, double* C)
{
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
static void cross3(const double* a, const double* b, double* c)
{
    c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
}
static double quad3(const double* M, const double* u, const double* v)
{   /* u^T M v */
    double s = 0;
    for (int i = 0; i < 3; ++i) s += u[i] * (M[3 * i] * v[0] + M[3 * i + 1] * v[1] + M[3 * i + 2] * v[2]);
    return s;
}

/* One-sided (Hestenes) Jacobi SVD, A (m x n, row-major, m >= n) = U diag(W) Vt, singular values
 * descending.  This is the algorithm OpenCV itself runs for matrices smaller than 25 x 25
 * (core/src/lapack.cpp JacobiSVDImpl_, eps = 10*DBL_EPSILON; LAPACK is only used above
 * that size), restated so that the SIGNS of the singular vectors match cv2's: EPnP's
 * control points c_k = c_0 + sqrt(d_k/n) u_k depend on the sign of u_k, and under noise the
 * EPnP pose depends on the control points at the 1e-4 level.  Checked against
 * cv2.SVDecomp in tests/test_oracle_pnp.py. */
void orc_jacobi_svd(const double* A, int m, int n, double* W, double* U, double* Vt)
{
    double At[12 * 12], Wd[12];
    const double eps = DBL_EPSILON * 10;
    for (int i = 0; i < n; ++i) {
        double sd = 0;
        for (int k = 0; k < m; ++k) { At[i * m + k] = A[k * n + i]; sd += At[i * m + k] * At[i * m + k]; }
        Wd[i] = sd;
        for (int k = 0; k < n; ++k) Vt[i * n + k] = (i == k);
    }
    int max_iter = m > 30 ? m : 30;
    for (int iter = 0; iter < max_iter; ++iter) {
        int changed = 0;
        for (int i = 0; i < n - 1; ++i)
            for (int j = i + 1; j < n; ++j) {
                double* Ai = At + i * m; double* Aj = At + j * m;
                double a = Wd[i], p = 0, b = Wd[j];
                for (int k = 0; k < m; ++k) p += Ai[k] * Aj[k];
                if (fabs(p) <= eps * sqrt(a * b)) continue;
                p *= 2;
                double beta = a - b, gamma = hypot(p, beta), c, s;
                if (beta < 0) {
                    double delta = (gamma - beta) * 0.5;
                    s = sqrt(delta / gamma);
                    c = p / (gamma * s * 2);
                } else {
                    c = sqrt((gamma + beta) / (gamma * 2));
                    s = p / (gamma * c * 2);
                }
                a = b = 0;
                for (int k = 0; k < m; ++k) {
                    double t0 = c * Ai[k] + s * Aj[k];
                    double t1 = -s * Ai[k] + c * Aj[k];
                    Ai[k] = t0; Aj[k] = t1;
                    a += t0 * t0; b += t1 * t1;
                }
                Wd[i] = a; Wd[j] = b;
                changed = 1;
                double* Vi = Vt + i * n; double* Vj = Vt + j * n;
                for (int k = 0; k < n; ++k) {
                    double t0 = c * Vi[k] + s * Vj[k];
                    double t1 = -s * Vi[k] + c * Vj[k];
                    Vi[k] = t0; Vj[k] = t1;
                }
            }
        if (!changed) break;
    }
    for (int i = 0; i < n; ++i) {
        double sd = 0;
        for (int k = 0; k < m; ++k) sd += At[i * m + k] * At[i * m + k];
        Wd[i] = sqrt(sd);
    }
    for (int i = 0; i < n - 1; ++i) {
        int j = i;
        for (int k = i + 1; k < n; ++k) if (Wd[j] < Wd[k]) j = k;
        if (i != j) {
            double tw = Wd[i]; Wd[i] = Wd[j]; Wd[j] = tw;
            for (int k = 0; k < m; ++k) { double tt = At[i * m + k]; At[i * m + k] = At[j * m + k]; At[j * m + k] = tt; }
            for (int k = 0; k < n; ++k) { double tt = Vt[i * n + k]; Vt[i * n + k] = Vt[j * n + k]; Vt[j * n + k] = tt; }
        }
    }
    for (int i = 0; i < n; ++i) {
        W[i] = Wd[i];
        double sc = Wd[i] > DBL_MIN ? 1 / Wd[i] : 0.;   /* (cv2 fills a random orthogonal vector for zero singular values) */
        for (int k = 0; k < m; ++k) U[k * n + i] = At[i * m + k] * sc;
    }
}

/* least squares min |A x - b| for small m x n (m >= n) by Householder QR */
static void qr_lstsq(const double* Ain, const double* bin, int m, int n, double* x)
{
    double A[6 * 5], b[6];
    memcpy(A, Ain, sizeof(double) * m * n);
    memcpy(b, bin, sizeof(double) * m);
    for (int k = 0; k < n; ++k) {
        double nrm = 0;
        for (int i = k; i < m; ++i) nrm += A[i * n + k] * A[i * n + k];
        nrm = sqrt(nrm);
        if (nrm == 0) continue;
        double alpha = A[k * n + k] > 0 ? -nrm : nrm;
        double v[6];
        for (int i = k; i < m; ++i) v[i] = A[i * n + k];
        v[k] -= alpha;
        double vn = 0;
        for (int i = k; i < m; ++i) vn += v[i] * v[i];
        if (vn == 0) continue;
        for (int j = k; j < n; ++j) {
            double s = 0;
            for (int i = k; i < m; ++i) s += v[i] * A[i * n + j];
            s = 2 * s / vn;
            for (int i = k; i < m; ++i) A[i * n + j] -= s * v[i];
        }
        double s = 0;
        for (int i = k; i < m; ++i) s += v[i] * b[i];
        s = 2 * s / vn;
        for (int i = k; i < m; ++i) b[i] -= s * v[i];
    }
    for (int k = n - 1; k >= 0; --k) {
        double s = b[k];
        for (int j = k + 1; j < n; ++j) s -= A[k * n + j] * x[j];
        x[k] = A[k * n + k] != 0 ? s / A[k * n + k] : 0;
    }
}

/* ------------------------------------------------------------------ Rodrigues */
void orc_rodrigues_to_R(const double r[3], double R[9])
{
    double theta = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    if (theta < DBL_EPSILON) {
        for (int i = 0; i < 9; ++i) R[i] = (i % 4) == 0;
        return;
    }
    double c = cos(theta), s = sin(theta), c1 = 1 - c, it = 1 / theta;
    double x = r[0] * it, y = r[1] * it, z = r[2] * it;
    R[0] = c + c1 * x * x; R[1] = c1 * x * y - s * z; R[2] = c1 * x * z + s * y;
    R[3] = c1 * x * y + s * z; R[4] = c + c1 * y * y; R[5] = c1 * y * z - s * x;
    R[6] = c1 * x * z - s * y; R[7] = c1 * y * z + s * x; R[8] = c + c1 * z * z;
}

void orc_R_to_rodrigues(const double R[9], double r[3])
{
    double rx = R[7] - R[5], ry = R[2] - R[6], rz = R[3] - R[1];
    double s = sqrt((rx * rx + ry * ry + rz * rz) * 0.25);
    double c = (R[0] + R[4] + R[8] - 1) * 0.5;
    c = c > 1 ? 1 : c < -1 ? -1 : c;
    double theta = acos(c);
    if (s < 1e-5) {
        if (c > 0) { r[0] = r[1] = r[2] = 0; return; }
        double t;
        t = (R[0] + 1) * 0.5; rx = sqrt(t > 0 ? t : 0);
        t = (R[4] + 1) * 0.5; ry = sqrt(t > 0 ? t : 0) * (R[1] < 0 ? -1. : 1.);
        t = (R[8] + 1) * 0.5; rz = sqrt(t > 0 ? t : 0) * (R[2] < 0 ? -1. : 1.);
        if (fabs(rx) < fabs(ry) && fabs(rx) < fabs(rz) && (R[5] > 0) != (ry * rz > 0)) rz = -rz;
        theta /= sqrt(rx * rx + ry * ry + rz * rz);
        r[0] = rx * theta; r[1] = ry * theta; r[2] = rz * theta;
        return;
    }
    double vth = 1 / (2 * s);
    vth *= theta;
    r[0] = rx * vth; r[1] = ry * vth; r[2] = rz * vth;
}

/* ------------------------------------------------------------------ P3P */
static int cubic_real_roots(double c3, double c2, double c1, double c0, double* roots)
{
    double scale = fabs(c3) + fabs(c2) + fabs(c1) + fabs(c0);
    if (!(scale > 0) || !isfinite(scale)) return 0;
    int n = 0;
    if (fabs(c3) < 1e-14 * scale) {
        if (fabs(c2) < 1e-14 * scale) {
            if (fabs(c1) < 1e-14 * scale) return 0;
            roots[0] = -c0 / c1;
            return 1;
        }
        double disc = c1 * c1 - 4 * c2 * c0;
        if (disc < 0) return 0;
        double q = -0.5 * (c1 + (c1 >= 0 ? 1 : -1) * sqrt(disc));
        roots[n++] = q / c2;
        if (q != 0) roots[n++] = c0 / q;
        return n;
    }
    double a = c2 / c3, b = c1 / c3, c = c0 / c3;
    double Q = (a * a - 3 * b) / 9, Rr = (2 * a * a * a - 9 * a * b + 27 * c) / 54;
    double Q3 = Q * Q * Q;
    if (Rr * Rr < Q3) {
        double th = acos(Rr / sqrt(Q3));
        double sq = -2 * sqrt(Q);
        roots[0] = sq * cos(th / 3) - a / 3;
        roots[1] = sq * cos((th + 2 * M_PI) / 3) - a / 3;
        roots[2] = sq * cos((th - 2 * M_PI) / 3) - a / 3;
        n = 3;
    } else {
        double A = -(Rr >= 0 ? 1 : -1) * cbrt(fabs(Rr) + sqrt(Rr * Rr - Q3));
        double B = A != 0 ? Q / A : 0;
        roots[0] = (A + B) - a / 3;
        n = 1;
    }
    for (int i = 0; i < n; ++i) {   /* Newton polish on the original polynomial */
        double x = roots[i];
        for (int it = 0; it < 3; ++it) {
            double f = ((c3 * x + c2) * x + c1) * x + c0;
            double df = (3 * c3 * x + 2 * c2) * x + c1;
            if (df == 0) break;
            double dx = f / df;
            if (!isfinite(dx)) break;
            x -= dx;
        }
        if (isfinite(x)) roots[i] = x;
    }
    return n;
}

static double det_cols(const double* A, const double* B, int pick)
{   /* det of the matrix whose column i is B's column i if i==pick else A's */
    double M[9];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) M[3 * r + c] = (c == pick) ? B[3 * r + c] : A[3 * r + c];
    return det3(M);
}

/* X: 3 object points (row-major 3x3); xn: 3 normalised image points (x,y) -> up to 4 poses */
int orc_p3p(const double X[9], const double xn[6], double R[4][9], double t[4][3])
{
    double b[3][3];
    for (int i = 0; i < 3; ++i) {
        double x = xn[2 * i], y = xn[2 * i + 1];
        double inv = 1.0 / sqrt(x * x + y * y + 1.0);
        b[i][0] = x * inv; b[i][1] = y * inv; b[i][2] = inv;
    }
    double d12[3], d13[3], d23[3];
    for (int k = 0; k < 3; ++k) { d12[k] = X[k] - X[3 + k]; d13[k] = X[k] - X[6 + k]; d23[k] = X[3 + k] - X[6 + k]; }
    double a12 = d12[0] * d12[0] + d12[1] * d12[1] + d12[2] * d12[2];
    double a13 = d13[0] * d13[0] + d13[1] * d13[1] + d13[2] * d13[2];
    double a23 = d23[0] * d23[0] + d23[1] * d23[1] + d23[2] * d23[2];
    double c12 = b[0][0] * b[1][0] + b[0][1] * b[1][1] + b[0][2] * b[1][2];
    double c13 = b[0][0] * b[2][0] + b[0][1] * b[2][1] + b[0][2] * b[2][2];
    double c23 = b[1][0] * b[2][0] + b[1][1] * b[2][1] + b[1][2] * b[2][2];
    double M12[9] = {1, -c12, 0, -c12, 1, 0, 0, 0, 0};
    double M13[9] = {1, 0, -c13, 0, 0, 0, -c13, 0, 1};
    double M23[9] = {0, 0, 0, 0, 1, -c23, 0, -c23, 1};
    double D1[9], D2[9];
    for (int i = 0; i < 9; ++i) { D1[i] = M12[i] * a23 - M23[i] * a12; D2[i] = M13[i] * a23 - M23[i] * a13; }
    double k3 = det3(D2), k0 = det3(D1);
    double k2 = det_cols(D2, D1, 0) + det_cols(D2, D1, 1) + det_cols(D2, D1, 2);
    double k1 = det_cols(D1, D2, 0) + det_cols(D1, D2, 1) + det_cols(D1, D2, 2);
    double roots[3];
    int nr = cubic_real_roots(k3, k2, k1, k0, roots);
    double best_sc = 0, g = 0, D0[9], B[9];
    int bi = -1;
    for (int r = 0; r < nr; ++r) {
        double T[9], Bt[9];
        double fro = 0;
        for (int i = 0; i < 9; ++i) { T[i] = D1[i] + roots[r] * D2[i]; fro += T[i] * T[i]; }
        /* B = -adj(T), T symmetric */
        Bt[0] = -(T[4] * T[8] - T[5] * T[5]); Bt[1] = -(T[2] * T[5] - T[1] * T[8]); Bt[2] = -(T[1] * T[5] - T[2] * T[4]);
        Bt[3] = Bt[1]; Bt[4] = -(T[0] * T[8] - T[2] * T[2]); Bt[5] = -(T[1] * T[2] - T[0] * T[5]);
        Bt[6] = Bt[2]; Bt[7] = Bt[5]; Bt[8] = -(T[0] * T[4] - T[1] * T[1]);
        int im = 0;
        if (fabs(Bt[4]) > fabs(Bt[0])) im = 1;
        if (fabs(Bt[8]) > fabs(Bt[4 * im])) im = 2;
        double sc = Bt[4 * im] / (fro + 1e-300);
        if (sc > best_sc) {
            best_sc = sc; g = roots[r]; bi = im;
            memcpy(D0, T, sizeof(T)); memcpy(B, Bt, sizeof(Bt));
        }
    }
    if (bi < 0) return 0;
    double sb = sqrt(B[4 * bi]);
    double p[3] = {B[bi] / sb, B[3 + bi] / sb, B[6 + bi] / sb};
    double C[9] = {D0[0], D0[1] - p[2], D0[2] + p[1], D0[3] + p[2], D0[4], D0[5] - p[0], D0[6] - p[1], D0[7] + p[0], D0[8]};
    int rm = 0, cm = 0;
    double cmax = -1;
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c)
            if (fabs(C[3 * r + c]) > cmax) { cmax = fabs(C[3 * r + c]); rm = r; cm = c; }
    double lines[2][3] = {{C[3 * rm], C[3 * rm + 1], C[3 * rm + 2]}, {C[cm], C[3 + cm], C[6 + cm]}};
    const double* Q = fabs(g) < 1 ? D2 : D1;
    const double* Ms; double as;
    if (a12 >= a13 && a12 >= a23) { Ms = M12; as = a12; }
    else if (a13 >= a23) { Ms = M13; as = a13; }
    else { Ms = M23; as = a23; }
    double Xm[9], Xi[9], cr[3];
    cross3(d12, d13, cr);
    for (int k = 0; k < 3; ++k) { Xm[3 * k] = d12[k]; Xm[3 * k + 1] = d13[k]; Xm[3 * k + 2] = cr[k]; }
    if (!inv3(Xm, Xi)) return 0;
    int ns = 0;
    for (int li = 0; li < 2; ++li) {
        const double* l = lines[li];
        int k = 0;
        if (fabs(l[1]) > fabs(l[0])) k = 1;
        if (fabs(l[2]) > fabs(l[k])) k = 2;
        if (l[k] == 0) continue;
        int i0 = k == 0 ? 1 : 0, i1 = k == 2 ? 1 : 2;
        double u[3] = {0, 0, 0}, v[3] = {0, 0, 0};
        u[i0] = 1; u[k] = -l[i0] / l[k];
        v[i1] = 1; v[k] = -l[i1] / l[k];
        double A = quad3(Q, v, v), Bq = quad3(Q, u, v), Cq = quad3(Q, u, u);
        double disc = Bq * Bq - A * Cq;
        if (!(disc >= 0)) continue;
        double sq = sqrt(disc);
        double qq = -(Bq + (Bq >= 0 ? sq : -sq));
        double taus[2];
        int nt = 0;
        if (A != 0) taus[nt++] = qq / A;
        if (qq != 0) taus[nt++] = Cq / qq;
        for (int ti = 0; ti < nt && ns < 4; ++ti) {
            double tau = taus[ti];
            if (!(tau > 0)) continue;
            double w[3] = {u[0] + tau * v[0], u[1] + tau * v[1], u[2] + tau * v[2]};
            double den = quad3(Ms, w, w);
            if (!(den > 0)) continue;
            double sc = sqrt(as / den);
            double lam[3] = {sc * w[0], sc * w[1], sc * w[2]};
            if (!(lam[0] > 0 && lam[1] > 0 && lam[2] > 0)) continue;
            for (int it = 0; it < 3; ++it) {   /* Gauss-Newton polish on the three distance equations */
                double f[3] = {quad3(M12, lam, lam) - a12, quad3(M13, lam, lam) - a13, quad3(M23, lam, lam) - a23};
                double J[9], Ji[9];
                for (int c = 0; c < 3; ++c) {
                    J[c] = 2 * (M12[c] * lam[0] + M12[3 + c] * lam[1] + M12[6 + c] * lam[2]);
                    J[3 + c] = 2 * (M13[c] * lam[0] + M13[3 + c] * lam[1] + M13[6 + c] * lam[2]);
                    J[6 + c] = 2 * (M23[c] * lam[0] + M23[3 + c] * lam[1] + M23[6 + c] * lam[2]);
                }
                if (!inv3(J, Ji)) break;
                for (int c = 0; c < 3; ++c) lam[c] -= Ji[3 * c] * f[0] + Ji[3 * c + 1] * f[1] + Ji[3 * c + 2] * f[2];
            }
            if (!(lam[0] > 0 && lam[1] > 0 && lam[2] > 0) || !isfinite(lam[0] + lam[1] + lam[2])) continue;
            double Y[3][3];
            for (int i = 0; i < 3; ++i)
                for (int c = 0; c < 3; ++c) Y[i][c] = lam[i] * b[i][c];
            double e1[3], e2[3], e3[3], Ym[9];
            for (int c = 0; c < 3; ++c) { e1[c] = Y[0][c] - Y[1][c]; e2[c] = Y[0][c] - Y[2][c]; }
            cross3(e1, e2, e3);
            for (int c = 0; c < 3; ++c) { Ym[3 * c] = e1[c]; Ym[3 * c + 1] = e2[c]; Ym[3 * c + 2] = e3[c]; }
            mat3mul(Ym, Xi, R[ns]);
            for (int c = 0; c < 3; ++c)
                t[ns][c] = Y[0][c] - (R[ns][3 * c] * X[0] + R[ns][3 * c + 1] * X[1] + R[ns][3 * c + 2] * X[2]);
            int fin = 1;
            for (int c = 0; c < 9; ++c) fin &= isfinite(R[ns][c]) != 0;
            for (int c = 0; c < 3; ++c) fin &= isfinite(t[ns][c]) != 0;
            if (fin) ++ns;
        }
    }
    return ns;
}

static void project_pt(const double R[9], const double t[3], const double K[9], double X, double Y, double Z,
                       double* u, double* v)
{   /* cvProjectPoints2 with zero distortion (operation order preserved) */
    double x = R[0] * X + R[1] * Y + R[2] * Z + t[0];
    double y = R[3] * X + R[4] * Y + R[5] * Z + t[1];
    double z = R[6] * X + R[7] * Y + R[8] * Z + t[2];
    z = z ? 1. / z : 1;
    x *= z; y *= z;
    *u = x * K[0] + K[2];
    *v = y * K[4] + K[5];
}

int orc_pnp_minimal(const float* obj4, const float* img4, const double K[9], double rvec[3], double tvec[3])
{
    double X[12], xn[8];
    const double ifx = 1. / K[0], ify = 1. / K[4];
    for (int i = 0; i < 4; ++i) {
        for (int c = 0; c < 3; ++c) X[3 * i + c] = obj4[3 * i + c];
        /* undistortPoints on CV_32F input: double arithmetic, result stored as float32 */
        xn[2 * i] = (double)(float)(((double)img4[2 * i] - K[2]) * ifx);
        xn[2 * i + 1] = (double)(float)(((double)img4[2 * i + 1] - K[5]) * ify);
    }
    double R[4][9], t[4][3];
    int ns = orc_p3p(X, xn, R, t);
    if (ns == 0) return 0;
    int best = -1;
    double best_e = 0;
    for (int s = 0; s < ns; ++s) {
        double e = 0;
        for (int i = 0; i < 4; ++i) {
            double u, v;
            project_pt(R[s], t[s], K, X[3 * i], X[3 * i + 1], X[3 * i + 2], &u, &v);
            double dx = (double)img4[2 * i] - u, dy = (double)img4[2 * i + 1] - v;
            e += dx * dx + dy * dy;
        }
        if (!isfinite(e)) continue;
        if (best < 0 || e < best_e) { best = s; best_e = e; }
    }
    if (best < 0) return 0;
    orc_R_to_rodrigues(R[best], rvec);
    for (int c = 0; c < 3; ++c) tvec[c] = t[best][c];
    return 1;
}

void orc_pnp_errors(const float* obj, const float* img, int n, const double K[9],
                    const double rvec[3], const double tvec[3], float* err)
{
    double R[9];
    orc_rodrigues_to_R(rvec, R);
    for (int i = 0; i < n; ++i) {
        double u, v;
        project_pt(R, tvec, K, obj[3 * i], obj[3 * i + 1], obj[3 * i + 2], &u, &v);
        float pu = (float)u, pv = (float)v;
        float dx = img[2 * i] - pu, dy = img[2 * i + 1] - pv;
        float xx = dx * dx, yy = dy * dy;
        err[i] = xx + yy;
    }
}

/* ------------------------------------------------------------------ EPnP */
static void svd3_sym_desc(const double* S, double* d, double* Ut)
{   /* cvSVD(S, D, Ut, 0, CV_SVD_U_T): rows of Ut = left singular vectors */
    double U[9], Vt[9];
    orc_jacobi_svd(S, 3, 3, d, U, Vt);
    for (int i = 0; i < 3; ++i) for (int k = 0; k < 3; ++k) Ut[3 * i + k] = U[3 * k + i];
}

static void svd3(const double* A, double* U, double* s, double* V)
{   /* A = U diag(s) V^T */
    double Vt[9];
    orc_jacobi_svd(A, 3, 3, s, U, Vt);
    for (int i = 0; i < 3; ++i) for (int k = 0; k < 3; ++k) V[3 * i + k] = Vt[3 * k + i];
}

typedef struct {
    int n;
    const double* pws;   /* n x 3 */
    double* us;          /* n x 2 */
    double* alphas;      /* n x 4 */
    double* pcs;         /* n x 3 */
    double cws[4][3], ccs[4][3];
    double fu, fv, uc, vc;
} epnp_t;

static double epnp_R_t(epnp_t* e, const double* ut, const double* betas, double R[9], double t[3])
{
    const double* v[4] = {ut + 12 * 11, ut + 12 * 10, ut + 12 * 9, ut + 12 * 8};
    for (int i = 0; i < 4; ++i)
        for (int k = 0; k < 3; ++k) {
            e->ccs[i][k] = 0;
            for (int j = 0; j < 4; ++j) e->ccs[i][k] += betas[j] * v[j][3 * i + k];
        }
    for (int i = 0; i < e->n; ++i)
        for (int k = 0; k < 3; ++k)
            e->pcs[3 * i + k] = e->alphas[4 * i] * e->ccs[0][k] + e->alphas[4 * i + 1] * e->ccs[1][k] +
                                e->alphas[4 * i + 2] * e->ccs[2][k] + e->alphas[4 * i + 3] * e->ccs[3][k];
    if (e->pcs[2] < 0) {
        for (int i = 0; i < 4; ++i) for (int k = 0; k < 3; ++k) e->ccs[i][k] = -e->ccs[i][k];
        for (int i = 0; i < 3 * e->n; ++i) e->pcs[i] = -e->pcs[i];
    }
    double pc0[3] = {0, 0, 0}, pw0[3] = {0, 0, 0};
    for (int i = 0; i < e->n; ++i)
        for (int k = 0; k < 3; ++k) { pc0[k] += e->pcs[3 * i + k]; pw0[k] += e->pws[3 * i + k]; }
    for (int k = 0; k < 3; ++k) { pc0[k] /= e->n; pw0[k] /= e->n; }
    double abt[9] = {0};
    for (int i = 0; i < e->n; ++i)
        for (int j = 0; j < 3; ++j)
            for (int k = 0; k < 3; ++k) abt[3 * j + k] += (e->pcs[3 * i + j] - pc0[j]) * (e->pws[3 * i + k] - pw0[k]);
    double U[9], s[3], V[9];
    svd3(abt, U, s, V);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R[3 * i + j] = U[3 * i] * V[3 * j] + U[3 * i + 1] * V[3 * j + 1] + U[3 * i + 2] * V[3 * j + 2];
    if (det3(R) < 0) { R[6] = -R[6]; R[7] = -R[7]; R[8] = -R[8]; }
    for (int k = 0; k < 3; ++k) t[k] = pc0[k] - (R[3 * k] * pw0[0] + R[3 * k + 1] * pw0[1] + R[3 * k + 2] * pw0[2]);
    double sum2 = 0;
    for (int i = 0; i < e->n; ++i) {
        const double* pw = e->pws + 3 * i;
        double Xc = R[0] * pw[0] + R[1] * pw[1] + R[2] * pw[2] + t[0];
        double Yc = R[3] * pw[0] + R[4] * pw[1] + R[5] * pw[2] + t[1];
        double iZ = 1.0 / (R[6] * pw[0] + R[7] * pw[1] + R[8] * pw[2] + t[2]);
        double ue = e->uc + e->fu * Xc * iZ, ve = e->vc + e->fv * Yc * iZ;
        double du = e->us[2 * i] - ue, dv = e->us[2 * i + 1] - ve;
        sum2 += sqrt(du * du + dv * dv);
    }
    return sum2 / e->n;
}

static void epnp_gauss_newton(const double* L, const double* rho, double* b)
{
    for (int it = 0; it < 5; ++it) {
        double A[24], B[6], x[4];
        for (int i = 0; i < 6; ++i) {
            const double* r = L + 10 * i;
            A[4 * i + 0] = 2 * r[0] * b[0] + r[1] * b[1] + r[3] * b[2] + r[6] * b[3];
            A[4 * i + 1] = r[1] * b[0] + 2 * r[2] * b[1] + r[4] * b[2] + r[7] * b[3];
            A[4 * i + 2] = r[3] * b[0] + r[4] * b[1] + 2 * r[5] * b[2] + r[8] * b[3];
            A[4 * i + 3] = r[6] * b[0] + r[7] * b[1] + r[8] * b[2] + 2 * r[9] * b[3];
            B[i] = rho[i] - (r[0] * b[0] * b[0] + r[1] * b[0] * b[1] + r[2] * b[1] * b[1] + r[3] * b[0] * b[2] +
                             r[4] * b[1] * b[2] + r[5] * b[2] * b[2] + r[6] * b[0] * b[3] + r[7] * b[1] * b[3] +
                             r[8] * b[2] * b[3] + r[9] * b[3] * b[3]);
        }
        qr_lstsq(A, B, 6, 4, x);
        for (int k = 0; k < 4; ++k) b[k] += x[k];
    }
}

int orc_epnp(const double* obj, const double* img, int n, const double K[9], double R[9], double t[3])
{
    if (n < 4) return 0;
    epnp_t e;
    e.n = n; e.pws = obj;
    e.fu = K[0]; e.fv = K[4]; e.uc = K[2]; e.vc = K[5];
    e.us = (double*)malloc(sizeof(double) * 2 * n);
    e.alphas = (double*)malloc(sizeof(double) * 4 * n);
    e.pcs = (double*)malloc(sizeof(double) * 3 * n);
    const double ifx = 1. / K[0], ify = 1. / K[4];
    for (int i = 0; i < n; ++i) {   /* undistortPoints (float64) then back to pixels, as cv2 does */
        e.us[2 * i] = ((img[2 * i] - K[2]) * ifx) * e.fu + e.uc;
        e.us[2 * i + 1] = ((img[2 * i + 1] - K[5]) * ify) * e.fv + e.vc;
    }
    /* control points */
    for (int k = 0; k < 3; ++k) e.cws[0][k] = 0;
    for (int i = 0; i < n; ++i) for (int k = 0; k < 3; ++k) e.cws[0][k] += obj[3 * i + k];
    for (int k = 0; k < 3; ++k) e.cws[0][k] /= n;
    double S[9] = {0};
    for (int i = 0; i < n; ++i)
        for (int a = 0; a < 3; ++a)
            for (int b2 = 0; b2 < 3; ++b2) S[3 * a + b2] += (obj[3 * i + a] - e.cws[0][a]) * (obj[3 * i + b2] - e.cws[0][b2]);
    double dc[3], uct[9];
    svd3_sym_desc(S, dc, uct);
    for (int i = 1; i < 4; ++i) {
        double k = sqrt((dc[i - 1] > 0 ? dc[i - 1] : 0) / n);
        for (int j = 0; j < 3; ++j) e.cws[i][j] = e.cws[0][j] + k * uct[3 * (i - 1) + j];
    }
    double cc[9], ci[9];
    for (int i = 0; i < 3; ++i) for (int j = 1; j < 4; ++j) cc[3 * i + j - 1] = e.cws[j][i] - e.cws[0][i];
    if (!inv3(cc, ci)) { free(e.us); free(e.alphas); free(e.pcs); return 0; }
    for (int i = 0; i < n; ++i) {
        const double* pi = obj + 3 * i;
        double* a = e.alphas + 4 * i;
        for (int j = 0; j < 3; ++j)
            a[1 + j] = ci[3 * j] * (pi[0] - e.cws[0][0]) + ci[3 * j + 1] * (pi[1] - e.cws[0][1]) + ci[3 * j + 2] * (pi[2] - e.cws[0][2]);
        a[0] = 1.0 - a[1] - a[2] - a[3];
    }
    /* MtM */
    double MtM[144] = {0};
    for (int i = 0; i < n; ++i) {
        double m1[12], m2[12];
        const double* a = e.alphas + 4 * i;
        for (int j = 0; j < 4; ++j) {
            m1[3 * j] = a[j] * e.fu; m1[3 * j + 1] = 0; m1[3 * j + 2] = a[j] * (e.uc - e.us[2 * i]);
            m2[3 * j] = 0; m2[3 * j + 1] = a[j] * e.fv; m2[3 * j + 2] = a[j] * (e.vc - e.us[2 * i + 1]);
        }
        for (int r = 0; r < 12; ++r)
            for (int c = 0; c < 12; ++c) MtM[12 * r + c] += m1[r] * m1[c] + m2[r] * m2[c];
    }
    double w[12], Um[144], Vtm[144], ut[144];
    orc_jacobi_svd(MtM, 12, 12, w, Um, Vtm);
    for (int i = 0; i < 12; ++i)
        for (int k = 0; k < 12; ++k) ut[12 * i + k] = Um[12 * k + i];   /* rows: descending singular value */
    /* L (6x10) and rho */
    const double* v[4] = {ut + 12 * 11, ut + 12 * 10, ut + 12 * 9, ut + 12 * 8};
    static const int pa[6] = {0, 0, 0, 1, 1, 2}, pb[6] = {1, 2, 3, 2, 3, 3};
    double dv[4][6][3];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 6; ++j)
            for (int k = 0; k < 3; ++k) dv[i][j][k] = v[i][3 * pa[j] + k] - v[i][3 * pb[j] + k];
#define DOT(p, q) ((p)[0] * (q)[0] + (p)[1] * (q)[1] + (p)[2] * (q)[2])
    double L[60], rho[6];
    for (int i = 0; i < 6; ++i) {
        double* r = L + 10 * i;
        r[0] = DOT(dv[0][i], dv[0][i]); r[1] = 2 * DOT(dv[0][i], dv[1][i]); r[2] = DOT(dv[1][i], dv[1][i]);
        r[3] = 2 * DOT(dv[0][i], dv[2][i]); r[4] = 2 * DOT(dv[1][i], dv[2][i]); r[5] = DOT(dv[2][i], dv[2][i]);
        r[6] = 2 * DOT(dv[0][i], dv[3][i]); r[7] = 2 * DOT(dv[1][i], dv[3][i]); r[8] = 2 * DOT(dv[2][i], dv[3][i]);
        r[9] = DOT(dv[3][i], dv[3][i]);
        double d[3] = {e.cws[pa[i]][0] - e.cws[pb[i]][0], e.cws[pa[i]][1] - e.cws[pb[i]][1], e.cws[pa[i]][2] - e.cws[pb[i]][2]};
        rho[i] = DOT(d, d);
    }
    double betas[4][4], Rs[4][9], ts[4][3], rep[4];
    {   /* approx 1: [B11 B12 B13 B14] */
        double A[24], b4[4];
        for (int i = 0; i < 6; ++i) { A[4 * i] = L[10 * i]; A[4 * i + 1] = L[10 * i + 1]; A[4 * i + 2] = L[10 * i + 3]; A[4 * i + 3] = L[10 * i + 6]; }
        qr_lstsq(A, rho, 6, 4, b4);
        double* be = betas[1];
        if (b4[0] < 0) { be[0] = sqrt(-b4[0]); be[1] = -b4[1] / be[0]; be[2] = -b4[2] / be[0]; be[3] = -b4[3] / be[0]; }
        else { be[0] = sqrt(b4[0]); be[1] = b4[1] / be[0]; be[2] = b4[2] / be[0]; be[3] = b4[3] / be[0]; }
    }
    {   /* approx 2: [B11 B12 B22] */
        double A[18], b3[3];
        for (int i = 0; i < 6; ++i) { A[3 * i] = L[10 * i]; A[3 * i + 1] = L[10 * i + 1]; A[3 * i + 2] = L[10 * i + 2]; }
        qr_lstsq(A, rho, 6, 3, b3);
        double* be = betas[2];
        if (b3[0] < 0) { be[0] = sqrt(-b3[0]); be[1] = (b3[2] < 0) ? sqrt(-b3[2]) : 0.0; }
        else { be[0] = sqrt(b3[0]); be[1] = (b3[2] > 0) ? sqrt(b3[2]) : 0.0; }
        if (b3[1] < 0) be[0] = -be[0];
        be[2] = 0; be[3] = 0;
    }
    {   /* approx 3: [B11 B12 B22 B13 B23] */
        double A[30], b5[5];
        for (int i = 0; i < 6; ++i) for (int k = 0; k < 5; ++k) A[5 * i + k] = L[10 * i + k];
        qr_lstsq(A, rho, 6, 5, b5);
        double* be = betas[3];
        if (b5[0] < 0) { be[0] = sqrt(-b5[0]); be[1] = (b5[2] < 0) ? sqrt(-b5[2]) : 0.0; }
        else { be[0] = sqrt(b5[0]); be[1] = (b5[2] > 0) ? sqrt(b5[2]) : 0.0; }
        if (b5[1] < 0) be[0] = -be[0];
        be[2] = b5[3] / be[0]; be[3] = 0;
    }
    for (int N = 1; N <= 3; ++N) {
        epnp_gauss_newton(L, rho, betas[N]);
        rep[N] = epnp_R_t(&e, ut, betas[N], Rs[N], ts[N]);
    }
    int N = 1;
    if (rep[2] < rep[1]) N = 2;
    if (rep[3] < rep[N]) N = 3;
    memcpy(R, Rs[N], sizeof(double) * 9);
    memcpy(t, ts[N], sizeof(double) * 3);
    free(e.us); free(e.alphas); free(e.pcs);
    int fin = 1;
    for (int c = 0; c < 9; ++c) fin &= isfinite(R[c]) != 0;
    return fin;
}

/* ------------------------------------------------------------------ full call */
int orc_solve_pnp_ransac_p3p(const float* obj, const float* img, int n, const double K[9], int iters,
                             float reproj_err, double conf, double rvec[3], double tvec[3],
                             int32_t* inliers, int* n_inliers, int* success, int* iters_run)
{
    *n_inliers = 0; *success = 0;
    if (iters_run) *iters_run = 0;
    if (n < 4) return -1;
    uint8_t* mask = (uint8_t*)calloc((size_t)n, 1);
    double brv[3] = {0, 0, 0}, btv[3] = {0, 0, 0};
    if (n == 4) {
        if (!orc_pnp_minimal(obj, img, K, rvec, tvec)) { free(mask); return 0; }
        for (int i = 0; i < 4; ++i) inliers[i] = i;
        *n_inliers = 4; *success = 1;
        free(mask);
        return 0;
    }
    int max_iters = iters > 1 ? iters : 1;
    int32_t* subsets = (int32_t*)malloc(sizeof(int32_t) * 4 * (size_t)max_iters);
    orc_ransac_subsets(n, 4, max_iters, subsets);
    float* err = (float*)malloc(sizeof(float) * (size_t)n);
    const float thr = (float)((double)reproj_err * (double)reproj_err);
    int niters = max_iters, max_good = 0, it;
    for (it = 0; it < niters; ++it) {
        float o4[12], i4[8];
        for (int k = 0; k < 4; ++k) {
            int s = subsets[4 * it + k];
            memcpy(o4 + 3 * k, obj + 3 * s, 12);
            memcpy(i4 + 2 * k, img + 2 * s, 8);
        }
        double rv[3], tv[3];
        if (!orc_pnp_minimal(o4, i4, K, rv, tv)) continue;
        orc_pnp_errors(obj, img, n, K, rv, tv, err);
        int good = 0;
        for (int i = 0; i < n; ++i) good += err[i] <= thr;
        if (good > (max_good > 3 ? max_good : 3)) {
            for (int i = 0; i < n; ++i) mask[i] = err[i] <= thr;
            memcpy(brv, rv, sizeof(rv)); memcpy(btv, tv, sizeof(tv));
            max_good = good;
            niters = orc_ransac_update_num_iters(conf, (double)(n - good) / n, 4, niters);
        }
    }
    if (iters_run) *iters_run = it;
    free(subsets); free(err);
    if (max_good <= 0) { free(mask); return 0; }
    double* o64 = (double*)malloc(sizeof(double) * 3 * (size_t)max_good);
    double* i64 = (double*)malloc(sizeof(double) * 2 * (size_t)max_good);
    int m = 0;
    for (int i = 0; i < n; ++i)
        if (mask[i]) {
            inliers[m] = i;
            for (int c = 0; c < 3; ++c) o64[3 * m + c] = obj[3 * i + c];
            for (int c = 0; c < 2; ++c) i64[2 * m + c] = img[2 * i + c];
            ++m;
        }
    double R[9], t[3];
    if (orc_epnp(o64, i64, m, K, R, t)) {
        orc_R_to_rodrigues(R, rvec);
        memcpy(tvec, t, sizeof(t));
    } else {
        memcpy(rvec, brv, sizeof(brv)); memcpy(tvec, btv, sizeof(btv));
    }
    *n_inliers = m; *success = 1;
    free(o64); free(i64); free(mask);
    return 0;
}
