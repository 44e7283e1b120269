// FP64 device helpers for the minimal solvers and the EPnP refit (compiled with
// --fmad=false: every multiply and add is a separate IEEE operation, as in OpenCV's
// x86-64 baseline build that the results are compared with).
#pragma once
#include <cuda_runtime.h>
#include <cfloat>
#include <cmath>

namespace vo {

// 1/sqrt(x) and 1/x to ~1 ulp without the library's IEEE routines (long dependent FP64 chains on this part): hardware
// float seed (22 bits) + two Newton steps in double.  Outside the float range of the seed the exact forms are used.
__device__ __forceinline__ double rsqrt_fast(double x)
{
    if (!(x > 1e-30 && x < 1e30)) return 1.0 / sqrt(x);
    float yf;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(yf) : "f"((float)x));
    double y = (double)yf;
    const double hx = 0.5 * x;
    y = fma(y, fma(-hx * y, y, 0.5), y);
    y = fma(y, fma(-hx * y, y, 0.5), y);
    return y;
}
__device__ __forceinline__ double rcp_fast(double x)
{
    const double ax = fabs(x);
    if (!(ax > 1e-30 && ax < 1e30)) return 1.0 / x;
    float yf;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(yf) : "f"((float)x));
    double y = (double)yf;
    y = fma(y, fma(-x, y, 1.0), y);
    y = fma(y, fma(-x, y, 1.0), y);
    return y;
}

__device__ __forceinline__ double det3(const double* M)
{
    return M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
}

__device__ __forceinline__ bool inv3(const double* M, double* I)
{
    const double d = det3(M);
    if (d == 0 || !isfinite(d)) return false;
    const double id = 1.0 / d;
    I[0] = (M[4] * M[8] - M[5] * M[7]) * id; I[1] = (M[2] * M[7] - M[1] * M[8]) * id; I[2] = (M[1] * M[5] - M[2] * M[4]) * id;
    I[3] = (M[5] * M[6] - M[3] * M[8]) * id; I[4] = (M[0] * M[8] - M[2] * M[6]) * id; I[5] = (M[2] * M[3] - M[0] * M[5]) * id;
    I[6] = (M[3] * M[7] - M[4] * M[6]) * id; I[7] = (M[1] * M[6] - M[0] * M[7]) * id; I[8] = (M[0] * M[4] - M[1] * M[3]) * id;
    return true;
}

__device__ __forceinline__ void cross3(const double* a, const double* b, double* c)
{
    c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
}

__device__ __forceinline__ double quad3(const double* M, const double* u, const double* v)
{
    double s = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i) s += u[i] * (M[3 * i] * v[0] + M[3 * i + 1] * v[1] + M[3 * i + 2] * v[2]);
    return s;
}

// cv::Rodrigues, vector -> matrix
__device__ inline void rodrigues_to_R(const double* r, double* R)
{
    const double theta = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    if (theta < DBL_EPSILON) {
        for (int i = 0; i < 9; ++i) R[i] = (i % 4) == 0 ? 1.0 : 0.0;
        return;
    }
    const double c = cos(theta), s = sin(theta), c1 = 1 - c, it = 1 / theta;
    const double x = r[0] * it, y = r[1] * it, z = r[2] * it;
    R[0] = c + c1 * x * x; R[1] = c1 * x * y - s * z; R[2] = c1 * x * z + s * y;
    R[3] = c1 * x * y + s * z; R[4] = c + c1 * y * y; R[5] = c1 * y * z - s * x;
    R[6] = c1 * x * z - s * y; R[7] = c1 * y * z + s * x; R[8] = c + c1 * z * z;
}

// cv::Rodrigues, matrix -> vector (R assumed orthonormal)
__device__ inline void R_to_rodrigues(const double* R, double* r)
{
    double rx = R[7] - R[5], ry = R[2] - R[6], rz = R[3] - R[1];
    const double s = sqrt((rx * rx + ry * ry + rz * rz) * 0.25);
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

// cvProjectPoints2 with zero distortion, operation order preserved
__device__ __forceinline__ void project_pt(const double* R, const double* t, double fx, double fy, double cx, double cy,
                                           double X, double Y, double Z, double& u, double& v)
{
    double x = R[0] * X + R[1] * Y + R[2] * Z + t[0];
    double y = R[3] * X + R[4] * Y + R[5] * Z + t[1];
    double z = R[6] * X + R[7] * Y + R[8] * Z + t[2];
    z = z != 0 ? 1. / z : 1;
    x *= z; y *= z;
    u = x * fx + cx;
    v = y * fy + cy;
}

// One-sided (Hestenes) Jacobi SVD of an N x N matrix, same rotation order, thresholds and
// final descending sort as the routine OpenCV uses below 25 x 25 -- singular-vector SIGNS
// therefore match cv2's (EPnP's control points depend on them).  A = U diag(W) Vt.
// At: N*N workspace (rows = columns of A on entry; rows = U columns scaled on exit).
template <int N>
__device__ inline void jacobi_svd(const double* A, double* W, double* U, double* Vt, double* At)
{
    const double eps = DBL_EPSILON * 10;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        double sd = 0;
#pragma unroll
        for (int k = 0; k < N; ++k) { At[i * N + k] = A[k * N + i]; sd += At[i * N + k] * At[i * N + k]; }
        W[i] = sd;
#pragma unroll
        for (int k = 0; k < N; ++k) Vt[i * N + k] = (i == k) ? 1.0 : 0.0;
    }
#pragma unroll 1
    for (int iter = 0; iter < 30; ++iter) {
        bool changed = false;
#pragma unroll
        for (int i = 0; i < N - 1; ++i)
#pragma unroll
            for (int j = i + 1; j < N; ++j) {
                double* Ai = At + i * N;
                double* Aj = At + j * N;
                double a = W[i], p = 0, b = W[j];
#pragma unroll
                for (int k = 0; k < N; ++k) p += Ai[k] * Aj[k];
                if (fabs(p) <= eps * sqrt(a * b)) continue;
                p *= 2;
                const double beta = a - b, gamma = hypot(p, beta);
                double c, s;
                if (beta < 0) {
                    const double delta = (gamma - beta) * 0.5;
                    s = sqrt(delta / gamma);
                    c = p / (gamma * s * 2);
                } else {
                    c = sqrt((gamma + beta) / (gamma * 2));
                    s = p / (gamma * c * 2);
                }
                a = b = 0;
#pragma unroll
                for (int k = 0; k < N; ++k) {
                    const double t0 = c * Ai[k] + s * Aj[k];
                    const double t1 = -s * Ai[k] + c * Aj[k];
                    Ai[k] = t0; Aj[k] = t1;
                    a += t0 * t0; b += t1 * t1;
                }
                W[i] = a; W[j] = b;
                changed = true;
                double* Vi = Vt + i * N;
                double* Vj = Vt + j * N;
#pragma unroll
                for (int k = 0; k < N; ++k) {
                    const double t0 = c * Vi[k] + s * Vj[k];
                    const double t1 = -s * Vi[k] + c * Vj[k];
                    Vi[k] = t0; Vj[k] = t1;
                }
            }
        if (!changed) break;
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
        double sd = 0;
#pragma unroll
        for (int k = 0; k < N; ++k) sd += At[i * N + k] * At[i * N + k];
        W[i] = sqrt(sd);
    }
#pragma unroll
    for (int i = 0; i < N - 1; ++i) {
        int j = i;
#pragma unroll
        for (int k = i + 1; k < N; ++k) if (W[j] < W[k]) j = k;
        if (i != j) {
            double tw = W[i]; W[i] = W[j]; W[j] = tw;
#pragma unroll
            for (int k = 0; k < N; ++k) { double tt = At[i * N + k]; At[i * N + k] = At[j * N + k]; At[j * N + k] = tt; }
#pragma unroll
            for (int k = 0; k < N; ++k) { double tt = Vt[i * N + k]; Vt[i * N + k] = Vt[j * N + k]; Vt[j * N + k] = tt; }
        }
    }
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const double sc = W[i] > DBL_MIN ? 1 / W[i] : 0.;
#pragma unroll
        for (int k = 0; k < N; ++k) U[k * N + i] = At[i * N + k] * sc;
    }
}

// least squares min |A x - b| (M x NC, M = 6) by Householder QR
template <int NC>
__device__ inline void qr_lstsq6(const double* Ain, const double* bin, double* x)
{
    constexpr int M = 6;
    double A[M * NC], b[M];
#pragma unroll
    for (int i = 0; i < M * NC; ++i) A[i] = Ain[i];
#pragma unroll
    for (int i = 0; i < M; ++i) b[i] = bin[i];
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        double nrm = 0;
#pragma unroll
        for (int i = k; i < M; ++i) nrm += A[i * NC + k] * A[i * NC + k];
        if (nrm == 0) continue;
        nrm = nrm * rsqrt_fast(nrm);      // sqrt
        const double alpha = A[k * NC + k] > 0 ? -nrm : nrm;
        double v[M];
#pragma unroll
        for (int i = k; i < M; ++i) v[i] = A[i * NC + k];
        v[k] -= alpha;
        double vn = 0;
#pragma unroll
        for (int i = k; i < M; ++i) vn += v[i] * v[i];
        if (vn == 0) continue;
        const double tau = 2.0 * rcp_fast(vn);      // H = I - tau v v^T: one reciprocal per reflector instead of a division per column
#pragma unroll
        for (int j = k; j < NC; ++j) {
            double s = 0;
#pragma unroll
            for (int i = k; i < M; ++i) s += v[i] * A[i * NC + j];
            s *= tau;
#pragma unroll
            for (int i = k; i < M; ++i) A[i * NC + j] -= s * v[i];
        }
        double s = 0;
#pragma unroll
        for (int i = k; i < M; ++i) s += v[i] * b[i];
        s *= tau;
#pragma unroll
        for (int i = k; i < M; ++i) b[i] -= s * v[i];
    }
#pragma unroll
    for (int k = NC - 1; k >= 0; --k) {
        double s = b[k];
#pragma unroll
        for (int j = k + 1; j < NC; ++j) s -= A[k * NC + j] * x[j];
        x[k] = A[k * NC + k] != 0 ? s * rcp_fast(A[k * NC + k]) : 0;
    }
}

__device__ inline int cubic_real_roots(double c3, double c2, double c1, double c0, double* roots)
{
    const double scale = fabs(c3) + fabs(c2) + fabs(c1) + fabs(c0);
    if (!(scale > 0) || !isfinite(scale)) return 0;
    int n = 0;
    if (fabs(c3) < 1e-14 * scale) {
        if (fabs(c2) < 1e-14 * scale) {
            if (fabs(c1) < 1e-14 * scale) return 0;
            roots[0] = -c0 / c1;
            return 1;
        }
        const double disc = c1 * c1 - 4 * c2 * c0;
        if (disc < 0) return 0;
        const double q = -0.5 * (c1 + (c1 >= 0 ? 1 : -1) * sqrt(disc));
        roots[n++] = q / c2;
        if (q != 0) roots[n++] = c0 / q;
        return n;
    }
    const double a = c2 / c3, b = c1 / c3, c = c0 / c3;
    const double Q = (a * a - 3 * b) / 9, Rr = (2 * a * a * a - 9 * a * b + 27 * c) / 54;
    const double Q3 = Q * Q * Q;
    if (Rr * Rr < Q3) {
        const double th = acos(Rr / sqrt(Q3));
        const double sq = -2 * sqrt(Q);
        const double two_pi = 6.283185307179586476925286766559;
        roots[0] = sq * cos(th / 3) - a / 3;
        roots[1] = sq * cos((th + two_pi) / 3) - a / 3;
        roots[2] = sq * cos((th - two_pi) / 3) - a / 3;
        n = 3;
    } else {
        const double A = -(Rr >= 0 ? 1 : -1) * cbrt(fabs(Rr) + sqrt(Rr * Rr - Q3));
        const double B = A != 0 ? Q / A : 0;
        roots[0] = (A + B) - a / 3;
        n = 1;
    }
    for (int i = 0; i < n; ++i) {
        double x = roots[i];
        for (int it = 0; it < 3; ++it) {
            const double f = ((c3 * x + c2) * x + c1) * x + c0;
            const double df = (3 * c3 * x + 2 * c2) * x + c1;
            if (df == 0) break;
            const double dx = f / df;
            if (!isfinite(dx)) break;
            x -= dx;
        }
        if (isfinite(x)) roots[i] = x;
    }
    return n;
}

__device__ __forceinline__ double det_cols(const double* A, const double* B, int pick)
{
    double M[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) M[3 * r + c] = (c == pick) ? B[3 * r + c] : A[3 * r + c];
    return det3(M);
}

// Exact P3P: X = 3 object points (row-major), xn = 3 normalised image points (x,y).
// Pencil of the two homogeneous depth quadrics -> one real root of the cubic -> degenerate
// conic split into two planes -> quadratic per plane -> Gauss-Newton polish -> (R,t) by
// aligning lambda_i b_i with X_i.  Returns the number of poses (<= 4) with positive depths.
__device__ inline int p3p(const double* X, const double* xn, double R[4][9], double t[4][3])
{
    double b[3][3];
    for (int i = 0; i < 3; ++i) {
        const double x = xn[2 * i], y = xn[2 * i + 1];
        const double inv = 1.0 / sqrt(x * x + y * y + 1.0);
        b[i][0] = x * inv; b[i][1] = y * inv; b[i][2] = inv;
    }
    double d12[3], d13[3], d23[3];
    for (int k = 0; k < 3; ++k) { d12[k] = X[k] - X[3 + k]; d13[k] = X[k] - X[6 + k]; d23[k] = X[3 + k] - X[6 + k]; }
    const double a12 = d12[0] * d12[0] + d12[1] * d12[1] + d12[2] * d12[2];
    const double a13 = d13[0] * d13[0] + d13[1] * d13[1] + d13[2] * d13[2];
    const double a23 = d23[0] * d23[0] + d23[1] * d23[1] + d23[2] * d23[2];
    const double c12 = b[0][0] * b[1][0] + b[0][1] * b[1][1] + b[0][2] * b[1][2];
    const double c13 = b[0][0] * b[2][0] + b[0][1] * b[2][1] + b[0][2] * b[2][2];
    const double c23 = b[1][0] * b[2][0] + b[1][1] * b[2][1] + b[1][2] * b[2][2];
    const double M12[9] = {1, -c12, 0, -c12, 1, 0, 0, 0, 0};
    const double M13[9] = {1, 0, -c13, 0, 0, 0, -c13, 0, 1};
    const double M23[9] = {0, 0, 0, 0, 1, -c23, 0, -c23, 1};
    double D1[9], D2[9];
    for (int i = 0; i < 9; ++i) { D1[i] = M12[i] * a23 - M23[i] * a12; D2[i] = M13[i] * a23 - M23[i] * a13; }
    const double k3 = det3(D2), k0 = det3(D1);
    const double k2 = det_cols(D2, D1, 0) + det_cols(D2, D1, 1) + det_cols(D2, D1, 2);
    const double k1 = det_cols(D1, D2, 0) + det_cols(D1, D2, 1) + det_cols(D1, D2, 2);
    double roots[3];
    const int nr = cubic_real_roots(k3, k2, k1, k0, roots);
    double best_sc = 0, g = 0, D0[9], B[9];
    int bi = -1;
    for (int r = 0; r < nr; ++r) {
        double T[9], Bt[9];
        double fro = 0;
        for (int i = 0; i < 9; ++i) { T[i] = D1[i] + roots[r] * D2[i]; fro += T[i] * T[i]; }
        Bt[0] = -(T[4] * T[8] - T[5] * T[5]); Bt[1] = -(T[2] * T[5] - T[1] * T[8]); Bt[2] = -(T[1] * T[5] - T[2] * T[4]);
        Bt[3] = Bt[1]; Bt[4] = -(T[0] * T[8] - T[2] * T[2]); Bt[5] = -(T[1] * T[2] - T[0] * T[5]);
        Bt[6] = Bt[2]; Bt[7] = Bt[5]; Bt[8] = -(T[0] * T[4] - T[1] * T[1]);
        int im = 0;
        if (fabs(Bt[4]) > fabs(Bt[0])) im = 1;
        if (fabs(Bt[8]) > fabs(Bt[4 * im])) im = 2;
        const double sc = Bt[4 * im] / (fro + 1e-300);
        if (sc > best_sc) {
            best_sc = sc; g = roots[r]; bi = im;
            for (int i = 0; i < 9; ++i) { D0[i] = T[i]; B[i] = Bt[i]; }
        }
    }
    if (bi < 0) return 0;
    const double sb = sqrt(B[4 * bi]);
    const double p[3] = {B[bi] / sb, B[3 + bi] / sb, B[6 + bi] / sb};
    const double C[9] = {D0[0], D0[1] - p[2], D0[2] + p[1], D0[3] + p[2], D0[4], D0[5] - p[0], D0[6] - p[1], D0[7] + p[0], D0[8]};
    int rm = 0, cm = 0;
    double cmax = -1;
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c)
            if (fabs(C[3 * r + c]) > cmax) { cmax = fabs(C[3 * r + c]); rm = r; cm = c; }
    const double lines[2][3] = {{C[3 * rm], C[3 * rm + 1], C[3 * rm + 2]}, {C[cm], C[3 + cm], C[6 + cm]}};
    const double* Q = fabs(g) < 1 ? D2 : D1;
    const double* Ms;
    double as;
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
        const int i0 = k == 0 ? 1 : 0, i1 = k == 2 ? 1 : 2;
        double u[3] = {0, 0, 0}, v[3] = {0, 0, 0};
        u[i0] = 1; u[k] = -l[i0] / l[k];
        v[i1] = 1; v[k] = -l[i1] / l[k];
        const double A = quad3(Q, v, v), Bq = quad3(Q, u, v), Cq = quad3(Q, u, u);
        const double disc = Bq * Bq - A * Cq;
        if (!(disc >= 0)) continue;
        const double sq = sqrt(disc);
        const double qq = -(Bq + (Bq >= 0 ? sq : -sq));
        double taus[2];
        int nt = 0;
        if (A != 0) taus[nt++] = qq / A;
        if (qq != 0) taus[nt++] = Cq / qq;
        for (int ti = 0; ti < nt && ns < 4; ++ti) {
            const double tau = taus[ti];
            if (!(tau > 0)) continue;
            const double w[3] = {u[0] + tau * v[0], u[1] + tau * v[1], u[2] + tau * v[2]};
            const double den = quad3(Ms, w, w);
            if (!(den > 0)) continue;
            const double sc = sqrt(as / den);
            double lam[3] = {sc * w[0], sc * w[1], sc * w[2]};
            if (!(lam[0] > 0 && lam[1] > 0 && lam[2] > 0)) continue;
            for (int it = 0; it < 3; ++it) {
                const double f[3] = {quad3(M12, lam, lam) - a12, quad3(M13, lam, lam) - a13, quad3(M23, lam, lam) - a23};
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
            double* Rn = R[ns];
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) Rn[3 * i + j] = Ym[3 * i] * Xi[j] + Ym[3 * i + 1] * Xi[3 + j] + Ym[3 * i + 2] * Xi[6 + j];
            bool fin = true;
            for (int c = 0; c < 3; ++c) {
                t[ns][c] = Y[0][c] - (Rn[3 * c] * X[0] + Rn[3 * c + 1] * X[1] + Rn[3 * c + 2] * X[2]);
                fin = fin && isfinite(t[ns][c]);
            }
            for (int c = 0; c < 9; ++c) fin = fin && isfinite(Rn[c]);
            if (fin) ++ns;
        }
    }
    return ns;
}

// RANSACUpdateNumIters
__device__ inline int ransac_update_num_iters(double p, double ep, int model_points, int max_iters)
{
    p = p < 0 ? 0 : p > 1 ? 1 : p;
    ep = ep < 0 ? 0 : ep > 1 ? 1 : ep;
    double num = 1 - p > DBL_MIN ? 1 - p : DBL_MIN;
    double denom = 1 - pow(1 - ep, (double)model_points);
    if (denom < DBL_MIN) return 0;
    num = log(num);
    denom = log(denom);
    return denom >= 0 || -num >= max_iters * (-denom) ? max_iters : (int)llrint(num / denom);
}

}  // namespace vo
