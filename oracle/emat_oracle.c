/*
 * oracle/emat_oracle.c -- TEST INFRASTRUCTURE ONLY (see vo_oracle.h).
 *
 * Restates cv2.findEssentialMat(p1, p2, K, method=RANSAC, prob, threshold) as called at reference
 * VisualOdometryPipeLine.py:308 (OpenCV modules/calib3d/src/five-point.cpp + ptsetreg.cpp; third
 * party, not vendored).  Spec: SURVEY.md A.6/A.7:
 *   - points to double, normalised ((u-cx)/fx, (v-cy)/fy); thr = threshold / ((fx+fy)/2);
 *   - per 5-subset the Nister five-point solver: null space {X,Y,Z,W} of the 5x9 epipolar matrix,
 *     E = xX + yY + zZ + W, ten cubic constraints (det E = 0, 2EE^tE - tr(EE^t)E = 0) as a 10x20
 *     matrix over the monomials [x3 y3 x2y xy2 x2z x2 y2z y2 xyz xy | xz2 xz x yz2 yz y z3 z2 z 1],
 *     Gauss-Jordan on the first ten columns, the 3x3 polynomial matrix B(z) from rows (x2z,x2),
 *     (y2z,y2), (xyz,xy), det B(z) = degree-10 polynomial, real roots (|Im| <= 1e-10), (x,y) from
 *     the null vector of B(z), E normalised to unit Frobenius norm;
 *   - Sampson error in double, stored as float32, inlier iff err <= (float)(thr*thr);
 *   - RANSAC loop of ransac_oracle.c with modelPoints = 5; result = winning minimal model (no refit).
 * This file derives the constraint coefficients by polynomial arithmetic (not OpenCV's expanded
 * expressions), takes any orthonormal null-space basis and finds roots by Aberth iteration, so
 * candidate ORDER inside a sample and the SIGN of E can differ from cv2 (cv2's own basis comes
 * from a seeded random completion); the candidate SET, the inlier mask and +-E are what is pinned
 * (tests/test_oracle_emat.py, tests/golden/emat.npz).
 */
#include "vo_oracle.h"
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* monomials of degree <= 3 in (x,y,z), Nister ordering */
static const int MON[20][3] = {
    {3,0,0},{0,3,0},{2,1,0},{1,2,0},{2,0,1},{2,0,0},{0,2,1},{0,2,0},{1,1,1},{1,1,0},
    {1,0,2},{1,0,1},{1,0,0},{0,1,2},{0,1,1},{0,1,0},{0,0,3},{0,0,2},{0,0,1},{0,0,0}};
static int mon_index(int a, int b, int c)
{
    for (int i = 0; i < 20; ++i) if (MON[i][0] == a && MON[i][1] == b && MON[i][2] == c) return i;
    return -1;
}
typedef struct { double c[20]; } poly_t;   /* indexed by MON */

static void pmul(const poly_t* a, const poly_t* b, poly_t* o)
{
    static int tab[20][20], init = 0;
    if (!init) {
        for (int i = 0; i < 20; ++i)
            for (int j = 0; j < 20; ++j) {
                int e0 = MON[i][0] + MON[j][0], e1 = MON[i][1] + MON[j][1], e2 = MON[i][2] + MON[j][2];
                tab[i][j] = (e0 + e1 + e2 <= 3) ? mon_index(e0, e1, e2) : -1;
            }
        init = 1;
    }
    memset(o, 0, sizeof(*o));
    for (int i = 0; i < 20; ++i) {
        if (a->c[i] == 0) continue;
        for (int j = 0; j < 20; ++j) {
            if (b->c[j] == 0 || tab[i][j] < 0) continue;
            o->c[tab[i][j]] += a->c[i] * b->c[j];
        }
    }
}
static void paxpy(poly_t* y, const poly_t* x, double s) { for (int i = 0; i < 20; ++i) y->c[i] += s * x->c[i]; }

/* orthonormal basis of the null space of the 5x9 matrix Q (rows), by Householder QR of Q^T */
static int null_space_5x9(const double Q[5][9], double N[4][9])
{
    double A[9][5];   /* Q^T */
    for (int i = 0; i < 5; ++i) for (int j = 0; j < 9; ++j) A[j][i] = Q[i][j];
    double V[5][9], beta[5];
    for (int k = 0; k < 5; ++k) {
        double nrm = 0;
        for (int i = k; i < 9; ++i) nrm += A[i][k] * A[i][k];
        nrm = sqrt(nrm);
        if (nrm < 1e-300) return 0;
        double alpha = A[k][k] > 0 ? -nrm : nrm;
        for (int i = 0; i < 9; ++i) V[k][i] = i < k ? 0 : A[i][k];
        V[k][k] -= alpha;
        double vn = 0;
        for (int i = k; i < 9; ++i) vn += V[k][i] * V[k][i];
        beta[k] = vn > 0 ? 2 / vn : 0;
        for (int j = k; j < 5; ++j) {
            double s = 0;
            for (int i = k; i < 9; ++i) s += V[k][i] * A[i][j];
            s *= beta[k];
            for (int i = k; i < 9; ++i) A[i][j] -= s * V[k][i];
        }
    }
    /* columns 5..8 of the orthogonal factor H0 H1 .. H4 */
    for (int c = 0; c < 4; ++c) {
        double e[9] = {0};
        e[5 + c] = 1;
        for (int k = 4; k >= 0; --k) {
            double s = 0;
            for (int i = k; i < 9; ++i) s += V[k][i] * e[i];
            s *= beta[k];
            for (int i = k; i < 9; ++i) e[i] -= s * V[k][i];
        }
        memcpy(N[c], e, sizeof(e));
    }
    return 1;
}

/* all complex roots of a real polynomial c[0] + c[1] z + .. + c[n] z^n (Aberth-Ehrlich) */
static int poly_roots(const double* c, int n, double* re, double* im)
{
    while (n > 0 && c[n] == 0) --n;
    if (n <= 0) return 0;
    double a[11];
    for (int i = 0; i <= n; ++i) a[i] = c[i] / c[n];
    double rad = 0;
    for (int i = 0; i < n; ++i) { double v = fabs(a[i]); if (v > rad) rad = v; }
    rad = 1 + rad;
    if (!isfinite(rad)) return 0;
    for (int k = 0; k < n; ++k) {
        double ang = 2 * M_PI * k / n + 0.4, r = rad * 0.5 * (1 + 0.1 * k / n);
        re[k] = r * cos(ang); im[k] = r * sin(ang);
    }
    /* Termination per root, as the CUDA solver and -- in effect -- cv2's solvePoly behave: a root is settled when its step
     * is below 1e-12 (relative), when |p(z)| has reached the rounding noise of its own evaluation (a root of a close
     * cluster cannot get closer: its imaginary part stays at ~1e-5 and the root is NOT taken as real, which is what
     * cv2's Durand-Kerner iteration does with such clusters; iterating on for the full 200 rounds lets the noise decide),
     * or when it is unmistakably complex and within 1e-6 of its limit. */
    for (int it = 0; it < 200; ++it) {
        int all_settled = 1;
        for (int k = 0; k < n; ++k) {
            /* p(z), p'(z) by Horner in complex arithmetic */
            double pr = 1, pi = 0, dr = 0, di = 0, zr = re[k], zi = im[k], sb = 1;
            const double az = sqrt(zr * zr + zi * zi);
            for (int i = n - 1; i >= 0; --i) {
                double ndr = dr * zr - di * zi + pr, ndi = dr * zi + di * zr + pi;
                double npr = pr * zr - pi * zi + a[i], npi = pr * zi + pi * zr;
                dr = ndr; di = ndi; pr = npr; pi = npi;
                sb = sb * az + fabs(a[i]);
            }
            double rel = 0;
            double den = dr * dr + di * di;
            if (den != 0) {
                double wr = (pr * dr + pi * di) / den, wi = (pi * dr - pr * di) / den;   /* p/p' */
                double sr = 0, si = 0;
                for (int j = 0; j < n; ++j) {
                    if (j == k) continue;
                    double er = zr - re[j], ei = zi - im[j], d2 = er * er + ei * ei;
                    if (d2 == 0) continue;
                    sr += er / d2; si -= ei / d2;
                }
                double qr = 1 - (wr * sr - wi * si), qi = -(wr * si + wi * sr);
                double qd = qr * qr + qi * qi;
                if (qd != 0) {
                    double stepr = (wr * qr + wi * qi) / qd, stepi = (wi * qr - wr * qi) / qd;
                    re[k] -= stepr; im[k] -= stepi;
                    rel = (fabs(stepr) + fabs(stepi)) / (fabs(re[k]) + fabs(im[k]) + 1e-300);
                }
            }
            const double mag = fabs(re[k]) + fabs(im[k]);
            const int noise = fabs(pr) + fabs(pi) <= 2e-14 * sb;
            const int settled = rel < 1e-12 || noise || (it >= 8 && rel < 1e-6 && fabs(im[k]) > 1e-4 * mag && fabs(im[k]) > 1e-7);
            if (!settled) all_settled = 0;
        }
        if (all_settled) break;   /* real roots are Newton-polished afterwards */
    }
    return n;
}

static void poly1_mul(const double* a, int na, const double* b, int nb, double* o)   /* ascending powers */
{
    for (int i = 0; i <= na + nb; ++i) o[i] = 0;
    for (int i = 0; i <= na; ++i) for (int j = 0; j <= nb; ++j) o[i + j] += a[i] * b[j];
}

/* x1, x2: 5 normalised points each (x,y).  E: up to 10 row-major 3x3 models (unit Frobenius norm). */
int orc_five_point(const double x1[10], const double x2[10], double E[10][9])
{
    double Q[5][9], N[4][9];
    for (int i = 0; i < 5; ++i) {
        const double a = x1[2 * i], b = x1[2 * i + 1], c = x2[2 * i], d = x2[2 * i + 1];
        const double r[9] = {c * a, c * b, c, d * a, d * b, d, a, b, 1.0};
        memcpy(Q[i], r, sizeof(r));
    }
    if (!null_space_5x9(Q, N)) return 0;
    /* E(i,j) as linear polynomials */
    poly_t Ep[3][3];
    const int ix = mon_index(1, 0, 0), iy = mon_index(0, 1, 0), iz = mon_index(0, 0, 1), i1 = mon_index(0, 0, 0);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            memset(&Ep[i][j], 0, sizeof(poly_t));
            Ep[i][j].c[ix] = N[0][3 * i + j]; Ep[i][j].c[iy] = N[1][3 * i + j];
            Ep[i][j].c[iz] = N[2][3 * i + j]; Ep[i][j].c[i1] = N[3][3 * i + j];
        }
    poly_t eq[10], t1, t2;
    /* det E */
    memset(&eq[0], 0, sizeof(poly_t));
    static const int perm[6][4] = {{0,1,2,1},{1,2,0,1},{2,0,1,1},{2,1,0,-1},{1,0,2,-1},{0,2,1,-1}};
    for (int p = 0; p < 6; ++p) {
        pmul(&Ep[0][perm[p][0]], &Ep[1][perm[p][1]], &t1);
        pmul(&t1, &Ep[2][perm[p][2]], &t2);
        paxpy(&eq[0], &t2, perm[p][3]);
    }
    /* 2 E E^t E - tr(E E^t) E */
    poly_t EEt[3][3], tr;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            memset(&EEt[i][j], 0, sizeof(poly_t));
            for (int k = 0; k < 3; ++k) { pmul(&Ep[i][k], &Ep[j][k], &t1); paxpy(&EEt[i][j], &t1, 1.0); }
        }
    memset(&tr, 0, sizeof(tr));
    for (int i = 0; i < 3; ++i) paxpy(&tr, &EEt[i][i], 1.0);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            poly_t* e = &eq[1 + 3 * i + j];
            memset(e, 0, sizeof(poly_t));
            for (int k = 0; k < 3; ++k) { pmul(&EEt[i][k], &Ep[k][j], &t1); paxpy(e, &t1, 2.0); }
            pmul(&tr, &Ep[i][j], &t1);
            paxpy(e, &t1, -1.0);
        }
    /* Gauss-Jordan with partial pivoting on the first 10 columns */
    double A[10][20];
    for (int r = 0; r < 10; ++r) memcpy(A[r], eq[r].c, sizeof(double) * 20);
    for (int c = 0; c < 10; ++c) {
        int piv = c;
        for (int r = c + 1; r < 10; ++r) if (fabs(A[r][c]) > fabs(A[piv][c])) piv = r;
        if (fabs(A[piv][c]) < 1e-300) return 0;
        if (piv != c) for (int k = 0; k < 20; ++k) { double t = A[c][k]; A[c][k] = A[piv][k]; A[piv][k] = t; }
        const double inv = 1.0 / A[c][c];
        for (int k = 0; k < 20; ++k) A[c][k] *= inv;
        for (int r = 0; r < 10; ++r) {
            if (r == c) continue;
            const double f = A[r][c];
            if (f == 0) continue;
            for (int k = 0; k < 20; ++k) A[r][k] -= f * A[c][k];
        }
    }
    /* B(z): rows from (x2z, x2), (y2z, y2), (xyz, xy); entries as ascending-power polynomials in z */
    double Bx[3][4], By[3][4], Bc[3][5];
    for (int i = 0; i < 3; ++i) {
        const double* r1 = A[2 * i + 4] + 10;   /* [xz2 xz x yz2 yz y z3 z2 z 1] */
        const double* r2 = A[2 * i + 5] + 10;
        /* row1 - z*row2, coefficient of x: descending (z3,z2,z,1) = (-r2[0], r1[0]-r2[1], r1[1]-r2[2], r1[2]) */
        Bx[i][3] = -r2[0]; Bx[i][2] = r1[0] - r2[1]; Bx[i][1] = r1[1] - r2[2]; Bx[i][0] = r1[2];
        By[i][3] = -r2[3]; By[i][2] = r1[3] - r2[4]; By[i][1] = r1[4] - r2[5]; By[i][0] = r1[5];
        Bc[i][4] = -r2[6]; Bc[i][3] = r1[6] - r2[7]; Bc[i][2] = r1[7] - r2[8]; Bc[i][1] = r1[8] - r2[9]; Bc[i][0] = r1[9];
    }
    /* det B(z) = Bx0 (By1 Bc2 - Bc1 By2) - By0 (Bx1 Bc2 - Bc1 Bx2) + Bc0 (Bx1 By2 - By1 Bx2) */
    double det[11] = {0}, m1[8], m2[8], m3[11];
    poly1_mul(By[1], 3, Bc[2], 4, m1); poly1_mul(Bc[1], 4, By[2], 3, m2);
    for (int k = 0; k < 8; ++k) m1[k] -= m2[k];
    poly1_mul(Bx[0], 3, m1, 7, m3);
    for (int k = 0; k < 11; ++k) det[k] += m3[k];
    poly1_mul(Bx[1], 3, Bc[2], 4, m1); poly1_mul(Bc[1], 4, Bx[2], 3, m2);
    for (int k = 0; k < 8; ++k) m1[k] -= m2[k];
    poly1_mul(By[0], 3, m1, 7, m3);
    for (int k = 0; k < 11; ++k) det[k] -= m3[k];
    double m4[7], m5[7];
    poly1_mul(Bx[1], 3, By[2], 3, m4); poly1_mul(By[1], 3, Bx[2], 3, m5);
    for (int k = 0; k < 7; ++k) m4[k] -= m5[k];
    poly1_mul(Bc[0], 4, m4, 6, m3);
    for (int k = 0; k < 11; ++k) det[k] += m3[k];
    double rr[10], ri[10];
    const int nroots = poly_roots(det, 10, rr, ri);
    /* real roots in ascending order (deterministic candidate order) */
    double zs[10];
    int nz = 0;
    for (int k = 0; k < nroots; ++k) {
        if (!(fabs(ri[k]) <= 1e-10)) continue;
        double z = rr[k];
        for (int it = 0; it < 2; ++it) {   /* Newton polish on the real polynomial */
            double p = det[10], dp = 0;
            for (int i = 9; i >= 0; --i) { dp = dp * z + p; p = p * z + det[i]; }
            if (dp != 0 && isfinite(p / dp)) z -= p / dp;
        }
        zs[nz++] = z;
    }
    for (int i = 1; i < nz; ++i) { double v = zs[i]; int j = i - 1; while (j >= 0 && zs[j] > v) { zs[j + 1] = zs[j]; --j; } zs[j + 1] = v; }
    int count = 0;
    for (int k = 0; k < nz; ++k) {
        const double z = zs[k];
        double Bz[3][3];
        for (int i = 0; i < 3; ++i) {
            Bz[i][0] = ((Bx[i][3] * z + Bx[i][2]) * z + Bx[i][1]) * z + Bx[i][0];
            Bz[i][1] = ((By[i][3] * z + By[i][2]) * z + By[i][1]) * z + By[i][0];
            Bz[i][2] = (((Bc[i][4] * z + Bc[i][3]) * z + Bc[i][2]) * z + Bc[i][1]) * z + Bc[i][0];
        }
        /* null vector of the (rank-2) matrix: the largest cross product of two rows */
        double best[3] = {0, 0, 0}, bn = -1;
        for (int a = 0; a < 3; ++a)
            for (int b = a + 1; b < 3; ++b) {
                double c[3] = {Bz[a][1] * Bz[b][2] - Bz[a][2] * Bz[b][1], Bz[a][2] * Bz[b][0] - Bz[a][0] * Bz[b][2],
                               Bz[a][0] * Bz[b][1] - Bz[a][1] * Bz[b][0]};
                double n2 = c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
                if (n2 > bn) { bn = n2; memcpy(best, c, sizeof(c)); }
            }
        if (!(bn > 0)) continue;
        const double nv = sqrt(bn);
        if (fabs(best[2] / nv) < 1e-10) continue;
        const double x = best[0] / best[2], y = best[1] / best[2];
        double Ev[9], nrm = 0;
        for (int i = 0; i < 9; ++i) { Ev[i] = x * N[0][i] + y * N[1][i] + z * N[2][i] + N[3][i]; nrm += Ev[i] * Ev[i]; }
        nrm = sqrt(nrm);
        if (!(nrm > 0) || !isfinite(nrm)) continue;
        for (int i = 0; i < 9; ++i) E[count][i] = Ev[i] / nrm;
        ++count;
    }
    return count;
}

void orc_sampson_errors(const double* x1n, const double* x2n, int n, const double E[9], float* err)
{
    for (int i = 0; i < n; ++i) {
        const double ax = x1n[2 * i], ay = x1n[2 * i + 1], bx = x2n[2 * i], by = x2n[2 * i + 1];
        const double Ex0 = E[0] * ax + E[1] * ay + E[2] * 1., Ex1 = E[3] * ax + E[4] * ay + E[5] * 1., Ex2 = E[6] * ax + E[7] * ay + E[8] * 1.;
        const double Et0 = E[0] * bx + E[3] * by + E[6] * 1., Et1 = E[1] * bx + E[4] * by + E[7] * 1.;
        const double s = bx * Ex0 + by * Ex1 + 1. * Ex2;
        const double a = Ex0 * Ex0, b = Ex1 * Ex1, c = Et0 * Et0, d = Et1 * Et1;
        err[i] = (float)(s * s / (a + b + c + d));
    }
}

int orc_find_essential_mat_ransac(const float* p1, const float* p2, int n, const double K[9],
                                  double prob, double thr, int max_iters, double E[9],
                                  uint8_t* mask, int* found, int* iters_run)
{
    *found = 0;
    if (iters_run) *iters_run = 0;
    if (n < 5) return 0;
    const double fx = K[0], fy = K[4], cx = K[2], cy = K[5];
    double* x1 = (double*)malloc(sizeof(double) * 2 * (size_t)n);
    double* x2 = (double*)malloc(sizeof(double) * 2 * (size_t)n);
    for (int i = 0; i < n; ++i) {
        x1[2 * i] = ((double)p1[2 * i] - cx) / fx; x1[2 * i + 1] = ((double)p1[2 * i + 1] - cy) / fy;
        x2[2 * i] = ((double)p2[2 * i] - cx) / fx; x2[2 * i + 1] = ((double)p2[2 * i + 1] - cy) / fy;
    }
    const double t = thr / ((fx + fy) / 2);
    const float t2 = (float)(t * t);
    float* err = (float*)malloc(sizeof(float) * (size_t)n);
    int niters = max_iters > 1 ? max_iters : 1, max_good = 0, it = 0;
    int32_t* subsets = (int32_t*)malloc(sizeof(int32_t) * 5 * (size_t)niters);
    if (n == 5) { for (int k = 0; k < 5; ++k) subsets[k] = k; niters = 1; }
    else orc_ransac_subsets(n, 5, niters, subsets);
    for (it = 0; it < niters; ++it) {
        double s1[10], s2[10], models[10][9];
        for (int k = 0; k < 5; ++k) {
            const int s = subsets[5 * it + k];
            s1[2 * k] = x1[2 * s]; s1[2 * k + 1] = x1[2 * s + 1];
            s2[2 * k] = x2[2 * s]; s2[2 * k + 1] = x2[2 * s + 1];
        }
        const int nm = orc_five_point(s1, s2, models);
        for (int m = 0; m < nm; ++m) {
            orc_sampson_errors(x1, x2, n, models[m], err);
            int good = 0;
            for (int i = 0; i < n; ++i) good += err[i] <= t2;
            if (good > (max_good > 4 ? max_good : 4)) {
                for (int i = 0; i < n; ++i) mask[i] = err[i] <= t2;
                memcpy(E, models[m], sizeof(double) * 9);
                max_good = good;
                niters = orc_ransac_update_num_iters(prob, (double)(n - good) / n, 5, niters);
            }
        }
    }
    if (iters_run) *iters_run = it;
    *found = max_good > 0;
    if (n == 5 && *found) for (int i = 0; i < n; ++i) mask[i] = 1;
    free(x1); free(x2); free(err); free(subsets);
    return 0;
}
