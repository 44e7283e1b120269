// Five-point essential-matrix RANSAC (replaces cv2.findEssentialMat(p1, p2, K, method=RANSAC, prob,
// threshold) at reference VisualOdometryPipeLine.py:308; spec SURVEY.md A.6/A.7 and the header of
// oracle/emat_oracle.c).
//
// All maxIters 5-subsets are drawn (bit-exact cv::RNG stream), solved and scored at once:
//   emat_normalize_kernel   pixels -> normalised double coordinates
//   ransac_samples_kernel<5> (ransac.cuh)
//   emat_solve_kernel       one thread per sample: Nister five-point in FP64 (Householder null
//                           space, cubic constraints by polynomial arithmetic, Gauss-Jordan,
//                           det B(z), Aberth roots, back-substitution) -> up to 10 models
//   emat_score_kernel       (256 points) x (8 samples x <=10 models): Sampson error in double ->
//                           float32, ballot+popc counts, one atomicAdd per warp
//   emat_update_kernel      sequential replay of cv2's loop over (sample, model) counts with the
//                           adaptive iteration bound; hypotheses go through in chunks of 128 and a
//                           chunk beyond the bound exits at once (cv2 stops after 10-60 samples)
//   emat_finish_kernel      the winner's E and mask
#include "internal.cuh"
#include "mathdev.cuh"
#include "ransac.cuh"

using namespace vo;

#define EM_MAXM 10

__constant__ signed char EM_TAB[20][20] = {
    {-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1, 0},
    {-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1, 1},
    {-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1, 2},
    {-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1, 3},
    {-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1, 4},
    {-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1, 0,-1,-1, 2,-1,-1, 4, 5},
    {-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1, 6},
    {-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1, 3,-1,-1, 1,-1,-1, 6, 7},
    {-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1, 8},
    {-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1, 2,-1,-1, 3,-1,-1, 8, 9},
    {-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,10},
    {-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1, 4,-1,-1, 8,-1,-1,10,11},
    {-1,-1,-1,-1,-1, 0,-1, 3,-1, 2,-1, 4, 5,-1, 8, 9,-1,10,11,12},
    {-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,13},
    {-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1, 8,-1,-1, 6,-1,-1,13,14},
    {-1,-1,-1,-1,-1, 2,-1, 1,-1, 3,-1, 8, 9,-1, 6, 7,-1,13,14,15},
    {-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,16},
    {-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,-1,10,-1,-1,13,-1,-1,16,17},
    {-1,-1,-1,-1,-1, 4,-1, 6,-1, 8,-1,10,11,-1,13,14,-1,16,17,18},
    { 0, 1, 2, 3, 4, 5, 6, 7, 8, 9,10,11,12,13,14,15,16,17,18,19}};
// monomial indices of x, y, z, 1 in the ordering [x3 y3 x2y xy2 x2z x2 y2z y2 xyz xy | xz2 xz x yz2 yz y z3 z2 z 1]
#define EM_IX 12
#define EM_IY 15
#define EM_IZ 18
#define EM_I1 19

struct Poly { double c[20]; };

// Products only ever pair (degree <= 1) x (degree <= 1) and (degree <= 2) x (degree <= 1): iterate
// over the 4 linear and 10 quadratic-or-lower monomials instead of all 20 x 20 pairs.
__constant__ signed char EM_LIN[4] = {12, 15, 18, 19};
__constant__ signed char EM_QUAD[10] = {5, 7, 9, 11, 12, 14, 15, 17, 18, 19};
template <int NA>
__device__ inline void pmul_n(const Poly& a, const Poly& b, Poly& o)
{
    for (int i = 0; i < 20; ++i) o.c[i] = 0;
    for (int ii = 0; ii < NA; ++ii) {
        const int i = NA == 4 ? EM_LIN[ii] : EM_QUAD[ii];
        const double ai = a.c[i];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const int j = EM_LIN[jj];
            o.c[EM_TAB[i][j]] += ai * b.c[j];
        }
    }
}
#define pmul_ll(a, b, o) pmul_n<4>(a, b, o)    /* linear x linear */
#define pmul_ql(a, b, o) pmul_n<10>(a, b, o)   /* quadratic x linear */
__device__ inline void paxpy(Poly& y, const Poly& x, double s)
{
    for (int i = 0; i < 20; ++i) y.c[i] += s * x.c[i];
}

__device__ inline void poly1_mul(const double* a, int na, const double* b, int nb, double* o)
{
    for (int i = 0; i <= na + nb; ++i) o[i] = 0;
    for (int i = 0; i <= na; ++i)
        for (int j = 0; j <= nb; ++j) o[i + j] += a[i] * b[j];
}

// all complex roots of c[0] + .. + c[n] z^n by Aberth-Ehrlich iteration
__device__ inline int poly_roots(const double* c, int n, double* re, double* im)
{
    while (n > 0 && c[n] == 0) --n;
    if (n <= 0) return 0;
    double a[11];
    for (int i = 0; i <= n; ++i) a[i] = c[i] / c[n];
    double rad = 0;   // Fujiwara-style root bound: max |a_i|^(1/(n-i))
    for (int i = 0; i < n; ++i) if (a[i] != 0) rad = fmax(rad, pow(fabs(a[i]), 1.0 / (n - i)));
    if (!isfinite(rad)) return 0;
    rad = rad > 0 ? 0.7 * rad : 1.0;
    for (int k = 0; k < n; ++k) {
        const double ang = 6.283185307179586476925286766559 * k / n + 0.4, r = rad * (1 + 0.1 * k / n);
        re[k] = r * cos(ang); im[k] = r * sin(ang);
    }
    for (int it = 0; it < 200; ++it) {
        double maxstep = 0;
        for (int k = 0; k < n; ++k) {
            double pr = 1, pi = 0, dr = 0, di = 0;
            const double zr = re[k], zi = im[k];
            for (int i = n - 1; i >= 0; --i) {
                const double ndr = dr * zr - di * zi + pr, ndi = dr * zi + di * zr + pi;
                const double npr = pr * zr - pi * zi + a[i], npi = pr * zi + pi * zr;
                dr = ndr; di = ndi; pr = npr; pi = npi;
            }
            const double den = dr * dr + di * di;
            if (den == 0) continue;
            const double wr = (pr * dr + pi * di) / den, wi = (pi * dr - pr * di) / den;
            double sr = 0, si = 0;
            for (int j = 0; j < n; ++j) {
                if (j == k) continue;
                const double er = zr - re[j], ei = zi - im[j], d2 = er * er + ei * ei;
                if (d2 == 0) continue;
                sr += er / d2; si -= ei / d2;
            }
            const double qr = 1 - (wr * sr - wi * si), qi = -(wr * si + wi * sr);
            const double qd = qr * qr + qi * qi;
            if (qd == 0) continue;
            const double stepr = (wr * qr + wi * qi) / qd, stepi = (wi * qr - wr * qi) / qd;
            re[k] -= stepr; im[k] -= stepi;
            const double st = fabs(stepr) + fabs(stepi), sc = fabs(re[k]) + fabs(im[k]) + 1e-300;
            maxstep = fmax(maxstep, st / sc);
        }
        if (maxstep < 1e-12) break;   // real roots are Newton-polished afterwards
    }
    return n;
}

// x1, x2: 5 normalised correspondences; E: up to 10 row-major models with unit Frobenius norm
__device__ int five_point(const double* x1, const double* x2, double (*E)[9])
{
    // ---- null space of the 5x9 epipolar matrix by Householder QR of its transpose ----
    double A9[9][5], V[5][9], beta[5], N[4][9];
    for (int i = 0; i < 5; ++i) {
        const double a = x1[2 * i], b = x1[2 * i + 1], c = x2[2 * i], d = x2[2 * i + 1];
        const double r[9] = {c * a, c * b, c, d * a, d * b, d, a, b, 1.0};
        for (int j = 0; j < 9; ++j) A9[j][i] = r[j];
    }
    for (int k = 0; k < 5; ++k) {
        double nrm = 0;
        for (int i = k; i < 9; ++i) nrm += A9[i][k] * A9[i][k];
        nrm = sqrt(nrm);
        if (nrm < 1e-300) return 0;
        const double alpha = A9[k][k] > 0 ? -nrm : nrm;
        for (int i = 0; i < 9; ++i) V[k][i] = i < k ? 0 : A9[i][k];
        V[k][k] -= alpha;
        double vn = 0;
        for (int i = k; i < 9; ++i) vn += V[k][i] * V[k][i];
        beta[k] = vn > 0 ? 2 / vn : 0;
        for (int j = k; j < 5; ++j) {
            double s = 0;
            for (int i = k; i < 9; ++i) s += V[k][i] * A9[i][j];
            s *= beta[k];
            for (int i = k; i < 9; ++i) A9[i][j] -= s * V[k][i];
        }
    }
    for (int c = 0; c < 4; ++c) {
        double e[9];
        for (int i = 0; i < 9; ++i) e[i] = (i == 5 + c) ? 1.0 : 0.0;
        for (int k = 4; k >= 0; --k) {
            double s = 0;
            for (int i = k; i < 9; ++i) s += V[k][i] * e[i];
            s *= beta[k];
            for (int i = k; i < 9; ++i) e[i] -= s * V[k][i];
        }
        for (int i = 0; i < 9; ++i) N[c][i] = e[i];
    }
    // ---- ten cubic constraints ----
    Poly Ep[3][3], t1, t2;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            for (int k = 0; k < 20; ++k) Ep[i][j].c[k] = 0;
            Ep[i][j].c[EM_IX] = N[0][3 * i + j]; Ep[i][j].c[EM_IY] = N[1][3 * i + j];
            Ep[i][j].c[EM_IZ] = N[2][3 * i + j]; Ep[i][j].c[EM_I1] = N[3][3 * i + j];
        }
    double A[10][20];
    {
        Poly acc;
        for (int k = 0; k < 20; ++k) acc.c[k] = 0;
        const int perm[6][4] = {{0, 1, 2, 1}, {1, 2, 0, 1}, {2, 0, 1, 1}, {2, 1, 0, -1}, {1, 0, 2, -1}, {0, 2, 1, -1}};
        for (int p = 0; p < 6; ++p) {
            pmul_ll(Ep[0][perm[p][0]], Ep[1][perm[p][1]], t1);
            pmul_ql(t1, Ep[2][perm[p][2]], t2);
            paxpy(acc, t2, (double)perm[p][3]);
        }
        for (int k = 0; k < 20; ++k) A[0][k] = acc.c[k];
    }
    {
        Poly EEt[3][3], tr;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                for (int k = 0; k < 20; ++k) EEt[i][j].c[k] = 0;
                for (int k = 0; k < 3; ++k) { pmul_ll(Ep[i][k], Ep[j][k], t1); paxpy(EEt[i][j], t1, 1.0); }
            }
        for (int k = 0; k < 20; ++k) tr.c[k] = EEt[0][0].c[k] + EEt[1][1].c[k] + EEt[2][2].c[k];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                Poly acc;
                for (int k = 0; k < 20; ++k) acc.c[k] = 0;
                for (int k = 0; k < 3; ++k) { pmul_ql(EEt[i][k], Ep[k][j], t1); paxpy(acc, t1, 2.0); }
                pmul_ql(tr, Ep[i][j], t1);
                paxpy(acc, t1, -1.0);
                for (int k = 0; k < 20; ++k) A[1 + 3 * i + j][k] = acc.c[k];
            }
    }
    // ---- Gauss-Jordan (partial pivoting) on the first ten columns ----
    for (int c = 0; c < 10; ++c) {
        int piv = c;
        for (int r = c + 1; r < 10; ++r) if (fabs(A[r][c]) > fabs(A[piv][c])) piv = r;
        if (fabs(A[piv][c]) < 1e-300) return 0;
        if (piv != c) for (int k = 0; k < 20; ++k) { const double t = A[c][k]; A[c][k] = A[piv][k]; A[piv][k] = t; }
        const double inv = 1.0 / A[c][c];
        for (int k = 0; k < 20; ++k) A[c][k] *= inv;
        for (int r = 0; r < 10; ++r) {
            if (r == c) continue;
            const double f = A[r][c];
            if (f == 0) continue;
            for (int k = 0; k < 20; ++k) A[r][k] -= f * A[c][k];
        }
    }
    // ---- B(z) and its determinant ----
    double Bx[3][4], By[3][4], Bc[3][5];
    for (int i = 0; i < 3; ++i) {
        const double* r1 = A[2 * i + 4] + 10;
        const double* r2 = A[2 * i + 5] + 10;
        Bx[i][3] = -r2[0]; Bx[i][2] = r1[0] - r2[1]; Bx[i][1] = r1[1] - r2[2]; Bx[i][0] = r1[2];
        By[i][3] = -r2[3]; By[i][2] = r1[3] - r2[4]; By[i][1] = r1[4] - r2[5]; By[i][0] = r1[5];
        Bc[i][4] = -r2[6]; Bc[i][3] = r1[6] - r2[7]; Bc[i][2] = r1[7] - r2[8]; Bc[i][1] = r1[8] - r2[9]; Bc[i][0] = r1[9];
    }
    double det[11], m1[8], m2[8], m3[11], m4[7], m5[7];
    for (int k = 0; k < 11; ++k) det[k] = 0;
    poly1_mul(By[1], 3, Bc[2], 4, m1); poly1_mul(Bc[1], 4, By[2], 3, m2);
    for (int k = 0; k < 8; ++k) m1[k] -= m2[k];
    poly1_mul(Bx[0], 3, m1, 7, m3);
    for (int k = 0; k < 11; ++k) det[k] += m3[k];
    poly1_mul(Bx[1], 3, Bc[2], 4, m1); poly1_mul(Bc[1], 4, Bx[2], 3, m2);
    for (int k = 0; k < 8; ++k) m1[k] -= m2[k];
    poly1_mul(By[0], 3, m1, 7, m3);
    for (int k = 0; k < 11; ++k) det[k] -= m3[k];
    poly1_mul(Bx[1], 3, By[2], 3, m4); poly1_mul(By[1], 3, Bx[2], 3, m5);
    for (int k = 0; k < 7; ++k) m4[k] -= m5[k];
    poly1_mul(Bc[0], 4, m4, 6, m3);
    for (int k = 0; k < 11; ++k) det[k] += m3[k];
    double rr[10], ri[10], zs[10];
    const int nroots = poly_roots(det, 10, rr, ri);
    int nz = 0;
    for (int k = 0; k < nroots; ++k) {
        if (!(fabs(ri[k]) <= 1e-10)) continue;
        double z = rr[k];
        for (int it = 0; it < 2; ++it) {
            double p = det[10], dp = 0;
            for (int i = 9; i >= 0; --i) { dp = dp * z + p; p = p * z + det[i]; }
            if (dp != 0 && isfinite(p / dp)) z -= p / dp;
        }
        zs[nz++] = z;
    }
    for (int i = 1; i < nz; ++i) {   // ascending z: deterministic candidate order
        const double v = zs[i];
        int j = i - 1;
        while (j >= 0 && zs[j] > v) { zs[j + 1] = zs[j]; --j; }
        zs[j + 1] = v;
    }
    int count = 0;
    for (int k = 0; k < nz; ++k) {
        const double z = zs[k];
        double Bz[3][3];
        for (int i = 0; i < 3; ++i) {
            Bz[i][0] = ((Bx[i][3] * z + Bx[i][2]) * z + Bx[i][1]) * z + Bx[i][0];
            Bz[i][1] = ((By[i][3] * z + By[i][2]) * z + By[i][1]) * z + By[i][0];
            Bz[i][2] = (((Bc[i][4] * z + Bc[i][3]) * z + Bc[i][2]) * z + Bc[i][1]) * z + Bc[i][0];
        }
        double best[3] = {0, 0, 0}, bn = -1;
        for (int a = 0; a < 3; ++a)
            for (int b = a + 1; b < 3; ++b) {
                const double c[3] = {Bz[a][1] * Bz[b][2] - Bz[a][2] * Bz[b][1], Bz[a][2] * Bz[b][0] - Bz[a][0] * Bz[b][2],
                                     Bz[a][0] * Bz[b][1] - Bz[a][1] * Bz[b][0]};
                const double n2 = c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
                if (n2 > bn) { bn = n2; best[0] = c[0]; best[1] = c[1]; best[2] = c[2]; }
            }
        if (!(bn > 0)) continue;
        if (fabs(best[2] / sqrt(bn)) < 1e-10) continue;
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

struct EmatArgs {
    int n, iters;
    const float* p1; const float* p2;     // [n][2] pixels
    double fx, fy, cx, cy;
    float thr_sq;
    double conf;
    double* x1; double* x2;               // [n][2] normalised
    int* n_dev;                            // [1] = n
    int* samples;                          // [iters][5]
    double* models;                        // [iters][10][9]
    int* nmodels;                          // [iters]
    int* counts;                           // [iters][10]
    int* flags;
    // outputs
    double* E; uint8_t* mask; int* result; // result[0] = found, [1] = iterations run, [2] = winner flat index
    int* state;                            // [0] current iteration bound (niters), [1] max_good, [2] winner, [3] iterations replayed
    int chunk_start, chunk_len;
};

__global__ void __launch_bounds__(256)
emat_normalize_kernel(EmatArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    a.x1[2 * i] = ((double)a.p1[2 * i] - a.cx) / a.fx; a.x1[2 * i + 1] = ((double)a.p1[2 * i + 1] - a.cy) / a.fy;
    a.x2[2 * i] = ((double)a.p2[2 * i] - a.cx) / a.fx; a.x2[2 * i + 1] = ((double)a.p2[2 * i + 1] - a.cy) / a.fy;
}

__global__ void __launch_bounds__(32)
emat_solve_kernel(EmatArgs a)
{
    if (a.chunk_start >= a.state[0]) return;   // cv2's adaptive bound was reached in an earlier chunk
    const int it = a.chunk_start + blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= a.iters || it >= a.chunk_start + a.chunk_len) return;
    for (int m = 0; m < EM_MAXM; ++m) a.counts[it * EM_MAXM + m] = 0;
    a.nmodels[it] = 0;
    const int* smp = a.samples + 5 * it;
    if (smp[0] < 0) return;
    double s1[10], s2[10];
    for (int k = 0; k < 5; ++k) {
        const int s = smp[k];
        s1[2 * k] = a.x1[2 * s]; s1[2 * k + 1] = a.x1[2 * s + 1];
        s2[2 * k] = a.x2[2 * s]; s2[2 * k + 1] = a.x2[2 * s + 1];
    }
    double E[EM_MAXM][9];
    const int nm = five_point(s1, s2, E);
    for (int m = 0; m < nm; ++m)
        for (int k = 0; k < 9; ++k) a.models[((size_t)it * EM_MAXM + m) * 9 + k] = E[m][k];
    a.nmodels[it] = nm;
}

__device__ __forceinline__ bool sampson_inlier(const double* E, double ax, double ay, double bx, double by, float thr_sq)
{
    const double Ex0 = E[0] * ax + E[1] * ay + E[2] * 1., Ex1 = E[3] * ax + E[4] * ay + E[5] * 1., Ex2 = E[6] * ax + E[7] * ay + E[8] * 1.;
    const double Et0 = E[0] * bx + E[3] * by + E[6] * 1., Et1 = E[1] * bx + E[4] * by + E[7] * 1.;
    const double s = bx * Ex0 + by * Ex1 + 1. * Ex2;
    const double aa = Ex0 * Ex0, bb = Ex1 * Ex1, cc = Et0 * Et0, dd = Et1 * Et1;
    const float e = (float)(s * s / (aa + bb + cc + dd));
    return e <= thr_sq;
}

#define EM_ST 8   // samples per scoring block
__global__ void __launch_bounds__(256)
emat_score_kernel(EmatArgs a)
{
    __shared__ double s_E[EM_ST * EM_MAXM * 9];
    __shared__ int s_nm[EM_ST];
    if (a.chunk_start >= a.state[0]) return;
    const int it0 = a.chunk_start + blockIdx.y * EM_ST;
    const int ns = min(EM_ST, min(a.iters, a.chunk_start + a.chunk_len) - it0);
    if (ns <= 0) return;
    for (int k = threadIdx.x; k < ns * EM_MAXM * 9; k += blockDim.x) s_E[k] = a.models[(size_t)it0 * EM_MAXM * 9 + k];
    if (threadIdx.x < ns) s_nm[threadIdx.x] = a.nmodels[it0 + threadIdx.x];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < a.n;
    double ax = 0, ay = 0, bx = 0, by = 0;
    if (live) { ax = a.x1[2 * i]; ay = a.x1[2 * i + 1]; bx = a.x2[2 * i]; by = a.x2[2 * i + 1]; }
    const int lane = threadIdx.x & 31;
    for (int s = 0; s < ns; ++s)
        for (int m = 0; m < s_nm[s]; ++m) {
            const bool in = live && sampson_inlier(s_E + (s * EM_MAXM + m) * 9, ax, ay, bx, by, a.thr_sq);
            const unsigned bm = __ballot_sync(0xffffffffu, in);
            if (lane == 0 && bm) atomicAdd(a.counts + (it0 + s) * EM_MAXM + m, __popc(bm));
        }
}

// replay of cv2's sequential loop over this chunk's (sample, model) counts; shrinks the bound
__global__ void emat_update_kernel(EmatArgs a)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int niters = a.state[0], max_good = a.state[1], win = a.state[2];
    if (a.chunk_start >= niters) return;
    const int N = a.n;
    int it = a.chunk_start;
    if (N == 5) {
        if (it == 0) { win = a.nmodels[0] > 0 ? 0 : -1; a.state[2] = win; a.state[3] = 1; a.state[0] = 1; }
        return;
    }
    const int end = min(a.chunk_start + a.chunk_len, a.iters);
    for (; it < end && it < niters; ++it) {
        const int nm = a.nmodels[it];
        for (int m = 0; m < nm; ++m) {
            const int good = a.counts[it * EM_MAXM + m];
            if (good > (max_good > 4 ? max_good : 4)) {
                win = it * EM_MAXM + m; max_good = good;
                niters = ransac_update_num_iters(a.conf, (double)(N - good) / N, 5, niters);
            }
        }
    }
    a.state[0] = niters; a.state[1] = max_good; a.state[2] = win; a.state[3] = it;
}

__global__ void __launch_bounds__(256)
emat_finish_kernel(EmatArgs a)
{
    __shared__ double s_E[9];
    const int win = a.state[2];
    if (threadIdx.x == 0) { a.result[0] = win >= 0; a.result[1] = a.state[3]; a.result[2] = win; }
    if (win < 0) {
        for (int i = threadIdx.x; i < a.n; i += blockDim.x) a.mask[i] = 0;
        return;
    }
    if (threadIdx.x < 9) { s_E[threadIdx.x] = a.models[(size_t)win * 9 + threadIdx.x]; a.E[threadIdx.x] = s_E[threadIdx.x]; }
    __syncthreads();
    for (int i = threadIdx.x; i < a.n; i += blockDim.x)
        a.mask[i] = (a.n == 5) ? 1 : (sampson_inlier(s_E, a.x1[2 * i], a.x1[2 * i + 1], a.x2[2 * i], a.x2[2 * i + 1], a.thr_sq) ? 1 : 0);
}

int vo_rng_table(b200vo_ctx* ctx, int n, const uint32_t** d_table);

extern "C" int b200vo_find_essential_mat_ransac(b200vo_ctx* ctx, const float* p1, const float* p2, int n, const double K[9],
                                                double prob, double thr, int max_iters, double E[9], uint8_t* mask, int* found)
{
    if (!ctx || !p1 || !p2 || !K || !E || !mask || !found) return B200VO_E_BADARG;
    *found = 0;
    if (n < 0) return vo_set_err(ctx, B200VO_E_BADARG, "npoints >= 0 && points2.checkVector(2) == npoints");
    if (n < 5) return 0;   // cv2 returns an empty matrix
    VO_CUDA(ctx, cudaSetDevice(ctx->device));
    VO_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    const int iters = max_iters > 1 ? max_iters : 1;
    EmatArgs a{};
    a.n = n; a.iters = iters;
    a.fx = K[0]; a.fy = K[4]; a.cx = K[2]; a.cy = K[5];
    const double t = thr / ((a.fx + a.fy) / 2);
    a.thr_sq = (float)(t * t);
    a.conf = prob;
    const int n_raw = 10 * iters + 256;
    const uint32_t* rng = nullptr;
    VO_TRY(vo_rng_table(ctx, n_raw, &rng));
    const size_t b_p = vo_align((size_t)n * 8, 256), b_x = vo_align((size_t)n * 16, 256);
    const size_t b_s = vo_align((size_t)iters * 5 * 4, 256), b_m = vo_align((size_t)iters * EM_MAXM * 9 * 8, 256);
    const size_t b_nm = vo_align((size_t)iters * 4, 256), b_c = vo_align((size_t)iters * EM_MAXM * 4, 256);
    const size_t b_mask = vo_align((size_t)n, 256);
    VO_TRY(vo_reserve(ctx, ctx->d_scratch[5], 2 * b_p + 2 * b_x + b_s + b_m + b_nm + b_c + b_mask + 1024));
    uint8_t* d = (uint8_t*)ctx->d_scratch[5].p;
    a.p1 = (const float*)d; a.p2 = (const float*)(d + b_p); d += 2 * b_p;
    a.x1 = (double*)d; a.x2 = (double*)(d + b_x); d += 2 * b_x;
    a.samples = (int*)d; d += b_s;
    a.models = (double*)d; d += b_m;
    a.nmodels = (int*)d; d += b_nm;
    a.counts = (int*)d; d += b_c;
    a.mask = d; d += b_mask;
    uint8_t* d_small = d;                       // E[9] | result[3] | n | flags
    a.E = (double*)d_small; a.result = (int*)(d_small + 128); a.n_dev = (int*)(d_small + 192); a.flags = (int*)(d_small + 256);
    a.state = (int*)(d_small + 320);
    VO_TRY(vo_reserve_pinned(ctx, 2 * b_p + b_mask + 1024));
    uint8_t* hp = (uint8_t*)ctx->h_pin;
    memcpy(hp, p1, (size_t)n * 8);
    memcpy(hp + b_p, p2, (size_t)n * 8);
    VO_CUDA(ctx, cudaMemcpyAsync((void*)a.p1, hp, 2 * b_p, cudaMemcpyHostToDevice, ctx->stream));
    VO_CUDA(ctx, cudaMemsetAsync(d_small, 0, 512, ctx->stream));
    int* h_small = (int*)(hp + 2 * b_p + b_mask);
    h_small[0] = n;
    h_small[1] = iters; h_small[2] = 0; h_small[3] = -1; h_small[4] = 0;   // state: bound, max_good, winner, replayed
    VO_CUDA(ctx, cudaMemcpyAsync(a.n_dev, h_small, 4, cudaMemcpyHostToDevice, ctx->stream));
    VO_CUDA(ctx, cudaMemcpyAsync(a.state, h_small + 1, 16, cudaMemcpyHostToDevice, ctx->stream));
    emat_normalize_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(a);
    {
        const size_t smem = (size_t)n_raw * sizeof(int);
        if (smem > 48 * 1024)
            VO_CUDA(ctx, cudaFuncSetAttribute(ransac_samples_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ransac_samples_kernel<5><<<1, 128, smem, ctx->stream>>>(rng, n_raw, a.n_dev, iters, a.samples, a.flags);
    }
    // chunks of hypotheses in stream order; a chunk whose first sample lies beyond cv2's adaptive
    // iteration bound (known on the device after the previous chunk) exits immediately
    const int CH = 128;
    ctx->launches += 2;
    for (int c0 = 0; c0 < iters; c0 += CH) {
        a.chunk_start = c0; a.chunk_len = CH;
        emat_solve_kernel<<<(CH + 31) / 32, 32, 0, ctx->stream>>>(a);
        emat_score_kernel<<<dim3((n + 255) / 256, (CH + EM_ST - 1) / EM_ST), 256, 0, ctx->stream>>>(a);
        emat_update_kernel<<<1, 32, 0, ctx->stream>>>(a);
        ctx->launches += 3;
    }
    emat_finish_kernel<<<1, 256, 0, ctx->stream>>>(a);
    ctx->launches += 1;
    VO_CUDA(ctx, cudaGetLastError());
    VO_CUDA(ctx, cudaMemcpyAsync(hp, a.mask, b_mask, cudaMemcpyDeviceToHost, ctx->stream));
    VO_CUDA(ctx, cudaMemcpyAsync(hp + b_mask, d_small, 512, cudaMemcpyDeviceToHost, ctx->stream));
    VO_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    VO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1);
    const int* res = (const int*)(hp + b_mask + 128);
    const int flg = *(const int*)(hp + b_mask + 256);
    if (flg & 1) return vo_set_err(ctx, B200VO_E_UNSUPPORTED, "RNG table exhausted while drawing subsets (n=%d)", n);
    *found = res[0];
    if (res[0]) {
        memcpy(E, hp + b_mask, 72);
        memcpy(mask, hp, (size_t)n);
    }
    return 0;
}
