/*
 * oracle/klt_oracle.c -- TEST INFRASTRUCTURE ONLY (see vo_oracle.h).
 *
 * Restates what cv2.calcOpticalFlowPyrLK computes for the reference's two call
 * sites (VisualOdometryPipeLine.py:281 and :287): OpenCV's buildOpticalFlowPyramid /
 * pyrDown, calcSharrDeriv and LKTrackerInvoker (OpenCV modules/video/src/lkpyramid.cpp,
 * modules/imgproc/src/pyramids.cpp; third party, not vendored).  Spec: SURVEY.md
 * Appendix A.1-A.3.  Pinned against the installed cv2 4.13.0 by tests/test_oracle_klt.py
 * and tests/golden/klt_*.npz.
 *
 * Compile with -ffp-contract=off: float32 operations must stay separate.
 */
#include "vo_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

static inline int reflect101(int p, int len)
{
    /* gfedcb|abcdefgh|gfedcba ; valid for |overshoot| < len */
    if (len == 1) return 0;
    while (p < 0 || p >= len) {
        if (p < 0) p = -p;
        else p = 2 * len - 2 - p;
    }
    return p;
}

/* A.1: 5x5 binomial, REFLECT_101, (sum+128)>>8, output ((w+1)/2, (h+1)/2) */
void orc_pyr_down_u8(const uint8_t* src, int w, int h, size_t sstep, uint8_t* dst, size_t dstep)
{
    int dw = (w + 1) / 2, dh = (h + 1) / 2;
    int* rowbuf = (int*)malloc(sizeof(int) * (size_t)dw * 5);
    for (int dy = 0; dy < dh; ++dy) {
        for (int k = 0; k < 5; ++k) {
            int sy = reflect101(2 * dy - 2 + k, h);
            const uint8_t* s = src + (size_t)sy * sstep;
            int* r = rowbuf + (size_t)k * dw;
            for (int dx = 0; dx < dw; ++dx) {
                int x = 2 * dx;
                r[dx] = s[reflect101(x - 2, w)] + 4 * s[reflect101(x - 1, w)] + 6 * s[x] +
                        4 * s[reflect101(x + 1, w)] + s[reflect101(x + 2, w)];
            }
        }
        uint8_t* d = dst + (size_t)dy * dstep;
        for (int dx = 0; dx < dw; ++dx) {
            int v = rowbuf[dx] + 4 * rowbuf[dw + dx] + 6 * rowbuf[2 * dw + dx] +
                    4 * rowbuf[3 * dw + dx] + rowbuf[4 * dw + dx];
            d[dx] = (uint8_t)((v + 128) >> 8);
        }
    }
    free(rowbuf);
}

/* A.2: unscaled 3/10/3 Scharr, REFLECT_101 inside the image; interleaved (Ix,Iy) int16 */
void orc_scharr_s16(const uint8_t* src, int w, int h, size_t sstep, int16_t* dst)
{
    for (int y = 0; y < h; ++y) {
        const uint8_t* r0 = src + (size_t)reflect101(y - 1, h) * sstep;
        const uint8_t* r1 = src + (size_t)y * sstep;
        const uint8_t* r2 = src + (size_t)reflect101(y + 1, h) * sstep;
        for (int x = 0; x < w; ++x) {
            int xm = reflect101(x - 1, w), xp = reflect101(x + 1, w);
            int t0m = (r0[xm] + r2[xm]) * 3 + r1[xm] * 10;
            int t0p = (r0[xp] + r2[xp]) * 3 + r1[xp] * 10;
            int t1m = r2[xm] - r0[xm];
            int t1c = r2[x] - r0[x];
            int t1p = r2[xp] - r0[xp];
            dst[((size_t)y * w + x) * 2 + 0] = (int16_t)(t0p - t0m);
            dst[((size_t)y * w + x) * 2 + 1] = (int16_t)((t1p + t1m) * 3 + t1c * 10);
        }
    }
}

int orc_pyr_levels(int w, int h, int win_w, int win_h, int max_level)
{
    int levels = 1;
    for (int l = 1; l <= max_level; ++l) {
        w = (w + 1) / 2;
        h = (h + 1) / 2;
        if (w <= win_w || h <= win_h) break;
        ++levels;
    }
    return levels;
}

typedef struct {
    int w, h;
    uint8_t* img;   /* w*h, tight */
    int16_t* der;   /* w*h*2 */
} level_t;

static inline int pix(const level_t* L, int x, int y)
{
    return L->img[(size_t)reflect101(y, L->h) * L->w + reflect101(x, L->w)];
}
static inline int dpix(const level_t* L, int x, int y, int c)
{
    if (x < 0 || y < 0 || x >= L->w || y >= L->h) return 0;
    return L->der[((size_t)y * L->w + x) * 2 + c];
}
static inline int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }
static inline int cv_round_f(float v) { return (int)lrintf(v); } /* round-half-even under default mode */

#define W_BITS 14

int orc_calc_optical_flow_pyr_lk(const uint8_t* prev, const uint8_t* next, int rows, int cols,
                                 size_t prev_step, size_t next_step,
                                 const float* prev_pts, int n, int win_w, int win_h, int max_level,
                                 int crit_type, int crit_max_count, double crit_eps,
                                 int flags, double min_eig_thr,
                                 float* next_pts, uint8_t* status, float* err, int32_t* iters_out)
{
    if (max_level < 0 || win_w <= 2 || win_h <= 2) return -1;
    if (flags != 0) return -2;
    int nlev = orc_pyr_levels(cols, rows, win_w, win_h, max_level);
    level_t* P = (level_t*)calloc((size_t)nlev, sizeof(level_t));
    level_t* Nx = (level_t*)calloc((size_t)nlev, sizeof(level_t));
    for (int l = 0, w = cols, h = rows; l < nlev; ++l) {
        P[l].w = Nx[l].w = w;
        P[l].h = Nx[l].h = h;
        P[l].img = (uint8_t*)malloc((size_t)w * h);
        Nx[l].img = (uint8_t*)malloc((size_t)w * h);
        if (l == 0) {
            for (int y = 0; y < h; ++y) {
                memcpy(P[0].img + (size_t)y * w, prev + (size_t)y * prev_step, (size_t)w);
                memcpy(Nx[0].img + (size_t)y * w, next + (size_t)y * next_step, (size_t)w);
            }
        } else {
            orc_pyr_down_u8(P[l - 1].img, P[l - 1].w, P[l - 1].h, (size_t)P[l - 1].w, P[l].img, (size_t)w);
            orc_pyr_down_u8(Nx[l - 1].img, Nx[l - 1].w, Nx[l - 1].h, (size_t)Nx[l - 1].w, Nx[l].img, (size_t)w);
        }
        P[l].der = (int16_t*)malloc(sizeof(int16_t) * 2 * (size_t)w * h);
        orc_scharr_s16(P[l].img, w, h, (size_t)w, P[l].der);
        w = (w + 1) / 2;
        h = (h + 1) / 2;
    }
    int maxCount = (crit_type & 1) ? (crit_max_count < 0 ? 0 : crit_max_count > 100 ? 100 : crit_max_count) : 30;
    double epsd = (crit_type & 2) ? (crit_eps < 0 ? 0 : crit_eps > 10 ? 10 : crit_eps) : 0.01;
    epsd *= epsd;
    const float FLT_SCALE = 1.f / (1 << 20);
    const float hwx = (win_w - 1) * 0.5f, hwy = (win_h - 1) * 0.5f;
    const float minEigThr = (float)min_eig_thr;
    int wsz = win_w * win_h;

#pragma omp parallel
    {
        int16_t* Iw = (int16_t*)malloc(sizeof(int16_t) * 3 * (size_t)wsz);
        int16_t* dIx = Iw + wsz;
        int16_t* dIy = Iw + 2 * wsz;
#pragma omp for schedule(dynamic, 16)
        for (int i = 0; i < n; ++i) {
            float npx = 0.f, npy = 0.f;
            uint8_t st = 1;
            float e = 0.f;
            int total_iters = 0;
            for (int level = nlev - 1; level >= 0; --level) {
                const level_t* I = &P[level];
                const level_t* J = &Nx[level];
                float sc = (float)(1. / (1 << level));
                float ppx = prev_pts[2 * i] * sc, ppy = prev_pts[2 * i + 1] * sc;
                float nx, ny;
                if (level == nlev - 1) { nx = ppx; ny = ppy; }
                else { nx = npx * 2.f; ny = npy * 2.f; }
                npx = nx; npy = ny;
                ppx -= hwx; ppy -= hwy;
                int ipx = (int)floorf(ppx), ipy = (int)floorf(ppy);
                if (ipx < -win_w || ipx >= I->w || ipy < -win_h || ipy >= I->h) {
                    if (level == 0) { st = 0; e = 0.f; }
                    continue;
                }
                float a = ppx - ipx, b = ppy - ipy;
                int iw00 = cv_round_f((1.f - a) * (1.f - b) * (1 << W_BITS));
                int iw01 = cv_round_f(a * (1.f - b) * (1 << W_BITS));
                int iw10 = cv_round_f((1.f - a) * b * (1 << W_BITS));
                int iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
                int64_t sA11 = 0, sA12 = 0, sA22 = 0;
                for (int y = 0; y < win_h; ++y)
                    for (int x = 0; x < win_w; ++x) {
                        int X = ipx + x, Y = ipy + y;
                        int iv = descale(pix(I, X, Y) * iw00 + pix(I, X + 1, Y) * iw01 +
                                         pix(I, X, Y + 1) * iw10 + pix(I, X + 1, Y + 1) * iw11, W_BITS - 5);
                        int ix = descale(dpix(I, X, Y, 0) * iw00 + dpix(I, X + 1, Y, 0) * iw01 +
                                         dpix(I, X, Y + 1, 0) * iw10 + dpix(I, X + 1, Y + 1, 0) * iw11, W_BITS);
                        int iy = descale(dpix(I, X, Y, 1) * iw00 + dpix(I, X + 1, Y, 1) * iw01 +
                                         dpix(I, X, Y + 1, 1) * iw10 + dpix(I, X + 1, Y + 1, 1) * iw11, W_BITS);
                        Iw[y * win_w + x] = (int16_t)iv;
                        dIx[y * win_w + x] = (int16_t)ix;
                        dIy[y * win_w + x] = (int16_t)iy;
                        sA11 += (int64_t)ix * ix;
                        sA12 += (int64_t)ix * iy;
                        sA22 += (int64_t)iy * iy;
                    }
                float A11 = (float)sA11 * FLT_SCALE, A12 = (float)sA12 * FLT_SCALE, A22 = (float)sA22 * FLT_SCALE;
                float D = A11 * A22 - A12 * A12;
                float minEig = (A22 + A11 - sqrtf((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) / (float)(2 * win_w * win_h);
                if (minEig < minEigThr || D < FLT_EPSILON) {
                    if (level == 0) st = 0;
                    continue;
                }
                D = 1.f / D;
                nx -= hwx; ny -= hwy;
                float pdx = 0.f, pdy = 0.f;
                for (int j = 0; j < maxCount; ++j) {
                    int inx = (int)floorf(nx), iny = (int)floorf(ny);
                    if (inx < -win_w || inx >= J->w || iny < -win_h || iny >= J->h) {
                        if (level == 0) st = 0;
                        break;
                    }
                    ++total_iters;
                    a = nx - inx; b = ny - iny;
                    iw00 = cv_round_f((1.f - a) * (1.f - b) * (1 << W_BITS));
                    iw01 = cv_round_f(a * (1.f - b) * (1 << W_BITS));
                    iw10 = cv_round_f((1.f - a) * b * (1 << W_BITS));
                    iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
                    int64_t sb1 = 0, sb2 = 0;
                    for (int y = 0; y < win_h; ++y)
                        for (int x = 0; x < win_w; ++x) {
                            int X = inx + x, Y = iny + y;
                            int diff = descale(pix(J, X, Y) * iw00 + pix(J, X + 1, Y) * iw01 +
                                               pix(J, X, Y + 1) * iw10 + pix(J, X + 1, Y + 1) * iw11, W_BITS - 5) -
                                       Iw[y * win_w + x];
                            sb1 += (int64_t)diff * dIx[y * win_w + x];
                            sb2 += (int64_t)diff * dIy[y * win_w + x];
                        }
                    float b1 = (float)sb1 * FLT_SCALE, b2 = (float)sb2 * FLT_SCALE;
                    float dx = (A12 * b2 - A22 * b1) * D;
                    float dy = (A12 * b1 - A11 * b2) * D;
                    nx += dx; ny += dy;
                    npx = nx + hwx; npy = ny + hwy;
                    if ((double)dx * dx + (double)dy * dy <= epsd) break;
                    if (j > 0 && fabsf(dx + pdx) < 0.01f && fabsf(dy + pdy) < 0.01f) {
                        npx -= dx * 0.5f;
                        npy -= dy * 0.5f;
                        break;
                    }
                    pdx = dx; pdy = dy;
                }
                if (st && level == 0) {
                    float qx = npx - hwx, qy = npy - hwy;
                    int inx = (int)floorf(qx), iny = (int)floorf(qy);
                    if (inx < -win_w || inx >= J->w || iny < -win_h || iny >= J->h) {
                        st = 0;
                        continue;
                    }
                    a = qx - inx; b = qy - iny;
                    iw00 = cv_round_f((1.f - a) * (1.f - b) * (1 << W_BITS));
                    iw01 = cv_round_f(a * (1.f - b) * (1 << W_BITS));
                    iw10 = cv_round_f((1.f - a) * b * (1 << W_BITS));
                    iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
                    int64_t se = 0;
                    for (int y = 0; y < win_h; ++y)
                        for (int x = 0; x < win_w; ++x) {
                            int X = inx + x, Y = iny + y;
                            int diff = descale(pix(J, X, Y) * iw00 + pix(J, X + 1, Y) * iw01 +
                                               pix(J, X, Y + 1) * iw10 + pix(J, X + 1, Y + 1) * iw11, W_BITS - 5) -
                                       Iw[y * win_w + x];
                            se += diff < 0 ? -diff : diff;
                        }
                    e = (float)se * 1.f / (float)(32 * win_w * win_h); /* cv2: errval * 1.f/(32*w*h) parses as a division */
                }
            }
            next_pts[2 * i] = npx;
            next_pts[2 * i + 1] = npy;
            status[i] = st;
            err[i] = e;
            if (iters_out) iters_out[i] = total_iters;
        }
        free(Iw);
    }
    for (int l = 0; l < nlev; ++l) {
        free(P[l].img); free(P[l].der); free(Nx[l].img);
    }
    free(P); free(Nx);
    return 0;
}
