/*
 * oracle/gftt_oracle.c -- TEST INFRASTRUCTURE ONLY (see vo_oracle.h).
 *
 * Restates cv2.goodFeaturesToTrack(img, maxCorners, quality, minDistance, blockSize=3,
 * useHarrisDetector=False) as called at reference VisualOdometryPipeLine.py:256 (OpenCV
 * modules/imgproc/src/{featureselect,corner}.cpp; third party, not vendored).  Spec: SURVEY.md
 * A.4.  Pinned live against cv2 and through tests/golden/gftt.npz: the float32 min-eigenvalue map
 * is BIT-EQUAL to cv2.cornerMinEigenVal and the ordered corner list (what the reference consumes)
 * is identical, including exact ties between symmetric corners.
 */
#include "vo_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

static inline int refl(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * len - 2 - p;
    return p;
}

/* eig = (a + c) - sqrt((a - c)^2 + b^2), a = 0.5*sum Dx^2, b = sum DxDy, c = 0.5*sum Dy^2 over a
 * block x block window (REFLECT_101 on the covariance image); Sobel 3x3 scaled by 1/(4*block*255).
 *
 * The float32 operation order below was pinned bit for bit against cv2 4.13.0 (x86-64 wheel, AVX2
 * dispatch) with cv2.Sobel / cv2.boxFilter / cv2.cornerMinEigenVal probes:
 *   Dx  = fma(S0 + S2, k0, f32(S1 * k1))            S = [-1 0 1] row differences (exact ints)
 *   r   = fma(k0, p2, fma(k1, p1, f32(k0 * p0)))    for x <  (w & ~31)   (vectorised row filter)
 *       = ((k0*p0) + (k1*p1)) + (k0*p2)             for x >= (w & ~31)   (scalar tail, unfused)
 *   Dy  = r[y+1] - r[y-1]
 *   with k0 = f32(scale), k1 = f32(2*scale);
 *   box sums accumulate in DOUBLE and round once to float32 (cv2's ColumnSum<double,float>);
 *   the final expression is evaluated without FMA.
 * Ties between mathematically symmetric corners survive this arithmetic exactly as in cv2. */
void orc_min_eig_map(const uint8_t* img, int w, int h, size_t step, int block, float* eig)
{
    const double scale = 1.0 / (4.0 * block * 255.0);
    const float k0 = (float)scale, k1 = (float)(2.0 * scale);
    const int body = w & ~31;
    float* cov = (float*)malloc(sizeof(float) * 3 * (size_t)w * h);
    float* rsm = (float*)malloc(sizeof(float) * (size_t)w * (h + 2));   /* smoothed rows y = -1..h */
    for (int yy = -1; yy <= h; ++yy) {
        const uint8_t* r = img + (size_t)refl(yy, h) * step;
        float* o = rsm + (size_t)(yy + 1) * w;
        for (int x = 0; x < w; ++x) {
            const float p0 = r[refl(x - 1, w)], p1 = r[x], p2 = r[refl(x + 1, w)];
            if (x < body) o[x] = fmaf(k0, p2, fmaf(k1, p1, k0 * p0));
            else { float t = k0 * p0; float u = k1 * p1; t = t + u; u = k0 * p2; o[x] = t + u; }
        }
    }
    for (int y = 0; y < h; ++y) {
        const uint8_t* r0 = img + (size_t)refl(y - 1, h) * step;
        const uint8_t* r1 = img + (size_t)y * step;
        const uint8_t* r2 = img + (size_t)refl(y + 1, h) * step;
        for (int x = 0; x < w; ++x) {
            const int xm = refl(x - 1, w), xp = refl(x + 1, w);
            const int s0 = r0[xp] - r0[xm], s1 = r1[xp] - r1[xm], s2 = r2[xp] - r2[xm];
            const float mid = (float)s1 * k1;
            const float dx = fmaf((float)(s0 + s2), k0, mid);
            const float dy = rsm[(size_t)(y + 2) * w + x] - rsm[(size_t)y * w + x];
            float* c = cov + 3 * ((size_t)y * w + x);
            c[0] = dx * dx; c[1] = dx * dy; c[2] = dy * dy;
        }
    }
    /* unnormalised box filter exactly as cv2 runs it for CV_32F: row sums in double (left to right),
     * then a RUNNING column sum in double (SUM += entering row; out = (float)SUM; SUM -= leaving row).
     * The running sum keeps ~1e-17 absolute residue from strong rows above, which decides the last
     * float32 bit of weak responses below them -- it has to be replayed, not re-derived. */
    const int r = block / 2;
    double* rows = (double*)malloc(sizeof(double) * 3 * (size_t)w * (h + block - 1));
    for (int yy = -r; yy < h + block - 1 - r; ++yy) {
        const int ys = refl(yy, h);
        double* o = rows + 3 * (size_t)w * (yy + r);
        for (int x = 0; x < w; ++x)
            for (int ch = 0; ch < 3; ++ch) {
                double acc = 0;
                for (int dx = -r; dx <= block - 1 - r; ++dx) acc += (double)cov[3 * ((size_t)ys * w + refl(x + dx, w)) + ch];
                o[3 * x + ch] = acc;
            }
    }
    double* SUM = (double*)calloc(3 * (size_t)w, sizeof(double));
    for (int k = 0; k < block - 1; ++k)
        for (size_t i = 0; i < 3 * (size_t)w; ++i) SUM[i] += rows[3 * (size_t)w * k + i];
    for (int y = 0; y < h; ++y) {
        const double* Sp = rows + 3 * (size_t)w * (y + block - 1);
        const double* Sm = rows + 3 * (size_t)w * y;
        for (int x = 0; x < w; ++x) {
            float bx[3];
            for (int ch = 0; ch < 3; ++ch) {
                const double s0 = SUM[3 * x + ch] + Sp[3 * x + ch];
                bx[ch] = (float)s0;
                SUM[3 * x + ch] = s0 - Sm[3 * x + ch];
            }
            const float a = bx[0] * 0.5f, b = bx[1], c = bx[2] * 0.5f;
            const float t = a - c;
            const float u = t * t, v = b * b;
            eig[(size_t)y * w + x] = (a + c) - sqrtf(u + v);
        }
    }
    free(rows); free(SUM);
    free(cov); free(rsm);
}

typedef struct { float v; int idx; } cand_t;
static int cand_cmp(const void* pa, const void* pb)
{
    const cand_t* a = (const cand_t*)pa; const cand_t* b = (const cand_t*)pb;
    if (a->v != b->v) return a->v > b->v ? -1 : 1;          /* value descending */
    return a->idx > b->idx ? -1 : a->idx < b->idx ? 1 : 0;  /* ties: larger address first */
}

int orc_good_features_to_track(const uint8_t* img, int rows, int cols, size_t step, int max_corners,
                               double quality, double min_dist, int block_size,
                               float* corners_xy, int* n_out)
{
    *n_out = 0;
    if (!(quality > 0) || min_dist < 0 || max_corners < 0) return -1;
    const int w = cols, h = rows;
    float* eig = (float*)malloc(sizeof(float) * (size_t)w * h);
    orc_min_eig_map(img, w, h, step, block_size, eig);
    double maxv = 0;
    for (size_t i = 0; i < (size_t)w * h; ++i) if (eig[i] > maxv) maxv = eig[i];
    const float thr = (float)(maxv * quality);
    cand_t* cand = (cand_t*)malloc(sizeof(cand_t) * (size_t)w * h);
    int nc = 0;
    for (int y = 1; y < h - 1; ++y)
        for (int x = 1; x < w - 1; ++x) {
            const float v = eig[(size_t)y * w + x];
            if (!(v > thr)) continue;
            int is_max = 1;
            for (int dy = -1; dy <= 1 && is_max; ++dy)
                for (int dx = -1; dx <= 1; ++dx)
                    if (eig[(size_t)(y + dy) * w + x + dx] > v) { is_max = 0; break; }
            if (is_max) { cand[nc].v = v; cand[nc].idx = y * w + x; ++nc; }
        }
    qsort(cand, (size_t)nc, sizeof(cand_t), cand_cmp);
    int n = 0;
    if (min_dist >= 1) {
        const int cell = (int)lrint(min_dist);
        const int gw = (w + cell - 1) / cell, gh = (h + cell - 1) / cell;
        const double md2 = min_dist * min_dist;
        int* cnt = (int*)calloc((size_t)gw * gh, sizeof(int));
        short* pts = (short*)malloc(sizeof(short) * 2 * 16 * (size_t)gw * gh);
        for (int i = 0; i < nc; ++i) {
            const int y = cand[i].idx / w, x = cand[i].idx - y * w;
            const int xc = x / cell, yc = y / cell;
            int x1 = xc - 1 < 0 ? 0 : xc - 1, y1 = yc - 1 < 0 ? 0 : yc - 1;
            int x2 = xc + 1 > gw - 1 ? gw - 1 : xc + 1, y2 = yc + 1 > gh - 1 ? gh - 1 : yc + 1;
            int good = 1;
            for (int yy = y1; yy <= y2 && good; ++yy)
                for (int xx = x1; xx <= x2 && good; ++xx) {
                    const int cidx = yy * gw + xx;
                    for (int j = 0; j < cnt[cidx]; ++j) {
                        float dx = (float)(x - pts[2 * (16 * cidx + j)]), dy = (float)(y - pts[2 * (16 * cidx + j) + 1]);
                        if ((double)(dx * dx + dy * dy) < md2) { good = 0; break; }
                    }
                }
            if (good) {
                const int cidx = yc * gw + xc;
                if (cnt[cidx] < 16) { pts[2 * (16 * cidx + cnt[cidx])] = (short)x; pts[2 * (16 * cidx + cnt[cidx]) + 1] = (short)y; cnt[cidx]++; }
                corners_xy[2 * n] = (float)x; corners_xy[2 * n + 1] = (float)y;
                ++n;
                if (max_corners > 0 && n == max_corners) break;
            }
        }
        free(cnt); free(pts);
    } else {
        for (int i = 0; i < nc; ++i) {
            const int y = cand[i].idx / w, x = cand[i].idx - y * w;
            corners_xy[2 * n] = (float)x; corners_xy[2 * n + 1] = (float)y;
            ++n;
            if (max_corners > 0 && n == max_corners) break;
        }
    }
    *n_out = n;
    free(cand); free(eig);
    return 0;
}
