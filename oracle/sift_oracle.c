/*
 * oracle/sift_oracle.c -- TEST INFRASTRUCTURE ONLY (see vo_oracle.h).
 *
 * CPU restatement of cv2.SIFT_create().detectAndCompute(img, None) with the default parameters
 * (reference VisualOdometryPipeLine.py:35, :226-227): nfeatures 0, nOctaveLayers 3,
 * contrastThreshold 0.04, edgeThreshold 10, sigma 1.6, float32 descriptors, image doubled first.
 * Follows OpenCV 4.x modules/features2d/src/sift.dispatch.cpp (createInitialImage, buildGaussianPyramid,
 * buildDoGPyramid, findScaleSpaceExtrema, detectAndCompute) and sift.simd.hpp (adjustLocalExtrema,
 * calcOrientationHist, calcSIFTDescriptor), core/matx.hpp (3x3 Cramer solve), imgproc smooth
 * (getGaussianKernel), imgproc resize (INTER_LINEAR x2, INTER_NEAREST /2), features2d keypoint.cpp
 * (KeyPointsFilter::removeDuplicatedSorted), core mathfuncs_core (fastAtan2 polynomial).
 *
 * Pin: floating point, tolerance-based (tests/test_oracle_sift.py against the installed cv2 4.13.0).
 * Bit-equal to cv2: the Gaussian kernels, the doubled image, cv2.GaussianBlur on float32 (hence the whole Gaussian
 * and DoG pyramids), fastAtan2, and -- with the FMA contraction GCC applies to sift.simd.hpp pinned for the sub-pixel
 * solve -- keypoint count, order, octave codes and x, y.  Not reproducible: IPP's exp (ippsExp_32f_A21, 1 ulp from expf
 * in ~20 % of the arguments), which moves orientation angles by an ulp or two in ~25 % of the keypoints and one
 * descriptor entry by 1 in ~0.1 % of the descriptors (measured on the synthetic KITTI / Parking / Malaga frames).
 * The sub-pixel refinement differentiates DoG values twice: one ulp of a Gaussian image moves a keypoint by ~1e-5 px,
 * so the bit-equal pyramid is what the keypoint parity rests on.
 */
#include "vo_oracle.h"
#include <float.h>
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define SIFT_LAYERS 3
#define SIFT_SIGMA 1.6
#define SIFT_CONTRAST_THR 0.04
#define SIFT_EDGE_THR 10.0
#define SIFT_IMG_BORDER 5
#define SIFT_MAX_INTERP_STEPS 5
#define SIFT_ORI_HIST_BINS 36
#define SIFT_ORI_SIG_FCTR 1.5f
#define SIFT_ORI_RADIUS 4.5f
#define SIFT_ORI_PEAK_RATIO 0.8f
#define SIFT_DESCR_WIDTH 4
#define SIFT_DESCR_HIST_BINS 8
#define SIFT_DESCR_SCL_FCTR 3.f
#define SIFT_DESCR_MAG_THR 0.2f
#define SIFT_INT_DESCR_FCTR 512.f
#define SIFT_FIRST_OCTAVE (-1)

typedef struct { int rows, cols; float* p; } fimg;

static int cv_round_d(double v) { return (int)lrint(v); }      /* round half to even, as cvRound */
static int cv_round_f(float v) { return (int)lrintf(v); }
static int cv_floor_f(float v) { int i = (int)v; return i - (i > v); }
static int cv_floor_d(double v) { int i = (int)v; return i - (i > v); }

static int reflect101(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) {
        if (p < 0) p = -p;
        else p = 2 * len - 2 - p;
    }
    return p;
}

/* cv::getGaussianKernel(n, sigma, CV_32F): x doubled, scale -0.125 / sigma^2, normalised, cast to float */
int orc_gaussian_kernel_f32(double sigma, float* k, int max_n)
{
    int n = cv_round_d(sigma * 4 * 2 + 1) | 1;
    if (n > max_n) return -n;
    const double scale2x = -0.125 / (sigma * sigma);
    const int n2 = (n - 1) / 2;
    double vals[128];
    double sum = 0;
    for (int i = 0, x = 1 - n; i < n2; ++i, x += 2) { vals[i] = exp((double)(x * x) * scale2x); sum += vals[i]; }
    sum *= 2; sum += 1.0;
    const double mul1 = 1.0 / sum;
    for (int i = 0; i < n2; ++i) { const float t = (float)(vals[i] * mul1); k[i] = t; k[n - 1 - i] = t; }
    k[n2] = (float)(1.0 * mul1);
    return n;
}

/* separable blur, BORDER_REFLECT_101; row pass k ascending, column pass centre first then symmetric pairs (what OpenCV's
 * RowVec_32f / SymmColumnVec_32f do: bit-equal to cv2.GaussianBlur on CV_32F in the build this was pinned on) */
void orc_gaussian_blur_f32(const float* src, int rows, int cols, double sigma, float* dst)
{
    float k[128];
    const int n = orc_gaussian_kernel_f32(sigma, k, 128), h = n / 2;
    /* where OpenCV's vector bodies contract to FMA (pinned on the installed build, AVX2 dispatch): the row filter for
     * x < (cols & ~3), the column filter for x < (cols & ~7); the remaining columns are mul + add */
    const int vec = cols & ~3, v8 = cols & ~7;
    float* tmp = (float*)malloc((size_t)rows * cols * sizeof(float));
    int* xi = (int*)malloc((size_t)(cols + n) * sizeof(int));
    for (int x = -h; x < cols + h; ++x) xi[x + h] = reflect101(x, cols);
    for (int y = 0; y < rows; ++y) {
        const float* s = src + (size_t)y * cols;
        float* t = tmp + (size_t)y * cols;
        for (int x = 0; x < cols; ++x) {
            float acc = k[0] * s[xi[x]];
            if (x < vec) for (int j = 1; j < n; ++j) acc = fmaf(s[xi[x + j]], k[j], acc);
            else for (int j = 1; j < n; ++j) acc += s[xi[x + j]] * k[j];
            t[x] = acc;
        }
    }
    for (int y = 0; y < rows; ++y) {
        float* d = dst + (size_t)y * cols;
        const float* c0 = tmp + (size_t)y * cols;
        for (int x = 0; x < cols; ++x) d[x] = k[h] * c0[x];
        for (int j = 1; j <= h; ++j) {
            const float* a = tmp + (size_t)reflect101(y - j, rows) * cols;
            const float* b = tmp + (size_t)reflect101(y + j, rows) * cols;
            for (int x = 0; x < v8; ++x) d[x] = fmaf(k[h + j], a[x] + b[x], d[x]);
            for (int x = v8; x < cols; ++x) d[x] += k[h + j] * (a[x] + b[x]);
        }
    }
    free(tmp); free(xi);
}

/* createInitialImage: u8 -> float, resize x2 INTER_LINEAR (weights 0.25 / 0.75, exact in float), blur sqrt(sigma^2 - 1) */
static void upscale2_linear(const uint8_t* img, int rows, int cols, size_t step, fimg* out)
{
    out->rows = rows * 2; out->cols = cols * 2;
    out->p = (float*)malloc((size_t)out->rows * out->cols * sizeof(float));
    int* sx = (int*)malloc((size_t)out->cols * sizeof(int));
    float* ax = (float*)malloc((size_t)out->cols * sizeof(float));
    for (int dx = 0; dx < out->cols; ++dx) {
        float fx = (float)((dx + 0.5) * 0.5 - 0.5);
        int s = cv_floor_f(fx);
        fx -= s;
        if (s < 0) { fx = 0; s = 0; }
        if (s >= cols - 1) { fx = 0; s = cols - 1; }
        sx[dx] = s; ax[dx] = fx;
    }
    for (int dy = 0; dy < out->rows; ++dy) {
        float fy = (float)((dy + 0.5) * 0.5 - 0.5);
        int s = cv_floor_f(fy);
        fy -= s;
        if (s < 0) { fy = 0; s = 0; }
        if (s >= rows - 1) { fy = 0; s = rows - 1; }
        const uint8_t* r0 = img + (size_t)s * step;
        const uint8_t* r1 = img + (size_t)(s + 1 < rows ? s + 1 : s) * step;
        float* d = out->p + (size_t)dy * out->cols;
        for (int dx = 0; dx < out->cols; ++dx) {
            const int x0 = sx[dx], x1 = x0 + 1 < cols ? x0 + 1 : x0;
            const float a = ax[dx];
            const float h0 = (float)r0[x0] * (1.f - a) + (float)r0[x1] * a;
            const float h1 = (float)r1[x0] * (1.f - a) + (float)r1[x1] * a;
            d[dx] = h0 * (1.f - fy) + h1 * fy;
        }
    }
    free(sx); free(ax);
}

/* resize(src, Size(cols/2, rows/2), INTER_NEAREST): sx = min(floor(x * (1 / (dcols / scols))), scols - 1) */
static void downsample_nearest(const fimg* src, fimg* dst)
{
    dst->rows = src->rows / 2; dst->cols = src->cols / 2;
    dst->p = (float*)malloc((size_t)(dst->rows > 0 ? dst->rows : 1) * (dst->cols > 0 ? dst->cols : 1) * sizeof(float));
    if (dst->rows <= 0 || dst->cols <= 0) return;
    const double ifx = 1.0 / ((double)dst->cols / src->cols), ify = 1.0 / ((double)dst->rows / src->rows);
    for (int y = 0; y < dst->rows; ++y) {
        int sy = cv_floor_d(y * ify);
        if (sy > src->rows - 1) sy = src->rows - 1;
        for (int x = 0; x < dst->cols; ++x) {
            int sxx = cv_floor_d(x * ifx);
            if (sxx > src->cols - 1) sxx = src->cols - 1;
            dst->p[(size_t)y * dst->cols + x] = src->p[(size_t)sy * src->cols + sxx];
        }
    }
}

typedef struct {
    int n_octaves;
    fimg* gauss;   /* n_octaves * (LAYERS + 3) */
    fimg* dog;     /* n_octaves * (LAYERS + 2) */
} sift_pyr;

static void pyr_free(sift_pyr* P)
{
    for (int i = 0; i < P->n_octaves * (SIFT_LAYERS + 3); ++i) free(P->gauss[i].p);
    for (int i = 0; i < P->n_octaves * (SIFT_LAYERS + 2); ++i) free(P->dog[i].p);
    free(P->gauss); free(P->dog);
}

static void pyr_build(const uint8_t* img, int rows, int cols, size_t step, sift_pyr* P)
{
    fimg up;
    upscale2_linear(img, rows, cols, step, &up);
    const float sg = (float)SIFT_SIGMA;                                   /* createInitialImage takes sigma as a float */
    const float sig_diff = sqrtf(fmaxf(sg * sg - 0.5f * 0.5f * 4, 0.01f));
    fimg base = {up.rows, up.cols, (float*)malloc((size_t)up.rows * up.cols * sizeof(float))};
    orc_gaussian_blur_f32(up.p, up.rows, up.cols, (double)sig_diff, base.p);
    free(up.p);
    const int mn = base.cols < base.rows ? base.cols : base.rows;
    P->n_octaves = cv_round_d(log((double)mn) / log(2.) - 2) - SIFT_FIRST_OCTAVE;
    const int L = SIFT_LAYERS;
    P->gauss = (fimg*)calloc((size_t)P->n_octaves * (L + 3), sizeof(fimg));
    P->dog = (fimg*)calloc((size_t)P->n_octaves * (L + 2), sizeof(fimg));
    double sig[SIFT_LAYERS + 3];
    sig[0] = SIFT_SIGMA;
    const double k = pow(2., 1. / L);
    for (int i = 1; i < L + 3; ++i) {
        const double sig_prev = pow(k, (double)(i - 1)) * SIFT_SIGMA, sig_total = sig_prev * k;
        sig[i] = sqrt(sig_total * sig_total - sig_prev * sig_prev);
    }
    for (int o = 0; o < P->n_octaves; ++o)
        for (int i = 0; i < L + 3; ++i) {
            fimg* dst = &P->gauss[o * (L + 3) + i];
            if (o == 0 && i == 0) *dst = base;
            else if (i == 0) downsample_nearest(&P->gauss[(o - 1) * (L + 3) + L], dst);
            else {
                const fimg* src = &P->gauss[o * (L + 3) + i - 1];
                dst->rows = src->rows; dst->cols = src->cols;
                dst->p = (float*)malloc((size_t)(src->rows > 0 ? src->rows : 1) * (src->cols > 0 ? src->cols : 1) * sizeof(float));
                if (src->rows > 0 && src->cols > 0) orc_gaussian_blur_f32(src->p, src->rows, src->cols, sig[i], dst->p);
            }
        }
    for (int o = 0; o < P->n_octaves; ++o)
        for (int i = 0; i < L + 2; ++i) {
            const fimg* a = &P->gauss[o * (L + 3) + i];
            const fimg* b = &P->gauss[o * (L + 3) + i + 1];
            fimg* d = &P->dog[o * (L + 2) + i];
            d->rows = a->rows; d->cols = a->cols;
            const size_t n = (size_t)(a->rows > 0 ? a->rows : 0) * (a->cols > 0 ? a->cols : 0);
            d->p = (float*)malloc((n ? n : 1) * sizeof(float));
            for (size_t q = 0; q < n; ++q) d->p[q] = b->p[q] - a->p[q];
        }
}

/* core/mathfuncs_core.simd.hpp fastAtan32f (vector body: Horner with FMA), degrees */
float orc_fast_atan2_deg(float y, float x)
{
    const float scale = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale;
    const float p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
    const float ax = fabsf(x), ay = fabsf(y);
    const float c = fminf(ax, ay) / (fmaxf(ax, ay) + (float)DBL_EPSILON);
    const float cc = c * c;
    float a = fmaf(fmaf(fmaf(cc, p7, p5), cc, p3), cc, p1) * c;
    if (!(ax >= ay)) a = 90.f - a;
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

/* exp / cos / sin / 2^x as float: OpenCV calls IPP's and libm's float routines here, whose results are the correctly
 * rounded ones in all but a few cases; the restatement (and the CUDA path, which must equal it) rounds the double result */
static float exp_f(float x) { return (float)exp((double)x); }
static float cos_f(float x) { return (float)cos((double)x); }
static float sin_f(float x) { return (float)sin((double)x); }
static float exp2_f(float x) { return (float)exp2((double)x); }

#define AT(im, r, c) ((im)->p[(size_t)(r) * (im)->cols + (c)])

typedef struct { float x, y, size, angle, response; int32_t octave; } sift_kp;

static int adjust_local_extrema(const sift_pyr* P, sift_kp* kpt, int octv, int* layer, int* r, int* c)
{
    const float img_scale = 1.f / 255, deriv_scale = img_scale * 0.5f, second_deriv_scale = img_scale, cross_deriv_scale = img_scale * 0.25f;
    float xi = 0, xr = 0, xc = 0, contr = 0;
    int i = 0;
    for (; i < SIFT_MAX_INTERP_STEPS; ++i) {
        const int idx = octv * (SIFT_LAYERS + 2) + *layer;
        const fimg *img = &P->dog[idx], *prev = &P->dog[idx - 1], *next = &P->dog[idx + 1];
        const int R = *r, C = *c;
        const float dD[3] = {(AT(img, R, C + 1) - AT(img, R, C - 1)) * deriv_scale, (AT(img, R + 1, C) - AT(img, R - 1, C)) * deriv_scale,
                             (AT(next, R, C) - AT(prev, R, C)) * deriv_scale};
        const float v2 = AT(img, R, C) * 2;
        const float dxx = (AT(img, R, C + 1) + AT(img, R, C - 1) - v2) * second_deriv_scale;
        const float dyy = (AT(img, R + 1, C) + AT(img, R - 1, C) - v2) * second_deriv_scale;
        const float dss = (AT(next, R, C) + AT(prev, R, C) - v2) * second_deriv_scale;
        const float dxy = (AT(img, R + 1, C + 1) - AT(img, R + 1, C - 1) - AT(img, R - 1, C + 1) + AT(img, R - 1, C - 1)) * cross_deriv_scale;
        const float dxs = (AT(next, R, C + 1) - AT(next, R, C - 1) - AT(prev, R, C + 1) + AT(prev, R, C - 1)) * cross_deriv_scale;
        const float dys = (AT(next, R + 1, C) - AT(next, R - 1, C) - AT(prev, R + 1, C) + AT(prev, R - 1, C)) * cross_deriv_scale;
        const float a00 = dxx, a01 = dxy, a02 = dxs, a10 = dxy, a11 = dyy, a12 = dys, a20 = dxs, a21 = dys, a22 = dss;
        float X[3] = {0, 0, 0};
        /* Matx33f H(dxx, dxy, dxs, dxy, dyy, dys, dxs, dys, dss); X = H.solve(dD, DECOMP_LU): Cramer's rule (matx.hpp), with
         * the contraction GCC applies in the AVX2 / AVX-512 copies of sift.simd.hpp: x*y - z*w = fma(x, y, -(z*w)), and the
         * three-term sums as nested FMAs (pinned: keypoint x, y bit-equal to cv2 on every test frame) */
#define DET2(x, y, z, w) fmaf((x), (y), -((z) * (w)))
        {
            const float P = DET2(a11, a22, a21, a12), Q = DET2(a10, a22, a20, a12), Rr = DET2(a10, a21, a20, a11);
            float d = fmaf(a02, Rr, fmaf(a00, P, -(a01 * Q)));
            if (d != 0) {
                d = 1 / d;
                const float b0 = dD[0], b1 = dD[1], b2 = dD[2];
                const float P0 = DET2(a11, a22, a12, a21), Q0 = DET2(b1, a22, a12, b2), R0 = DET2(b1, a21, a11, b2);
                const float Q1 = DET2(a10, a22, a12, a20), R1 = DET2(a10, b2, b1, a20);
                const float P2 = DET2(a11, b2, b1, a21), R2 = DET2(a10, a21, a11, a20);
                X[0] = d * fmaf(a02, R0, fmaf(b0, P0, -(a01 * Q0)));
                X[1] = d * fmaf(a02, R1, fmaf(a00, Q0, -(b0 * Q1)));
                X[2] = d * fmaf(b0, R2, fmaf(a00, P2, -(a01 * R1)));
            }
        }
        xi = -X[2]; xr = -X[1]; xc = -X[0];
        if (fabsf(xi) < 0.5f && fabsf(xr) < 0.5f && fabsf(xc) < 0.5f) break;
        if (fabsf(xi) > (float)(INT_MAX / 3) || fabsf(xr) > (float)(INT_MAX / 3) || fabsf(xc) > (float)(INT_MAX / 3)) return 0;
        *c += cv_round_f(xc); *r += cv_round_f(xr); *layer += cv_round_f(xi);
        if (*layer < 1 || *layer > SIFT_LAYERS || *c < SIFT_IMG_BORDER || *c >= img->cols - SIFT_IMG_BORDER || *r < SIFT_IMG_BORDER ||
            *r >= img->rows - SIFT_IMG_BORDER)
            return 0;
    }
    if (i >= SIFT_MAX_INTERP_STEPS) return 0;
    {
        const int idx = octv * (SIFT_LAYERS + 2) + *layer;
        const fimg *img = &P->dog[idx], *prev = &P->dog[idx - 1], *next = &P->dog[idx + 1];
        const int R = *r, C = *c;
        const float dD[3] = {(AT(img, R, C + 1) - AT(img, R, C - 1)) * deriv_scale, (AT(img, R + 1, C) - AT(img, R - 1, C)) * deriv_scale,
                             (AT(next, R, C) - AT(prev, R, C)) * deriv_scale};
        const float t = dD[0] * xc + dD[1] * xr + dD[2] * xi;
        contr = fmaf(AT(img, R, C), img_scale, t * 0.5f);      /* contracted form, pinned like the solve */
        if (fabsf(contr) * SIFT_LAYERS < (float)SIFT_CONTRAST_THR) return 0;
        const float v2 = AT(img, R, C) * 2.f;
        const float dxx = (AT(img, R, C + 1) + AT(img, R, C - 1) - v2) * second_deriv_scale;
        const float dyy = (AT(img, R + 1, C) + AT(img, R - 1, C) - v2) * second_deriv_scale;
        const float dxy = (AT(img, R + 1, C + 1) - AT(img, R + 1, C - 1) - AT(img, R - 1, C + 1) + AT(img, R - 1, C - 1)) * cross_deriv_scale;
        const float tr = dxx + dyy, det = dxx * dyy - dxy * dxy;
        const float et = (float)SIFT_EDGE_THR;
        if (det <= 0 || tr * tr * et >= (et + 1) * (et + 1) * det) return 0;
    }
    kpt->x = (*c + xc) * (1 << octv);
    kpt->y = (*r + xr) * (1 << octv);
    kpt->octave = octv + (*layer << 8) + (cv_round_d((xi + 0.5) * 255) << 16);
    kpt->size = (float)SIFT_SIGMA * exp2_f((*layer + xi) / SIFT_LAYERS) * (1 << octv) * 2;
    kpt->response = fabsf(contr);
    return 1;
}

static float calc_orientation_hist(const fimg* img, int px, int py, int radius, float sigma, float* hist, int n)
{
    const float expf_scale = -1.f / (2.f * sigma * sigma);
    float tbuf[SIFT_ORI_HIST_BINS + 4];
    float* temphist = tbuf + 2;
    for (int i = 0; i < n; ++i) temphist[i] = 0.f;
    for (int i = -radius; i <= radius; ++i) {
        const int y = py + i;
        if (y <= 0 || y >= img->rows - 1) continue;
        for (int j = -radius; j <= radius; ++j) {
            const int x = px + j;
            if (x <= 0 || x >= img->cols - 1) continue;
            const float dx = AT(img, y, x + 1) - AT(img, y, x - 1);
            const float dy = AT(img, y - 1, x) - AT(img, y + 1, x);
            const float w = exp_f((float)(i * i + j * j) * expf_scale);
            const float ori = orc_fast_atan2_deg(dy, dx);
            const float mag = sqrtf(dx * dx + dy * dy);
            int bin = cv_round_f((n / 360.f) * ori);
            if (bin >= n) bin -= n;
            if (bin < 0) bin += n;
            temphist[bin] += w * mag;
        }
    }
    temphist[-1] = temphist[n - 1]; temphist[-2] = temphist[n - 2];
    temphist[n] = temphist[0]; temphist[n + 1] = temphist[1];
    for (int i = 0; i < n; ++i) {   /* vector body (32 of the 36 bins) with FMAs, scalar tail without */
        if (i < 32)
            hist[i] = fmaf(temphist[i - 2] + temphist[i + 2], 1.f / 16.f, fmaf(temphist[i - 1] + temphist[i + 1], 4.f / 16.f, temphist[i] * (6.f / 16.f)));
        else
            hist[i] = (temphist[i - 2] + temphist[i + 2]) * (1.f / 16.f) + (temphist[i - 1] + temphist[i + 1]) * (4.f / 16.f) + temphist[i] * (6.f / 16.f);
    }
    float mx = hist[0];
    for (int i = 1; i < n; ++i) mx = fmaxf(mx, hist[i]);
    return mx;
}

static void calc_sift_descriptor(const fimg* img, float ptx, float pty, float ori, float scl, float* dst)
{
    const int d = SIFT_DESCR_WIDTH, n = SIFT_DESCR_HIST_BINS;
    const int px = cv_round_f(ptx), py = cv_round_f(pty);
    float cos_t = cos_f(ori * (float)(3.14159265358979323846 / 180));
    float sin_t = sin_f(ori * (float)(3.14159265358979323846 / 180));
    const float bins_per_rad = n / 360.f;
    const float exp_scale = -1.f / (d * d * 0.5f);
    const float hist_width = SIFT_DESCR_SCL_FCTR * scl;
    int radius = cv_round_f(hist_width * 1.4142135623730951f * (d + 1) * 0.5f);
    const int diag = (int)sqrt((double)img->cols * img->cols + (double)img->rows * img->rows);
    if (radius > diag) radius = diag;
    cos_t /= hist_width; sin_t /= hist_width;
    const int rows = img->rows, cols = img->cols;
    float hist[(SIFT_DESCR_WIDTH + 2) * (SIFT_DESCR_WIDTH + 2) * (SIFT_DESCR_HIST_BINS + 2)];
    memset(hist, 0, sizeof(hist));
    for (int i = -radius; i <= radius; ++i)
        for (int j = -radius; j <= radius; ++j) {
            const float c_rot = j * cos_t - i * sin_t;
            const float r_rot = j * sin_t + i * cos_t;
            float rbin = r_rot + d / 2 - 0.5f;
            float cbin = c_rot + d / 2 - 0.5f;
            const int r = py + i, c = px + j;
            if (rbin > -1 && rbin < d && cbin > -1 && cbin < d && r > 0 && r < rows - 1 && c > 0 && c < cols - 1) {
                const float dx = AT(img, r, c + 1) - AT(img, r, c - 1);
                const float dy = AT(img, r - 1, c) - AT(img, r + 1, c);
                const float w = exp_f((c_rot * c_rot + r_rot * r_rot) * exp_scale);
                const float o = orc_fast_atan2_deg(dy, dx);
                const float mag = sqrtf(dx * dx + dy * dy) * w;
                float obin = (o - ori) * bins_per_rad;
                const int r0 = cv_floor_f(rbin), c0 = cv_floor_f(cbin);
                int o0 = cv_floor_f(obin);
                rbin -= r0; cbin -= c0; obin -= o0;
                if (o0 < 0) o0 += n;
                if (o0 >= n) o0 -= n;
                const float v_r1 = mag * rbin, v_r0 = mag - v_r1;
                const float v_rc11 = v_r1 * cbin, v_rc10 = v_r1 - v_rc11;
                const float v_rc01 = v_r0 * cbin, v_rc00 = v_r0 - v_rc01;
                const float v_rco111 = v_rc11 * obin, v_rco110 = v_rc11 - v_rco111;
                const float v_rco101 = v_rc10 * obin, v_rco100 = v_rc10 - v_rco101;
                const float v_rco011 = v_rc01 * obin, v_rco010 = v_rc01 - v_rco011;
                const float v_rco001 = v_rc00 * obin, v_rco000 = v_rc00 - v_rco001;
                const int idx = ((r0 + 1) * (d + 2) + c0 + 1) * (n + 2) + o0;
                hist[idx] += v_rco000; hist[idx + 1] += v_rco001;
                hist[idx + (n + 2)] += v_rco010; hist[idx + (n + 3)] += v_rco011;
                hist[idx + (d + 2) * (n + 2)] += v_rco100; hist[idx + (d + 2) * (n + 2) + 1] += v_rco101;
                hist[idx + (d + 3) * (n + 2)] += v_rco110; hist[idx + (d + 3) * (n + 2) + 1] += v_rco111;
            }
        }
    float raw[SIFT_DESCR_WIDTH * SIFT_DESCR_WIDTH * SIFT_DESCR_HIST_BINS];
    for (int i = 0; i < d; ++i)
        for (int j = 0; j < d; ++j) {
            const int idx = ((i + 1) * (d + 2) + (j + 1)) * (n + 2);
            hist[idx] += hist[idx + n];
            hist[idx + 1] += hist[idx + n + 1];
            for (int k = 0; k < n; ++k) raw[(i * d + j) * n + k] = hist[idx + k];
        }
    const int len = d * d * n;
    float nrm2 = 0;
    for (int k = 0; k < len; ++k) nrm2 += raw[k] * raw[k];
    const float thr = sqrtf(nrm2) * SIFT_DESCR_MAG_THR;
    nrm2 = 0;
    for (int k = 0; k < len; ++k) { const float v = fminf(raw[k], thr); raw[k] = v; nrm2 += v * v; }
    nrm2 = SIFT_INT_DESCR_FCTR / fmaxf(sqrtf(nrm2), FLT_EPSILON);
    for (int k = 0; k < len; ++k) {
        int v = cv_round_f(raw[k] * nrm2);
        dst[k] = (float)(v < 0 ? 0 : v > 255 ? 255 : v);
    }
}

static int kp_less(const void* pa, const void* pb)
{
    const sift_kp* a = (const sift_kp*)pa;
    const sift_kp* b = (const sift_kp*)pb;
    if (a->x != b->x) return a->x < b->x ? -1 : 1;
    if (a->y != b->y) return a->y < b->y ? -1 : 1;
    if (a->size != b->size) return a->size > b->size ? -1 : 1;
    if (a->angle != b->angle) return a->angle < b->angle ? -1 : 1;
    if (a->response != b->response) return a->response > b->response ? -1 : 1;
    if (a->octave != b->octave) return a->octave > b->octave ? -1 : 1;
    return 0;
}

/* Gaussian pyramid image (octave, layer) of the oracle's own pyramid, for stage-wise checks; returns rows*cols or 0 */
int orc_sift_gauss_image(const uint8_t* img, int rows, int cols, size_t step, int octave, int layer, float* out, int* orows, int* ocols)
{
    sift_pyr P;
    pyr_build(img, rows, cols, step, &P);
    int n = 0;
    if (octave >= 0 && octave < P.n_octaves && layer >= 0 && layer < SIFT_LAYERS + 3) {
        const fimg* g = &P.gauss[octave * (SIFT_LAYERS + 3) + layer];
        *orows = g->rows; *ocols = g->cols;
        n = g->rows * g->cols;
        if (out && n > 0) memcpy(out, g->p, (size_t)n * sizeof(float));
    }
    pyr_free(&P);
    return n;
}

/* kps: max_kp x {x, y, size, angle, response, octave (int32 bits)}; desc: max_kp x 128.  Returns the number of keypoints
 * found (may exceed max_kp: then only the first max_kp, in cv2's order, are written). */
int orc_sift_detect_and_compute(const uint8_t* img, int rows, int cols, size_t step, int max_kp, float* kps, float* desc)
{
    sift_pyr P;
    pyr_build(img, rows, cols, step, &P);
    const int L = SIFT_LAYERS, n = SIFT_ORI_HIST_BINS;
    const int threshold = cv_floor_d(0.5 * SIFT_CONTRAST_THR / L * 255);
    int cap = 4096, cnt = 0;
    sift_kp* K = (sift_kp*)malloc((size_t)cap * sizeof(sift_kp));
    for (int o = 0; o < P.n_octaves; ++o)
        for (int i = 1; i <= L; ++i) {
            const int idx = o * (L + 2) + i;
            const fimg *img1 = &P.dog[idx], *prev = &P.dog[idx - 1], *next = &P.dog[idx + 1];
            const int rws = img1->rows, cls = img1->cols;
            for (int r = SIFT_IMG_BORDER; r < rws - SIFT_IMG_BORDER; ++r)
                for (int c = SIFT_IMG_BORDER; c < cls - SIFT_IMG_BORDER; ++c) {
                    const float val = AT(img1, r, c);
                    if (!(fabsf(val) > threshold)) continue;
                    int ext = 1;
                    if (val > 0) {
                        for (int dr = -1; dr <= 1 && ext; ++dr)
                            for (int dc = -1; dc <= 1; ++dc)
                                if (!(val >= AT(img1, r + dr, c + dc) && val >= AT(prev, r + dr, c + dc) && val >= AT(next, r + dr, c + dc))) { ext = 0; break; }
                    } else if (val < 0) {
                        for (int dr = -1; dr <= 1 && ext; ++dr)
                            for (int dc = -1; dc <= 1; ++dc)
                                if (!(val <= AT(img1, r + dr, c + dc) && val <= AT(prev, r + dr, c + dc) && val <= AT(next, r + dr, c + dc))) { ext = 0; break; }
                    } else ext = 0;
                    if (!ext) continue;
                    sift_kp kpt;
                    int r1 = r, c1 = c, layer = i;
                    if (!adjust_local_extrema(&P, &kpt, o, &layer, &r1, &c1)) continue;
                    const float scl_octv = kpt.size * 0.5f / (1 << o);
                    float hist[SIFT_ORI_HIST_BINS];
                    const float omax = calc_orientation_hist(&P.gauss[o * (L + 3) + layer], c1, r1, cv_round_f(SIFT_ORI_RADIUS * scl_octv),
                                                             SIFT_ORI_SIG_FCTR * scl_octv, hist, n);
                    const float mag_thr = (float)(omax * SIFT_ORI_PEAK_RATIO);
                    for (int j = 0; j < n; ++j) {
                        const int l = j > 0 ? j - 1 : n - 1, r2 = j < n - 1 ? j + 1 : 0;
                        if (hist[j] > hist[l] && hist[j] > hist[r2] && hist[j] >= mag_thr) {
                            float bin = j + 0.5f * (hist[l] - hist[r2]) / (hist[l] - 2 * hist[j] + hist[r2]);
                            bin = bin < 0 ? n + bin : bin >= n ? bin - n : bin;
                            kpt.angle = 360.f - (float)((360.f / n) * bin);
                            if (fabsf(kpt.angle - 360.f) < FLT_EPSILON) kpt.angle = 0.f;
                            if (cnt == cap) { cap *= 2; K = (sift_kp*)realloc(K, (size_t)cap * sizeof(sift_kp)); }
                            K[cnt++] = kpt;
                        }
                    }
                }
        }
    /* KeyPointsFilter::removeDuplicatedSorted */
    if (cnt >= 2) {
        qsort(K, (size_t)cnt, sizeof(sift_kp), kp_less);
        int i = 0;
        for (int j = 1; j < cnt; ++j)
            if (K[i].x != K[j].x || K[i].y != K[j].y || K[i].size != K[j].size || K[i].angle != K[j].angle) K[++i] = K[j];
        cnt = i + 1;
    }
    /* firstOctave < 0: back to the input image's coordinates */
    for (int i = 0; i < cnt; ++i) {
        const float scale = 1.f / (float)(1 << -SIFT_FIRST_OCTAVE);
        K[i].octave = (K[i].octave & ~255) | ((K[i].octave + SIFT_FIRST_OCTAVE) & 255);
        K[i].x *= scale; K[i].y *= scale; K[i].size *= scale;
    }
    const int n_w = cnt < max_kp ? cnt : max_kp;
    for (int i = 0; i < n_w; ++i) {
        const sift_kp* k = &K[i];
        int octave = k->octave & 255;
        const int layer = (k->octave >> 8) & 255;
        octave = octave < 128 ? octave : (-128 | octave);
        const float scale = octave >= 0 ? 1.f / (1 << octave) : (float)(1 << -octave);
        const float size = k->size * scale;
        const fimg* g = &P.gauss[(octave - SIFT_FIRST_OCTAVE) * (L + 3) + layer];
        float angle = 360.f - k->angle;
        if (fabsf(angle - 360.f) < FLT_EPSILON) angle = 0.f;
        if (desc) calc_sift_descriptor(g, k->x * scale, k->y * scale, angle, size * 0.5f, desc + (size_t)i * 128);
        float* o = kps + (size_t)i * 6;
        o[0] = k->x; o[1] = k->y; o[2] = k->size; o[3] = k->angle; o[4] = k->response;
        memcpy(&o[5], &k->octave, 4);
    }
    free(K);
    pyr_free(&P);
    return cnt;
}
