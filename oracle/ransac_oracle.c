/*
 * oracle/ransac_oracle.c -- TEST INFRASTRUCTURE ONLY (see vo_oracle.h).
 *
 * Restates OpenCV's RANSACPointSetRegistrator (modules/calib3d/src/ptsetreg.cpp; third
 * party, not vendored) as used by cv2.solvePnPRansac (reference
 * VisualOdometryPipeLine.py:343) and cv2.findEssentialMat (:308): the cv::RNG((uint64)-1)
 * sample stream, getSubset's duplicate rejection, RANSACUpdateNumIters, and the sequential
 * "first strictly better model wins, shrink niters" loop.  Spec: SURVEY.md A.6.
 * Known answers (checked in tests/test_oracle_ransac.py): first 4-subset for
 * count=2000 is [1605,1004,940,1173]; count=1000 -> [605,4,940,173].
 */
#include "vo_oracle.h"
#include <float.h>
#include <math.h>

typedef struct { uint64_t state; } cv_rng;
static inline uint32_t rng_next(cv_rng* r)
{
    r->state = (uint64_t)(uint32_t)r->state * 4164903690u + (r->state >> 32);
    return (uint32_t)r->state;
}

void orc_ransac_subsets(int count, int model_points, int n_iters, int32_t* subsets)
{
    cv_rng r = { 0xFFFFFFFFFFFFFFFFull };
    for (int it = 0; it < n_iters; ++it) {
        int32_t* idx = subsets + (size_t)it * model_points;
        for (int i = 0; i < model_points; ++i) {
            int v;
            for (;;) {
                v = (int)(rng_next(&r) % (uint32_t)count);
                int j = 0;
                for (; j < i; ++j) if (idx[j] == v) break;
                if (j == i) break;
            }
            idx[i] = v;
        }
    }
}

int orc_ransac_update_num_iters(double p, double ep, int model_points, int max_iters)
{
    p = p < 0 ? 0 : p > 1 ? 1 : p;
    ep = ep < 0 ? 0 : ep > 1 ? 1 : ep;
    double num = 1 - p > DBL_MIN ? 1 - p : DBL_MIN;
    double denom = 1 - pow(1 - ep, model_points);
    if (denom < DBL_MIN) return 0;
    num = log(num);
    denom = log(denom);
    return denom >= 0 || -num >= max_iters * (-denom) ? max_iters : (int)lrint(num / denom);
}

int orc_ransac_select(const int32_t* counts, const int32_t* nmodels, int max_models, int max_iters,
                      int n_points, int model_points, double conf, int* iters_run)
{
    int niters = max_iters > 1 ? max_iters : 1;
    int max_good = 0, best = -1, iter = 0;
    for (iter = 0; iter < niters; ++iter) {
        int nm = nmodels ? nmodels[iter] : 1;
        for (int m = 0; m < nm; ++m) {
            int good = counts[(size_t)iter * max_models + m];
            int floor_ = max_good > model_points - 1 ? max_good : model_points - 1;
            if (good > floor_) {
                best = iter * max_models + m;
                max_good = good;
                niters = orc_ransac_update_num_iters(conf, (double)(n_points - good) / n_points, model_points, niters);
            }
        }
    }
    if (iters_run) *iters_run = iter;
    return best;
}
