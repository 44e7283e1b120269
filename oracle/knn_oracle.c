/*
 * oracle/knn_oracle.c -- TEST INFRASTRUCTURE ONLY (see vo_oracle.h).
 *
 * Restates cv2.BFMatcher(NORM_L2).knnMatch(q, t, k=2) (OpenCV modules/features2d/src/matchers.cpp
 * -> core batch_distance.cpp; third party, not vendored) and the reference's ratio loop
 * (VisualOdometryPipeLine.py:218-224, :229).  Spec: SURVEY.md A.5: scan train rows in ascending
 * index, keep (best, second) with strict '<' (on ties the lower train index ranks first),
 * distance = sqrtf((float)d^2), accept iff (double)d1 < ratio * (double)d2.
 * Pinned against cv2 in tests/test_oracle_knn.py and tests/golden/knn.npz.
 */
#include "vo_oracle.h"
#include <float.h>
#include <math.h>

int orc_knn2_ratio(const float* q, int nq, const float* t, int nt, int dim, double ratio,
                   int32_t* idx2, float* dist2, uint8_t* accept)
{
    if (dim <= 0) return -1;
#pragma omp parallel for
    for (int i = 0; i < nq; ++i) {
        const float* a = q + (size_t)i * dim;
        double b1 = DBL_MAX, b2 = DBL_MAX;
        int i1 = -1, i2 = -1;
        for (int j = 0; j < nt; ++j) {
            const float* b = t + (size_t)j * dim;
            double s = 0;
            for (int k = 0; k < dim; ++k) { double d = (double)a[k] - (double)b[k]; s += d * d; }
            if (s < b1) { b2 = b1; i2 = i1; b1 = s; i1 = j; }
            else if (s < b2) { b2 = s; i2 = j; }
        }
        const float d1 = i1 >= 0 ? sqrtf((float)b1) : FLT_MAX, d2 = i2 >= 0 ? sqrtf((float)b2) : FLT_MAX;
        idx2[2 * i] = i1; idx2[2 * i + 1] = i2;
        dist2[2 * i] = d1; dist2[2 * i + 1] = d2;
        accept[i] = (i1 >= 0 && i2 >= 0 && (double)d1 < ratio * (double)d2) ? 1 : 0;
    }
    return 0;
}
